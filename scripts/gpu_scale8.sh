#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N "$@" 2>&1 | grep '^{' | tail -1; }
run --workload dist --size 32768 --steps 2 --warmup 1 | tee gpurun_out/scale_dist_32768_N$N.json | cut -c1-260
run --workload dist --size 65536 --steps 1 --warmup 1 | tee gpurun_out/scale_dist_65536_N$N.json | cut -c1-260
run --workload batched --steps 3 --warmup 2 | tee gpurun_out/scale_batched_N$N.json | cut -c1-260
run --workload lu --steps 3 --warmup 2 --no-e2e --no-cpu-baseline | tee gpurun_out/scale_lu_N$N.json | cut -c1-200
