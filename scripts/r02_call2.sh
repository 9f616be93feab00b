#!/bin/bash
# Round 2, GPU call 2 (TWO B200s): the multi-GPU engine on real peers — one process per GPU under torchrun
# (cudaIpc windows + peer stores, and the NCCL fallback), one process driving both GPUs (team handle), the
# new batched warp kernel, and the default bench line at N = 2.
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8 > gpurun_out/r02c2_topo.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py tests/test_gpu_batched.py tests/test_gpu_widen.py tests/test_gpu_interface.py -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/r02c2_tests.log
timeout 300 python bench.py --workload batched --steps 5 --warmup 3 > gpurun_out/r02c2_batched.json 2> gpurun_out/r02c2_batched.err
tail -2 gpurun_out/r02c2_batched.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02c2_bench_N2.json 2> gpurun_out/r02c2_bench_N2.err
tail -5 gpurun_out/r02c2_bench_N2.err
B200LU_DIST_MODE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c2_bench_N2_nccl.json 2> gpurun_out/r02c2_bench_N2_nccl.err
tail -3 gpurun_out/r02c2_bench_N2_nccl.err
head -c 1200 gpurun_out/r02c2_bench_N2.json; echo; head -c 600 gpurun_out/r02c2_batched.json
