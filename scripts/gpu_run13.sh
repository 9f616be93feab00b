#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, d.get('backward_error'))" || tail -5 $1; }
for mode in 0 1; do echo "== mixed n=16384 sgemm mode $mode"; timeout 300 python bench.py --workload mixed --n 16384 --nrhs 4 --sgemm-mode $mode --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_mixed_16384_sg$mode.log 2>&1; show gpurun_out/bench_mixed_16384_sg$mode.log; done
