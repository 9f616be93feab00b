#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, round(d['roofline']['achieved'],2), round(d['roofline']['peak'],2), round(d['roofline']['gemm_share_of_getrf'],3), d.get('backward_error'))" || tail -5 $1; }
for n in 8192 16384; do for mode in 0 1; do echo "== n=$n panel mode $mode"; timeout 300 python bench.py --n $n --panel-mode $mode --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_${n}_pm$mode.log 2>&1; show gpurun_out/bench_${n}_pm$mode.log; done; done
echo "== n=32768"; timeout 300 python bench.py --n 32768 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_32768.log 2>&1; show gpurun_out/bench_32768.log
