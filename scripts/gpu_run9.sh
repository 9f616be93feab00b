#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, round(d['roofline']['achieved'],2), round(d['roofline']['peak'],2), round(d['roofline']['gemm_share_of_getrf'],3), d.get('backward_error'))" || tail -5 $1; }
for n in 8192 16384; do echo "== n=$n"; timeout 300 python bench.py --n $n --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_${n}_pm0.log 2>&1; show gpurun_out/bench_${n}_pm0.log; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 3000 --csv --log-file gpurun_out/launches_8192_v5.csv python scripts/prof_driver.py 8192 lu > gpurun_out/ncu_v5.log 2>&1
tail -2 gpurun_out/ncu_v5.log
