#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_gemm.py -m gpu -q -x --timeout 600 2>&1 | tail -3
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, round(d['roofline']['achieved'],2), round(d['roofline']['peak'],2))" || tail -5 $1; }
for stg in 0 15000 30000 45000; do for n in 16384; do echo "== n=$n stagger $stg"; timeout 300 python bench.py --n $n --gemm-stagger $stg --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_${n}_st$stg.log 2>&1; show gpurun_out/bench_${n}_st$stg.log; done; done
for stg in 0 30000; do echo "== n=32768 stagger $stg"; timeout 300 python bench.py --n 32768 --gemm-stagger $stg --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_32768_st$stg.log 2>&1; show gpurun_out/bench_32768_st$stg.log; done
echo "== n=8192"; timeout 300 python bench.py --n 8192 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_8192_x.log 2>&1; show gpurun_out/bench_8192_x.log
