#!/bin/bash
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, round(d['roofline']['achieved'],2))" || tail -3 $1; }
for cfg in 0 1; do for n in 16384 32768; do echo "== n=$n gemm cfg $cfg"; timeout 300 python bench.py --n $n --gemm-cfg $cfg --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_${n}_cfg$cfg.log 2>&1; show gpurun_out/bench_${n}_cfg$cfg.log; done; done
