#!/bin/bash
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for n in 8192 16384; do
timeout 300 python bench.py --n $n --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_$n.log 2>&1; tail -1 gpurun_out/bench_$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['gemm_share_of_getrf'])"
done
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name regex skip
  timeout 600 $NCU -k regex:"$2" -s $3 -c 1 -o /tmp/prof/$1 -f python scripts/prof_driver.py ${4:-8192} ${5:-lu} > gpurun_out/prof_$1.log 2>&1
  if [ -f /tmp/prof/$1.ncu-rep ]; then
    ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
    ncu -i /tmp/prof/$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
    ncu -i /tmp/prof/$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
    ls -la /tmp/prof/$1.ncu-rep gpurun_out/prof_$1_source.csv
  else tail -3 gpurun_out/prof_$1.log; fi
}
cap panel "panel_base" 20
cap trsm8 "trsm_lunit_kernel<double, 8" 3
cap gemm "dgemm_sub" 31
cap trsv1 "trsv_block_kernel<double, 1, 0" 0
cap trsv8 "trsv_block_kernel<double, 8, 0" 0
cap plan "laswp_plan" 1
cap batched "getrf_batched" 1 64 batched
du -sh gpurun_out
