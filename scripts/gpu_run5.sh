#!/bin/bash
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','getrf_ms','getrs_ms','getrf_gflops')}, round(d['roofline']['achieved'],2), round(d['roofline']['peak'],2), round(d['roofline']['gemm_share_of_getrf'],3), d.get('e2e',{}).get('value'))"; }
for cfg in 0 1 2; do echo "== n=8192 gemm cfg $cfg"; timeout 300 python bench.py --n 8192 --gemm-cfg $cfg --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_8192_g$cfg.log 2>&1; show gpurun_out/bench_8192_g$cfg.log; done
for cfg in 0 1 2; do echo "== n=16384 gemm cfg $cfg"; timeout 300 python bench.py --n 16384 --gemm-cfg $cfg --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_16384_g$cfg.log 2>&1; show gpurun_out/bench_16384_g$cfg.log; done
echo "== n=8192 rpt2"; timeout 300 python bench.py --n 8192 --rpt 2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_8192_rpt2.log 2>&1; show gpurun_out/bench_8192_rpt2.log
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 900 --csv --log-file gpurun_out/launches_8192.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name regex skip
  timeout 600 $NCU -k regex:"$2" -s $3 -c 1 -o /tmp/prof/$1 -f python scripts/prof_driver.py ${4:-8192} ${5:-lu} > gpurun_out/prof_$1.log 2>&1
  if [ -f /tmp/prof/$1.ncu-rep ]; then
    ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
    ncu -i /tmp/prof/$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
    ncu -i /tmp/prof/$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
    ls -la gpurun_out/prof_$1_source.csv
  else tail -3 gpurun_out/prof_$1.log; fi
}
cap panel "panel_base" 20
du -sh gpurun_out
