#!/bin/bash
mkdir -p gpurun_out /tmp/prof
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name regex skip n what
  timeout 600 $NCU -k regex:"$2" -s $3 -c 1 -o /tmp/prof/$1 -f python scripts/prof_driver.py $4 $5 > gpurun_out/prof_$1.log 2>&1
  ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
  ls -la gpurun_out/prof_$1_raw.csv
}
cap tcbig "sgemm3x_tc" 1 16384 mixed
cap trsv2 "trsv2_kernel" 2 8192 lu
cap pclv5 "panel_cluster" 20 8192 lu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 3000 --csv --log-file gpurun_out/launches_8192_v6.csv python scripts/prof_driver.py 8192 lu > gpurun_out/ncu_v6.log 2>&1
