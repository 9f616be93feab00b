#!/bin/bash
mkdir -p gpurun_out /tmp/prof
NCU="ncu --set full --clock-control none --import-source on --warp-sampling-interval 0"
cap() { # name regex skip n what
  timeout 600 $NCU -k regex:"$2" -s $3 -c 1 -o /tmp/prof/$1 -f python scripts/prof_driver.py ${4:-8192} ${5:-lu} > gpurun_out/prof_$1.log 2>&1
  if [ -f /tmp/prof/$1.ncu-rep ]; then
    ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
    ncu -i /tmp/prof/$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
    ncu -i /tmp/prof/$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
    ls -la gpurun_out/prof_$1_source.csv
  else tail -3 gpurun_out/prof_$1.log; fi
}
cap pcl "panel_cluster" 20
