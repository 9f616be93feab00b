#!/bin/bash
mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 6000 --csv --log-file gpurun_out/launches_mixed_16384.csv python scripts/prof_driver.py 16384 mixed > gpurun_out/ncu_mixed.log 2>&1
tail -2 gpurun_out/ncu_mixed.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:"sgemm3x_tc" -s 2 -c 1 -o /tmp/prof/tc -f python scripts/prof_driver.py 16384 mixed > gpurun_out/prof_tc.log 2>&1
ncu -i /tmp/prof/tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
ncu -i /tmp/prof/tc.ncu-rep --page details > gpurun_out/prof_tc_details.txt 2>/dev/null
ls -la gpurun_out/prof_tc_raw.csv
