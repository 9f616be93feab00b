#!/bin/bash
mkdir -p gpurun_out /tmp/prof
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:"getrf_batched" -s 1 -c 1 -o /tmp/prof/bat -f python scripts/prof_driver.py 8192 batched > gpurun_out/prof_bat.log 2>&1
ncu -i /tmp/prof/bat.ncu-rep --page raw --csv > gpurun_out/prof_bat_raw.csv 2>/dev/null
ncu -i /tmp/prof/bat.ncu-rep --page source --csv > gpurun_out/prof_bat_source.csv 2>/dev/null
ncu -i /tmp/prof/bat.ncu-rep --page details > gpurun_out/prof_bat_details.txt 2>/dev/null
ls -la gpurun_out/prof_bat_source.csv
