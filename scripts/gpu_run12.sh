#!/bin/bash
mkdir -p gpurun_out /tmp/prof
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:"dgemm_sub" -s 15 -c 1 -o /tmp/prof/gemm -f python scripts/prof_driver.py 8192 lu > gpurun_out/prof_gemm.log 2>&1
ncu -i /tmp/prof/gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu -i /tmp/prof/gemm.ncu-rep --page source --csv > gpurun_out/prof_gemm_source.csv 2>/dev/null
ncu -i /tmp/prof/gemm.ncu-rep --page details > gpurun_out/prof_gemm_details.txt 2>/dev/null
ls -la gpurun_out/prof_gemm_source.csv
