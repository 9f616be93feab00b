#!/bin/bash
# round-end validation: smoke, full GPU suite, the bench lines, one ncu capture of the cluster getrs kernel
bash scripts/gpu_final2.sh
mkdir -p /tmp/prof
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"trsv3_kernel" -s 2 -c 1 -o /tmp/prof/trsv3 -f python scripts/prof_driver.py 8192 lu > gpurun_out/prof_trsv3.log 2>&1
ncu -i /tmp/prof/trsv3.ncu-rep --page raw --csv > gpurun_out/prof_trsv3_raw.csv 2>/dev/null
ncu -i /tmp/prof/trsv3.ncu-rep --page details > gpurun_out/prof_trsv3_details.txt 2>/dev/null
ls -la gpurun_out/prof_trsv3_raw.csv
