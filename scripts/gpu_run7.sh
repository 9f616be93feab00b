#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 3000 --csv --log-file gpurun_out/launches_8192_v4.csv python scripts/prof_driver.py 8192 lu > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
