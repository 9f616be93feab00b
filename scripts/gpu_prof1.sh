#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:panel_base -s 300 -c 2 -o gpurun_out/prof_panel -f python scripts/prof_driver.py 8192 lu > gpurun_out/prof_panel.log 2>&1; tail -2 gpurun_out/prof_panel.log
timeout 600 $NCU -k regex:trsm_lunit -s 1250 -c 3 -o gpurun_out/prof_trsm -f python scripts/prof_driver.py 8192 lu > gpurun_out/prof_trsm.log 2>&1; tail -2 gpurun_out/prof_trsm.log
timeout 600 $NCU -k regex:dgemm_sub -s 1140 -c 40 -o gpurun_out/prof_gemm -f python scripts/prof_driver.py 8192 lu > gpurun_out/prof_gemm.log 2>&1; tail -2 gpurun_out/prof_gemm.log
timeout 600 $NCU -k regex:"trsv_block|laswp_plan|laswp_apply<double" -s 60 -c 8 -o gpurun_out/prof_trsv -f python scripts/prof_driver.py 8192 lu > gpurun_out/prof_trsv.log 2>&1; tail -2 gpurun_out/prof_trsv.log
timeout 600 $NCU -k regex:batched -s 2 -c 2 -o gpurun_out/prof_batched -f python scripts/prof_driver.py 64 batched > gpurun_out/prof_batched.log 2>&1; tail -2 gpurun_out/prof_batched.log
ls -la gpurun_out/*.ncu-rep
