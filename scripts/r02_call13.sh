#!/bin/bash
# Round 2, GPU call 13 (one B200): prologue stamps of the fused panel kernel (instrumented build)
mkdir -p gpurun_out
B200LU_LIB=$PWD/linearsolve.jl_b200/csrc/libb200lu_timing.so B200LU_PANEL_DBG=1 timeout 300 python scripts/dist_one.py 32768 2> gpurun_out/r02c13_panel_stamps_32768.txt | tail -1
grep -A2 "pdbg\] launch" gpurun_out/r02c13_panel_stamps_32768.txt | awk 'NR%30<3' | head -40
