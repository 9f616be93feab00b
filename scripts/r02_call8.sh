#!/bin/bash
# Round 2, GPU call 8 (one B200): clock64 stamps of the fused cluster panel kernel (instrumented build, -DPCL_TIMING)
mkdir -p gpurun_out
B200LU_LIB=$PWD/linearsolve.jl_b200/csrc/libb200lu_timing.so B200LU_PANEL_DBG=1 timeout 300 python scripts/dist_one.py 32768 2> gpurun_out/r02c8_panel_stamps_32768.txt | tail -1
B200LU_LIB=$PWD/linearsolve.jl_b200/csrc/libb200lu_timing.so B200LU_PANEL_DBG=1 timeout 300 python scripts/dist_one.py 8192 2> gpurun_out/r02c8_panel_stamps_8192.txt | tail -1
grep -c pdbg gpurun_out/r02c8_panel_stamps_32768.txt
