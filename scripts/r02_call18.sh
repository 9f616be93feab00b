#!/bin/bash
# diagnose the mixed n=16384 failure, full GPU suite without -x, the wavefront getrs (next library) with one rank,
# then the ncu captures with the text export done on the box (the .ncu-rep files together exceed the 64 MiB pull limit)
mkdir -p gpurun_out
timeout 400 python scripts/diag_mixed.py 2>&1 | tail -12
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_final_tests.log 2>&1; tail -12 gpurun_out/r02_final_tests.log | cut -c1-300
NEXT=$PWD/linearsolve.jl_b200/csrc/libb200lu_next.so
echo "--- wavefront getrs, one rank"
B200LU_LIB=$NEXT timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -3
timeout 200 python scripts/dist_one.py 32768 128 solve 2>&1 | tail -2
B200LU_LIB=$NEXT timeout 200 python scripts/dist_one.py 32768 128 solve 2>&1 | tail -2
N="ncu --set full --clock-control none --import-source on -f"
cap() {   # name, regex, skip, driver args...
    local name=$1 rx=$2 skip=$3; shift 3
    timeout 400 $N -k regex:$rx --launch-skip $skip --launch-count 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
    if [ -f gpurun_out/$name.ncu-rep ]; then
        python scripts/ncu_extract.py gpurun_out/$name.ncu-rep gpurun_out/${name}_metrics.txt
        ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
        ls -la gpurun_out/$name.ncu-rep
    else
        tail -5 gpurun_out/$name.log
    fi
}
# launch 1 of a factorization = the first full trailing update (launch 0 is the look-ahead block's)
cap r02_ncu_dgemm_n32768 dgemm_sub_kernel 1 python scripts/prof_driver.py 32768 lu
cap r02_ncu_dgemm_n8192 dgemm_sub_kernel 1 python scripts/prof_driver.py 8192 lu
cap r02_ncu_panel_fused_8x8 panel_cluster_kernel 1 python scripts/dist_one.py 32768
cap r02_ncu_panel_fused_32x1 panel_cluster_kernel 4 python scripts/dist_one.py 4096
cap r02_ncu_batched_warp getrf_batched_warp_kernel 1 python scripts/prof_driver.py 0 batched
cap r02_ncu_dist_step dist_step_kernel 300 python scripts/dist_one.py 32768 256 solve
ncu -i gpurun_out/r02_ncu_batched_warp.ncu-rep --page source --csv > gpurun_out/r02_ncu_batched_warp_source.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_dgemm_n8192.ncu-rep gpurun_out/r02_ncu_panel_fused_32x1.ncu-rep gpurun_out/r02_ncu_dist_step.ncu-rep gpurun_out/r02_ncu_dgemm_n32768.ncu-rep
du -sh gpurun_out
