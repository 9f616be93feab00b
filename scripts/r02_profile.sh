#!/bin/bash
# Round 2 ncu evidence (one B200): full captures of the dominant kernels, the launch list of the default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_final_tests.log 2>&1; tail -3 gpurun_out/r02_final_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_bench_N1.json 2> gpurun_out/r02_final_bench_N1.err; tail -c 600 gpurun_out/r02_final_bench_N1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
N="ncu --set full --clock-control none --import-source on -f"
# launch 1 of a factorization = the first full trailing update (launch 0 is the look-ahead block's)
timeout 400 $N -k regex:dgemm_sub_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r02_ncu_dgemm_n32768 python scripts/prof_driver.py 32768 lu > gpurun_out/r02_ncu_dgemm.log 2>&1
timeout 300 $N -k regex:dgemm_sub_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r02_ncu_dgemm_n8192 python scripts/prof_driver.py 8192 lu > gpurun_out/r02_ncu_dgemm2.log 2>&1
timeout 300 $N -k regex:panel_cluster_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r02_ncu_panel_fused_8x8 python scripts/dist_one.py 32768 > gpurun_out/r02_ncu_panel1.log 2>&1
timeout 300 $N -k regex:panel_cluster_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r02_ncu_panel_fused_32x1 python scripts/dist_one.py 4096 > gpurun_out/r02_ncu_panel2.log 2>&1
timeout 300 $N -k regex:getrf_batched_warp_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r02_ncu_batched_warp python scripts/prof_driver.py 0 batched > gpurun_out/r02_ncu_batched.log 2>&1
timeout 300 $N -k regex:dist_step_kernel --launch-skip 300 --launch-count 1 -o gpurun_out/r02_ncu_dist_step python scripts/dist_one.py 32768 256 solve > gpurun_out/r02_ncu_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-extras --no-cpu-baseline --no-comparator > gpurun_out/r02_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_default_bench.csv 16
ls -la gpurun_out/*.ncu-rep
