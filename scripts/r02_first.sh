#!/bin/bash
# First GPU call of round 2 (one B200): the state the round starts from.
#   1. the whole -m gpu suite, smoke()
#   2. the default bench line + reference arm
#   3. where a single-RHS solve's host time goes
#   4. ncu: launch list of the default bench, full captures of the kernels added late in round 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r02_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2>&1
timeout 120 python scripts/solve_overhead.py 4096 8192 16384 2>&1 | tee gpurun_out/r02_solve_overhead.log
timeout 120 python scripts/late_timings.py 2>&1 | tee gpurun_out/r02_late_timings.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:getrf_batched_smem_kernel -c 1 \
    -o gpurun_out/r02_batched_smem python scripts/late_timings.py > gpurun_out/r02_ncu_smem.log 2>&1
