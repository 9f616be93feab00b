#!/bin/bash
# Round 2, GPU call 1 (one B200): regression of the whole -m gpu suite after the multi-GPU engine rewrite
# (the single-rank engine tests run here), the new default bench line, the reference arm, the dist engine
# with one rank at n = 8192 / 32768, and the cluster panel's clock64 stamps for tall panels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02c1_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02c1_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c1_bench_default.json 2> gpurun_out/r02c1_bench_default.err
tail -3 gpurun_out/r02c1_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c1_bench_reference.json 2>&1
timeout 300 python bench.py --workload dist --size 8192 --steps 3 --warmup 2 --no-e2e --no-extras > gpurun_out/r02c1_dist1_8192.json 2> gpurun_out/r02c1_dist1_8192.err
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c1_dist1_32768.json 2> gpurun_out/r02c1_dist1_32768.err
tail -3 gpurun_out/r02c1_dist1_8192.err gpurun_out/r02c1_dist1_32768.err
B200LU_PANEL_DBG=1 timeout 300 python scripts/prof_driver.py 32768 lu 2> gpurun_out/r02c1_panel_stamps_32768.txt | tail -1
B200LU_PANEL_DBG=1 timeout 300 python scripts/prof_driver.py 16384 lu 2> gpurun_out/r02c1_panel_stamps_16384.txt | tail -1
head -c 1500 gpurun_out/r02c1_bench_default.json
