#!/bin/bash
B200LU_TRSV_DBG=1 timeout 120 python scripts/trsv_bench.py 8192 2>&1 | tail -12
timeout 120 python scripts/trsv_bench.py 8192 2>&1 | tail -2
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv
