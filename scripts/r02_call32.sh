#!/bin/bash
# Round 2, GPU call 32 (one B200): panel kernel with parity-double-buffered candidate staging: full suite, stress, n = 8192 timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/r02_final_tests.log 2>&1; tail -3 gpurun_out/r02_final_tests.log | cut -c1-400
timeout 200 python scripts/stress_determinism.py 10 2>&1 | grep -v "rep [0-9]*:" | tail -6
timeout 120 python bench.py --workload lu --n 8192 --nrhs 100 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-comparator --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('ms_per_step','getrf_ms','getrs_ms')})"
