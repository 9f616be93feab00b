#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_interface.py -x -q -m gpu 2>&1 | tail -4
timeout 200 python scripts/e2e_breakdown.py 8192 2>&1 | grep -v "^\[stream\]" | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparator > gpurun_out/bench_default_mapped.log 2>&1; tail -1 gpurun_out/bench_default_mapped.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])" || tail -5 gpurun_out/bench_default_mapped.log
