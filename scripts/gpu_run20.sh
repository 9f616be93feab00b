#!/bin/bash
# streamed upload: parity test, e2e breakdown, default bench
timeout 300 python -m pytest tests/test_gpu_getrf.py -x -q -m gpu -k "streamed or large_n or full_size" 2>&1 | tail -5
timeout 200 python scripts/e2e_breakdown.py 8192 16384 2>&1 | tail -12
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparator > gpurun_out/bench_default_stream.log 2>&1; tail -1 gpurun_out/bench_default_stream.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])" || tail -5 gpurun_out/bench_default_stream.log
