#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_interface.py tests/test_gpu_mixed.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['getrf_ms'], d['getrs_ms'], d.get('e2e'))"
timeout 600 python bench.py --no-cpu-baseline --nrhs 1 --no-e2e > gpurun_out/bench_nrhs1.log 2>&1; tail -1 gpurun_out/bench_nrhs1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('nrhs=1', d['getrf_ms'], d['getrs_ms'], d['roofline']['getrs'])"
