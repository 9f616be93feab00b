#!/bin/bash
# Round 2, GPU call 23 (one B200): B200LU_OPT_HOST_REGISTER — test, full suite, default bench with e2e_pageable_registered
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_widen.py -q -m gpu -x -k host_register 2>&1 | tail -15 | cut -c1-400
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_final_tests.log 2>&1; tail -4 gpurun_out/r02_final_tests.log | cut -c1-600
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_bench_N1.json 2> gpurun_out/r02_final_bench_N1.err
tail -3 gpurun_out/r02_final_bench_N1.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final_bench_N1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "gpu_launches")})
for k in ("e2e", "e2e_pageable", "e2e_pageable_registered"):
    print("  ", k, d.get(k))
print("  roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic", "traffic_algorithmic_of_that_launch", "traffic_source")})
PY
