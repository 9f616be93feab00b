#!/bin/bash
# Round 2, GPU call 25 (one B200): fixed-order residual / norm reductions of the mixed mode: determinism stress, full suite
mkdir -p gpurun_out
timeout 600 python scripts/stress_determinism.py 20 2>&1 | grep -v "rep [0-9]*:" | tail -8
timeout 600 python scripts/stress_determinism.py 20 2>&1 | grep -c "differs" 
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_final_tests.log 2>&1; tail -4 gpurun_out/r02_final_tests.log | cut -c1-600
timeout 300 python bench.py --workload mixed --n 16384 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02c25_mixed16384.json 2> gpurun_out/r02c25_mixed.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c25_mixed16384.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")})
PY
