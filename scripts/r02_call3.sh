#!/bin/bash
# Round 2, GPU call 3 (one B200): batched warp kernel v2 (rows never move), fused distributed-getrs step kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_multi.py tests/test_gpu_widen.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02c3_tests.log
timeout 300 python bench.py --workload batched --steps 5 --warmup 3 > gpurun_out/r02c3_batched.json 2> gpurun_out/r02c3_batched.err
tail -2 gpurun_out/r02c3_batched.err
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c3_dist1_32768.json 2> gpurun_out/r02c3_dist1_32768.err
tail -2 gpurun_out/r02c3_dist1_32768.err
python - <<'PY'
import json
for f in ("gpurun_out/r02c3_batched.json", "gpurun_out/r02c3_dist1_32768.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")}, d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
