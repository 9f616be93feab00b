#!/bin/bash
# Round 2, GPU call 6 (one B200): the left-looking fused cluster panel kernel (default) — whole GPU suite,
# n = 8192 / 32768 timings against the round-1 kernels (PANEL_MODE 2), chain profile of the dist engine.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r02c6_tests.log
for mode in 0 2; do
  timeout 300 python bench.py --workload lu --size 8192 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-comparator --panel-mode $mode > gpurun_out/r02c6_lu8192_mode$mode.json 2> gpurun_out/r02c6_lu8192_mode$mode.err
  tail -2 gpurun_out/r02c6_lu8192_mode$mode.err
done
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c6_dist1_32768.json 2> gpurun_out/r02c6_dist1_32768.err
tail -2 gpurun_out/r02c6_dist1_32768.err
timeout 300 python bench.py --workload dist --size 16384 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c6_dist1_16384.json 2> gpurun_out/r02c6_dist1_16384.err
python - <<'PY'
import json
for f in ("lu8192_mode0", "lu8192_mode2", "dist1_32768", "dist1_16384"):
    try:
        d = json.loads(open(f"gpurun_out/r02c6_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")}, d.get("roofline", {}).get("chain_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
