#!/bin/bash
# Round 2, GPU call 7 (one B200): fused panel kernel with pipelined multiplier loads — getrf tests + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_widen.py tests/test_gpu_mixed.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r02c7_tests.log
for mode in 0 2; do
  timeout 300 python bench.py --workload lu --size 8192 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-comparator --panel-mode $mode > gpurun_out/r02c7_lu8192_mode$mode.json 2> gpurun_out/r02c7_lu8192_mode$mode.err
done
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c7_dist1_32768.json 2> gpurun_out/r02c7_dist1_32768.err
timeout 300 python bench.py --workload dist --size 32768 --nb 128 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c7_dist1_32768_nb128.json 2> gpurun_out/r02c7_dist1_32768_nb128.err
timeout 300 python bench.py --workload mixed --size 16384 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02c7_mixed16384.json 2> gpurun_out/r02c7_mixed16384.err
python - <<'PY'
import json
for f in ("lu8192_mode0", "lu8192_mode2", "dist1_32768", "dist1_32768_nb128", "mixed16384"):
    try:
        d = json.loads(open(f"gpurun_out/r02c7_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")}, d.get("roofline", {}).get("chain_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
