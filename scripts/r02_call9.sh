#!/bin/bash
# Round 2, GPU call 9 (one B200): fused panel with cp.async-staged multipliers — tests, timings, stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_widen.py tests/test_gpu_mixed.py tests/test_gpu_multi.py tests/test_gpu_interface.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r02c9_tests.log
timeout 300 python bench.py --workload lu --size 8192 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-comparator > gpurun_out/r02c9_lu8192.json 2> gpurun_out/r02c9_lu8192.err
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c9_dist1_32768.json 2> gpurun_out/r02c9_dist1_32768.err
timeout 300 python bench.py --workload dist --size 32768 --nb 128 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c9_dist1_32768_nb128.json 2> gpurun_out/r02c9_dist1_32768_nb128.err
B200LU_LIB=$PWD/linearsolve.jl_b200/csrc/libb200lu_timing.so B200LU_PANEL_DBG=1 timeout 300 python scripts/dist_one.py 32768 2> gpurun_out/r02c9_panel_stamps_32768.txt | tail -1
python - <<'PY'
import json
for f in ("lu8192", "dist1_32768", "dist1_32768_nb128"):
    try:
        d = json.loads(open(f"gpurun_out/r02c9_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")}, d.get("roofline", {}).get("chain_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
grep "pdbg\] launch" gpurun_out/r02c9_panel_stamps_32768.txt | awk 'NR%10==1' | head -12
