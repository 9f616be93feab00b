#!/bin/bash
# Round 2, GPU call 15 (one B200): multipliers staged by TMA bulk copies — getrf tests, P = 1 chain, stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_multi.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02c15_tests.log
timeout 300 python bench.py --workload dist --size 32768 --nb 128 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c15_dist1_32768_nb128.json 2> gpurun_out/r02c15_dist1.err
timeout 300 python bench.py --workload lu --size 8192 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-comparator > gpurun_out/r02c15_lu8192.json 2> gpurun_out/r02c15_lu8192.err
B200LU_LIB=$PWD/linearsolve.jl_b200/csrc/libb200lu_timing.so B200LU_PANEL_DBG=1 timeout 300 python scripts/dist_one.py 32768 2> gpurun_out/r02c15_panel_stamps_32768.txt | tail -1
python - <<'PY'
import json
for f in ("dist1_32768_nb128", "lu8192"):
    try:
        d = json.loads(open(f"gpurun_out/r02c15_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")}, d.get("roofline", {}).get("chain_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
grep -A1 "pdbg\] launch" gpurun_out/r02c15_panel_stamps_32768.txt | awk 'NR%45<2' | cut -c1-250 | head -24
