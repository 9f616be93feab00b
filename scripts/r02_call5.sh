#!/bin/bash
# Round 2, GPU call 5 (one B200): headline-size parity tests, the distributed getrs after the latency fix,
# and the ncu launch list of the TALL outer panels (n = 32768, first 3000 launches of one factorization).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_headline.py -q -m gpu -s 2>&1 | tail -15 | tee gpurun_out/r02c5_tests.log
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c5_dist1_32768.json 2> gpurun_out/r02c5_dist1_32768.err
tail -2 gpurun_out/r02c5_dist1_32768.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 3000 --csv \
    --log-file gpurun_out/r02c5_launches_tall_n32768.csv python scripts/dist_one.py 32768 > gpurun_out/r02c5_ncu.log 2>&1
tail -2 gpurun_out/r02c5_ncu.log
python scripts/launch_summary.py gpurun_out/r02c5_launches_tall_n32768.csv 20
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c5_dist1_32768.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "backward_error")})
print(d.get("roofline", {}).get("chain_ms"))
PY
