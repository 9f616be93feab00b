#!/bin/bash
# Round 2, GPU call 20 (TWO B200s): the wavefront getrs (one tag per row block) with real peers:
# multi-process and one-process tests, then the default bench line at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r02c20_N2.json 2> gpurun_out/r02c20_N2.err
tail -2 gpurun_out/r02c20_N2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c20_N2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")})
print("   e2e", d.get("e2e"), d.get("e2e_error"))
PY
