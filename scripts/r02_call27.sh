#!/bin/bash
# Round 2, GPU call 27 (TWO B200s): final library — multi-process and one-process multi-GPU tests, N = 2 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py tests/test_gpu_widen.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r02c27_N2.json 2> gpurun_out/r02c27_N2.err
tail -2 gpurun_out/r02c27_N2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c27_N2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")}, d["config"]["nb"])
print("   e2e", d.get("e2e"), d.get("e2e_error"))
print("   roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic")})
PY
