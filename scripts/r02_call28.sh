#!/bin/bash
# Round 2, GPU call 28 (EIGHT B200s): the default bench line at N = 8 with the final library (panel width chosen by the library)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02c28_N8.json 2> gpurun_out/r02c28_N8.err
tail -2 gpurun_out/r02c28_N8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c28_N8.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")})
print("   chain", d.get("roofline", {}).get("chain_ms"), "gemm frac", d.get("roofline", {}).get("frac"))
print("   e2e", d.get("e2e"), d.get("e2e_error"), "batched", (d.get("batched_65536x64") or {}).get("value"))
PY
