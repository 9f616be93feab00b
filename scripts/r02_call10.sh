#!/bin/bash
# Round 2, GPU call 10 (EIGHT B200s): default bench line at N = 8 after the fused panel kernel / distributed getrs
# fixes (nb = 128 by default from four GPUs on), nb = 256 and nb = 64 for comparison.
mkdir -p gpurun_out
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $np "$@" > gpurun_out/r02c10_$name.json 2> gpurun_out/r02c10_$name.err
  tail -2 gpurun_out/r02c10_$name.err | cut -c1-300
}
run N8 8 --steps 3 --warmup 2
run N8_nb256 8 --steps 2 --warmup 1 --nb 256 --no-e2e --no-extras
run N8_nb64 8 --steps 2 --warmup 1 --nb 64 --no-e2e --no-extras
python - <<'PY'
import json
for f in ("N8", "N8_nb256", "N8_nb64"):
    try:
        d = json.loads(open(f"gpurun_out/r02c10_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")})
        print("   chain", d.get("roofline", {}).get("chain_ms"), "gemm frac", d.get("roofline", {}).get("frac"))
        print("   e2e", d.get("e2e"), d.get("e2e_error"), "batched", (d.get("batched_65536x64") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
