#!/bin/bash
# Round 2, GPU call 4 (EIGHT B200s): the default bench line at N = 8 (block-cyclic n = 32768, peer-store panel
# hand-off, chain profile, sharded batch, e2e through one 8-GPU handle), nb = 128 variant, N = 4, and the multi-GPU tests.
mkdir -p gpurun_out
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $np "$@" > gpurun_out/r02c4_$name.json 2> gpurun_out/r02c4_$name.err
  tail -2 gpurun_out/r02c4_$name.err | cut -c1-300
}
run N8 8 --steps 3 --warmup 2
run N8_nb128 8 --steps 2 --warmup 1 --nb 128 --no-e2e --no-extras
run N4 4 --steps 2 --warmup 1 --no-e2e --no-extras
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02c4_tests.log
python - <<'PY'
import json
for f in ("N8", "N8_nb128", "N4"):
    try:
        d = json.loads(open(f"gpurun_out/r02c4_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")})
        print("   chain", d.get("roofline", {}).get("chain_ms"), "gemm frac", d.get("roofline", {}).get("frac"))
        print("   e2e", d.get("e2e"), d.get("e2e_error"), "batched", (d.get("batched_65536x64") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
