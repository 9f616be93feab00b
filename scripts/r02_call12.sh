#!/bin/bash
# Round 2, GPU call 12 (EIGHT B200s): default bench line at N = 8 with the merged receive / push kernels, and config 5
# (n = 65536 block-cyclic over 8 GPUs)
mkdir -p gpurun_out
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $np "$@" > gpurun_out/r02c12_$name.json 2> gpurun_out/r02c12_$name.err
  tail -2 gpurun_out/r02c12_$name.err | cut -c1-300
}
run N8 8 --steps 3 --warmup 2
run N8_65536 8 --workload dist --size 65536 --steps 2 --warmup 1 --no-e2e --no-extras
python - <<'PY'
import json
for f in ("N8", "N8_65536"):
    try:
        d = json.loads(open(f"gpurun_out/r02c12_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")})
        print("   chain", d.get("roofline", {}).get("chain_ms"), "gemm frac", d.get("roofline", {}).get("frac"))
        print("   e2e", d.get("e2e"), d.get("e2e_error"), "batched", (d.get("batched_65536x64") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
