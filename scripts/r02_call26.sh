#!/bin/bash
# Round 2, GPU call 26 (FOUR B200s): panel width at N = 4 — nb = 256 against the default 128 of call 24 (264.8 ms)
mkdir -p gpurun_out
for NB in 256 192; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 4 --steps 4 --warmup 3 --nb $NB --no-e2e --no-extras > gpurun_out/r02c26_N4_nb$NB.json 2> gpurun_out/r02c26_N4_nb$NB.err
python - $NB <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02c26_N4_nb{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("nb", sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check")}, d["roofline"].get("chain_ms", {}).get("panel_factorizations"), d["roofline"].get("frac"))
PY
done
