#!/bin/bash
# Round 2, GPU call 11 (TWO B200s): merged receive / push kernels and the 8-sub-block tall panel class on real peers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py tests/test_gpu_getrf.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r02c11_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c11_N2.json 2> gpurun_out/r02c11_N2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 2 --warmup 1 --nb 128 --no-e2e --no-extras > gpurun_out/r02c11_N2_nb128.json 2> gpurun_out/r02c11_N2_nb128.err
timeout 300 python bench.py --workload dist --size 32768 --steps 2 --warmup 1 --no-e2e --no-extras > gpurun_out/r02c11_dist1_32768.json 2> gpurun_out/r02c11_dist1_32768.err
python - <<'PY'
import json
for f in ("N2", "N2_nb128", "dist1_32768"):
    try:
        d = json.loads(open(f"gpurun_out/r02c11_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "dist_check", "backward_error", "transport")})
        print("   chain", d.get("roofline", {}).get("chain_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
