#!/bin/bash
# Round 2, GPU call 22 (one B200): full suite + smoke + default bench with the final library, then the ncu capture of
# the trailing-update DGEMM as a stand-alone launch of the headline's first full update (and of the 8-rank shape)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_final_tests.log 2>&1; tail -4 gpurun_out/r02_final_tests.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_bench_N1.json 2> gpurun_out/r02_final_bench_N1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final_bench_N1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "getrf_ms", "getrs_ms", "gpu_launches")})
print("  e2e", d.get("e2e"))
print("  roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic", "getrs")})
print("  config2", d.get("config2_n8192_100rhs"), "batched", d.get("batched_65536x64"))
PY
N="ncu --set full --clock-control none --import-source on -f"
cap() {   # name, regex, skip, shape, algorithmic bytes, driver args...
    local name=$1 rx=$2 skip=$3 shape=$4 alg=$5; shift 5
    timeout 400 $N -k regex:$rx --launch-skip $skip --launch-count 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
    if [ -f gpurun_out/$name.ncu-rep ]; then
        python scripts/ncu_extract.py gpurun_out/$name.ncu-rep gpurun_out/${name}_metrics.txt shape $shape algorithmic_bytes $alg
        ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
        grep -E "gpu__time_duration|dram__bytes_(read|write).sum |launch__grid_size" gpurun_out/${name}_metrics.txt
    else
        tail -5 gpurun_out/$name.log
    fi
}
# C[32512 x 32256] -= L21[32512 x 256] U12[256 x 32256]: 2 M N 8 + (M + N) K 8 bytes
cap r02_ncu_dgemm_n32768 dgemm_sub_kernel 2 32512x32256x256 16911958016 python scripts/prof_gemm.py 32768 256
rm -f gpurun_out/r02_ncu_dgemm_n32768.ncu-rep
timeout 300 python scripts/gemm_shapes.py 2>&1 | tee gpurun_out/r02_gemm_shapes.txt | tail -8
du -sh gpurun_out
