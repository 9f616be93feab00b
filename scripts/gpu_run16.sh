#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_interface.py tests/test_gpu_mixed.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['getrf_ms'], d['getrs_ms'], d.get('e2e'))"
timeout 600 python bench.py --no-cpu-baseline --nrhs 1 --no-e2e > gpurun_out/bench_nrhs1.log 2>&1; tail -1 gpurun_out/bench_nrhs1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('nrhs=1', d['getrf_ms'], d['getrs_ms'], d['roofline']['getrs'])"
timeout 300 python - <<'P'
import sys; sys.path.insert(0,'.')
import torch, time, linearsolve_jl_b200 as ls
C=ls._capi; dev=torch.device('cuda',0)
for n in (8192, 16384):
    h=ls.Handle(C.F64); A=torch.empty((n,n),dtype=torch.float64,device=dev); b=torch.empty((1,n),dtype=torch.float64,device=dev); x=torch.empty_like(b)
    h.fill_uniform_device(A.data_ptr(),n,n,n,seed=1); h.fill_uniform_device(b.data_ptr(),n,n,1,seed=2)
    h.factor_device(A.data_ptr(),n,n)
    for mode in (0,1):
        h.set_option(C.OPT_TRSV_MODE, mode)
        for _ in range(3): h.solve_device(b.data_ptr(),n,x.data_ptr(),n,1)
        ts=[]
        for _ in range(10):
            h.solve_device(b.data_ptr(),n,x.data_ptr(),n,1); ts.append(h.timing(C.T_SOLVE))
        print(f"n={n} trsv mode {mode}: solve {min(ts)*1e3:.1f} us (median {sorted(ts)[5]*1e3:.1f}) -> {8*n*n/min(ts)/1e6:.0f} GB/s")
P
