"""small driver for ncu captures: one FP64 factor + solves at n, one batched call"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import linearsolve_jl_b200 as ls
C = ls._capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
what = sys.argv[2] if len(sys.argv) > 2 else "lu"
dev = torch.device("cuda", 0)
h = ls.Handle(C.MIXED if what == "mixed" else C.F64)
if what in ("lu", "mixed"):
    A = torch.empty((n, n), dtype=torch.float64, device=dev)
    B = torch.empty((16, n), dtype=torch.float64, device=dev)
    X = torch.empty_like(B)
    h.fill_uniform_device(A.data_ptr(), n, n, n, seed=1, diag_shift=5.0 if what == "mixed" else 0.0)
    h.fill_uniform_device(B.data_ptr(), n, n, 16, seed=2)
    for _ in range(2):
        h.factor_device(A.data_ptr(), n, n)
        h.solve_device(B.data_ptr(), n, X.data_ptr(), n, 1)
        h.solve_device(B.data_ptr(), n, X.data_ptr(), n, 16)
else:
    per = 16384
    A = torch.empty((per, 64, 64), dtype=torch.float64, device=dev)
    b = torch.empty((per, 64), dtype=torch.float64, device=dev)
    x = torch.empty_like(b)
    h.fill_uniform_device(A.data_ptr(), 64, 64, per * 64, seed=5)
    h.fill_uniform_device(b.data_ptr(), 64, 64, per, seed=6)
    for _ in range(2):
        h.factor_batched_device(A.data_ptr(), per, 64)
        h.solve_batched_device(b.data_ptr(), x.data_ptr(), 1)
torch.cuda.synchronize()
print("done")
