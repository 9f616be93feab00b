"""one distributed-engine factorization with ONE rank (no peers): the workload ncu wraps to list the kernels of
the tall outer panels (n = 32768: the first 64 panels are taller than one cluster and take the L2-mailbox kernel)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linearsolve_jl_b200 as ls

C = ls._capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
h = C.Handle(C.F64)
h.set_option(C.OPT_NB, nb)
h.comm_init(None, 0, 1)
A = torch.empty((n, n), dtype=torch.float64, device="cuda:0")
h.fill_uniform_device(A.data_ptr(), n, n, n, seed=123)
print("info", h.factor_dist(A.data_ptr(), n, n), "ms", h.timing(C.T_FACTOR))
if len(sys.argv) > 3 and sys.argv[3] == "solve":
    b = torch.empty(n, dtype=torch.float64, device="cuda:0")
    x = torch.empty_like(b)
    h.fill_uniform_device(b.data_ptr(), n, n, 1, seed=7)
    for _ in range(2):
        h.solve_dist(b.data_ptr(), n, x.data_ptr(), n, 1)
    print("solve ms", h.timing(C.T_SOLVE))
