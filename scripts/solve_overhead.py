"""Where the host-side time of ONE right-hand-side solve goes (config 2 applies 100 of them):
wall clock of b200lu_solve through ctypes vs the CUDA-event time of its kernels, with the mapped
staging on and off, at n = 4096 / 8192 / 16384.  No torch.  Usage: python scripts/solve_overhead.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linearsolve_jl_b200 as ls  # noqa: E402

C = ls._capi
sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
rng = np.random.default_rng(1)
for n in sizes:
    A = np.asfortranarray(rng.random((n, n)))
    h = ls.Handle(C.F64)
    h.factor(A, want_ipiv=False)
    bs = [rng.random(n) for _ in range(100)]
    out = np.empty(n)
    for mapped in (1, 0):
        h.set_option(C.OPT_MAPPED_RHS, mapped)
        h.solve(bs[0], out=out)
        dev = []
        t0 = time.perf_counter()
        for b in bs:
            h.solve(b, out=out)
            dev.append(h.timing(C.T_SOLVE))
        wall = (time.perf_counter() - t0) / len(bs) * 1e3
        print(f"n={n} mapped_rhs={mapped}: wall {wall:.3f} ms per solve, kernels {np.median(dev):.3f} ms "
              f"(min {min(dev):.3f}), host overhead {wall - np.median(dev):.3f} ms", flush=True)
    h.close()
