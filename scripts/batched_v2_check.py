"""Round-2 first check of the experimental batched getrf (B200LU_OPT_BATCHED_MODE = 1): bitwise
comparison with the default kernel on 4096 systems, then device times of both on BASELINE config 4
(65536 systems of 64x64).  No torch.  Usage: python scripts/batched_v2_check.py [batch]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linearsolve_jl_b200 as ls  # noqa: E402

C = ls._capi
rng = np.random.default_rng(0)
n, batch = 64, 4096
A = rng.random((batch, n, n)) + n * np.eye(n)
A[17] = 0.0
hs = []
for mode in (0, 1):
    h = ls.Handle(C.F64)
    h.set_option(C.OPT_BATCHED_MODE, mode)
    ipiv, info = h.factor_batched(A)
    LU, _, _ = h.get_factors_batched()
    hs.append((h, ipiv, info, LU))
same = all(np.array_equal(hs[0][i], hs[1][i]) for i in (1, 2, 3))
print(f"mode 1 == mode 0 (ipiv, info, factors) on {batch} systems: {same}", flush=True)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
A = rng.random((batch, n, n)) + n * np.eye(n)
for mode, (h, *_r) in enumerate(hs):
    ts = []
    for _ in range(4):
        h.factor_batched(A)
        ts.append(h.timing(C.T_FACTOR))
    t = min(ts)
    print(f"mode {mode}: getrf of {batch} x {n}x{n}: {t:.3f} ms = {batch / t / 1e3:.2f} M systems/s "
          f"({66816 * batch / t / 1e6:.0f} GB/s algorithmic)", flush=True)
