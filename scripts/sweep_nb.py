"""Outer panel width sweep of the device-resident getrf (B200LU_OPT_NB) — which nb wins at which n.
No torch: host matrix -> b200lu_factor with the streamed upload off, so B200LU_T_FACTOR is the
device time of the factorization alone.  Usage: python scripts/sweep_nb.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linearsolve_jl_b200 as ls  # noqa: E402

C = ls._capi
sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
rng = np.random.default_rng(123)
for n in sizes:
    A = np.asfortranarray(rng.random((n, n)))
    h = ls.Handle(C.F64)
    h.set_option(C.OPT_STREAM_H2D, 0)
    ref = None
    for nb in (256, 224, 192, 160, 128, 96, 64):
        h.set_option(C.OPT_NB, nb)
        ts = []
        for _ in range(3):
            ipiv, info = h.factor(A)
            ts.append(h.timing(C.T_FACTOR))
        if ref is None:
            ref = ipiv.copy()
        same = bool(np.array_equal(ref, ipiv))
        t = min(ts)
        print(f"n={n} nb={nb}: getrf {t:.3f} ms = {2 / 3 * n ** 3 / t / 1e9:.2f} TF/s  (runs {', '.join(f'{x:.2f}' for x in ts)}) ipiv==nb256 {same}",
              flush=True)
    h.close()
