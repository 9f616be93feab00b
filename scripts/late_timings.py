"""Late round-1 timings (no torch): the shared-memory batched class and the device residual check."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linearsolve_jl_b200 as ls  # noqa: E402

C = ls._capi
rng = np.random.default_rng(0)
for n, batch in ((96, 4096), (128, 4096), (160, 2048)):
    A = rng.random((batch, n, n)) + n * np.eye(n)
    b = rng.random((batch, n))
    h = ls.Handle(C.F64)
    for _ in range(2):
        h.factor_batched(A)
        tf = h.timing(C.T_FACTOR)
        h.solve_batched(b)
        ts = h.timing(C.T_SOLVE)
    print(f"batched smem class n={n} batch={batch}: getrf {tf:.3f} ms = {batch / tf / 1e3:.3f} M systems/s "
          f"({2 / 3 * n ** 3 * batch / tf / 1e9:.2f} TF/s), getrs {ts:.3f} ms", flush=True)
    h.close()
n = 8192
A = np.asfortranarray(rng.random((n, n)))
b = rng.random(n)
h = ls.Handle(C.F64)
h.set_option(C.OPT_KEEP_A, 1)
h.factor(A)
t_keep = h.timing(C.T_FACTOR)
x = h.solve(b)
for _ in range(2):
    t0 = time.perf_counter()
    r, bn = h.residual_norms(b, x)
    t1 = time.perf_counter()
t2 = time.perf_counter()
rh = np.linalg.norm(A @ x - b)
t3 = time.perf_counter()
print(f"residual check n={n}: device {1e3 * (t1 - t0):.3f} ms wall (kernels {h.timing(C.T_SOLVE):.3f} ms), "
      f"host numpy {1e3 * (t3 - t2):.1f} ms; resid {r[0]:.3e} vs {rh:.3e}; getrf with the kept copy {t_keep:.2f} ms", flush=True)
