"""Single-right-hand-side getrs timing on device-resident factors (CUDA-event phase of the library).
python scripts/trsv_bench.py [n ...]; env B200LU_TRSV3_NEAR / B200LU_TRSV3_CS / B200LU_TRSV_DBG select
and instrument the cluster-chain kernel; mode 2 = the 2-D work-item kernel for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import linearsolve_jl_b200 as ls

C = ls._capi


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [8192]
    dev = torch.device("cuda", 0)
    for n in sizes:
        A = torch.empty((n, n), dtype=torch.float64, device=dev)
        b = torch.empty((1, n), dtype=torch.float64, device=dev)
        x = torch.empty_like(b)
        for mode in (0, 2):
            h = ls.Handle(C.F64)
            h.set_option(C.OPT_TRSV_MODE, mode)
            h.fill_uniform_device(A.data_ptr(), n, n, n, seed=7)
            h.fill_uniform_device(b.data_ptr(), n, n, 1, seed=8)
            A0 = A.clone()
            assert h.factor_device(A.data_ptr(), n, n) == 0
            ts = []
            for _ in range(30):
                h.solve_device(b.data_ptr(), n, x.data_ptr(), n)
                ts.append(h.timing(2) * 1e3)
            r = b[0] - A0.T @ x[0]      # A0 holds the matrix column-major: entry (i, j) = A0[j, i]
            berr = (r.norm() / (A0.norm() * x[0].norm())).item()
            ts.sort()
            gbs = 8.0 * n * n / (ts[len(ts) // 2] * 1e-6) / 1e9
            print(f"n={n} trsv_mode={mode}: getrs median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us "
                  f"({gbs:.0f} GB/s), backward error {berr:.2e}", flush=True)
            del h


if __name__ == "__main__":
    main()
