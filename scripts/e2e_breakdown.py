"""Where the end-to-end time of config 2 goes: b200lu_factor from a pinned host matrix with the
streamed upload on/off (wall clock + the library's CUDA-event phases), and 100 sequential host solves.
Run on a GPU box: python scripts/e2e_breakdown.py [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import linearsolve_jl_b200 as ls
C = ls._capi


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [8192]
    for n in sizes:
        g = torch.Generator().manual_seed(n)
        A_pin = torch.rand((n, n), dtype=torch.float64, generator=g).pin_memory()
        A = A_pin.numpy().T                      # Fortran-ordered view
        B_pin = torch.rand((100, n), dtype=torch.float64, generator=g).pin_memory()
        B = B_pin.numpy()
        for stream in (0, 1, 0, 1):
            h = ls.Handle(C.F64)
            h.set_option(C.OPT_STREAM_H2D, stream)
            h.factor(A, want_ipiv=False)         # warm-up (allocations)
            ts = []
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                h.factor(A, want_ipiv=False)
                ts.append((time.perf_counter() - t0) * 1e3)
            print(f"n={n} stream_h2d={stream}: factor wall ms {['%.2f' % t for t in ts]}  "
                  f"phases: h2d {h.timing(0):.2f} ms, factor {h.timing(1):.2f} ms", flush=True)
            if stream == 1:
                x = np.empty(n)
                h.solve(B[0], out=x)
                t0 = time.perf_counter()
                for r in range(100):
                    h.solve(B[r], out=x)
                print(f"n={n}: 100 sequential host solves {(time.perf_counter() - t0) * 1e3:.2f} ms "
                      f"(last: h2d {h.timing(0) * 1e3:.0f} us, solve {h.timing(2) * 1e3:.0f} us, d2h {h.timing(3) * 1e3:.0f} us)", flush=True)
            del h


if __name__ == "__main__":
    main()
