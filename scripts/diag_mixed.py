"""which path of the mixed-precision n = 16384 run fails: try MIXED / F32 handles with every panel mode"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linearsolve_jl_b200 as ls
C = ls._capi
dev = torch.device("cuda", 0)
for n in (16384, 8192):
    A = torch.empty((n, n), dtype=torch.float64, device=dev)
    b = torch.empty((1, n), dtype=torch.float64, device=dev)
    x = torch.empty_like(b)
    for mode in (0, 2, 1):
        h = ls.Handle(C.MIXED)
        h.set_option(C.OPT_PANEL_MODE, mode)
        h.fill_uniform_device(A.data_ptr(), n, n, n, seed=16384, diag_shift=5.0)
        h.fill_uniform_device(b.data_ptr(), n, n, 1, seed=7)
        try:
            info = h.factor_device(A.data_ptr(), n, n)
            tf = h.timing(C.T_FACTOR)
            h.solve_device(b.data_ptr(), n, x.data_ptr(), n, 1)
            torch.cuda.synchronize()
            r = torch.mv(A.t(), x[0]) - b[0]
            berr = (r.norm() / (A.norm() * x[0].norm())).item()
            print(f"n={n} panel_mode={mode}: info {info} getrf {tf:.2f} ms sweeps {int(h.counter(C.C_REFINE_ITERS))} berr {berr:.3e}", flush=True)
        except Exception as e:
            print(f"n={n} panel_mode={mode}: EXCEPTION {type(e).__name__}: {e}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("  context dead:", e2, flush=True)
                sys.exit(0)
        h.close()
