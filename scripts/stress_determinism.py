"""repeat factor + solve on fresh handles and compare the solutions BITWISE across repetitions: a race in any
kernel of the path shows up as a differing x (or as an exception, printed with its message)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linearsolve_jl_b200 as ls
C = ls._capi
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for code, name, n, shift in ((C.MIXED, "mixed", 16384, 5.0), (C.MIXED, "mixed", 12000, 0.0), (C.F32, "f32", 16384, 0.0),
                             (C.F32, "f32", 6000, 0.0), (C.F64, "f64", 16384, 0.0), (C.F64, "f64", 8192, 0.0)):
    edt = torch.float32 if code == C.F32 else torch.float64
    A = torch.empty((n, n), dtype=edt, device=dev)
    b = torch.empty((1, n), dtype=edt, device=dev)
    x = torch.empty_like(b)
    x0 = None
    bad = exc = 0
    t0 = time.time()
    for it in range(reps):
        h = ls.Handle(code)
        h.fill_uniform_device(A.data_ptr(), n, n, n, seed=16384, diag_shift=shift)
        h.fill_uniform_device(b.data_ptr(), n, n, 1, seed=7)
        try:
            info = h.factor_device(A.data_ptr(), n, n)
            h.solve_device(b.data_ptr(), n, x.data_ptr(), n, 1)
            torch.cuda.synchronize()
            if x0 is None:
                x0 = x.clone()
            elif not torch.equal(x0, x):
                bad += 1
                d = (x0 - x).abs().max().item()
                print(f"  {name} n={n} rep {it}: x differs from rep 0, max |dx| = {d:.3e} (|x| max {x0.abs().max().item():.3e})", flush=True)
        except Exception as e:
            exc += 1
            print(f"  {name} n={n} rep {it}: EXCEPTION {type(e).__name__}: {e}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("  context dead:", e2, flush=True)
                sys.exit(0)
        h.close()
    print(f"{name} n={n}: {reps} repetitions, {bad} differing solutions, {exc} exceptions, {time.time() - t0:.1f} s", flush=True)
