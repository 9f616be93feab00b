"""small instances of every device path (compute-sanitizer wraps this: memcheck / initcheck / racecheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import linearsolve_jl_b200 as ls
C = ls._capi
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
# memcheck rejects st.async into the CTA's own shared::cluster window when the cluster has ONE CTA ("Cluster needs to
# have at least 2 blocks": every factorization ends with such panels; the hardware executes them): panel mode 1
# (the L2-mailbox kernel, no DSMEM) lets memcheck see the rest of the path
PANEL_MODE = int(os.environ.get("SAN_PANEL_MODE", "0"))
rng = np.random.default_rng(0)


def dense(code, n, nrhs=(1, 3)):
    dt = np.float32 if code == C.F32 else np.float64
    A = np.asfortranarray(rng.random((n, n)).astype(dt) + (5.0 * np.eye(n, dtype=dt) if code == C.MIXED else 0))
    h = C.Handle(code)
    h.set_option(C.OPT_PANEL_MODE, PANEL_MODE)
    ipiv, info = h.factor(A)
    for k in nrhs:
        b = rng.random((n, k)).astype(dt) if k > 1 else rng.random(n).astype(dt)
        x = h.solve(b)
        r = np.linalg.norm(A.astype(np.float64) @ x.reshape(n, -1) - b.reshape(n, -1)) / (np.linalg.norm(A) * np.linalg.norm(x))
        assert r < 1e-3 if code == C.F32 else r < 1e-12, r
    if code != C.MIXED and nrhs:
        x = h.solve(rng.random(n).astype(dt), trans="T")
    h.close()
    print("dense", code, n, "ok", flush=True)


if which in ("all", "dense"):
    dense(C.F64, 1500)      # fused cluster panels (32 x 1 rows), DMMA update, trsv2
    dense(C.F64, 4500)      # 32 x 2 class
    dense(C.F32, 2500)      # FP32 fused panels + tcgen05 update
    dense(C.MIXED, 2500)    # + refinement (residual partial sums, norms)
if which == "panel1":
    dense(C.F64, 1500, nrhs=())
if which == "panels":       # every fused panel class (racecheck: shared-memory hazards inside the cluster kernels)
    dense(C.F64, 1500, nrhs=())
    dense(C.F32, 2500, nrhs=())
    dense(C.F64, 4500, nrhs=())     # 32 x 2
    dense(C.F32, 9000, nrhs=())     # FP32 32 x 4
    dense(C.F64, 9000, nrhs=())     # 16 x 4
    dense(C.F64, 17000, nrhs=())    # 8 x 8
if which in ("all", "batched"):
    A = rng.random((300, 64, 64)) + 64 * np.eye(64)
    b = rng.random((300, 64))
    h = C.Handle(C.F64)
    x, ipiv, info = h.factor_solve_batched(A, b)            # A[s] is column-major: entry (i, j) at A[s, j, i]
    assert np.abs(np.einsum("bji,bj->bi", A, x) - b).max() < 1e-10
    h.solve_batched(b)
    h.solve_batched(b, trans="T")
    A2 = rng.random((40, 100, 100)) + 100 * np.eye(100)
    h.factor_batched(A2)
    h.solve_batched(rng.random((40, 100)))
    h.close()
    print("batched ok", flush=True)
if which in ("all", "dist"):
    n, nb = 1000, 128
    h = C.Handle(C.F64)
    h.set_option(C.OPT_NB, nb)
    h.set_option(C.OPT_PANEL_MODE, PANEL_MODE)
    h.comm_init(None, 0, 1)
    Aloc = torch.rand((n, n), dtype=torch.float64, device=dev)
    assert h.factor_dist(Aloc.data_ptr(), n, n) == 0
    bb = torch.rand((3, n), dtype=torch.float64, device=dev)
    xx = torch.empty_like(bb)
    h.solve_dist(bb.data_ptr(), n, xx.data_ptr(), n, 1)
    h.solve_dist(bb.data_ptr(), n, xx.data_ptr(), n, 3)
    torch.cuda.synchronize()
    h.close()
    print("dist ok", flush=True)
