"""FP32-factor modes (reference test/Core/test_mixed_precision.jl:9-31: 100x100
rand+5I, rel. error and residual < 1e-5) and the FP64 refinement superset
(north star: backward error <= 10 n eps64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mixed_reference_behaviour_no_refine(gpu_required, ls):
    rng = np.random.default_rng(123)
    n = 100
    A = rng.random((n, n)) + 5 * np.eye(n)
    b = rng.random(n)
    sol = ls.solve(ls.LinearProblem(A, b), ls.B200LU32MixedLUFactorization(refine=False))
    assert sol.retcode == ls.ReturnCode.Success
    x_ref = np.linalg.solve(A, b)
    assert np.linalg.norm(sol.u - x_ref) / np.linalg.norm(x_ref) < 1e-5
    assert np.linalg.norm(A @ sol.u - b) / np.linalg.norm(b) < 1e-5


@pytest.mark.parametrize("n,shift", [(100, 5.0), (1000, 5.0), (2048, 0.0), (3000, 5.0)])
def test_mixed_refinement_reaches_fp64(gpu_required, ls, oracle, n, shift):
    rng = np.random.default_rng(n)
    A = rng.random((n, n)) + shift * np.eye(n)
    b = rng.random(n)
    B = rng.random((n, 3))
    alg = ls.B200LU32MixedLUFactorization()
    cache = ls.init(ls.LinearProblem(A, b), alg)
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success
    eps = np.finfo(np.float64).eps
    assert oracle.backward_error(A, sol.u, b) <= 10 * n * eps
    # tighter than the bar: refinement should get to a few eps
    assert oracle.backward_error(A, sol.u, b) <= 50 * eps
    cache.b = B
    cache.u = np.zeros_like(B)
    sol = ls.solve_(cache)
    assert oracle.backward_error(A, sol.u, B) <= 10 * n * eps
    # FP32 factor pivots == sgetrf pivots (up to ties)
    ipiv = cache.cacheval.handle.get_ipiv()
    _, ipiv_ref, _ = oracle.lapack_getrf(A.astype(np.float32))
    assert oracle.compare_ipiv(A.astype(np.float32), ipiv, ipiv_ref)[1] in ("exact", "tie")


def test_mixed_singular(gpu_required, ls):
    A = np.ones((50, 50))
    sol = ls.solve(ls.LinearProblem(A, np.ones(50)), ls.B200LU32MixedLUFactorization())
    assert sol.retcode == ls.ReturnCode.Failure


def test_mixed_full_size_config3(gpu_required, ls):
    """BASELINE config 3 at full size: FP32 factor (cluster panels of 16384 rows with 4 rows per thread,
    tcgen05 / TMA / TMEM trailing update) + FP64 refinement at n = 16384, well-conditioned rand + 5 I
    (reference idiom, test/Core/test_mixed_precision.jl:16-17).  Size-independent property: backward
    error <= 10 n eps64, computed on the device."""
    import torch
    C = ls._capi
    n = 16384
    dev = torch.device("cuda", 0)
    h = ls.Handle(C.MIXED)
    A = torch.empty((n, n), dtype=torch.float64, device=dev)        # column-major: A[j, i] = entry (i, j)
    b = torch.empty((1, n), dtype=torch.float64, device=dev)
    x = torch.empty_like(b)
    h.fill_uniform_device(A.data_ptr(), n, n, n, seed=31, diag_shift=5.0)
    h.fill_uniform_device(b.data_ptr(), n, n, 1, seed=32)
    assert h.factor_device(A.data_ptr(), n, n) == 0
    h.solve_device(b.data_ptr(), n, x.data_ptr(), n, 1)
    torch.cuda.synchronize()
    r = A.T @ x[0] - b[0]
    berr = (r.norm() / (A.norm() * x[0].norm())).item()
    assert berr <= 10 * n * np.finfo(np.float64).eps, berr
    assert h.counter(C.C_REFINE_ITERS) >= 1


@pytest.mark.parametrize("n,nrhs", [(1000, 20), (3000, 100), (1002, 7)])
def test_mixed_matrix_rhs_block_refinement(gpu_required, ls, oracle, n, nrhs):
    """MIXED with a matrix right-hand side: the whole block is refined together (FP32 blocked TRSM,
    one FP64 GEMM residual per sweep); every column reaches backward error <= 10 n eps64."""
    rng = np.random.default_rng(n + nrhs)
    A = np.asfortranarray(rng.random((n, n)) + 5.0 * np.eye(n))
    B = np.asfortranarray(rng.random((n, nrhs)))
    h = ls.Handle(ls._capi.MIXED)
    _, info = h.factor(A)
    assert info == 0
    X = h.solve(B)
    eps = np.finfo(np.float64).eps
    R = A @ X - B
    nA = np.linalg.norm(A)
    worst = max(np.linalg.norm(R[:, c]) / (nA * np.linalg.norm(X[:, c])) for c in range(nrhs))
    assert worst <= 10 * n * eps, worst
    assert worst <= 100 * eps, worst       # refinement gets to a few eps, not just under the bar
