"""Parity of the CUDA getrf/getrs (through the C ABI) with the oracle.

Bars (north star): ipiv bit-exact except where ties occur; normwise backward
error ||Ax-b||/(||A|| ||x||) <= 10*n*eps(eltype); scaled LU residual < 20
(reference test/Core/blocked_lufact.jl:8-28)."""
import numpy as np
import pytest

from conftest import decisive_matrix

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 5, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256, 257,
         300, 500, 512, 513, 1000, 1025]


def _handle(ls, dtype=np.float64, **opts):
    h = ls.Handle(ls._capi.F64 if dtype == np.float64 else ls._capi.F32)
    for k, v in opts.items():
        h.set_option(getattr(ls._capi, k), v)
    return h


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_residual_and_ipiv_random(gpu_required, ls, oracle, dtype):
    """reference: "residual, square, default params" (test/Core/blocked_lufact.jl:44-54)"""
    rng = np.random.default_rng(1234)
    h = _handle(ls, dtype)
    for n in SIZES:
        A = np.asfortranarray(rng.standard_normal((n, n)).astype(dtype))
        ipiv, info = h.factor(A)
        assert info == 0
        LU = h.get_factors()
        assert oracle.scaled_residual(A, LU, ipiv) < 20, n
        _, ipiv_ref, _ = oracle.lapack_getrf(A)
        k, status = oracle.compare_ipiv(A, ipiv, ipiv_ref)
        assert status in ("exact", "tie"), (n, k, status)
        if dtype == np.float64 and n <= 600:
            assert status == "exact", (n, k)


@pytest.mark.parametrize("nb,la", [(16, 0), (16, 1), (32, 1), (48, 1), (64, 0), (128, 1), (256, 1)])
def test_forced_panel_widths(gpu_required, ls, oracle, nb, la):
    """reference: "forced blocked driver, remainder paths" (test/Core/blocked_lufact.jl:56-70)"""
    rng = np.random.default_rng(7)
    h = _handle(ls, OPT_NB=nb, OPT_LOOKAHEAD=la)
    for n in (11, 41, 67, 70, 97, 130, 190, 257, 530, 777):
        A = np.asfortranarray(rng.standard_normal((n, n)))
        ipiv, info = h.factor(A)
        assert info == 0
        assert oracle.scaled_residual(A, h.get_factors(), ipiv) < 20, (n, nb)
        _, ipiv_ref, _ = oracle.lapack_getrf(A)
        assert oracle.compare_ipiv(A, ipiv, ipiv_ref)[1] in ("exact", "tie"), (n, nb)


def test_pivots_match_lapack_decisive_margins(gpu_required, ls, oracle):
    """reference test/Core/blocked_lufact.jl:148-161"""
    rng = np.random.default_rng(42)
    h = _handle(ls)
    for n in (17, 64, 129, 300, 700):
        for _ in range(3):
            A = decisive_matrix(rng, n)
            ipiv, info = h.factor(A)
            _, ipiv_ref, _ = oracle.lapack_getrf(A)
            _, ipiv_jl, _ = oracle.ref_lufact(A)
            assert info == 0
            assert np.array_equal(ipiv, ipiv_ref), n
            assert np.array_equal(ipiv, ipiv_jl), n
            assert oracle.scaled_residual(A, h.get_factors(), ipiv) < 20


def test_permutation_matrix_exact(gpu_required, ls, oracle):
    """reference test/Core/blocked_lufact.jl:163-173: factors and ipiv exactly LAPACK's"""
    rng = np.random.default_rng(3)
    h = _handle(ls)
    for n in (16, 65, 200, 600):
        A = np.asfortranarray(np.eye(n)[rng.permutation(n), :])
        ipiv, info = h.factor(A)
        lu_ref, ipiv_ref, info_ref = oracle.lapack_getrf(A)
        assert info == 0 == info_ref
        assert np.array_equal(ipiv, ipiv_ref)
        assert np.array_equal(h.get_factors(), lu_ref)


def test_wilkinson_growth(gpu_required, ls, oracle):
    """reference test/Core/blocked_lufact.jl:175-184: ipiv == 1:n, U[n,n] == 2^(n-1) exactly"""
    h = _handle(ls)
    for n in (24, 53):
        A = np.eye(n) - np.tril(np.ones((n, n)), -1)
        A[:, n - 1] = 1.0
        A = np.asfortranarray(A)
        ipiv, info = h.factor(A)
        LU = h.get_factors()
        assert info == 0
        assert np.array_equal(ipiv, np.arange(1, n + 1))
        assert LU[n - 1, n - 1] == 2.0 ** (n - 1)
        assert oracle.scaled_residual(A, LU, ipiv) < 20


def test_singularity_info(gpu_required, ls, oracle):
    """reference test/Core/blocked_lufact.jl:186-214: info == LAPACK's first zero pivot;
    zero matrix => info == 1; NaN propagates"""
    rng = np.random.default_rng(11)
    h = _handle(ls)
    # n = 1500: a 4-CTA cluster — the zero-pivot path then crosses CTAs (remote shared-memory stores
    # of the row at position k + a cluster barrier), and a zero column inside the second 32-wide block
    for n in (10, 50, 130, 300, 1500):
        for zc in (0, 3, n - 1) + ((40, 700) if n == 1500 else ()):
            A = np.asfortranarray(rng.standard_normal((n, n)))
            A[:, zc] = 0.0
            _, info = h.factor(A)
            _, _, info_ref = oracle.lapack_getrf(A)
            _, _, info_jl = oracle.ref_lufact(A)
            assert info > 0
            assert info == info_ref == info_jl, (n, zc)
            with pytest.raises(ls.B200LUError):
                h.solve(np.ones(n))
    Z = np.zeros((50, 50), order="F")
    _, info = h.factor(Z)
    assert info == 1
    An = np.asfortranarray(rng.standard_normal((30, 30)))
    An[1, 1] = np.nan
    h.factor(An)
    assert np.isnan(h.get_factors()).any()
    # zero matrix / all-NaN column on a multi-CTA cluster
    Z = np.zeros((1100, 1100), order="F")
    _, info = h.factor(Z)
    assert info == 1
    Ac = np.asfortranarray(rng.standard_normal((1100, 1100)))
    Ac[:, 5] = np.nan
    ipiv_c, _ = h.factor(Ac)
    _, ipiv_ref_c, _ = oracle.lapack_getrf(Ac)
    assert np.array_equal(np.asarray(ipiv_c)[:6], np.asarray(ipiv_ref_c)[:6])   # NaN never wins: kp = k at the NaN column
    assert np.isnan(h.get_factors()).any()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_solve_backward_error(gpu_required, ls, oracle, dtype):
    """getrs parity: backward error <= 10 n eps; agreement with LAPACK getrs and the
    reference's _naive_lu_ldiv! (test/Core/genericlu_naive_ldiv.jl:28-46)"""
    rng = np.random.default_rng(5)
    h = _handle(ls, dtype)
    eps = np.finfo(dtype).eps
    for n in (1, 2, 7, 51, 64, 65, 100, 257, 640, 1000, 2000):
        A = np.asfortranarray((rng.random((n, n)) + n * np.eye(n)).astype(dtype))
        b = rng.random(n).astype(dtype)
        B = np.asfortranarray(rng.random((n, 4)).astype(dtype))
        ipiv, info = h.factor(A)
        assert info == 0
        x = h.solve(b)
        X = h.solve(B)
        assert oracle.backward_error(A, x, b) <= 10 * n * eps
        assert oracle.backward_error(A, X, B) <= 10 * n * eps
        lu_ref, ipiv_ref, _ = oracle.lapack_getrf(A)
        x_ref = oracle.lapack_getrs(lu_ref, ipiv_ref, b)
        np.testing.assert_allclose(x, x_ref, rtol=100 * eps * n)
        if n <= 257:
            F, ip, _ = oracle.ref_lufact(A)
            np.testing.assert_allclose(x, oracle.ref_ldiv(F, ip, b), rtol=100 * eps * n)
            np.testing.assert_allclose(X, oracle.ref_ldiv(F, ip, B), rtol=100 * eps * n)


def test_uniform_random_solve_2000(gpu_required, ls, oracle):
    """BASELINE config 1: rand(2000,2000), rand(2000) Float64 vs LAPACK"""
    rng = np.random.default_rng(123)
    n = 2000
    A = np.asfortranarray(rng.random((n, n)))
    b = rng.random(n)
    h = _handle(ls)
    ipiv, info = h.factor(A)
    _, ipiv_ref, _ = oracle.lapack_getrf(A)
    assert info == 0
    assert oracle.compare_ipiv(A, ipiv, ipiv_ref)[1] in ("exact", "tie")
    x = h.solve(b)
    assert oracle.backward_error(A, x, b) <= 10 * n * np.finfo(np.float64).eps
    # 20 re-solves on one factorization (test/Core/genericlu_naive_ldiv.jl:49-67)
    for _ in range(20):
        bi = rng.random(n)
        assert oracle.backward_error(A, h.solve(bi), bi) <= 10 * n * np.finfo(np.float64).eps


def test_strided_leading_dimension(gpu_required, ls, oracle):
    """unit row stride with lda > n (test/Core/blocked_lufact.jl:98-104,134-146)"""
    rng = np.random.default_rng(9)
    h = _handle(ls)
    for n in (17, 80, 130):
        P = np.asfortranarray(rng.standard_normal((n + 9, n + 4)))
        V = P[:n, :n]
        ipiv, info = h.factor(V)
        assert info == 0
        assert oracle.scaled_residual(V.copy(), h.get_factors(), ipiv) < 20


def test_large_n_properties(gpu_required, ls, oracle):
    """n = 4096: size-independent properties (backward error, P A = L U on a row
    sample, ipiv vs LAPACK up to ties)"""
    rng = np.random.default_rng(2024)
    n = 4096
    A = np.asfortranarray(rng.random((n, n)))
    b = rng.random(n)
    h = _handle(ls)
    ipiv, info = h.factor(A)
    assert info == 0
    x = h.solve(b)
    assert oracle.backward_error(A, x, b) <= 10 * n * np.finfo(np.float64).eps
    _, ipiv_ref, _ = oracle.lapack_getrf(A)
    assert oracle.compare_ipiv(A, ipiv, ipiv_ref)[1] in ("exact", "tie")
    LU = h.get_factors()
    perm = oracle.ipiv_to_perm(ipiv)
    rows = rng.choice(n, 64, replace=False)
    L = np.tril(LU, -1) + np.eye(n)
    U = np.triu(LU)
    np.testing.assert_allclose(L[rows, :] @ U, A[perm[rows], :], atol=1e-10 * n)


@pytest.mark.parametrize("dtype,n", [(np.float64, 8192), (np.float64, 9000), (np.float32, 9000)])
def test_full_size_config2_and_tall_panels(gpu_required, ls, oracle, dtype, n):
    """BASELINE config 2 at full size (n = 8192, FP64: 16-CTA cluster panels of 8192 rows) and the
    tall-panel variants of the cluster kernel (8193..16384 rows: 16-wide FP64 blocks with 4 rows per
    thread, FP32 blocks with 4 rows per thread): ipiv equal to LAPACK's up to admissible ties,
    backward error <= 10 n eps, P A = L U on a row sample."""
    rng = np.random.default_rng(77 + n)
    A = np.asfortranarray(rng.random((n, n)).astype(dtype))
    b = rng.random(n).astype(dtype)
    h = _handle(ls, dtype)
    ipiv, info = h.factor(A)
    assert info == 0
    x = h.solve(b)
    assert oracle.backward_error(A, x, b) <= 10 * n * np.finfo(dtype).eps
    if dtype == np.float64:
        # FP64 pivots are decisive on this matrix: equal to LAPACK's outright.  (In FP32 a
        # rounding-level tie appears after a few thousand columns and the oracle's tie check
        # replays that many unblocked elimination steps in numpy — minutes; the FP32 pivot rule is
        # pinned at n <= 1025 above and here by P A = L U below.)
        _, ipiv_ref, _ = oracle.lapack_getrf(A)
        assert np.array_equal(np.asarray(ipiv), np.asarray(ipiv_ref))
    LU = h.get_factors()
    perm = oracle.ipiv_to_perm(ipiv)
    rows = rng.choice(n, 32, replace=False)
    L = np.tril(LU, -1) + np.eye(n, dtype=dtype)
    U = np.triu(LU)
    tol = (1e-10 if dtype == np.float64 else 2e-3) * n
    np.testing.assert_allclose((L[rows, :].astype(np.float64) @ U.astype(np.float64)), A[perm[rows], :].astype(np.float64), atol=tol)


def test_l2_mailbox_fallback_panels(gpu_required, ls):
    """n = 16640 > 16384: the first outer panels are taller than one cluster and take the L2-mailbox
    kernel (panel.cuh), later ones the 16-wide and 32-wide cluster kernels — all three panel paths in
    one factorization.  Size-independent property (LAPACK at this size is too slow for a test):
    backward error of the solve, computed on the device."""
    import torch
    C = ls._capi
    n = 16640
    dev = torch.device("cuda", 0)
    h = ls.Handle(C.F64)
    A = torch.empty((n, n), dtype=torch.float64, device=dev)        # column-major: A[j, i] = entry (i, j)
    b = torch.empty((1, n), dtype=torch.float64, device=dev)
    x = torch.empty_like(b)
    h.fill_uniform_device(A.data_ptr(), n, n, n, seed=4242)
    h.fill_uniform_device(b.data_ptr(), n, n, 1, seed=4243)
    A0 = A.clone()
    assert h.factor_device(A.data_ptr(), n, n) == 0
    h.solve_device(b.data_ptr(), n, x.data_ptr(), n, 1)
    torch.cuda.synchronize()
    r = A0.T @ x[0] - b[0]
    berr = (r.norm() / (A0.norm() * x[0].norm())).item()
    assert berr <= 10 * n * np.finfo(np.float64).eps, berr


@pytest.mark.parametrize("dtype,n,pad", [(np.float64, 2048, 0), (np.float64, 3000, 7), (np.float64, 4096, 0),
                                         (np.float32, 3000, 0)])
def test_streamed_upload_matches_resident(gpu_required, ls, oracle, dtype, n, pad):
    """b200lu_factor from a host matrix uploads A in column chunks and factors underneath the copy
    (B200LU_OPT_STREAM_H2D, default on for n >= 2048); late chunks are caught up left-looking.  Every
    element sees the same arithmetic in the same order as in the copy-then-factor path: FP64 factors
    and pivots are bitwise equal (FP32 may take the FFMA instead of the tcgen05 kernel for a small
    catch-up block, so it is checked through the residual instead)."""
    rng = np.random.default_rng(4100 + n)
    P = np.asfortranarray(rng.standard_normal((n + pad, n)).astype(dtype))
    A = P[:n, :]
    h0 = _handle(ls, dtype, OPT_STREAM_H2D=0)
    h1 = _handle(ls, dtype, OPT_STREAM_H2D=1)
    ipiv0, info0 = h0.factor(A)
    ipiv1, info1 = h1.factor(A)
    assert info0 == 0 and info1 == 0
    LU1 = h1.get_factors()
    if dtype == np.float64:
        assert np.array_equal(np.asarray(ipiv0), np.asarray(ipiv1))
        assert np.array_equal(h0.get_factors(), LU1)
    assert oracle.scaled_residual(A.copy(order="F"), LU1, ipiv1) < 20
    # twice on the same handle: chunk events and plans are reused
    ipiv2, _ = h1.factor(A)
    assert np.array_equal(np.asarray(ipiv1), np.asarray(ipiv2))
    assert np.array_equal(LU1, h1.get_factors())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_single_rhs_getrs_all_paths(gpu_required, ls, oracle, dtype):
    """The three single-right-hand-side getrs kernels on the same factors: the cluster chain over
    DSMEM (mode 3; the default from n = 6144), the 2-D work items, one CTA per block row.  Each meets the
    backward-error bar and agrees with LAPACK getrs; ragged last blocks (n not a multiple of 64),
    general (not diagonally dominant) matrices, repeated solves (ring / epoch reuse)."""
    rng = np.random.default_rng(606)
    eps = np.finfo(dtype).eps
    C = ls._capi
    for n in (1024, 1500, 2048, 4097, 6200):
        A = np.asfortranarray(rng.standard_normal((n, n)).astype(dtype))
        lu_ref, ipiv_ref, _ = oracle.lapack_getrf(A)
        xs = []
        for mode in (3, 2, 1, 0):
            h = _handle(ls, dtype, OPT_TRSV_MODE=mode)
            _, info = h.factor(A)
            assert info == 0
            for rep in range(3):
                b = rng.standard_normal(n).astype(dtype)
                x = h.solve(b)
                assert oracle.backward_error(A, x, b) <= 10 * n * eps, (n, mode, rep)
            x_ref = oracle.lapack_getrs(lu_ref, ipiv_ref, b)
            scale = np.linalg.norm(x_ref, np.inf)
            assert np.linalg.norm(x - x_ref, np.inf) <= 1e4 * n * eps * scale, (n, mode)
            xs.append(x)
            # deterministic: the same solve again gives the same bits
            assert np.array_equal(x, h.solve(b)), (n, mode)
            # ... and so does the copy-engine path (default: b and x in device-mapped host memory)
            h.set_option(C.OPT_MAPPED_RHS, 0)
            assert np.array_equal(x, h.solve(b)), (n, mode)
