"""Batched small systems (BlockDiagonal surface): per-block parity with the
reference's per-block lu!/ldiv! (ext/LinearSolveBlockDiagonalsExt.jl:119-125,
183-205; test/Core/basictests.jl:1168-1223)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 3, 8, 16, 17, 31, 32, 33, 48, 63, 64])
def test_batched_factor_solve_vs_oracle(gpu_required, ls, oracle, n):
    rng = np.random.default_rng(n)
    batch = 37
    A = rng.random((batch, n, n)) + (n if n % 2 else 0) * np.eye(n)   # [s, col, row]
    b = rng.random((batch, n))
    h = ls.Handle(ls._capi.F64)
    ipiv, info = h.factor_batched(A)
    assert not info.any()
    x = h.solve_batched(b)
    LU, ipiv2, info2 = h.get_factors_batched()
    assert np.array_equal(ipiv, ipiv2)
    F_ref, ipiv_ref, info_ref, x_ref = oracle.ref_batched(A, b)
    eps = np.finfo(np.float64).eps
    for s in range(batch):
        As = A[s].T
        assert oracle.compare_ipiv(As, ipiv[s], ipiv_ref[s])[1] in ("exact", "tie"), (n, s)
        assert oracle.scaled_residual(As, LU[s].T, ipiv[s]) < 20
        assert oracle.backward_error(As, x[s], b[s]) <= 10 * n * eps
        _, ipiv_lapack, _ = oracle.lapack_getrf(As)
        assert oracle.compare_ipiv(As, ipiv[s], ipiv_lapack)[1] in ("exact", "tie")
    np.testing.assert_allclose(x, x_ref, rtol=1e-9)


def test_batched_singular_and_permutation(gpu_required, ls, oracle):
    rng = np.random.default_rng(0)
    n, batch = 64, 8
    A = rng.random((batch, n, n))
    A[2] = 0.0                         # zero matrix => info 1
    A[5][:, :] = A[5][:, :]            # keep
    A[5].T[:, 3] = 0.0                 # zero column 4 of system 5 (A[s] is column-major)
    P = np.eye(n)[rng.permutation(n), :]
    A[6] = P.T                         # permutation matrix
    h = ls.Handle(ls._capi.F64)
    ipiv, info = h.factor_batched(A)
    assert info[2] == 1
    _, _, info_ref = oracle.lapack_getrf(A[5].T)
    assert info[5] == info_ref > 0
    lu_ref, ipiv_ref, _ = oracle.lapack_getrf(P)
    LU, _, _ = h.get_factors_batched()
    assert np.array_equal(ipiv[6], ipiv_ref)
    assert np.array_equal(LU[6].T, lu_ref)
    assert info[[0, 1, 3, 4, 6, 7]].sum() == 0


def test_blockdiagonal_problem(gpu_required, ls):
    """test/Core/basictests.jl:1168-1223: blocks [3,3,3,3], [2,3,4], a singular block => Failure"""
    rng = np.random.default_rng(1)
    for sizes in ([3, 3, 3, 3], [2, 3, 4], [64] * 5 + [70]):
        blocks = [rng.random((k, k)) + k * np.eye(k) for k in sizes]
        A = ls.BlockDiagonal(blocks)
        n = sum(sizes)
        for b in (rng.random(n), rng.random((n, 3))):
            sol = ls.solve(ls.LinearProblem(A, b), ls.B200LUFactorization())
            assert sol.retcode == ls.ReturnCode.Success
            np.testing.assert_allclose(sol.u, np.linalg.solve(A.to_dense(), b), rtol=1e-10)
    blocks = [rng.random((3, 3)) + 3 * np.eye(3) for _ in range(3)] + [np.ones((3, 3))]
    sol = ls.solve(ls.LinearProblem(ls.BlockDiagonal(blocks), rng.random(12)), ls.B200LUFactorization())
    assert sol.retcode == ls.ReturnCode.Failure


def test_batched_full_size_properties(gpu_required, ls, oracle):
    """BASELINE config 4 geometry (reduced batch for the host copy): 8192 systems of
    64x64 + 64 I; property checks on a sample + residual norm over all."""
    rng = np.random.default_rng(64)
    n, batch = 64, 8192
    A = rng.random((batch, n, n)) + n * np.eye(n)
    b = rng.random((batch, n))
    h = ls.Handle(ls._capi.F64)
    ipiv, info = h.factor_batched(A)
    assert not info.any()
    x = h.solve_batched(b)
    r = np.einsum("sji,sj->si", A, x) - b          # A[s].T @ x[s]
    assert np.abs(r).max() < 1e-12 * n
    for s in rng.choice(batch, 16, replace=False):
        _, ipiv_ref, _ = oracle.lapack_getrf(A[s].T)
        assert np.array_equal(ipiv[s], ipiv_ref)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 5, 8, 9, 16, 17, 32, 33, 40, 63, 64])
def test_warp_kernel_matches_reference_algorithm_bitwise(gpu_required, ls, oracle, n, dtype):
    """The warp-per-system kernel (OPT_BATCHED_MODE 0, blocked with 8-column register panels) and the round-1
    row-per-thread kernel (mode 1) both run the FMA sequence of the unblocked right-looking algorithm, and so
    does the C oracle (`_blocked_lu_unblocked!`, src/blocked_lufact.jl:58-90: pivot rule of :38-54, FMA updates): factors, pivots and
    info are compared BIT FOR BIT, also with a singular and a NaN-carrying system in the batch."""
    C = ls._capi
    rng = np.random.default_rng(100 + n)
    batch = 23
    A = rng.random((batch, n, n)).astype(dtype)
    if n > 2:
        A[3].T[:, 1] = 0          # zero column 2 of system 3: info = 2, the factorization runs on
        A[7][0, min(2, n - 1)] = np.nan   # a NaN entry: never chosen while a finite candidate exists
    code = C.F64 if dtype == np.float64 else C.F32
    h0, h1 = ls.Handle(code), ls.Handle(code)
    h1.set_option(C.OPT_BATCHED_MODE, 1)
    ip0, info0 = h0.factor_batched(A)
    ip1, info1 = h1.factor_batched(A)
    LU0, _, _ = h0.get_factors_batched()
    LU1, _, _ = h1.get_factors_batched()
    assert np.array_equal(ip0, ip1) and np.array_equal(info0, info1)
    assert np.array_equal(LU0, LU1, equal_nan=True)
    if dtype == np.float64:
        for s in range(batch):
            F, ipr, infr = oracle.ref_lufact(np.asfortranarray(A[s].T), variant="unblocked")
            assert infr == info0[s], (s, infr, info0[s])
            assert np.array_equal(ipr, ip0[s]), s
            assert np.array_equal(F, LU0[s].T, equal_nan=True), s


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [3, 16, 31, 32, 47, 64])
def test_fused_factor_solve_batched(gpu_required, ls, oracle, n, dtype):
    """b200lu_factor_solve_batched: getrf + the first getrs in one kernel, factors kept: same pivots and factors
    as b200lu_factor_batched, x within the backward-error bar, later solves from the cached factors agree."""
    C = ls._capi
    rng = np.random.default_rng(200 + n)
    batch = 41
    A = (rng.random((batch, n, n)) + n * np.eye(n)).astype(dtype)
    b = rng.random((batch, n)).astype(dtype)
    code = C.F64 if dtype == np.float64 else C.F32
    h, hr = ls.Handle(code), ls.Handle(code)
    x, ipiv, info = h.factor_solve_batched(A, b)
    ipr, infr = hr.factor_batched(A)
    assert np.array_equal(ipiv, ipr) and not info.any() and not infr.any()
    assert np.array_equal(h.get_factors_batched()[0], hr.get_factors_batched()[0])
    eps = np.finfo(dtype).eps
    for s in range(batch):
        assert oracle.backward_error(A[s].T.astype(np.float64), x[s].astype(np.float64), b[s].astype(np.float64)) <= 10 * n * eps
        _, ipiv_lapack, _ = oracle.lapack_getrf(A[s].T)
        assert oracle.compare_ipiv(A[s].T, ipiv[s], ipiv_lapack)[1] in ("exact", "tie")
    assert np.array_equal(h.solve_batched(b), hr.solve_batched(b))
    np.testing.assert_allclose(h.solve_batched(b), x, rtol=0, atol=100 * n * eps * np.abs(x).max())
    # device entry with the same data
    import torch
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    Ad = torch.from_numpy(A).cuda()
    bd = torch.from_numpy(b).cuda()
    xd = torch.empty_like(bd)
    assert h.factor_solve_batched_device(Ad.data_ptr(), bd.data_ptr(), xd.data_ptr(), batch, n) == 0
    torch.cuda.synchronize()
    assert np.array_equal(xd.cpu().numpy(), x)
    # a singular system in the batch is reported, the others are solved
    A2 = A.copy()
    A2[5] = 1.0
    x2, _, info2 = h.factor_solve_batched(A2, b)
    assert info2[5] > 0 and not np.delete(info2, 5).any()
    assert np.array_equal(np.delete(x2, 5, axis=0), np.delete(x, 5, axis=0))
