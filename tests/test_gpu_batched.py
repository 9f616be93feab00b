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
