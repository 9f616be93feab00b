"""The reference's interface tests, re-read for B200LUFactorization
(test/Core/basictests.jl:64-92 test_interface, lu_refactorization.jl,
direct_blas_refactorization.jl, batch.jl, retcodes.jl, resolve.jl, Trim)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _interface(ls, alg, prob1, prob2, rtol=1e-10):
    """test_interface(alg, prob1, prob2): solve, cache reuse with new A, new A+b"""
    A1, b1 = prob1.A, prob1.b
    A2, b2 = prob2.A, prob2.b
    x1 = np.linalg.solve(A1.astype(np.float64), b1.astype(np.float64))
    x2 = np.linalg.solve(A2.astype(np.float64), b2.astype(np.float64))
    sol = ls.solve(prob1, alg)
    np.testing.assert_allclose(A1 @ sol.u, b1, rtol=rtol, atol=rtol)
    cache = ls.init(prob1, alg)
    sol = ls.solve_(cache)
    np.testing.assert_allclose(sol.u, x1, rtol=rtol)
    cache.A = A2.copy()
    sol = ls.solve_(cache)
    np.testing.assert_allclose(sol.u, np.linalg.solve(A2.astype(np.float64), b1.astype(np.float64)), rtol=rtol)
    cache.b = b2.copy()
    sol = ls.solve_(cache)
    np.testing.assert_allclose(sol.u, x2, rtol=rtol)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-10), (np.float32, 2e-3)])
def test_interface_8x8(gpu_required, ls, dtype, rtol):
    rng = np.random.default_rng(0)
    n = 8
    A1 = (rng.random((n, n)) / 10 + np.eye(n)).astype(dtype)
    A2 = (rng.random((n, n)) / 10 + np.eye(n)).astype(dtype)
    b1, b2 = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
    _interface(ls, ls.B200LUFactorization(), ls.LinearProblem(A1, b1), ls.LinearProblem(A2, b2), rtol)


def test_known_answers_2x2(gpu_required, ls):
    """test/Core/retcodes.jl:17-18,41-42; test/Core/resolve.jl:84-96; test/Trim/runtests.jl:7"""
    alg = ls.B200LUFactorization()
    sol = ls.solve(ls.LinearProblem(np.array([[2.0, 1.0], [-1.0, 1.0]]), np.array([-1.0, 1.0])), alg)
    assert sol.retcode == ls.ReturnCode.Success
    np.testing.assert_allclose(sol.u, np.linalg.solve([[2.0, 1.0], [-1.0, 1.0]], [-1.0, 1.0]))
    sol = ls.solve(ls.LinearProblem(np.ones((2, 2)), np.ones(2)), alg)
    assert sol.retcode == ls.ReturnCode.Failure
    A = np.array([[1.0, 2.0], [3.0, 4.0]])
    sol = ls.solve(ls.LinearProblem(A.T @ A, np.array([1.0, 2.0])), alg)
    np.testing.assert_allclose(sol.u, [-2.0, 1.5], rtol=1e-12)
    sol = ls.solve(ls.LinearProblem(np.array([[4.0, 1.0], [1.0, 3.0]]), np.array([1.0, 2.0])), alg)
    np.testing.assert_allclose(sol.u, [1 / 11, 7 / 11], rtol=1e-14)
    # integer inputs are promoted (src/common.jl:448-502)
    sol = ls.solve(ls.LinearProblem(np.array([[4, 1], [1, 3]]), np.array([1, 2])), alg)
    np.testing.assert_allclose(sol.u, [1 / 11, 7 / 11], rtol=1e-14)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_refactorization_sequence(gpu_required, ls, dtype):
    """test/Core/direct_blas_refactorization.jl:15-46: A1 -> A2 -> A1 -> A2 -> singular -> A1"""
    rng = np.random.default_rng(42)
    n = 51
    rtol = 1e-10 if dtype == np.float64 else 1e-3
    A1 = (rng.random((n, n)) + n * np.eye(n)).astype(dtype)
    A2 = (rng.random((n, n)) + n * np.eye(n)).astype(dtype)
    Asing = A1.copy()
    Asing[:, 0] = 0
    b = rng.random(n).astype(dtype)
    cache = ls.init(ls.LinearProblem(A1.copy(), b.copy()), ls.B200LUFactorization())
    np.testing.assert_allclose(ls.solve_(cache).u, np.linalg.solve(A1, b), rtol=rtol)
    handle_before = None
    for Ak in (A2, A1, A2):
        cache.A = Ak.copy()
        sol = ls.solve_(cache)
        assert sol.retcode == ls.ReturnCode.Success
        np.testing.assert_allclose(sol.u, np.linalg.solve(Ak, b), rtol=rtol)
        if handle_before is not None:
            assert cache.cacheval.handle is handle_before   # the cached device buffers are reused
        handle_before = cache.cacheval.handle
    cache.A = Asing.copy()
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Failure
    assert cache.isfresh                                     # Failure leaves isfresh set
    cache.A = A1.copy()
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success
    np.testing.assert_allclose(sol.u, np.linalg.solve(A1, b), rtol=rtol)


def test_isfresh_protocol(gpu_required, ls):
    """cache.b = b2 must NOT refactor; cache.A = A2 must (src/common.jl:330-348)"""
    rng = np.random.default_rng(1)
    n = 300
    A = rng.random((n, n)) + n * np.eye(n)
    cache = ls.init(ls.LinearProblem(A, rng.random(n)), ls.B200LUFactorization())
    ls.solve_(cache)
    assert not cache.isfresh
    launches = ls.launch_count()
    cache.b = rng.random(n)
    sol = ls.solve_(cache)
    solve_only = ls.launch_count() - launches
    np.testing.assert_allclose(A @ sol.u, cache.b, rtol=1e-10)
    assert not cache.isfresh
    cache.A = A + np.eye(n)
    assert cache.isfresh
    launches = ls.launch_count()
    ls.solve_(cache)
    assert ls.launch_count() - launches > solve_only        # a factorization ran again


@pytest.mark.parametrize("nrhs", [1, 2, 4, 5, 8, 17, 100])
def test_matrix_rhs(gpu_required, ls, oracle, nrhs):
    """test/Core/batch.jl:9-49,118-133: matrix right-hand side vs A \\ B, reuse with new B"""
    rng = np.random.default_rng(nrhs)
    n = 200
    A = rng.random((n, n)) + n * np.eye(n)
    B = rng.random((n, nrhs))
    cache = ls.init(ls.LinearProblem(A, B), ls.B200LUFactorization())
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success
    np.testing.assert_allclose(sol.u, np.linalg.solve(A, B), rtol=1e-10, atol=1e-14)
    B2 = rng.random((n, nrhs))
    cache.b = B2
    np.testing.assert_allclose(ls.solve_(cache).u, np.linalg.solve(A, B2), rtol=1e-10, atol=1e-14)
    # singular => failure retcode (batch.jl:135-147)
    cache.A = np.zeros((n, n))
    assert ls.solve_(cache).retcode == ls.ReturnCode.Failure


def test_residualsafety(gpu_required, ls):
    rng = np.random.default_rng(2)
    n = 64
    A = rng.random((n, n)) + n * np.eye(n)
    b = rng.random(n)
    sol = ls.solve(ls.LinearProblem(A, b), ls.B200LUFactorization(residualsafety=True))
    assert sol.retcode == ls.ReturnCode.Success


def test_default_algorithm_selects_b200(gpu_required, ls):
    """the new arm sits above n = 600 and is gated on availability"""
    rng = np.random.default_rng(3)
    n = 1100
    A = rng.random((n, n)) + n * np.eye(n)
    b = rng.random(n)
    assert ls.defaultalg(A, b).alg == ls.DefaultAlgorithmChoice.B200LUFactorization
    sol = ls.solve(ls.LinearProblem(A, b))
    assert sol.retcode == ls.ReturnCode.Success
    np.testing.assert_allclose(A @ sol.u, b, rtol=1e-9)


def test_threads_independent_caches(gpu_required, ls):
    """independent caches on host threads (test/Core/basictests.jl:1331-1358)"""
    import threading
    rng = np.random.default_rng(4)
    n = 128
    probs = [(rng.random((n, n)) + n * np.eye(n), rng.random(n)) for _ in range(4)]
    out = [None] * 4

    def work(i):
        A, b = probs[i]
        out[i] = ls.solve(ls.LinearProblem(A, b), ls.B200LUFactorization()).u

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for (A, b), u in zip(probs, out):
        np.testing.assert_allclose(A @ u, b, rtol=1e-10)


@pytest.mark.parametrize("dtype,n", [(np.float64, 5), (np.float64, 64), (np.float64, 65), (np.float64, 300),
                                     (np.float64, 1000), (np.float64, 2500), (np.float32, 700)])
def test_adjoint_solve_reuses_factorization(gpu_required, ls, dtype, n):
    """`solve!(cache; adjoint = true)` (reference src/common.jl:1012-1027, test/Core/adjoint.jl): the
    cached factorization of A solves adjoint(A) u = b — getrs with trans = 'T' through the C ABI;
    backward error of the transposed system <= 10 n eps; vector and matrix right-hand sides."""
    rng = np.random.default_rng(400 + n)
    A = np.asfortranarray(rng.random((n, n)).astype(dtype) + (0.5 * np.eye(n)).astype(dtype))
    b = rng.random(n).astype(dtype)
    cache = ls.init(ls.LinearProblem(A, b), ls.B200LUFactorization())
    x = ls.solve_(cache).u.copy()
    eps = np.finfo(dtype).eps
    assert np.linalg.norm(A.astype(np.float64) @ x - b) / (np.linalg.norm(A) * np.linalg.norm(x)) <= 10 * n * eps
    sol = ls.solve_(cache, adjoint=True)
    assert sol.retcode == ls.ReturnCode.Success
    xt = sol.u.copy()
    r = A.T.astype(np.float64) @ xt - b
    assert np.linalg.norm(r) / (np.linalg.norm(A) * np.linalg.norm(xt)) <= 10 * n * eps
    # the factorization was reused, and a normal solve still works afterwards
    assert not cache.isfresh
    x2 = ls.solve_(cache).u
    assert np.array_equal(x, x2)
    # matrix right-hand side through the handle
    B = np.asfortranarray(rng.random((n, 3)).astype(dtype))
    Xt = cache.cacheval.handle.solve(B, trans="T")
    R = A.T.astype(np.float64) @ Xt - B
    assert np.linalg.norm(R) / (np.linalg.norm(A) * np.linalg.norm(Xt)) <= 10 * n * eps


@pytest.mark.parametrize("dtype,n,nrhs", [(np.float64, 1000, 100), (np.float64, 1028, 16), (np.float64, 2050, 33),
                                          (np.float64, 1001, 40), (np.float32, 1536, 64), (np.float32, 6144, 64),
                                          (np.float64, 4096, 100)])
def test_matrix_rhs_blocked_trsm(gpu_required, ls, dtype, n, nrhs):
    """getrs with many right-hand sides (BASELINE config 2 applies 100): the blocked TRSM path —
    diagonal 256-blocks by the block-row kernel, off-diagonal updates by the tensor-core GEMM — and
    its fallbacks (n not a multiple of 4: block-row kernel only).  Column-wise backward error
    <= 10 n eps; agreement with the one-right-hand-side path."""
    rng = np.random.default_rng(n + nrhs)
    A = np.asfortranarray(rng.random((n, n)).astype(dtype))
    B = np.asfortranarray(rng.random((n, nrhs)).astype(dtype))
    h = ls.Handle(ls._capi.F64 if dtype == np.float64 else ls._capi.F32)
    _, info = h.factor(A)
    assert info == 0
    X = h.solve(B)
    eps = np.finfo(dtype).eps
    R = A.astype(np.float64) @ X.astype(np.float64) - B
    nA = np.linalg.norm(A)
    for c in range(nrhs):
        assert np.linalg.norm(R[:, c]) / (nA * np.linalg.norm(X[:, c])) <= 10 * n * eps, c
    x0 = h.solve(np.ascontiguousarray(B[:, 0]))
    assert np.linalg.norm(x0 - X[:, 0]) <= 100 * n * eps * np.linalg.norm(x0)
