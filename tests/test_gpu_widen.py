"""Rows SURVEY §8 marks "next", at the same parity bar as the hot path: adjoint solves for
every handle type (`solve!(cache; adjoint = true)`, reference src/common.jl:1012-1027 and
test/Core/adjoint.jl) and ragged BlockDiagonal problems in one batched launch per kernel class
(ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205; test/Core/basictests.jl:1168-1223)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps


def _berr(M, x, b):
    return np.linalg.norm(M @ x - b) / (np.linalg.norm(M) * np.linalg.norm(x))


@pytest.mark.parametrize("n", [100, 1000, 2500])
def test_mixed_adjoint_refines_transposed_system(gpu_required, ls, n):
    """FP32 factors of A + FP64 refinement of A^T x = b: transposed FP32 sweeps, residual
    b - A^T x in FP64; backward error of the TRANSPOSED system <= 10 n eps64 (north-star bar)."""
    rng = np.random.default_rng(900 + n)
    A = np.asfortranarray(rng.random((n, n)) + 5.0 * np.eye(n))
    b = rng.random(n)
    cache = ls.init(ls.LinearProblem(A, b), ls.B200LU32MixedLUFactorization())
    x = ls.solve_(cache).u.copy()
    assert _berr(A, x, b) <= 10 * n * EPS
    sol = ls.solve_(cache, adjoint=True)
    assert sol.retcode == ls.ReturnCode.Success and not cache.isfresh
    xt = sol.u.copy()
    assert _berr(A.T, xt, b) <= 10 * n * EPS
    assert _berr(A.T, xt, b) <= 50 * EPS          # refinement reaches a few eps
    # matrix right-hand side, column by column
    B = np.asfortranarray(rng.random((n, 3)))
    Xt = cache.cacheval.handle.solve(B, trans="T")
    for c in range(3):
        assert _berr(A.T, Xt[:, c], B[:, c]) <= 10 * n * EPS
    # the untransposed solve still works afterwards (not bitwise: its residual kernel sums column
    # chunks with atomics)
    x2 = ls.solve_(cache).u
    assert _berr(A, x2, b) <= 10 * n * EPS
    np.testing.assert_allclose(x2, x, rtol=1e-9)
    # without refinement: the reference's *32Mixed accuracy class for the transposed system
    h = ls.Handle(ls._capi.MIXED)
    h.set_option(ls._capi.OPT_REFINE_MAXIT, 0)
    _, info = h.factor(A)
    assert info == 0
    x32 = h.solve(b, trans="T")
    # FP32 accuracy class: the backward error bar of the FP32 element type (the forward error is that times the
    # condition number: ~1e-3 at n = 2500 for this matrix, and it moves with the rounding order of the kernels)
    assert _berr(A.T, x32, b) <= 10 * n * np.finfo(np.float32).eps
    assert np.linalg.norm(x32 - xt) / np.linalg.norm(xt) < 1e-2


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 3, 16, 17, 32, 33, 50, 64])
def test_batched_transposed_solve(gpu_required, ls, dtype, n):
    rng = np.random.default_rng(70 + n)
    batch = 19
    A = (rng.random((batch, n, n)) + 0.5 * n * np.eye(n)).astype(dtype)    # [s, col, row]: matrix s is A[s].T
    h = ls.Handle(ls._capi.F64 if dtype == np.float64 else ls._capi.F32)
    _, info = h.factor_batched(A)
    assert not info.any()
    eps = np.finfo(dtype).eps
    for nrhs in (1, 3):
        b = rng.random((batch, nrhs, n)).astype(dtype)
        xt = h.solve_batched(b, trans="T")
        x = h.solve_batched(b)
        for s in range(batch):
            M = A[s].T.astype(np.float64)
            for r in range(nrhs):
                assert _berr(M.T, xt[s, r].astype(np.float64), b[s, r]) <= 10 * n * eps, (s, r)
                assert _berr(M, x[s, r].astype(np.float64), b[s, r]) <= 10 * n * eps, (s, r)
    # 'C' == 'T' for real element types; a bad trans is an argument error, not a crash
    assert np.array_equal(h.solve_batched(b, trans="C"), xt)
    with pytest.raises(ls.B200LUError):
        h.solve_batched(b, trans="X")


@pytest.mark.parametrize("sizes", [[2, 3, 4], [5, 64, 17, 33, 70, 16, 1], [3, 3, 3, 3], [40, 9, 130], [100, 161, 97, 300]])
def test_ragged_blockdiagonal_one_launch_per_class(gpu_required, ls, oracle, sizes):
    rng = np.random.default_rng(sum(sizes))
    blocks = [rng.random((k, k)) + k * np.eye(k) for k in sizes]
    A = ls.BlockDiagonal(blocks)
    n = sum(sizes)
    D = A.to_dense()
    cache = ls.init(ls.LinearProblem(A, rng.random(n)), ls.B200LUFactorization())
    before = ls.launch_count()
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success
    np.testing.assert_allclose(sol.u, np.linalg.solve(D, cache.b), rtol=1e-10)
    plan = ls.plan_blockdiag(sizes)
    assert len(cache.cacheval.groups) == len(plan) <= 6 + sum(k > 160 for k in sizes)
    n_batched = sum(kind == "batched" for kind, *_ in plan)
    if all(k <= 160 for k in sizes):
        # classes of up to 64 rows: ONE launch (the first getrs rides in the getrf kernel); above: getrf + getrs
        n_fused = sum(kind == "batched" and m <= 64 for kind, _, m in plan)
        assert ls.launch_count() - before == 2 * n_batched - n_fused
    # pivots of the padded systems are the pivots of the blocks themselves (LAPACK, up to ties)
    for kind, h, idx, m in cache.cacheval.groups:
        if kind != "batched":
            continue
        _, ipiv, info = h.get_factors_batched()
        assert not info.any()
        for s, i in enumerate(idx):
            k = sizes[i]
            _, ipiv_ref, _ = oracle.lapack_getrf(blocks[i])
            assert oracle.compare_ipiv(blocks[i], ipiv[s, :k], ipiv_ref)[1] in ("exact", "tie")
            assert np.array_equal(ipiv[s, k:], np.arange(k + 1, m + 1))
    # matrix right-hand side, re-solve only
    B = rng.random((n, 3))
    cache.b = B
    cache.u = np.zeros_like(B)
    np.testing.assert_allclose(ls.solve_(cache).u, np.linalg.solve(D, B), rtol=1e-10)
    # adjoint with the same per-block factors
    sol = ls.solve_(cache, adjoint=True)
    assert sol.retcode == ls.ReturnCode.Success and not cache.isfresh
    np.testing.assert_allclose(sol.u, np.linalg.solve(D.T, B), rtol=1e-10)
    # a singular block anywhere => Failure, isfresh stays set (retcode protocol, src/openblas.jl:362-459)
    bad = [b.copy() for b in blocks]
    bad[1][:, 0] = 0.0
    cache.A = ls.BlockDiagonal(bad)
    assert ls.solve_(cache).retcode == ls.ReturnCode.Failure and cache.isfresh


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [65, 70, 96, 100, 128, 131, 159, 160])
def test_batched_shared_memory_class(gpu_required, ls, oracle, dtype, n):
    """blocks of 65 ... 160 rows: one CTA per system with the system in shared memory — the same
    parity bar as the register kernels (tests/test_gpu_batched.py): pivots equal to LAPACK's up to
    ties, info, scaled residual < 20, backward error <= 10 n eps, transposed solves"""
    rng = np.random.default_rng(1000 + n)
    batch = 11
    A = rng.random((batch, n, n)).astype(dtype)                  # [s, col, row]
    A[3] = ls.pad_blocks([np.eye(n)[rng.permutation(n), :]], [0], n, dtype)[0]   # a permutation matrix
    A[5].T[:, 6] = 0.0                                           # zero column 7 => info of LAPACK
    A[7] = 0.0                                                   # zero matrix => info 1
    h = ls.Handle(ls._capi.F64 if dtype == np.float64 else ls._capi.F32)
    ipiv, info = h.factor_batched(A)
    LU, ipiv2, info2 = h.get_factors_batched()
    assert np.array_equal(ipiv, ipiv2) and np.array_equal(info, info2)
    eps = np.finfo(dtype).eps
    for s in range(batch):
        M = np.asfortranarray(A[s].T)
        lu_ref, ipiv_ref, info_ref = oracle.lapack_getrf(M)
        assert info[s] == info_ref, s
        if s == 7:
            assert info[s] == 1 and np.array_equal(ipiv[s], np.arange(1, n + 1))
            continue
        if s == 5:
            assert info[s] == 7 and np.array_equal(ipiv[s, :6], ipiv_ref[:6])
            continue
        assert oracle.compare_ipiv(M, ipiv[s], ipiv_ref)[1] in ("exact", "tie"), s
        if s == 3:
            assert np.array_equal(ipiv[s], ipiv_ref) and np.array_equal(LU[s].T, lu_ref)
        assert oracle.scaled_residual(M, LU[s].T, ipiv[s]) < 20
    # solves on a non-singular batch
    A = (rng.random((batch, n, n)) + 0.25 * n * np.eye(n)).astype(dtype)
    _, info = h.factor_batched(A)
    assert not info.any()
    b = rng.random((batch, 2, n)).astype(dtype)
    x = h.solve_batched(b)
    xt = h.solve_batched(b, trans="T")
    for s in range(batch):
        M = A[s].T.astype(np.float64)
        for r in range(2):
            assert _berr(M, x[s, r].astype(np.float64), b[s, r]) <= 10 * n * eps, (s, r)
            assert _berr(M.T, xt[s, r].astype(np.float64), b[s, r]) <= 10 * n * eps, (s, r)


@pytest.mark.parametrize("dtype,n,nrhs", [(np.float64, 64, 1), (np.float64, 777, 3), (np.float64, 3000, 2),
                                          (np.float32, 500, 2)])
def test_device_residual_norms(gpu_required, ls, dtype, n, nrhs):
    """b200lu_residual_norms: the reference's a-posteriori check (`_check_residual_safety`,
    src/factorization.jl:127-156) with the norms accumulated in FP64 on the device from the copy of
    A kept by B200LU_OPT_KEEP_A"""
    rng = np.random.default_rng(5000 + n)
    A = np.asfortranarray((rng.random((n, n)) + 0.1 * n * np.eye(n)).astype(dtype))
    B = np.asfortranarray(rng.random((n, nrhs)).astype(dtype))
    code = ls._capi.F64 if dtype == np.float64 else ls._capi.F32
    h = ls.Handle(code)
    h.set_option(ls._capi.OPT_KEEP_A, 1)
    ipiv, info = h.factor(A)
    assert info == 0
    X = h.solve(B)
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    resid, bnorm = h.residual_norms(B, X)
    np.testing.assert_allclose(bnorm, np.linalg.norm(B64, axis=0), rtol=1e-12)
    eps = np.finfo(dtype).eps
    assert np.all(resid <= 10 * n * eps * np.linalg.norm(A64) * np.linalg.norm(X.astype(np.float64), axis=0))
    # a perturbed "solution": the residual is well above rounding and must match the host value
    X2 = (X + 1e-2 * rng.random(X.shape)).astype(dtype)
    resid2, _ = h.residual_norms(B, X2)
    np.testing.assert_allclose(resid2, np.linalg.norm(A64 @ X2.astype(np.float64) - B64, axis=0), rtol=1e-9)
    # keeping A bypasses the streamed upload, not the arithmetic: same pivots as a default handle,
    # which in turn refuses the check (no copy of A on the device)
    h2 = ls.Handle(code)
    ipiv2, _ = h2.factor(A)
    assert np.array_equal(ipiv, ipiv2)
    with pytest.raises(ls.B200LUError):
        h2.residual_norms(B, X)
    # a refactorization with the option switched off invalidates the kept copy
    h.set_option(ls._capi.OPT_KEEP_A, 0)
    h.factor(A)
    with pytest.raises(ls.B200LUError):
        h.residual_norms(B, X)


def test_residualsafety_on_device(gpu_required, ls):
    rng = np.random.default_rng(11)
    n = 1200
    A = rng.random((n, n)) + n * np.eye(n)
    for b in (rng.random(n), rng.random((n, 4))):
        for alg in (ls.B200LUFactorization(residualsafety=True), ls.B200LU32MixedLUFactorization(residualsafety=True)):
            cache = ls.init(ls.LinearProblem(A, b), alg)
            sol = ls.solve_(cache)
            assert sol.retcode == ls.ReturnCode.Success
            assert np.linalg.norm(A @ sol.u - b) <= cache.abstol + cache.reltol * np.linalg.norm(b)
            if not isinstance(alg, ls.B200LU32MixedLUFactorization):
                assert cache.cacheval.handle.get_option(ls._capi.OPT_KEEP_A) == 1
            # an impossible tolerance turns the same solve into a Failure (the check is really evaluated)
            strict = ls.init(ls.LinearProblem(A, b), alg, abstol=0.0, reltol=1e-300)
            assert ls.solve_(strict).retcode == ls.ReturnCode.APosterioriSafetyFailure
    # BlockDiagonal: block-by-block check
    blocks = [rng.random((k, k)) + k * np.eye(k) for k in (3, 70, 20)]
    bd = ls.BlockDiagonal(blocks)
    sol = ls.solve(ls.LinearProblem(bd, rng.random(93)), ls.B200LUFactorization(residualsafety=True))
    assert sol.retcode == ls.ReturnCode.Success



def _is_pinned(arr):
    """cudaPointerGetAttributes through cuda-python: is this host buffer page-locked / registered?"""
    from cuda.bindings import runtime as cudart
    err, at = cudart.cudaPointerGetAttributes(arr.ctypes.data)
    assert int(err) == 0
    return int(at.type) == int(cudart.cudaMemoryType.cudaMemoryTypeHost)


def test_host_register_option(gpu_required, ls):
    """B200LU_OPT_HOST_REGISTER: the library page-locks the caller's pageable matrix once (the cache's own copy of
    A with alias_A = false, reference src/common.jl:818-842), keeps it registered across refactorizations, moves
    the registration when another buffer arrives, releases it with the handle; factors are the same either way."""
    C = ls._capi
    n = 3000
    rng = np.random.default_rng(31)
    A = np.asfortranarray(rng.random((n, n)))
    A2 = np.asfortranarray(rng.random((n, n)))
    h0 = C.Handle(C.F64)
    ipiv0, info0 = h0.factor(A)
    LU0 = h0.get_factors()
    assert not _is_pinned(A)                       # default: the caller's buffer is left alone
    h = C.Handle(C.F64)
    h.set_option(C.OPT_HOST_REGISTER, 1)
    assert h.get_option(C.OPT_HOST_REGISTER) == 1
    for _ in range(2):                             # second call: same buffer, already registered
        ipiv, info = h.factor(A)
        assert info == info0 == 0 and np.array_equal(ipiv, ipiv0) and np.array_equal(h.get_factors(), LU0)
        assert _is_pinned(A)
    h.factor(A2)                                   # another buffer: the registration moves
    assert _is_pinned(A2) and not _is_pinned(A)
    h.close()
    assert not _is_pinned(A2)
    # through the public interface
    b = rng.random(n)
    cache = ls.init(ls.LinearProblem(A, b), ls.B200LUFactorization(host_register=True))
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success and _berr(A, sol.u, b) <= 10 * n * EPS


@pytest.mark.parametrize("n", [1500, 6000])
def test_mixed_refinement_is_deterministic(gpu_required, ls, n):
    """FP32 factors + FP64 refinement give the SAME bits on every run: the FP64 residual and the norms that decide
    the sweep count are reduced in one fixed order (no atomics), like every other kernel of the path."""
    import torch
    C = ls._capi
    dev = torch.device("cuda", 0)
    A = torch.empty((n, n), dtype=torch.float64, device=dev)
    b = torch.empty((2, n), dtype=torch.float64, device=dev)
    xs = []
    for rep in range(4):
        h = ls.Handle(C.MIXED)
        h.fill_uniform_device(A.data_ptr(), n, n, n, seed=77)
        h.fill_uniform_device(b.data_ptr(), n, n, 2, seed=78)
        assert h.factor_device(A.data_ptr(), n, n) == 0
        x1 = torch.empty((1, n), dtype=torch.float64, device=dev)
        x2 = torch.empty((2, n), dtype=torch.float64, device=dev)
        h.solve_device(b.data_ptr(), n, x1.data_ptr(), n, 1)          # vector right-hand side
        h.solve_device(b.data_ptr(), n, x2.data_ptr(), n, 2)          # matrix right-hand side
        torch.cuda.synchronize()
        xs.append((x1.clone(), x2.clone(), int(h.counter(C.C_REFINE_ITERS))))
        h.close()
    r = torch.mv(A.t(), xs[0][0][0]) - b[0]
    assert float(r.norm() / (A.norm() * xs[0][0][0].norm())) <= 10 * n * EPS
    for x1, x2, it in xs[1:]:
        assert torch.equal(x1, xs[0][0]) and torch.equal(x2, xs[0][1]) and it == xs[0][2]
