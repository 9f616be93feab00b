"""Multi-GPU engine through the C ABI.

* The distributed getrf / getrs engine with ONE rank (runs on a 1-GPU box): same scheduling code
  as with P ranks (look-ahead on the panel stream, slots, tags), no peers — factors must be
  BITWISE those of the single-GPU getrf, pivots equal, getrs within the backward-error bar.
* One process, P >= 2 GPUs (`b200lu_create(ngpus = P)`; skipped on a 1-GPU box): host matrix in,
  pivots / factors bitwise equal to the single-GPU handle, distributed getrs, sharded batches,
  and the same through `init / solve!` with `B200LUFactorization(devices = ...)`.
The reference has no dense multi-device path: parity is against the single-GPU result, which is
itself pinned against LAPACK (tests/test_gpu_getrf.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _berr(A, x, b):
    return np.linalg.norm(A @ x - b) / (np.linalg.norm(A) * np.linalg.norm(x))


@pytest.mark.parametrize("n,nb,dtype", [(2048, 256, np.float64), (1000, 64, np.float64), (1001, 128, np.float64),
                                        (777, 256, np.float64), (1536, 256, np.float32)])
def test_dist_engine_single_rank(gpu_required, ls, n, nb, dtype):
    import torch
    C = ls._capi
    code = C.F64 if dtype == np.float64 else C.F32
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.random((n, n)).astype(dtype))
    hs = C.Handle(code)
    hs.set_option(C.OPT_NB, nb)
    ipiv_s, info_s = hs.factor(A)
    LU_s = hs.get_factors()
    hd = C.Handle(code)
    hd.set_option(C.OPT_NB, nb)
    hd.comm_init(None, 0, 1)
    assert hd.dist_local_cols(n) == n
    lda = ((n + 3) // 4) * 4                       # 16-byte multiple for both element types
    Aloc = torch.zeros((n, lda), dtype=tdt, device=dev)      # [col, row]
    Aloc[:, :n] = torch.from_numpy(np.ascontiguousarray(A.T)).to(dev)
    info_d = hd.factor_dist(Aloc.data_ptr(), n, lda)
    torch.cuda.synchronize()
    assert info_s == info_d == 0
    LU_d = Aloc[:, :n].cpu().numpy().T
    if dtype == np.float64:
        # same kernels, same per-element FMA order (FP32 differs: its look-ahead update runs on the FFMA
        # kernel here and on the tcgen05 3xTF32 kernel in the single-GPU driver)
        assert np.array_equal(hd.get_ipiv(), ipiv_s)
        assert np.array_equal(LU_d, LU_s), np.abs(LU_d - LU_s).max()
    else:
        from oracle import lu_oracle
        assert lu_oracle.scaled_residual(A, LU_d, hd.get_ipiv()) < 20
    nrhs = 3
    B = rng.random((n, nrhs)).astype(dtype)
    Bd = torch.from_numpy(np.ascontiguousarray(B.T)).to(dev)  # [rhs, row], ld = n
    Xd = torch.empty_like(Bd)
    for _ in range(2):                                        # warm call: same buffers, new tags
        hd.solve_dist(Bd.data_ptr(), n, Xd.data_ptr(), n, nrhs)
    torch.cuda.synchronize()
    X = Xd.cpu().numpy().T
    eps = np.finfo(dtype).eps
    A64 = A.astype(np.float64)
    for c in range(nrhs):
        assert _berr(A64, X[:, c].astype(np.float64), B[:, c].astype(np.float64)) <= 10 * n * eps
    Xs = hs.solve(np.asfortranarray(B))
    assert np.allclose(X, Xs, rtol=0, atol=1e3 * n * eps * np.abs(Xs).max())
    # ONE right-hand side takes the fused per-step kernel (wait, inverted diagonal block, update, tag)
    x1 = torch.empty(n, dtype=tdt, device=dev)
    for _ in range(2):
        hd.solve_dist(Bd[1].data_ptr(), n, x1.data_ptr(), n, 1)
    torch.cuda.synchronize()
    assert _berr(A64, x1.cpu().numpy().astype(np.float64), B[:, 1].astype(np.float64)) <= 10 * n * eps
    assert np.allclose(x1.cpu().numpy(), Xs[:, 1], rtol=0, atol=1e3 * n * eps * np.abs(Xs).max())
    # a second factorization on the same handle (slots and tags are reused), singular this time
    Z = np.asfortranarray(rng.random((n, n)).astype(dtype))
    Z[:, 5] = 0
    Aloc[:, :n] = torch.from_numpy(np.ascontiguousarray(Z.T)).to(dev)
    assert hd.factor_dist(Aloc.data_ptr(), n, lda) == 6
    with pytest.raises(ls.B200LUError):
        hd.solve_dist(Bd.data_ptr(), n, Xd.data_ptr(), n, nrhs)


def test_dist_rejects_misaligned_arguments(gpu_required, ls):
    import torch
    C = ls._capi
    hd = C.Handle(C.F64)
    with pytest.raises(ls.B200LUError):
        hd.factor_dist(0, 64, 64)                 # no communicator yet
    hd.comm_init(None, 0, 1)
    A = torch.zeros((65, 65), dtype=torch.float64, device="cuda:0")
    with pytest.raises(ls.B200LUError):
        hd.factor_dist(A.data_ptr(), 65, 65)      # odd leading dimension: 16-byte operand chunks
    with pytest.raises(ls.B200LUError):
        hd.factor_dist(A.data_ptr() + 8, 64, 66)  # misaligned base


@pytest.mark.parametrize("n,nb,dtype", [(3000, 256, np.float64), (1001, 64, np.float64), (2048, 128, np.float32)])
def test_one_process_multi_gpu_dense(gpu_required, ls, oracle, n, nb, dtype):
    P = _ngpu()
    if P < 2:
        pytest.skip("needs >= 2 GPUs in one process")
    C = ls._capi
    code = C.F64 if dtype == np.float64 else C.F32
    rng = np.random.default_rng(n + 1)
    A = np.asfortranarray(rng.random((n, n)).astype(dtype))
    hs = C.Handle(code)
    hs.set_option(C.OPT_NB, nb)
    ipiv_s, info_s = hs.factor(A)
    LU_s = hs.get_factors()
    for devs in ([0, 1], list(range(min(P, 8)))):
        hm = C.Handle(code, devices=devs)
        assert hm.ngpus == len(devs)
        hm.set_option(C.OPT_NB, nb)
        assert hm.get_option(C.OPT_NB) == nb
        ipiv_m, info_m = hm.factor(A)
        assert info_m == info_s == 0
        assert np.array_equal(hm.get_ipiv(), ipiv_m)
        if dtype == np.float64:
            assert np.array_equal(ipiv_m, ipiv_s)
            assert np.array_equal(hm.get_factors(), LU_s)
        else:
            assert oracle.scaled_residual(A, hm.get_factors(), ipiv_m) < 20
        eps = np.finfo(dtype).eps
        A64 = A.astype(np.float64)
        for B in (rng.random(n).astype(dtype), rng.random((n, 5)).astype(dtype)):
            X = hm.solve(B)
            Xm = X.reshape(n, -1).astype(np.float64)
            Bm = B.reshape(n, -1).astype(np.float64)
            for c in range(Bm.shape[1]):
                assert _berr(A64, Xm[:, c], Bm[:, c]) <= 10 * n * eps
        # refactor with a singular matrix: info like LAPACK's, solve refuses
        Z = A.copy(order="F")
        Z[:, 10] = 0
        _, info_z = hm.factor(Z)
        assert info_z == 11
        with pytest.raises(ls.B200LUError):
            hm.solve(rng.random(n).astype(dtype))
        with pytest.raises(ls.B200LUError):
            hm.factor_device(0, n, n)             # device pointers belong to single-GPU handles
        hm.close()


def test_one_process_multi_gpu_batched_and_interface(gpu_required, ls):
    P = _ngpu()
    if P < 2:
        pytest.skip("needs >= 2 GPUs in one process")
    C = ls._capi
    rng = np.random.default_rng(5)
    batch, n = 1001, 64                            # ragged shards
    A = rng.random((batch, n, n)) + n * np.eye(n)
    b = rng.random((batch, n))
    hs = C.Handle(C.F64)
    ipiv_s, info_s = hs.factor_batched(A)
    xs = hs.solve_batched(b)
    hm = C.Handle(C.F64, devices=list(range(min(P, 8))))
    ipiv_m, info_m = hm.factor_batched(A)
    assert np.array_equal(ipiv_m, ipiv_s) and np.array_equal(info_m, info_s)
    assert np.array_equal(hm.solve_batched(b), xs)
    assert np.array_equal(hm.solve_batched(b, trans="T"), hs.solve_batched(b, trans="T"))
    LUm, ipm, _ = hm.get_factors_batched()
    LUs, _, _ = hs.get_factors_batched()
    assert np.array_equal(LUm, LUs) and np.array_equal(ipm, ipiv_s)
    # the public interface: one LinearCache, all GPUs
    n = 2500
    M = rng.random((n, n))
    rhs = rng.random(n)
    alg = ls.B200LUFactorization(devices=tuple(range(min(P, 8))))
    cache = ls.init(ls.LinearProblem(M, rhs), alg)
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success and not cache.isfresh
    assert _berr(M, sol.u, rhs) <= 10 * n * np.finfo(np.float64).eps
    cache.b = rng.random(n)                        # cache reuse: getrs only
    sol = ls.solve_(cache)
    assert _berr(M, sol.u, cache.b) <= 10 * n * np.finfo(np.float64).eps
    cache.A = np.ones((n, n))
    assert ls.solve_(cache).retcode == ls.ReturnCode.Failure and cache.isfresh
    blocks = [rng.random((k, k)) + k * np.eye(k) for k in (3, 64, 64, 20, 64, 7)]
    bd = ls.BlockDiagonal(blocks)
    rb = rng.random(sum(B.shape[0] for B in blocks))
    sol = ls.solve(ls.LinearProblem(bd, rb), alg)
    assert sol.retcode == ls.ReturnCode.Success
    assert np.allclose(bd.to_dense() @ sol.u, rb, atol=1e-10)
