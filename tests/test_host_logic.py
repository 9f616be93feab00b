"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports
every symbol include/b200lu.h declares; the product path fails loudly without a
GPU (no CPU fallback); the default-algorithm bands of the reference are unchanged
(test/Core/default_algs.jl:4-66); the LinearCache isfresh protocol."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu(ls):
    return not ls.useb200()


def test_header_symbols_are_exported(ls):
    hdr = open(os.path.join(ROOT, "include", "b200lu.h")).read()
    declared = sorted(set(re.findall(r"\b(b200lu_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no prototypes found"
    lib = ls._capi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200lu.h but not exported"
    assert sorted(ls._capi.SYMBOLS) == declared
    assert lib.b200lu_version() >= 100


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "b200lu.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)          # prototypes only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "linearsolve.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "lu_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
                assert "scipy" not in src and "numpy.linalg.solve" not in src, f


def test_fails_loudly_without_gpu(ls):
    if not _no_gpu(ls):
        pytest.skip("a GPU is present")
    with pytest.raises(ls.B200LUError):
        ls.Handle(ls._capi.F64)
    with pytest.raises(RuntimeError):
        ls.B200LUFactorization()                       # throwerror=True
    alg = ls.B200LUFactorization(throwerror=False)     # speculative construction is allowed
    with pytest.raises(ls.B200LUError):
        ls.solve(ls.LinearProblem(np.eye(4), np.ones(4)), alg)   # ... but solving is not


def test_bad_arguments_do_not_need_a_device(ls):
    lib = ls._capi.load()
    h = ctypes.c_void_p()
    assert lib.b200lu_create(ctypes.byref(h), 7, 1, None) == -2       # bad dtype
    assert lib.b200lu_create(ctypes.byref(h), 0, 0, None) == -3       # ngpus out of range
    assert lib.b200lu_create(ctypes.byref(h), 0, 17, None) == -3      # at most 16 GPUs behind one handle
    assert lib.b200lu_create(ctypes.byref(h), 2, 2, None) == -2       # MIXED has no multi-GPU handle
    assert lib.b200lu_create(ctypes.byref(h), 0, 3, None) == 1        # ngpus > 1 needs that many devices (none here)
    assert lib.b200lu_create(None, 0, 1, None) == -1
    assert lib.b200lu_last_timing(None, 0) == -1.0
    assert lib.b200lu_set_option(None, 0, 64) == -1
    # every compute entry point answers a null handle with status -1 (LAPACK style: first argument), no crash
    assert lib.b200lu_factor(None, 4, None, 4, None, None) == -1
    assert lib.b200lu_solve(None, b"N", 1, None, 4, None, 4) == -1
    assert lib.b200lu_residual_norms(None, 1, None, 4, None, 4, None, None) == -1
    assert lib.b200lu_factor_batched(None, 1, 4, None, 4, 16, None, None) == -1
    assert lib.b200lu_solve_batched(None, 1, None, 4, 4, None, 4, 4) == -1
    assert lib.b200lu_solve_batched_trans(None, b"T", 1, None, 4, 4, None, 4, 4) == -1
    assert lib.b200lu_solve_batched_trans_device(None, b"T", 1, None, 4, 4, None, 4, 4) == -1
    assert lib.b200lu_get_factors(None, None, 4) == -1 and lib.b200lu_get_ipiv(None, None) == -1
    assert lib.b200lu_last_error(None) == b"null handle"


def test_header_enums_match_the_python_binding(ls):
    """include/b200lu.h is the contract: every B200LU_OPT_* / element-type constant the ctypes
    binding uses has the value the header declares, and the header's OPT_COUNT covers them all."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200lu.h")).read()
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\bB200LU_([A-Z0-9_]+)\s*=\s*(\d+)", hdr)}
    C = ls._capi
    opts = [k for k in enums if k.startswith("OPT_") and k != "OPT_COUNT"]
    assert len(opts) == enums["OPT_COUNT"] and sorted(enums[k] for k in opts) == list(range(enums["OPT_COUNT"]))
    for k in opts:
        assert getattr(C, k) == enums[k], k
    for k in ("F64", "F32", "MIXED"):
        assert getattr(C, k) == enums[k], k


def test_default_algorithm_bands_unchanged(ls):
    """reference test/Core/default_algs.jl:4-66 with and without the new arm"""
    C = ls.DefaultAlgorithmChoice
    pick = lambda n, **kw: ls.defaultalg(np.zeros((n, n)), np.zeros(n), b200_available=kw.pop("gpu", False), **kw).alg
    assert pick(3) == C.GenericLUFactorization
    assert pick(10) == C.GenericLUFactorization
    assert pick(50) == C.RFLUFactorization
    assert pick(400) == C.RFLUFactorization                 # OpenBLAS band <= 500
    assert pick(600) == C.LUFactorization
    assert pick(600, usemkl=True) == C.MKLLUFactorization
    assert pick(150, isopenblas=False, usemkl=True) == C.RFLUFactorization
    assert pick(300, isopenblas=False, usemkl=True) == C.MKLLUFactorization
    assert pick(200, userecursivefactorization=False) == C.GenericLUFactorization   # OpenBLAS <= 256
    # the new arm: only when available, only above the break-even, never at n = 600
    for n in (3, 50, 400, 600, 1000):
        assert pick(n, gpu=True) == pick(n, gpu=False)
    assert pick(1024, gpu=True) == C.B200LUFactorization
    assert pick(8192, gpu=True) == C.B200LUFactorization
    assert pick(8192, gpu=False) == C.LUFactorization
    assert ls.defaultalg(np.zeros((2000, 2000), dtype=np.complex128), np.zeros(2000), b200_available=True).alg \
        == C.LUFactorization
    ill = ls.OperatorAssumptions(condition="VeryIllConditioned")
    assert ls.defaultalg(np.zeros((2000, 2000)), np.zeros(2000), ill, b200_available=True).alg == C.QRFactorization
    bd = ls.BlockDiagonal([np.eye(3)] * 4)
    assert ls.defaultalg(bd, np.zeros(12), b200_available=False).alg == C.LUFactorization
    assert ls.defaultalg(bd, np.zeros(12), b200_available=True).alg == C.B200LUFactorization


def test_cache_protocol_without_compute(ls):
    """src/common.jl:313-360: cache.A= sets isfresh, cache.b= does not; init copies A and b
    (alias_A = false default for dense factorizations, src/common.jl:525,830-842)"""
    A = np.arange(16.0).reshape(4, 4) + 10 * np.eye(4)
    b = np.ones(4)
    alg = ls.B200LUFactorization(throwerror=False)
    cache = ls.init(ls.LinearProblem(A, b), alg)
    assert cache.isfresh
    assert cache.A is not A and np.array_equal(cache.A, A) and cache.A.flags.f_contiguous
    assert cache.b is not b
    assert np.array_equal(cache.u, np.zeros(4))
    cache.isfresh = False
    cache.b = np.zeros(4)
    assert not cache.isfresh
    cache.A = A
    assert cache.isfresh
    aliased = ls.init(ls.LinearProblem(A, b), alg, alias_A=True, alias_b=True)
    assert aliased.A is A and aliased.b is b
    with pytest.raises(ValueError):
        ls.init(ls.LinearProblem(np.zeros((3, 4)), np.zeros(3)), alg)      # needs_square_A
    with pytest.raises(ValueError):
        ls.init(ls.LinearProblem(np.eye(3), np.zeros(4)), alg)
    # integer promotion (src/common.jl:448-502)
    ci = ls.init(ls.LinearProblem(np.eye(3, dtype=np.int64), np.ones(3, dtype=np.int32)), alg)
    assert ci.A.dtype == np.float64 and ci.b.dtype == np.float64
    with pytest.raises(NotImplementedError):
        ls.init(ls.LinearProblem(np.eye(3), np.ones(3)))       # n = 3 -> reference CPU algorithm, out of scope


def test_block_diagonal_container(ls):
    bd = ls.BlockDiagonal([np.full((2, 2), 1.0), np.full((3, 3), 2.0)])
    assert bd.shape == (5, 5) and bd.all_square()
    D = bd.to_dense()
    assert D[0, 0] == 1 and D[4, 4] == 2 and D[0, 4] == 0
    assert list(bd.offsets) == [0, 2, 5]


def test_blockdiag_plan_classes(ls):
    """ragged BlockDiagonal blocks (reference case [2, 3, 4], test/Core/basictests.jl:1168-1223, and the
    variable-size supernode blocks of SURVEY 8(f)3): blocks <= 160 share the batched launch of their
    size class, padded to its largest member; blocks above 160 rows go one by one; empty blocks vanish"""
    assert ls.plan_blockdiag([3, 3, 3, 3]) == [("batched", [0, 1, 2, 3], 3)]
    assert ls.plan_blockdiag([2, 3, 4]) == [("batched", [0, 1, 2], 4)]
    plan = ls.plan_blockdiag([64, 5, 0, 70, 17, 33, 16, 32, 1000, 96, 130, 161])
    assert plan == [("batched", [1, 6], 16), ("batched", [4, 7], 32), ("batched", [0, 5], 64),
                    ("batched", [3, 9], 96), ("batched", [10], 130), ("single", [8], 1000), ("single", [11], 161)]
    assert ls.plan_blockdiag([]) == [] and ls.plan_blockdiag([0, 0]) == []
    # uniform sizes are never padded
    assert ls.plan_blockdiag([64] * 7) == [("batched", list(range(7)), 64)]


def test_blockdiag_padding_is_exact(ls, oracle):
    """diag(B, I) has the pivots, info and leading factors of B itself, bit for bit, and solving it
    with a zero-padded right-hand side gives B's solution — checked with the oracle's per-block
    lu!/ldiv! restatement (ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205)"""
    rng = np.random.default_rng(7)
    sizes = [2, 3, 4, 7, 1, 4]
    blocks = [rng.standard_normal((k, k)) for k in sizes]
    blocks[3][:, 2] = 0.0                              # singular block: zero column 3 => info = 3
    blocks[5] = np.eye(4)[[2, 0, 3, 1], :]             # permutation block
    (kind, idx, m), = ls.plan_blockdiag(sizes)
    assert kind == "batched" and m == 7
    P = ls.pad_blocks(blocks, idx, m, np.float64)
    assert P.shape == (6, 7, 7)
    rhs = np.zeros((6, m))
    for s, k in enumerate(sizes):
        rhs[s, :k] = rng.random(k)
    Fp, ipiv_p, info_p, x_p = oracle.ref_batched(P, rhs)
    for s, k in enumerate(sizes):
        Bs = np.ascontiguousarray(blocks[s].T)[None]   # [1, col, row]
        F, ipiv, info, x = oracle.ref_batched(Bs, rhs[s:s + 1, :k].copy())
        assert np.array_equal(ipiv_p[s, :k], ipiv[0]) and info_p[s] == info[0]
        assert np.array_equal(Fp[s, :k, :k], F[0])
        assert np.array_equal(ipiv_p[s, k:], np.arange(k + 1, m + 1))      # the padding never interchanges
        if info[0] == 0:
            assert np.array_equal(x_p[s, :k], x[0]) and not x_p[s, k:].any()
    assert info_p[3] == 3 and not info_p[[0, 1, 2, 4, 5]].any()


def _header_prototypes():
    """name -> list of C parameter type strings, parsed from include/b200lu.h"""
    hdr = open(os.path.join(ROOT, "include", "b200lu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(b200lu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        args = " ".join(m.group(2).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[m.group(1)] = params
    return protos


def _ctype_class(c_param):
    """coarse class of a C parameter: pointer / int64 / uint64 / int / char / double"""
    t = c_param.rsplit(" ", 1)[0] if " " in c_param else c_param
    if "*" in c_param:
        return "ptr"
    for key, cls in (("uint64_t", "u64"), ("int64_t", "i64"), ("double", "f64"), ("char", "char"), ("int", "int")):
        if key in t:
            return cls
    raise AssertionError(f"unclassified parameter {c_param!r}")


def test_ctypes_binding_matches_header_prototypes(ls):
    """every argtypes list in _capi.py has the arity and the coarse types of the C prototype"""
    protos = _header_prototypes()
    lib = ls._capi.load()
    cls_of = {ctypes.c_int64: "i64", ctypes.c_uint64: "u64", ctypes.c_int: "int", ctypes.c_double: "f64",
              ctypes.c_char: "char", ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr"}
    assert sorted(protos) == sorted(ls._capi.SYMBOLS)
    for name, params in protos.items():
        argtypes = getattr(lib, name).argtypes
        assert argtypes is not None and len(argtypes) == len(params), (name, params, argtypes)
        for ct, cp in zip(argtypes, params):
            got = cls_of.get(ct, "ptr")       # POINTER(...) types are pointers
            assert got == _ctype_class(cp), (name, cp, ct)


def test_julia_glue_ccalls_match_header_prototypes():
    """the Julia file a maintainer adds (not runnable here) binds only symbols the header declares,
    with the right number of arguments and pointer / integer / char / double in the right places"""
    protos = _header_prototypes()
    src = open(os.path.join(ROOT, "linearsolve.jl_b200", "julia", "B200LUFactorization.jl")).read()
    jl_cls = {"Int64": "i64", "UInt64": "u64", "Cint": "int", "UInt8": "char", "Cdouble": "f64", "Float64": "f64"}
    calls = re.findall(r"ccall\(\(:(b200lu_[a-z0-9_]+),\s*libb200lu\[\]\),\s*(\w+),\s*\(([^)]*)\)", src)
    assert len(calls) >= 8
    seen = set()
    for name, ret, types in calls:
        assert name in protos, f"{name} is not declared in include/b200lu.h"
        seen.add(name)
        tl = [t.strip() for t in types.split(",") if t.strip()]
        params = protos[name]
        assert len(tl) == len(params), (name, tl, params)
        for jt, cp in zip(tl, params):
            want = _ctype_class(cp)
            got = "ptr" if jt.startswith(("Ptr{", "Ref{")) or jt == "Cstring" else jl_cls[jt]
            assert got == want, (name, jt, cp)
    # the FFI surface INTEGRATION.md names is actually bound
    for must in ("b200lu_create", "b200lu_destroy", "b200lu_factor", "b200lu_solve", "b200lu_factor_batched",
                 "b200lu_solve_batched_trans", "b200lu_residual_norms", "b200lu_last_error", "b200lu_set_option"):
        assert must in seen, must


class _OracleHandle:
    """Stand-in for _capi.Handle backed by the CPU oracle: lets the BlockDiagonal host logic (grouping,
    padding, packing of right-hand sides, adjoint wiring, retcodes) run without a GPU.  Test
    infrastructure only — the product has no such path."""
    created = 0

    def __init__(self, dtype=0, device=0, devices=None):
        type(self).created += 1
        self.devices = devices
        self.dtype, self.np_dtype, self.n = dtype, np.float64, 0
        self.options = {}

    def set_option(self, opt, value):
        self.options[opt] = value

    def factor_batched(self, A):
        from oracle import lu_oracle
        self.A = np.array(A, copy=True)                      # [s, col, row]
        _, ipiv, info, _ = lu_oracle.ref_batched(self.A)
        self.b_batch, self.b_n = A.shape[0], A.shape[1]
        return ipiv, info

    def solve_batched(self, B, trans="N"):
        X = np.empty_like(B)
        for s in range(B.shape[0]):
            M = self.A[s].T if trans == "N" else self.A[s]
            X[s] = np.linalg.solve(M, B[s].T).T
        return X

    def factor(self, A, want_ipiv=True):
        self.nfactor = getattr(self, "nfactor", 0) + 1
        self.M, self.n = np.array(A, copy=True), A.shape[0]
        return None, int(np.linalg.matrix_rank(self.M) < self.n)

    def solve(self, B, out=None, trans="N"):
        X = np.linalg.solve(self.M if trans == "N" else self.M.T, B)
        if out is not None:
            out[...] = X
            return out
        return X

    def residual_norms(self, B, X):
        Bm, Xm = B.reshape(self.n, -1), X.reshape(self.n, -1)
        return np.linalg.norm(Bm - self.M @ Xm, axis=0), np.linalg.norm(Bm, axis=0)


def test_blockdiagonal_host_logic_with_oracle_backend(ls, monkeypatch):
    """ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205 through interface.py with the device
    replaced by the oracle: ragged blocks, vector and matrix right-hand sides, adjoint, a singular
    block => Failure with isfresh kept, handle reuse across refactorizations"""
    monkeypatch.setattr(ls._capi, "Handle", _OracleHandle)
    _OracleHandle.created = 0
    rng = np.random.default_rng(3)
    sizes = [2, 3, 4, 70, 130, 200, 0, 16]
    blocks = [rng.random((k, k)) + k * np.eye(k) for k in sizes]
    A = ls.BlockDiagonal(blocks)
    D = A.to_dense()
    n = sum(sizes)
    alg = ls.B200LUFactorization(throwerror=False, residualsafety=True)
    cache = ls.init(ls.LinearProblem(A, rng.random(n)), alg)
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success and not cache.isfresh
    np.testing.assert_allclose(sol.u, np.linalg.solve(D, cache.b), rtol=1e-11)
    plan = ls.plan_blockdiag(sizes)
    assert _OracleHandle.created == len(plan) == 4          # classes 16, 96, 160 and the 200-row block
    assert all(h.options.get(ls._capi.OPT_KEEP_A) == 1 for _, h, _, _ in cache.cacheval.groups)
    B = rng.random((n, 3))
    cache.b = B
    cache.u = np.zeros_like(B)
    np.testing.assert_allclose(ls.solve_(cache).u, np.linalg.solve(D, B), rtol=1e-11)
    np.testing.assert_allclose(ls.solve_(cache, adjoint=True).u, np.linalg.solve(D.T, B), rtol=1e-11)
    # new values, same block sizes: the handles are reused
    cache.A = ls.BlockDiagonal([2.0 * b for b in blocks])
    np.testing.assert_allclose(ls.solve_(cache).u, np.linalg.solve(2.0 * D, B), rtol=1e-11)
    assert _OracleHandle.created == 4
    # a singular block anywhere => Failure, isfresh stays set
    bad = [b.copy() for b in blocks]
    bad[2][:, 1] = 0.0
    cache.A = ls.BlockDiagonal(bad)
    assert ls.solve_(cache).retcode == ls.ReturnCode.Failure and cache.isfresh
    # different block sizes: a new plan
    cache.A = ls.BlockDiagonal(blocks[:3] + [np.eye(n - 9)])
    assert ls.solve_(cache).retcode == ls.ReturnCode.Success
    assert [g[0] for g in cache.cacheval.groups] == ["batched", "single"]


def test_dense_solve_protocol_with_oracle_backend(ls, monkeypatch):
    """solve! protocol of src/openblas.jl:362-459 through interface.py with the device replaced by the
    oracle: factor only when fresh, Failure keeps isfresh, getrs straight into cache.u, adjoint reuse,
    residual safety (src/factorization.jl:127-156) and the option wiring of both algorithms"""
    monkeypatch.setattr(ls._capi, "Handle", _OracleHandle)
    _OracleHandle.created = 0
    rng = np.random.default_rng(8)
    n = 40
    A = rng.random((n, n)) + n * np.eye(n)
    b = rng.random(n)
    alg = ls.B200LUFactorization(throwerror=False, residualsafety=True, nb=128, lookahead=False)
    cache = ls.init(ls.LinearProblem(A, b), alg)
    u_buffer = cache.u
    sol = ls.solve_(cache)
    assert sol.retcode == ls.ReturnCode.Success and not cache.isfresh
    assert sol.u is u_buffer                                   # written in place
    np.testing.assert_allclose(sol.u, np.linalg.solve(A, b), rtol=1e-12)
    h = cache.cacheval.handle
    assert h.options == {ls._capi.OPT_NB: 128, ls._capi.OPT_LOOKAHEAD: 0, ls._capi.OPT_KEEP_A: 1}
    cache.b = rng.random(n)                                    # re-solve only: no new handle, no refactor
    np.testing.assert_allclose(ls.solve_(cache).u, np.linalg.solve(A, cache.b), rtol=1e-12)
    np.testing.assert_allclose(ls.solve_(cache, adjoint=True).u, np.linalg.solve(A.T, cache.b), rtol=1e-12)
    assert _OracleHandle.created == 1 and h.nfactor == 1
    cache.A = np.ones((n, n))                                  # singular => Failure, isfresh stays set
    assert cache.isfresh
    assert ls.solve_(cache).retcode == ls.ReturnCode.Failure and cache.isfresh
    cache.A = A
    assert ls.solve_(cache).retcode == ls.ReturnCode.Success
    strict = ls.init(ls.LinearProblem(A, b), alg, abstol=0.0, reltol=1e-300)
    assert ls.solve_(strict).retcode == ls.ReturnCode.APosterioriSafetyFailure   # src/factorization.jl:150-153
    mixed = ls.B200LU32MixedLUFactorization(throwerror=False, refine=False)
    cm = ls.init(ls.LinearProblem(A, b), mixed)
    assert ls.solve_(cm).retcode == ls.ReturnCode.Success
    assert cm.cacheval.handle.dtype == ls._capi.MIXED and cm.cacheval.handle.options == {ls._capi.OPT_REFINE_MAXIT: 0}
    with pytest.raises(TypeError):
        ls.init(ls.LinearProblem(A.astype(np.float32), b.astype(np.float32)), mixed).alg.handle_dtype(np.float32)
    # host_register: the library is told to page-lock the cache's own copy of A (B200LU_OPT_HOST_REGISTER); default off
    ch = ls.init(ls.LinearProblem(A, b), ls.B200LUFactorization(throwerror=False, host_register=True))
    assert ls.solve_(ch).retcode == ls.ReturnCode.Success
    assert ch.cacheval.handle.options == {ls._capi.OPT_HOST_REGISTER: 1}
    with pytest.raises(ValueError):                            # the multi-GPU handle keeps no copy of A
        ls.B200LUFactorization(throwerror=False, residualsafety=True, devices=(0, 1))


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "512",
                          "--nrhs", "4", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_roofline_traffic_comes_from_a_committed_capture():
    """`roofline.traffic` is parsed from the newest ncu capture of the DGEMM under profiles/ (dram bytes read +
    written of ONE launch), with that launch's shape and algorithmic bytes beside it — not typed into bench.py."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tr = bench.newest_profile_traffic("*ncu_dgemm*metrics*.txt")
    assert tr is not None and os.path.exists(os.path.join(ROOT, tr["file"]))
    M, N, K = (int(v) for v in tr["shape"].split("x"))
    assert tr["algorithmic"] == 2 * M * N * 8 + (M + N) * K * 8
    # C is read and written once, the panels come from L2: within 1.25x of the algorithmic bytes
    assert tr["algorithmic"] <= tr["bytes"] <= 1.25 * tr["algorithmic"]
    assert bench.newest_profile_traffic("*no_such_capture*.txt") is None
