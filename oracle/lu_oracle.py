"""CPU oracle for the dense-LU path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``linearsolve.jl_b200``) never does: it fails loudly without its CUDA library.

Two oracles, both pinned in ``tests/test_oracle.py`` against the reference's own
known-answer cases:

* ``lapack_*``   — LAPACK ``?getrf/?getrs`` from scipy's OpenBLAS: the very
  arithmetic behind the reference's ``LUFactorization`` /
  ``OpenBLASLUFactorization`` (reference src/factorization.jl:632-637,
  src/openblas.jl:144-151,267-275).  This is the ipiv oracle.
* ``ref_*``      — ``liblu_oracle.so``: a C restatement of the reference's in-tree
  kernels ``generic_lufact!`` (src/generic_lufact.jl:71-141), the blocked
  ``_blocked_lufact!`` (src/blocked_lufact.jl:38-54,93-178,658-679,711-743) and
  ``_naive_lu_ldiv!`` (src/factorization.jl:433-491).

The reference itself (Julia) cannot run here: no Julia binary, no network.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/lu_oracle.c -> oracle/liblu_oracle.so with gcc."""
    so = os.path.join(_HERE, "liblu_oracle.so")
    src = os.path.join(_HERE, "lu_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"]
        )
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        for name in ("generic", "unblocked", "reference"):
            f = getattr(lib, f"oracle_{name}_lufact_d")
            f.restype, f.argtypes = i64, [p, i64, i64, i64, p]
        lib.oracle_blocked_lufact_d.restype = i64
        lib.oracle_blocked_lufact_d.argtypes = [p, i64, i64, i64, i64, p]
        lib.oracle_blocked_lufact_s.restype = i64
        lib.oracle_blocked_lufact_s.argtypes = [p, i64, i64, i64, i64, p]
        lib.oracle_reference_lufact_s.restype = i64
        lib.oracle_reference_lufact_s.argtypes = [p, i64, i64, i64, p]
        for f in (lib.oracle_lu_ldiv_d, lib.oracle_lu_ldiv_s):
            f.restype, f.argtypes = None, [p, i64, i64, p, p, i64, i64, ctypes.c_int]
        lib.oracle_batched_lufact_solve_d.restype = i64
        lib.oracle_batched_lufact_solve_d.argtypes = [p, i64, i64, p, p, p, i64]
        _LIB = lib
    return _LIB


def _f(a, dtype):
    return np.asfortranarray(np.array(a, dtype=dtype, copy=True))


# ----------------------------------------------------------------- ref_* ----
def ref_lufact(A, variant: str = "reference", nb: int | None = None):
    """In-tree reference kernel. Returns (factors, ipiv[1-based int64], info)."""
    A = np.asarray(A)
    dt = np.float32 if A.dtype == np.float32 else np.float64
    F = _f(A, dt)
    m, n = F.shape
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    lib = _lib()
    sfx = "s" if dt == np.float32 else "d"
    lda = max(1, F.strides[1] // F.itemsize) if F.ndim == 2 and n > 0 else max(1, m)
    if variant == "blocked":
        fn = getattr(lib, f"oracle_blocked_lufact_{sfx}")
        info = fn(F.ctypes.data, lda, m, n, int(nb), ipiv.ctypes.data)
    elif variant in ("generic", "unblocked"):
        if sfx == "s":
            raise ValueError("float32 restatement exists for blocked/reference only")
        fn = getattr(lib, f"oracle_{variant}_lufact_d")
        info = fn(F.ctypes.data, lda, m, n, ipiv.ctypes.data)
    else:
        fn = getattr(lib, f"oracle_reference_lufact_{sfx}")
        info = fn(F.ctypes.data, lda, m, n, ipiv.ctypes.data)
    return F, ipiv, int(info)


def ref_ldiv(factors, ipiv, B):
    """`_naive_lu_ldiv!`: vector form for 1-D B, matrix form for 2-D B."""
    F = np.asfortranarray(factors)
    dt = F.dtype
    n = F.shape[0]
    X = _f(B, dt)
    matrix_form = X.ndim == 2
    Xm = X.reshape(n, -1, order="F") if not matrix_form else X
    Xm = np.asfortranarray(Xm)
    ip = np.ascontiguousarray(ipiv, dtype=np.int64)
    fn = _lib().oracle_lu_ldiv_s if dt == np.float32 else _lib().oracle_lu_ldiv_d
    fn(F.ctypes.data, max(1, n), n, ip.ctypes.data, Xm.ctypes.data, max(1, n), Xm.shape[1], int(matrix_form))
    return Xm if matrix_form else Xm[:, 0].copy()


def ref_batched(A, B=None):
    """Per-block lu! + ldiv! over A[batch, n, n] given as Fortran blocks
    (A[s] column-major).  Returns (factors, ipiv, info, X)."""
    A = np.array(A, dtype=np.float64, copy=True)  # [batch][col][row] layout: A[s].T is the matrix
    batch, n, _ = A.shape
    ipiv = np.zeros((batch, n), dtype=np.int64)
    info = np.zeros(batch, dtype=np.int64)
    X = None if B is None else np.array(B, dtype=np.float64, copy=True)
    nrhs = 0 if X is None else (1 if X.ndim == 2 else X.shape[1])
    _lib().oracle_batched_lufact_solve_d(
        A.ctypes.data, n, batch, ipiv.ctypes.data, info.ctypes.data,
        None if X is None else X.ctypes.data, nrhs)
    return A, ipiv, info, X


# -------------------------------------------------------------- lapack_* ----
def lapack_getrf(A):
    """LAPACK getrf via scipy (OpenBLAS). Returns (lu, ipiv 1-based int64, info)."""
    from scipy.linalg import lapack

    A = np.asarray(A)
    fn = lapack.sgetrf if A.dtype == np.float32 else lapack.dgetrf
    lu, piv, info = fn(np.asfortranarray(A), overwrite_a=False)
    return lu, piv.astype(np.int64) + 1, int(info)


def lapack_getrs(lu, ipiv, B, trans: int = 0):
    from scipy.linalg import lapack

    fn = lapack.sgetrs if lu.dtype == np.float32 else lapack.dgetrs
    x, info = fn(lu, (np.asarray(ipiv) - 1).astype(np.int32), np.asfortranarray(B), trans=trans)
    assert info == 0
    return x


# --------------------------------------------------------------- metrics ----
def scaled_residual(A, factors, ipiv):
    """||P A - L U||_1 / (||L||_1 ||U||_1 eps max(m,n)); the reference's test
    metric (test/Core/blocked_lufact.jl:8-28), bound < 20 there."""
    A = np.asarray(A, dtype=np.float64)
    F = np.asarray(factors, dtype=np.float64)
    m, n = A.shape
    k = min(m, n)
    L = np.tril(F[:, :k], -1) + np.eye(m, k)
    U = np.triu(F[:k, :])
    PA = A.copy()
    for i, p in enumerate(np.asarray(ipiv) - 1):
        if p != i:
            PA[[i, p], :] = PA[[p, i], :]
    eps = np.finfo(np.asarray(factors).dtype).eps
    den = np.linalg.norm(L, 1) * np.linalg.norm(U, 1) * eps * max(m, n)
    return np.linalg.norm(PA - L @ U, 1) / den


def backward_error(A, x, b):
    """Normwise backward error ||A x - b|| / (||A|| ||x||) (north star), with
    Frobenius/2-norms; per column for matrix right-hand sides (max returned)."""
    A = np.asarray(A, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    r = A @ x - b
    na = np.linalg.norm(A)
    if x.ndim == 1:
        return np.linalg.norm(r) / (na * np.linalg.norm(x))
    return float(np.max(np.linalg.norm(r, axis=0) / (na * np.linalg.norm(x, axis=0))))


def ipiv_to_perm(ipiv):
    ip = np.asarray(ipiv) - 1
    perm = np.arange(len(ip))
    for i, p in enumerate(ip):
        if p != i:
            perm[i], perm[p] = perm[p], perm[i]
    return perm


def compare_ipiv(A, ipiv_test, ipiv_ref, factors_ref=None, tie_rtol: float = None):
    """ipiv parity "bit-exact except where ties occur" (north star).

    Returns (n_equal_prefix, status) with status in {"exact", "tie", "mismatch"}.
    A first difference at step k is an admissible TIE iff, in the ORACLE's own
    partially factored column k, the two candidate rows' magnitudes agree to
    within tie_rtol * growth (rounding level); after an admissible tie the two
    factorizations legitimately diverge and only the backward error is compared.
    """
    it = np.asarray(ipiv_test)
    ir = np.asarray(ipiv_ref)
    neq = np.nonzero(it != ir)[0]
    if len(neq) == 0:
        return len(ir), "exact"
    k = int(neq[0])
    A = np.asarray(A)
    dt = A.dtype
    eps = np.finfo(dt).eps
    if tie_rtol is None:
        tie_rtol = 64 * eps * max(A.shape)
    # recompute the oracle's state at step k: apply the first k steps of LAPACK-style LU
    W = np.array(A, dtype=np.float64, order="F", copy=True)
    n = W.shape[0]
    for j in range(k):
        p = ir[j] - 1
        if p != j:
            W[[j, p], :] = W[[p, j], :]
        if W[j, j] != 0:
            W[j + 1:, j] /= W[j, j]
            W[j + 1:, j + 1:] -= np.outer(W[j + 1:, j], W[j, j + 1:])
    col = np.abs(W[:, k])
    a_ref, a_test = col[ir[k] - 1], col[it[k] - 1]
    scale = max(np.max(np.abs(W[k:, k:])), np.finfo(np.float64).tiny)
    if abs(a_ref - a_test) <= tie_rtol * scale:
        return k, "tie"
    return k, "mismatch"
