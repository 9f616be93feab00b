"""ctypes binding of libb200lu.so (include/b200lu.h).

This is the Python twin of the `ccall` layer a Julia maintainer would write
(see INTEGRATION.md and julia/B200LUFactorization.jl); it mirrors how the
reference binds LAPACK in src/openblas.jl:81-311 — thin, no logic, integer
status codes turned into exceptions.  There is NO fallback: if the shared
library is missing or no B200 is visible, `load()` / `Handle()` raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200LU_LIB: an instrumented build of the same library (e.g. -DPCL_TIMING for the panel kernels' clock64 stamps)
LIB_PATH = os.environ.get("B200LU_LIB") or os.path.join(_HERE, "csrc", "libb200lu.so")

F64, F32, MIXED = 0, 1, 2
T_H2D, T_FACTOR, T_SOLVE, T_D2H, T_GEMM, T_PANEL, T_LOOKAHEAD, T_PUSH = range(8)
OPT_NB, OPT_LOOKAHEAD, OPT_REFINE_MAXIT, OPT_PANEL_CTAS, OPT_SOLVE_NRHS_TILE, OPT_PROFILE, OPT_PANEL_RPT, OPT_GEMM_CFG, OPT_PANEL_MODE, OPT_SGEMM_MODE, OPT_TRSV_MODE, OPT_STREAM_H2D, OPT_MAPPED_RHS, OPT_KEEP_A, OPT_BATCHED_MODE, OPT_HOST_REGISTER = range(16)
C_GEMM_FLOPS, C_GEMM_LAUNCHES, C_REFINE_ITERS = range(3)
PEAK_FP64_DMMA, PEAK_FP64_DFMA, PEAK_HBM_COPY = range(3)

# every symbol include/b200lu.h declares (tests check the export list against this)
SYMBOLS = [
    "b200lu_version", "b200lu_launch_count", "b200lu_create", "b200lu_destroy",
    "b200lu_last_error", "b200lu_last_timing", "b200lu_last_counter", "b200lu_probe_peak", "b200lu_debug_gemm_sub",
    "b200lu_set_option", "b200lu_get_option",
    "b200lu_factor", "b200lu_solve", "b200lu_factor_device", "b200lu_solve_device",
    "b200lu_get_factors", "b200lu_get_ipiv", "b200lu_residual_norms",
    "b200lu_factor_batched", "b200lu_solve_batched", "b200lu_factor_batched_device",
    "b200lu_solve_batched_device", "b200lu_get_factors_batched",
    "b200lu_solve_batched_trans", "b200lu_solve_batched_trans_device",
    "b200lu_factor_solve_batched", "b200lu_factor_solve_batched_device",
    "b200lu_comm_unique_id", "b200lu_comm_init", "b200lu_dist_local_cols", "b200lu_dist_transport",
    "b200lu_factor_dist", "b200lu_solve_dist", "b200lu_fill_uniform_device",
]


class B200LUError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"b200lu status {status}: {msg}")
        self.status = status


_lib = None


def load():
    """dlopen libb200lu.so and declare prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LUError(
            -1000, f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    i64, vp, ci, cd = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    pi64 = ctypes.POINTER(ctypes.c_int64)
    P = lambda name, res, args: (setattr(getattr(lib, name), "restype", res),
                                 setattr(getattr(lib, name), "argtypes", args))
    P("b200lu_version", ci, [])
    P("b200lu_launch_count", i64, [])
    P("b200lu_create", ci, [ctypes.POINTER(vp), ci, ci, ctypes.POINTER(ci)])
    P("b200lu_destroy", None, [vp])
    P("b200lu_last_error", ctypes.c_char_p, [vp])
    P("b200lu_last_timing", cd, [vp, ci])
    P("b200lu_last_counter", cd, [vp, ci])
    P("b200lu_probe_peak", ci, [vp, ci, ctypes.POINTER(cd)])
    P("b200lu_debug_gemm_sub", ci, [vp, i64, i64, i64, vp, i64, vp, i64, vp, i64])
    P("b200lu_set_option", ci, [vp, ci, i64])
    P("b200lu_get_option", i64, [vp, ci])
    P("b200lu_factor", ci, [vp, i64, vp, i64, vp, pi64])
    P("b200lu_solve", ci, [vp, ctypes.c_char, i64, vp, i64, vp, i64])
    P("b200lu_factor_device", ci, [vp, i64, vp, i64, pi64])
    P("b200lu_solve_device", ci, [vp, ctypes.c_char, i64, vp, i64, vp, i64])
    P("b200lu_residual_norms", ci, [vp, i64, vp, i64, vp, i64, vp, vp])
    P("b200lu_get_factors", ci, [vp, vp, i64])
    P("b200lu_get_ipiv", ci, [vp, vp])
    P("b200lu_factor_batched", ci, [vp, i64, i64, vp, i64, i64, vp, vp])
    P("b200lu_solve_batched", ci, [vp, i64, vp, i64, i64, vp, i64, i64])
    P("b200lu_factor_batched_device", ci, [vp, i64, i64, vp, i64, i64, pi64])
    P("b200lu_solve_batched_device", ci, [vp, i64, vp, i64, i64, vp, i64, i64])
    P("b200lu_solve_batched_trans", ci, [vp, ctypes.c_char, i64, vp, i64, i64, vp, i64, i64])
    P("b200lu_solve_batched_trans_device", ci, [vp, ctypes.c_char, i64, vp, i64, i64, vp, i64, i64])
    P("b200lu_get_factors_batched", ci, [vp, vp, i64, i64, vp, vp])
    P("b200lu_factor_solve_batched", ci, [vp, i64, i64, vp, i64, i64, vp, i64, vp, i64, vp, vp])
    P("b200lu_factor_solve_batched_device", ci, [vp, i64, i64, vp, i64, i64, vp, i64, vp, i64, pi64])
    P("b200lu_comm_unique_id", ci, [vp])
    P("b200lu_comm_init", ci, [vp, vp, ci, ci])
    P("b200lu_dist_local_cols", ci, [vp, i64, pi64])
    P("b200lu_dist_transport", ci, [vp])
    P("b200lu_factor_dist", ci, [vp, i64, vp, i64, pi64])
    P("b200lu_solve_dist", ci, [vp, i64, vp, i64, vp, i64])
    P("b200lu_fill_uniform_device", ci, [vp, vp, i64, i64, i64, i64, i64, i64, ctypes.c_uint64, cd])
    _lib = lib
    return lib


def is_available() -> bool:
    """`useb200()` hook (pattern: reference src/LinearSolve.jl:741-743): library
    built AND a usable device present.  Never raises."""
    try:
        lib = load()
        h = ctypes.c_void_p()
        rc = lib.b200lu_create(ctypes.byref(h), F64, 1, None)
        if rc != 0:
            return False
        lib.b200lu_destroy(h)
        return True
    except Exception:
        return False


def launch_count() -> int:
    return int(load().b200lu_launch_count())


_NP = {F64: np.float64, F32: np.float32, MIXED: np.float64}
_NPF = {F64: np.float64, F32: np.float32, MIXED: np.float32}


class Handle:
    """RAII wrapper of a b200lu_handle (freed like the reference frees AMGX
    handles by finalizer, src/extension_algs.jl:1555-1556)."""

    def __init__(self, dtype=F64, device=0, devices=None):
        """`devices`: a sequence of >= 2 device indices creates ONE handle that drives all of them from
        this process (b200lu_create with ngpus > 1); otherwise a single-GPU handle on `device`."""
        self.lib = load()
        self.dtype = dtype
        self.np_dtype = _NP[dtype]          # interface element type
        self.factor_dtype = _NPF[dtype]     # element type of the stored factors
        self._h = ctypes.c_void_p()
        devs = [int(d) for d in devices] if devices is not None and len(devices) > 1 else [int(device)]
        self.ngpus = len(devs)
        dev = (ctypes.c_int * len(devs))(*devs)
        rc = self.lib.b200lu_create(ctypes.byref(self._h), dtype, len(devs), dev)
        if rc != 0:
            self._h = None
            why = ("the devices are not all peers of each other" if rc == 5 else
                   "no usable sm_100 CUDA device (this library has no CPU fallback)")
            raise B200LUError(rc, f"b200lu_create(ngpus={len(devs)}) failed: {why}")
        self.n = 0

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200lu_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise B200LUError(rc, self.lib.b200lu_last_error(self._h).decode())

    def set_option(self, opt, value):
        self._check(self.lib.b200lu_set_option(self._h, opt, int(value)))

    def get_option(self, opt):
        return int(self.lib.b200lu_get_option(self._h, opt))

    def timing(self, phase):
        return float(self.lib.b200lu_last_timing(self._h, phase))

    def counter(self, which):
        return float(self.lib.b200lu_last_counter(self._h, which))

    def debug_gemm_sub(self, M, N, K, dA, lda, dB, ldb, dC, ldc):
        """C -= A @ B on device pointers with the handle's trailing-update kernel (test hook)."""
        self._check(self.lib.b200lu_debug_gemm_sub(self._h, M, N, K, dA, lda, dB, ldb, dC, ldc))

    def probe_peak(self, kind):
        out = ctypes.c_double(0.0)
        self._check(self.lib.b200lu_probe_peak(self._h, kind, ctypes.byref(out)))
        return float(out.value)

    # ---- host-buffer calls (what a `ccall` from Julia does) -----------------
    def _colmajor(self, A):
        A = np.asarray(A)
        if A.dtype != self.np_dtype:
            raise TypeError(f"expected {self.np_dtype}, got {A.dtype}")
        if A.ndim == 2 and not A.flags.f_contiguous:
            # strided views: unit row stride with a wider leading dimension is accepted as is
            if A.strides[0] == A.itemsize and A.strides[1] >= A.shape[0] * A.itemsize:
                return A, A.strides[1] // A.itemsize
            A = np.asfortranarray(A)
        lda = max(1, A.strides[1] // A.itemsize) if (A.ndim == 2 and A.shape[1] > 1) else max(1, A.shape[0])
        return A, lda

    def factor(self, A, want_ipiv=True):
        A, lda = self._colmajor(A)
        n = A.shape[0]
        if A.ndim != 2 or A.shape[1] != n:
            raise ValueError("square matrix required")
        ipiv = np.zeros(n, dtype=np.int64) if want_ipiv else None
        info = ctypes.c_int64(0)
        rc = self.lib.b200lu_factor(self._h, n, A.ctypes.data, lda,
                                    ipiv.ctypes.data if want_ipiv else None, ctypes.byref(info))
        self._check(rc)
        self.n = n
        return ipiv, int(info.value)

    def solve(self, B, out=None, trans="N"):
        B = np.asarray(B)
        if B.dtype != self.np_dtype:
            raise TypeError(f"expected {self.np_dtype}, got {B.dtype}")
        vec = B.ndim == 1
        Bm = np.asfortranarray(B.reshape(self.n, -1, order="F") if vec else B)
        nrhs = Bm.shape[1]
        X = np.empty_like(Bm, order="F") if out is None else out.reshape(self.n, -1, order="F")
        ld = max(1, self.n)
        rc = self.lib.b200lu_solve(self._h, trans.encode(), nrhs, Bm.ctypes.data, ld, X.ctypes.data, ld)
        self._check(rc)
        return X[:, 0] if vec else X

    def residual_norms(self, B, X):
        """(||B[:, c] - A X[:, c]||_2, ||B[:, c]||_2) per column, computed on the device with the copy of
        A kept by OPT_KEEP_A (or the FP64 copy of a MIXED handle)."""
        B, X = np.asarray(B), np.asarray(X)
        if B.dtype != self.np_dtype or X.dtype != self.np_dtype:
            raise TypeError(f"expected {self.np_dtype}, got {B.dtype} / {X.dtype}")
        Bm = np.asfortranarray(B.reshape(self.n, -1, order="F"))
        Xm = np.asfortranarray(X.reshape(self.n, -1, order="F"))
        nrhs = Bm.shape[1]
        resid = np.zeros(nrhs)
        bnorm = np.zeros(nrhs)
        ld = max(1, self.n)
        self._check(self.lib.b200lu_residual_norms(self._h, nrhs, Bm.ctypes.data, ld, Xm.ctypes.data, ld,
                                                   resid.ctypes.data, bnorm.ctypes.data))
        return resid, bnorm

    def get_factors(self):
        LU = np.zeros((self.n, self.n), dtype=self.factor_dtype, order="F")
        self._check(self.lib.b200lu_get_factors(self._h, LU.ctypes.data, max(1, self.n)))
        return LU

    def get_ipiv(self):
        ipiv = np.zeros(self.n, dtype=np.int64)
        self._check(self.lib.b200lu_get_ipiv(self._h, ipiv.ctypes.data))
        return ipiv

    # ---- device-pointer calls ------------------------------------------------
    def factor_device(self, ptr, n, lda):
        info = ctypes.c_int64(0)
        self._check(self.lib.b200lu_factor_device(self._h, n, ctypes.c_void_p(ptr), lda, ctypes.byref(info)))
        self.n = n
        return int(info.value)

    def solve_device(self, b_ptr, ldb, x_ptr, ldx, nrhs=1, trans="N"):
        self._check(self.lib.b200lu_solve_device(self._h, trans.encode(), nrhs, ctypes.c_void_p(b_ptr), ldb,
                                                 ctypes.c_void_p(x_ptr), ldx))

    def fill_uniform_device(self, ptr, lda, n, ncols, seed, diag_shift=0.0, first_global_col=0,
                            col_block=None, col_block_stride=None):
        cb = ncols if col_block is None else col_block
        cs = cb if col_block_stride is None else col_block_stride
        self._check(self.lib.b200lu_fill_uniform_device(self._h, ctypes.c_void_p(ptr), lda, n, ncols,
                                                        first_global_col, cb, cs, seed, diag_shift))

    # ---- one-process-per-GPU distributed mode (1D block-cyclic columns) --------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = load().b200lu_comm_unique_id(buf)
        if rc != 0:
            raise B200LUError(rc, "b200lu_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, id_bytes: bytes, rank: int, nranks: int):
        buf = ctypes.create_string_buffer(id_bytes or b"", 128)
        self._check(self.lib.b200lu_comm_init(self._h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def dist_transport(self) -> str:
        """how the factored panel travels between the ranks: 'p2p' (peer stores into mapped windows) or 'nccl'"""
        return {1: "p2p", 0: "nccl"}.get(int(self.lib.b200lu_dist_transport(self._h)), "none")

    def dist_local_cols(self, n: int) -> int:
        out = ctypes.c_int64(0)
        self._check(self.lib.b200lu_dist_local_cols(self._h, n, ctypes.byref(out)))
        return int(out.value)

    def factor_dist(self, ptr, n, lda):
        info = ctypes.c_int64(0)
        self._check(self.lib.b200lu_factor_dist(self._h, n, ctypes.c_void_p(ptr), lda, ctypes.byref(info)))
        self.n = n
        return int(info.value)

    def solve_dist(self, b_ptr, ldb, x_ptr, ldx, nrhs=1):
        self._check(self.lib.b200lu_solve_dist(self._h, nrhs, ctypes.c_void_p(b_ptr), ldb,
                                               ctypes.c_void_p(x_ptr), ldx))

    # ---- batched -------------------------------------------------------------
    def factor_batched(self, A):
        """A: (batch, n, n) with A[s] holding system s COLUMN-major, i.e. the
        array element [s, j, i] is entry (i, j) of system s."""
        A = np.ascontiguousarray(A)
        if A.dtype != self.np_dtype:
            raise TypeError(f"expected {self.np_dtype}, got {A.dtype}")
        batch, n, n2 = A.shape
        assert n == n2
        ipiv = np.zeros((batch, n), dtype=np.int64)
        info = np.zeros(batch, dtype=np.int64)
        self._check(self.lib.b200lu_factor_batched(self._h, batch, n, A.ctypes.data, n, n * n,
                                                   ipiv.ctypes.data, info.ctypes.data))
        self.b_batch, self.b_n = batch, n
        return ipiv, info

    def factor_solve_batched(self, A, b):
        """getrf of every system AND getrs of its right-hand side b[s] in one kernel (n <= 64); the
        factors stay cached.  A: (batch, n, n) column-major per system, b: (batch, n).  Returns (x, ipiv, info)."""
        A = np.ascontiguousarray(A)
        b = np.ascontiguousarray(b)
        if A.dtype != self.np_dtype or b.dtype != self.np_dtype:
            raise TypeError(f"expected {self.np_dtype}, got {A.dtype} / {b.dtype}")
        batch, n, n2 = A.shape
        assert n == n2 and b.shape == (batch, n)
        ipiv = np.zeros((batch, n), dtype=np.int64)
        info = np.zeros(batch, dtype=np.int64)
        x = np.empty_like(b)
        self._check(self.lib.b200lu_factor_solve_batched(self._h, batch, n, A.ctypes.data, n, n * n, b.ctypes.data, n,
                                                         x.ctypes.data, n, ipiv.ctypes.data, info.ctypes.data))
        self.b_batch, self.b_n = batch, n
        return x, ipiv, info

    def factor_solve_batched_device(self, a_ptr, b_ptr, x_ptr, batch, n):
        bad = ctypes.c_int64(0)
        self._check(self.lib.b200lu_factor_solve_batched_device(self._h, batch, n, ctypes.c_void_p(a_ptr), n, n * n,
                                                                ctypes.c_void_p(b_ptr), n, ctypes.c_void_p(x_ptr), n,
                                                                ctypes.byref(bad)))
        self.b_batch, self.b_n = batch, n
        return int(bad.value)

    def solve_batched(self, B, trans="N"):
        """B: (batch, n) or (batch, nrhs, n) (each right-hand side contiguous).
        trans = 'T'/'C': op(A_i) = A_i^T with the same cached factors."""
        B = np.ascontiguousarray(B)
        if B.dtype != self.np_dtype:
            raise TypeError(f"expected {self.np_dtype}, got {B.dtype}")
        vec = B.ndim == 2
        Bm = B.reshape(self.b_batch, 1, self.b_n) if vec else B
        nrhs = Bm.shape[1]
        X = np.empty_like(Bm)
        n = self.b_n
        if trans in ("N", "n"):
            self._check(self.lib.b200lu_solve_batched(self._h, nrhs, Bm.ctypes.data, n, n * nrhs,
                                                      X.ctypes.data, n, n * nrhs))
        else:
            self._check(self.lib.b200lu_solve_batched_trans(self._h, trans.encode(), nrhs, Bm.ctypes.data, n,
                                                            n * nrhs, X.ctypes.data, n, n * nrhs))
        return X.reshape(B.shape)

    def get_factors_batched(self):
        LU = np.zeros((self.b_batch, self.b_n, self.b_n), dtype=self.factor_dtype)
        ipiv = np.zeros((self.b_batch, self.b_n), dtype=np.int64)
        info = np.zeros(self.b_batch, dtype=np.int64)
        n = self.b_n
        self._check(self.lib.b200lu_get_factors_batched(self._h, LU.ctypes.data, n, n * n,
                                                        ipiv.ctypes.data, info.ctypes.data))
        return LU, ipiv, info

    def factor_batched_device(self, ptr, batch, n, lda=None, stride=None):
        bad = ctypes.c_int64(0)
        lda = n if lda is None else lda
        stride = n * n if stride is None else stride
        self._check(self.lib.b200lu_factor_batched_device(self._h, batch, n, ctypes.c_void_p(ptr), lda, stride,
                                                          ctypes.byref(bad)))
        self.b_batch, self.b_n = batch, n
        return int(bad.value)

    def solve_batched_device(self, b_ptr, x_ptr, nrhs=1, trans="N"):
        n = self.b_n
        if trans in ("N", "n"):
            self._check(self.lib.b200lu_solve_batched_device(self._h, nrhs, ctypes.c_void_p(b_ptr), n, n * nrhs,
                                                             ctypes.c_void_p(x_ptr), n, n * nrhs))
        else:
            self._check(self.lib.b200lu_solve_batched_trans_device(self._h, trans.encode(), nrhs,
                                                                   ctypes.c_void_p(b_ptr), n, n * nrhs,
                                                                   ctypes.c_void_p(x_ptr), n, n * nrhs))
