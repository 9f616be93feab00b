"""Host-side mirror of the reference's SciML interface for the dense-LU path.

The reference is Julia; no Julia toolchain exists in this environment, so the
host side above the C ABI is written in Python with the SAME names, argument
meaning and error behaviour as the reference, so that the parity tests read
like the reference's own tests.  The file a Julia maintainer would add is
`julia/B200LUFactorization.jl` (same logic, `ccall` instead of ctypes).

Mirrored (reference file:line):
  LinearProblem(A, b; u0)                        SciMLBase, src/LinearSolve.jl:21-22
  init(prob, alg; alias_A, alias_b, ...)         src/common.jl:700-712,758-931
  LinearCache + `cache.A =` / `cache.b =`        src/common.jl:281-306,313-360
  solve!(cache) / solve(prob, alg)               src/common.jl:966-1017
  reinit!(cache; A, b)                           src/common.jl:933-964
  ReturnCode.Success / Failure / APosterioriSafetyFailure   src/factorization.jl:714-722,150-153
  B200LUFactorization <: AbstractFactorization   (new; shaped like CudaOffloadLUFactorization,
                                                  src/extension_algs.jl:344-354, and the solve!
                                                  protocol of OpenBLASLUFactorization,
                                                  src/openblas.jl:362-459)
  B200LU32MixedLUFactorization                   (like OpenBLAS32MixedLUFactorization,
                                                  src/openblas.jl:470-543, plus FP64 refinement)
  BlockDiagonal + blockwise LU                   ext/LinearSolveBlockDiagonalsExt.jl:49-205
  defaultalg(A, b, assumptions)                  src/default.jl:411-520
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field

import numpy as np

from . import _capi


class ReturnCode(enum.Enum):
    """the members of SciMLBase.ReturnCode this path produces"""
    Default = 0
    Success = 1
    Failure = 2
    # the a-posteriori residual check failed (src/factorization.jl:150-153); the default solver
    # treats it like Failure and runs its QR fallback (src/default.jl:946-969)
    APosterioriSafetyFailure = 3


def successful_retcode(sol) -> bool:
    return sol.retcode == ReturnCode.Success


@dataclass
class OperatorAssumptions:
    """src/common.jl:184-201"""
    issq: bool = True
    condition: str = "IllConditioned"


class LinearProblem:
    def __init__(self, A, b, u0=None, p=None):
        self.A, self.b, self.u0, self.p = A, b, u0, p


@dataclass
class LinearSolution:
    u: np.ndarray
    retcode: ReturnCode
    alg: object
    resid: object = None
    iters: int = 0


# ------------------------------------------------------------- algorithms ----
class AbstractFactorization:
    """needs_concrete_A = true, needs_square_A = true (src/LinearSolve.jl:299,720-731)"""
    needs_concrete_A = True
    needs_square_A = True


class B200LUFactorization(AbstractFactorization):
    """Dense partially pivoted LU on one B200 through libb200lu.so.

    `throwerror=False` lets the default solver build the algorithm speculatively
    on machines without the library (src/extension_algs.jl:346-353).
    `residualsafety` enables the a-posteriori residual check of
    src/factorization.jl:127-156 after a fresh factorization.
    """
    _dtype_code = {np.dtype(np.float64): _capi.F64, np.dtype(np.float32): _capi.F32}

    def __init__(self, throwerror: bool = True, residualsafety: bool = False, device: int = 0,
                 nb: int | None = None, lookahead: bool | None = None, devices=None, host_register: bool = False):
        """`devices = (0, 1, ..., 7)`: ONE cache drives all these GPUs from this process (Julia:
        `B200LUFactorization(; devices = 0:7)`): a dense A is factored 1-D block-cyclic over them, the
        blocks of a BlockDiagonal are sharded over them.  Float32/Float64; `solve!(cache; adjoint = true)`
        and `residualsafety` need the single-GPU handle.
        `host_register = True`: the library page-locks the matrix it is handed (B200LU_OPT_HOST_REGISTER) — the
        cache's private copy of A with `alias_A = false`, src/common.jl:818-842 — once, so that every
        refactorization from that buffer gets the streamed upload a pinned buffer gets."""
        if throwerror and not useb200():
            raise RuntimeError("B200LUFactorization requires libb200lu.so and a B200 (sm_100) GPU; "
                               "there is no CPU fallback")
        self.residualsafety = residualsafety
        self.device = device
        self.devices = tuple(int(d) for d in devices) if devices is not None and len(devices) > 1 else None
        if self.devices is not None and residualsafety:
            raise ValueError("residualsafety needs the single-GPU handle (the multi-GPU handle keeps no copy of A)")
        self.nb = nb
        self.lookahead = lookahead
        self.host_register = bool(host_register)

    def handle_dtype(self, eltype):
        try:
            return self._dtype_code[np.dtype(eltype)]
        except KeyError:
            raise TypeError(f"B200LUFactorization supports Float32/Float64, got {eltype}") from None


class B200LU32MixedLUFactorization(B200LUFactorization):
    """FP64 interface, FP32 factorization on the GPU, FP64 iterative refinement.
    With `refine=False` it is exactly the reference's *32Mixed behaviour
    (cast, sgetrf, sgetrs, cast back; src/openblas.jl:487-543)."""

    def __init__(self, refine: bool = True, maxiters: int = 10, **kw):
        super().__init__(**kw)
        self.refine = refine
        self.maxiters = maxiters

    def handle_dtype(self, eltype):
        if np.dtype(eltype) != np.float64:
            raise TypeError("B200LU32MixedLUFactorization expects a Float64 problem")
        if self.devices is not None:
            raise TypeError("B200LU32MixedLUFactorization runs on one GPU")
        return _capi.MIXED


def useb200() -> bool:
    """availability hook, overridden when the library loads
    (pattern: src/LinearSolve.jl:741-743, ext/LinearSolveCUDAExt.jl:16)"""
    return _capi.is_available()


# -------------------------------------------------------- block diagonal ----
class BlockDiagonal:
    """Minimal stand-in for BlockDiagonals.BlockDiagonal (square blocks)."""

    def __init__(self, blocks):
        self.blocks = [np.asarray(B) for B in blocks]
        for B in self.blocks:
            if B.ndim != 2:
                raise ValueError("blocks must be matrices")
        self.offsets = np.concatenate([[0], np.cumsum([B.shape[0] for B in self.blocks])])

    @property
    def shape(self):
        return (int(self.offsets[-1]), int(sum(B.shape[1] for B in self.blocks)))

    @property
    def dtype(self):
        return self.blocks[0].dtype

    def all_square(self):
        return all(B.shape[0] == B.shape[1] for B in self.blocks)

    def to_dense(self):
        n = self.shape[0]
        M = np.zeros((n, n), dtype=self.dtype)
        for B, o in zip(self.blocks, self.offsets[:-1]):
            M[o:o + B.shape[0], o:o + B.shape[1]] = B
        return M

    def copy(self):
        return BlockDiagonal([B.copy() for B in self.blocks])


# ------------------------------------------------------------------ cache ----
class _B200LUCache:
    """cacheval: (handle, ipiv, info) — the analogue of OpenBLASLUCache
    (src/openblas.jl:312-316).  The factors stay on the device."""

    def __init__(self):
        self.handle = None
        self.ipiv = None
        self.info = 0
        self.groups = None  # BlockDiagonal: list of (kind, handle, block indices, padded size)
        self.group_sizes = None


class LinearCache:
    """src/common.jl:281-306.  Assigning `cache.A` marks the cache fresh
    (refactor on the next solve!); assigning `cache.b` does not."""

    def __init__(self, A, b, u, alg, cacheval, assumptions, abstol, reltol, verbose):
        object.__setattr__(self, "_A", A)
        object.__setattr__(self, "_b", b)
        self.u = u
        self.alg = alg
        self.cacheval = cacheval
        self.isfresh = True
        self.assumptions = assumptions
        self.abstol, self.reltol = abstol, reltol
        self.verbose = verbose
        self.fell_back_to_qr = False

    @property
    def A(self):
        return self._A

    @A.setter
    def A(self, val):
        object.__setattr__(self, "_A", val)
        self.isfresh = True            # src/common.jl:330-348
        self.fell_back_to_qr = False

    @property
    def b(self):
        return self._b

    @b.setter
    def b(self, val):
        object.__setattr__(self, "_b", val)


def _promote(x):
    """int arrays are promoted to float like __promote_int_arrays (src/common.jl:448-502)"""
    x = np.asarray(x) if not isinstance(x, BlockDiagonal) else x
    if isinstance(x, np.ndarray) and x.dtype.kind in "iub":
        return x.astype(np.float64)
    return x


def init(prob: LinearProblem, alg=None, alias_A: bool = False, alias_b: bool = False,
         abstol=None, reltol=None, verbose: bool = False, assumptions: OperatorAssumptions | None = None):
    A, b = _promote(prob.A), _promote(prob.b)
    assumptions = assumptions or OperatorAssumptions(issq=(A.shape[0] == A.shape[1]))
    if alg is None:
        alg = defaultalg(A, b, assumptions)
    if isinstance(alg, DefaultLinearSolver):
        alg = alg.to_alg()
    if not isinstance(alg, B200LUFactorization):
        raise NotImplementedError(
            f"{type(alg).__name__} is a reference (CPU) algorithm outside this library's hot path")
    if A.shape[0] != A.shape[1]:
        raise ValueError("B200LUFactorization needs a square A (needs_square_A)")
    if b.shape[0] != A.shape[0]:
        raise ValueError(f"b has leading dimension {b.shape[0]}, but needs {A.shape[0]}")
    # dense factorizations default to alias_A = false: the cache owns a private copy
    if not alias_A:
        A = A.copy() if isinstance(A, BlockDiagonal) else np.array(A, order="F", copy=True)
    if not alias_b:
        b = np.array(b, copy=True)
    u = np.zeros_like(b) if prob.u0 is None else np.array(prob.u0, copy=True)
    eps = np.finfo(b.dtype).eps if b.dtype.kind == "f" else np.finfo(np.float64).eps
    abstol = np.sqrt(eps) if abstol is None else abstol
    reltol = np.sqrt(eps) if reltol is None else reltol
    return LinearCache(A, b, u, alg, _B200LUCache(), assumptions, abstol, reltol, verbose)


def reinit(cache: LinearCache, A=None, b=None, u=None):
    """reinit!(cache; A, b, u) — src/common.jl:933-964"""
    if A is not None:
        cache.A = A
    if b is not None:
        cache.b = b
    if u is not None:
        cache.u = u
    return cache


def _configure(handle, alg):
    if alg.nb is not None:
        handle.set_option(_capi.OPT_NB, alg.nb)
    if alg.lookahead is not None:
        handle.set_option(_capi.OPT_LOOKAHEAD, int(alg.lookahead))
    if getattr(alg, "host_register", False):
        handle.set_option(_capi.OPT_HOST_REGISTER, 1)
    if isinstance(alg, B200LU32MixedLUFactorization):
        handle.set_option(_capi.OPT_REFINE_MAXIT, alg.maxiters if alg.refine else 0)
    elif alg.residualsafety:
        handle.set_option(_capi.OPT_KEEP_A, 1)    # the residual check runs on the device


# size classes of the batched getrf/getrs: up to 64 rows the register kernels (template NMAX in
# csrc/batched.cuh), 65 ... 160 rows the shared-memory kernel (any n, classes only bound the padding);
# larger blocks take the single-system path, one handle each
_BATCHED_CLASSES = (16, 32, 64, 96, 128, 160)


def plan_blockdiag(sizes):
    """Launch plan for a BlockDiagonal with blocks of `sizes` (ragged allowed, the reference's
    `[2, 3, 4]` case and the variable-size supernode blocks of SURVEY §8(f)3): every block of up
    to 160 rows joins the batched launch of its size class, embedded in the class's largest
    member size m as diag(B, I) — at most six launches however ragged the sizes are; blocks
    above 160 are factored one by one.  Returns [(kind, block indices, m)], kind in
    {"batched", "single"}; empty blocks appear nowhere."""
    by_class = {}
    singles = []
    for i, n in enumerate(sizes):
        if n == 0:
            continue
        cls = next((c for c in _BATCHED_CLASSES if n <= c), None)
        if cls is None:
            singles.append(i)
        else:
            by_class.setdefault(cls, []).append(i)
    plan = [("batched", idx, max(sizes[i] for i in idx)) for _, idx in sorted(by_class.items())]
    plan += [("single", [i], sizes[i]) for i in singles]
    return plan


def pad_blocks(blocks, idx, m, dtype):
    """Stack blocks[idx] as (len(idx), m, m) column-major systems, block B embedded as diag(B, I).
    Partial pivoting never looks at the padding: in column k < n the padding rows hold exact zeros
    (an all-zero subcolumn keeps kp = k, like the unpadded block), their multipliers are 0 and the
    rank-1 updates leave them untouched, and for k >= n the pivot is the unit diagonal.  So
    ipiv[:n], info and the leading n x n factors are those of B itself, bit for bit."""
    stack = np.zeros((len(idx), m, m), dtype=dtype)
    pad = np.arange(m)
    for s, i in enumerate(idx):
        B = blocks[i]
        n = B.shape[0]
        stack[s, :n, :n] = B.T          # [s, col, row]
        stack[s, pad[n:], pad[n:]] = 1
    return stack


def _factor_blockdiag(cache, alg, fuse_b=None):
    """Blockwise LU (ext/LinearSolveBlockDiagonalsExt.jl:119-125): see plan_blockdiag.
    success = all(issuccess) (:121-124).  `fuse_b` (a vector right-hand side): batched groups of
    blocks of up to 64 rows solve it inside the factorization kernel (b200lu_factor_solve_batched);
    returns (info, {group index: solution rows})."""
    A = cache.A
    cv = cache.cacheval
    if not A.all_square():
        raise ValueError("B200LUFactorization needs square diagonal blocks")
    sizes = [B.shape[0] for B in A.blocks]
    dt = alg.handle_dtype(A.dtype)
    if cv.groups is None or cv.group_sizes != (sizes, dt):
        cv.groups = []
        for kind, idx, m in plan_blockdiag(sizes):
            h = _capi.Handle(dt, alg.device, devices=alg.devices if kind == "batched" else None)
            _configure(h, alg)
            cv.groups.append((kind, h, idx, m))
        cv.group_sizes = (sizes, dt)
    ok = True
    fused = {}
    for gi, (kind, h, idx, m) in enumerate(cv.groups):
        if kind == "batched":
            stack = pad_blocks(A.blocks, idx, m, h.np_dtype)
            if fuse_b is not None and m <= 64 and hasattr(h, "factor_solve_batched"):
                rhs = np.zeros((len(idx), m), dtype=h.np_dtype)
                for s_, i in enumerate(idx):
                    o, n = A.offsets[i], A.blocks[i].shape[0]
                    rhs[s_, :n] = fuse_b[o:o + n]
                x, _, info = h.factor_solve_batched(stack, rhs)
                fused[gi] = x
            else:
                _, info = h.factor_batched(stack)
            ok = ok and not np.any(info != 0)
        else:
            _, info = h.factor(np.asfortranarray(A.blocks[idx[0]], dtype=h.np_dtype), want_ipiv=False)
            ok = ok and info == 0
    cv.info = 0 if ok else 1
    cv.fused = fused if ok else {}
    return cv.info


def _solve_blockdiag(cache, adjoint=False, fused=None):
    """per-block ldiv! on views of b (ext/LinearSolveBlockDiagonalsExt.jl:183-205); `adjoint`: the
    transposed blocks with the same factors (block-diagonal structure is its own transpose);
    `fused`: solutions the factorization kernel already produced for some groups"""
    A, cv = cache.A, cache.cacheval
    b = np.asarray(cache.b)
    vec = b.ndim == 1
    Bm = b.reshape(b.shape[0], -1)
    X = np.empty_like(Bm)
    trans = "T" if adjoint else "N"
    for gi, (kind, h, idx, m) in enumerate(cv.groups):
        if fused and gi in fused:
            for s, i in enumerate(idx):
                o, n = A.offsets[i], A.blocks[i].shape[0]
                X[o:o + n, 0] = fused[gi][s, :n]
            continue
        if kind == "batched":
            # (batch, nrhs, m): each right-hand side contiguous, zero in the padding rows
            rhs = np.zeros((len(idx), Bm.shape[1], m), dtype=h.np_dtype)
            for s, i in enumerate(idx):
                o, n = A.offsets[i], A.blocks[i].shape[0]
                rhs[s, :, :n] = Bm[o:o + n, :].T
            sol = h.solve_batched(rhs, trans=trans)
            for s, i in enumerate(idx):
                o, n = A.offsets[i], A.blocks[i].shape[0]
                X[o:o + n, :] = sol[s, :, :n].T
        else:
            o, n = A.offsets[idx[0]], m
            X[o:o + n, :] = h.solve(np.asfortranarray(Bm[o:o + n, :], dtype=h.np_dtype), trans=trans)
    return X[:, 0] if vec else X


def _check_residual_safety(cache, A_original, u):
    """a-posteriori check ‖A u − b‖ <= abstol + reltol‖b‖ (src/factorization.jl:127-156).  Dense
    problems: the norms come from the device (b200lu_residual_norms on the copy of A the handle
    kept); BlockDiagonal problems: block by block on the host (O(sum n_i^2) work on data the host
    already holds)."""
    b = np.asarray(cache.b)
    if isinstance(A_original, BlockDiagonal):
        r = u.copy()
        for B, o in zip(A_original.blocks, A_original.offsets[:-1]):
            k = B.shape[0]
            r[o:o + k] = B @ u[o:o + k]
        rn, bn = np.linalg.norm(r - b), np.linalg.norm(b)
    else:
        dt = cache.cacheval.handle.np_dtype
        resid, bnorm = cache.cacheval.handle.residual_norms(b.astype(dt, copy=False), np.asarray(u, dtype=dt))
        rn, bn = np.sqrt(np.sum(resid ** 2)), np.sqrt(np.sum(bnorm ** 2))    # Frobenius over the columns
    return rn <= cache.abstol + cache.reltol * bn


def solve_(cache: LinearCache, alg=None, adjoint: bool = False) -> LinearSolution:
    """`solve!(cache)` — protocol of src/openblas.jl:362-459: factor only when
    `cache.isfresh`; on info != 0 return ReturnCode.Failure and LEAVE isfresh
    set; otherwise getrs from cache.b into cache.u.

    `adjoint=True` mirrors `solve!(cache; adjoint = true)` (src/common.jl:1012-1027):
    adjoint(A) u = b with the SAME cached factorization (getrs with trans = 'T'; the
    element types here are real, so adjoint == transpose)."""
    alg = alg or cache.alg
    cv = cache.cacheval
    A = cache.A
    if adjoint and getattr(alg, "devices", None) is not None:
        raise NotImplementedError("solve!(cache; adjoint = true) needs the single-GPU handle")
    check_safety = alg.residualsafety and cache.isfresh
    fused = None
    if cache.isfresh:
        if isinstance(A, BlockDiagonal):
            b0 = np.asarray(cache.b)
            info = _factor_blockdiag(cache, alg, fuse_b=b0 if (b0.ndim == 1 and not adjoint) else None)
            fused = getattr(cv, "fused", None)
        else:
            A = np.asarray(A)
            if cv.handle is None or cv.handle.dtype != alg.handle_dtype(A.dtype):
                cv.handle = _capi.Handle(alg.handle_dtype(A.dtype), alg.device, devices=alg.devices)
                _configure(cv.handle, alg)
            cv.ipiv, info = cv.handle.factor(A)
            cv.info = info
        if info != 0:
            if cache.verbose:
                print("Solver failed")          # @SciMLMessage("Solver failed", ..., :solver_failure)
            return LinearSolution(cache.u, ReturnCode.Failure, alg)
        cache.isfresh = False
    if isinstance(A, BlockDiagonal):
        x = _solve_blockdiag(cache, adjoint, fused)
    else:
        u, b = cache.u, np.asarray(cache.b)
        # a plain vector u, or a column-major matrix u: getrs writes straight into cache.u
        direct = (isinstance(u, np.ndarray) and u.shape == b.shape and u.dtype == cv.handle.np_dtype and
                  ((u.ndim == 1 and u.flags.c_contiguous) or (u.ndim == 2 and u.flags.f_contiguous)) and
                  not np.shares_memory(u, b))
        x = cv.handle.solve(b, out=u if direct else None, trans="T" if adjoint else "N")
        if direct:
            x = None   # getrs wrote straight into cache.u
    if x is not None:
        cache.u[...] = x
    if adjoint:
        return LinearSolution(cache.u, ReturnCode.Success, alg)
    if check_safety and not _check_residual_safety(cache, A, cache.u):
        if cache.verbose:
            print("Residual safety check failed")   # @SciMLMessage(cache.verbose, :residual_safety)
        return LinearSolution(cache.u, ReturnCode.APosterioriSafetyFailure, alg)
    return LinearSolution(cache.u, ReturnCode.Success, alg)


def solve(prob: LinearProblem, alg=None, **kw) -> LinearSolution:
    return solve_(init(prob, alg, **kw))


# ----------------------------------------------------------- polyalgorithm ----
class DefaultAlgorithmChoice(enum.Enum):
    """the members of src/LinearSolve.jl:330-357 that the dense arm can return,
    plus the new slot"""
    GenericLUFactorization = 1
    RFLUFactorization = 2
    LUFactorization = 3
    MKLLUFactorization = 4
    QRFactorization = 5
    SVDFactorization = 6
    B200LUFactorization = 100


@dataclass
class DefaultLinearSolver:
    alg: DefaultAlgorithmChoice
    safetyfallback: bool = True

    def to_alg(self):
        """algchoice_to_alg (src/default.jl:522-581) for the slot this library owns"""
        if self.alg == DefaultAlgorithmChoice.B200LUFactorization:
            return B200LUFactorization(throwerror=False)
        return _ReferenceCPUAlgorithm(self.alg)


@dataclass
class _ReferenceCPUAlgorithm:
    choice: DefaultAlgorithmChoice


# Break-even n above which the GPU path is selected when available.  The
# reference documents "around 1,000 x 1,000" for CUDA offload
# (docs/src/tutorials/gpu.md:19-22); must stay above n = 600 so the reference's
# selection tests (test/Core/default_algs.jl:4-66) are unchanged.
B200_DEFAULT_MIN_N = 1024


def defaultalg(A, b, assump: OperatorAssumptions | None = None, *, isopenblas: bool = True,
               usemkl: bool = False, userecursivefactorization: bool = True,
               b200_available: bool | None = None) -> DefaultLinearSolver:
    """Dense square BLAS-eltype arm of src/default.jl:411-520 with the new
    availability-gated B200 arm inserted above n = 600."""
    assump = assump or OperatorAssumptions()
    n = b.shape[0]
    C = DefaultAlgorithmChoice
    if isinstance(A, BlockDiagonal):
        # ext/LinearSolveBlockDiagonalsExt.jl:205-217 -> LU; the GPU arm takes it when available
        avail = useb200() if b200_available is None else b200_available
        return DefaultLinearSolver(C.B200LUFactorization if avail else C.LUFactorization)
    if assump.condition == "VeryIllConditioned":
        return DefaultLinearSolver(C.QRFactorization)
    if assump.condition == "SuperIllConditioned":
        return DefaultLinearSolver(C.SVDFactorization)
    real_float = np.dtype(A.dtype) in (np.dtype(np.float32), np.dtype(np.float64))
    if n <= 10:
        return DefaultLinearSolver(C.GenericLUFactorization)
    if real_float and n >= B200_DEFAULT_MIN_N:
        avail = useb200() if b200_available is None else b200_available
        if avail:
            return DefaultLinearSolver(C.B200LUFactorization)
    if (n <= 100 or (isopenblas and n <= 500) or (usemkl and n <= 200)) and real_float \
            and userecursivefactorization:
        return DefaultLinearSolver(C.RFLUFactorization)
    if (n <= 32 or (isopenblas and n <= 256)) and real_float:
        return DefaultLinearSolver(C.GenericLUFactorization)
    if usemkl:
        return DefaultLinearSolver(C.MKLLUFactorization)
    return DefaultLinearSolver(C.LUFactorization)


# ------------------------------------------------------------- multi-GPU host logic ----
def shard_batch(batch: int, rank: int, nranks: int):
    """Contiguous range of the batch index owned by `rank` (independent systems shard
    with no communication).  Returns (start, stop)."""
    base, rem = divmod(batch, nranks)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def block_cyclic_columns(n: int, nb: int, rank: int, nranks: int):
    """1D block-cyclic column distribution used by b200lu_factor_dist: global column
    block g (width nb, ragged tail) lives on rank g % nranks at local block g // nranks.
    Returns the global column indices owned by `rank`, in local storage order."""
    cols = []
    nblk = -(-n // nb)
    for g in range(rank, nblk, nranks):
        cols.extend(range(g * nb, min(n, (g + 1) * nb)))
    return np.asarray(cols, dtype=np.int64)


def reduce_info(infos):
    """BlockDiagonalFactorization.success = all(issuccess) (ext/LinearSolveBlockDiagonalsExt.jl:121-124)
    across ranks: the job fails if any shard reported a zero pivot."""
    infos = np.asarray(infos)
    return ReturnCode.Success if not np.any(infos != 0) else ReturnCode.Failure
