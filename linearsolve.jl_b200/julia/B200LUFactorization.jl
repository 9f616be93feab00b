# B200LUFactorization.jl — the glue a LinearSolve.jl maintainer adds to reach
# libb200lu.so.  NOT runnable in the build environment (no Julia); it is kept
# declarative and follows, line for line, patterns that already exist in the
# reference (v5.12.0):
#   * struct + `throwerror=false` constructor  : src/extension_algs.jl:344-354
#   * direct-`ccall` getrf/getrs wrappers        : src/openblas.jl:131-154,247-278
#   * cacheval struct + init_cacheval            : src/openblas.jl:312-360
#   * solve! protocol (isfresh / Failure / u<-b) : src/openblas.jl:362-459
#   * 32Mixed variant                            : src/openblas.jl:470-543
#   * availability hook                          : src/LinearSolve.jl:741-743
# The Python package `linearsolve.jl_b200` (interface.py/_capi.py) is the tested
# twin of this file: same names, same control flow, ctypes instead of ccall.
#
# Everything below lives in `src/b200lu.jl` (included from src/LinearSolve.jl
# next to `include("openblas.jl")`); the enum wiring is listed in INTEGRATION.md.

const libb200lu = Ref{String}(get(ENV, "LINEARSOLVE_B200LU_LIB", "libb200lu.so"))
const _b200lu_handle_ok = Ref{Union{Nothing, Bool}}(nothing)

# dtype codes of include/b200lu.h
const B200LU_F64 = Cint(0)
const B200LU_F32 = Cint(1)
const B200LU_MIXED = Cint(2)

"""
    useb200()

`true` when libb200lu.so can be dlopen'ed and creates a handle on an sm_100
device.  Cached after the first call.  Mirrors `usecuda`/`usemetal`
(src/LinearSolve.jl:741-743): the default algorithm only routes here when it
returns `true`, so machines without the library see no behaviour change.
"""
function useb200()
    ok = _b200lu_handle_ok[]
    ok === nothing || return ok
    ok = try
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:b200lu_create, libb200lu[]), Cint,
            (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cint}), h, B200LU_F64, 1, C_NULL)
        rc == 0 && ccall((:b200lu_destroy, libb200lu[]), Cvoid, (Ptr{Cvoid},), h[])
        rc == 0
    catch
        false
    end
    _b200lu_handle_ok[] = ok
    return ok
end

"""
    B200LUFactorization(; throwerror = true, residualsafety = false, device = 0)

Dense partially pivoted LU on one NVIDIA B200 through `libb200lu.so`
(hand-written sm_100a kernels; no CUDA.jl, no cuSOLVER, no CPU fallback).
Host arrays in, host arrays out; the factors stay on the device inside the
cache, so `cache.b = b2; solve!(cache)` only runs getrs.
"""
struct B200LUFactorization <: AbstractFactorization
    residualsafety::Bool
    device::Int
    devices::Vector{Cint}      # more than one entry: ONE cache drives all these GPUs (b200lu_create with ngpus > 1)
    host_register::Bool        # the library page-locks the cache's copy of A once (B200LU_OPT_HOST_REGISTER)
    function B200LUFactorization(; throwerror = true, residualsafety::Bool = false, device::Int = 0,
            devices = nothing, host_register::Bool = false)
        if throwerror && !useb200()
            error("B200LUFactorization requires libb200lu.so and an NVIDIA B200 (sm_100) GPU")
        end
        devs = devices === nothing ? Cint[device] : collect(Cint, devices)    # e.g. devices = 0:7
        length(devs) > 1 && residualsafety &&
            error("residualsafety needs the single-GPU handle (the multi-GPU handle keeps no copy of A)")
        return new(residualsafety, Int(first(devs)), devs, host_register)
    end
end

# smallest n at which `defaultalg` prefers the GPU path when it is available: above every CPU band of
# src/default.jl:444-474 (the reference's own guidance for GPU offload is "around 1,000 x 1,000",
# docs/src/tutorials/gpu.md:19-22), so machines without the library select exactly what they select today
const B200LU_DEFAULT_MIN_N = 1024

"""
    B200LU32MixedLUFactorization(; refine = true, maxiters = 10, throwerror = true)

Float64 interface, Float32 factorization on the GPU.  With `refine = false` this
is the behaviour of the other `*32MixedLUFactorization` algorithms (cast, sgetrf,
sgetrs, cast back); with `refine = true` an FP64 residual/correction loop runs on
the device until the normwise backward error reaches FP64 working accuracy.
"""
struct B200LU32MixedLUFactorization <: AbstractFactorization
    refine::Bool
    maxiters::Int
    device::Int
    function B200LU32MixedLUFactorization(; refine::Bool = true, maxiters::Int = 10,
            throwerror = true, device::Int = 0)
        if throwerror && !useb200()
            error("B200LU32MixedLUFactorization requires libb200lu.so and an NVIDIA B200 (sm_100) GPU")
        end
        return new(refine, maxiters, device)
    end
end

# traits live next to the struct, never in an extension (src/interface.jl:203-209)
needs_concrete_A(::B200LUFactorization) = true
needs_concrete_A(::B200LU32MixedLUFactorization) = true
default_alias_A(::B200LUFactorization, ::Any, ::Any) = false
default_alias_b(::B200LUFactorization, ::Any, ::Any) = false
default_alias_A(::B200LU32MixedLUFactorization, ::Any, ::Any) = false
default_alias_b(::B200LU32MixedLUFactorization, ::Any, ::Any) = false
_get_residualsafety(alg::B200LUFactorization) = alg.residualsafety

# ------------------------------------------------------------------ cacheval --
mutable struct B200LUCache
    handle::Ptr{Cvoid}          # b200lu_handle*, C_NULL until the first fresh solve
    dtype::Cint
    ipiv::Vector{BlasInt}       # 1-based LAPACK interchange sequence (BlasInt == Int64)
    info::Base.RefValue{BlasInt}
    n::Int
    registered::Any             # the host matrix the library page-locked (host_register): kept alive with the handle
end

function _b200lu_finalize(c::B200LUCache)
    if c.handle != C_NULL
        ccall((:b200lu_destroy, libb200lu[]), Cvoid, (Ptr{Cvoid},), c.handle)   # also releases the registration
        c.handle = C_NULL
    end
    c.registered = nothing
    return nothing
end

function _b200lu_cache(dtype::Cint)
    c = B200LUCache(C_NULL, dtype, Vector{BlasInt}(undef, 0), Ref{BlasInt}(0), 0, nothing)
    finalizer(_b200lu_finalize, c)      # pattern: AMGX handles, src/extension_algs.jl:1555-1556
    return c
end

_b200lu_dtype(::B200LUFactorization, ::Type{Float64}) = B200LU_F64
_b200lu_dtype(::B200LUFactorization, ::Type{Float32}) = B200LU_F32
_b200lu_dtype(::B200LU32MixedLUFactorization, ::Type{Float64}) = B200LU_MIXED

# init_cacheval must return the SAME concrete type solve! later stores
# (docs/src/advanced/algorithm_interface.md:68); no device allocation here —
# the default solver builds every slot eagerly (src/default.jl:645-694).
function init_cacheval(
        alg::Union{B200LUFactorization, B200LU32MixedLUFactorization}, A, b, u, Pl, Pr,
        maxiters::Int, abstol, reltol, verbose::Union{LinearVerbosity, Bool},
        assumptions::OperatorAssumptions
    )
    T = A === nothing ? Float64 : eltype(A)
    dtype = T === Float32 && alg isa B200LUFactorization ? B200LU_F32 :
            (alg isa B200LU32MixedLUFactorization ? B200LU_MIXED : B200LU_F64)
    return _b200lu_cache(dtype)
end

function _b200lu_error(c::B200LUCache, rc::Cint)
    msg = unsafe_string(ccall((:b200lu_last_error, libb200lu[]), Cstring, (Ptr{Cvoid},), c.handle))
    return error("libb200lu status $rc: $msg")
end

function _b200lu_ensure_handle!(c::B200LUCache, alg)
    c.handle != C_NULL && return c
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dev = alg isa B200LUFactorization ? alg.devices : Cint[alg.device]
    # ngpus > 1: the library distributes the host matrix 1-D block-cyclic over the GPUs itself (each GPU pulls
    # its column blocks over its own PCIe link) and shards BlockDiagonal batches by index
    rc = ccall((:b200lu_create, libb200lu[]), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cint}), h, c.dtype, length(dev), dev)
    rc == 5 && error("b200lu_create: the devices $(Int.(dev)) are not all peers of each other (NVLink / NVSwitch)")
    rc == 0 || error("b200lu_create failed with status $rc (no usable sm_100 device; there is no CPU fallback)")
    c.handle = h[]
    if alg isa B200LU32MixedLUFactorization
        ccall((:b200lu_set_option, libb200lu[]), Cint, (Ptr{Cvoid}, Cint, Int64),
            c.handle, 2, alg.refine ? alg.maxiters : 0)          # B200LU_OPT_REFINE_MAXIT
    elseif alg.residualsafety
        ccall((:b200lu_set_option, libb200lu[]), Cint, (Ptr{Cvoid}, Cint, Int64),
            c.handle, 13, 1)                                     # B200LU_OPT_KEEP_A
    end
    if alg isa B200LUFactorization && alg.host_register
        # cache.A is the cache's own copy (alias_A = false, src/common.jl:818-842) and lives as long as the cache:
        # pinned once, every refactorization streams its upload under the factorization (the cacheval's finalizer
        # destroys the handle — and with it the registration — so it must run before cache.A is collected: the
        # cacheval keeps a reference to that array, see B200LUCache.registered)
        ccall((:b200lu_set_option, libb200lu[]), Cint, (Ptr{Cvoid}, Cint, Int64),
            c.handle, 15, 1)                                     # B200LU_OPT_HOST_REGISTER
    end
    return c
end

# getrf: same call shape as openblas_getrf! (src/openblas.jl:131-154)
@inline function _direct_lu_factorize!(c::B200LUCache, A::StridedMatrix{T}, alg) where {T}
    chkstride1(A)
    n = checksquare(A)
    _b200lu_ensure_handle!(c, alg)
    length(c.ipiv) == n || resize!(c.ipiv, n)
    rc = ccall((:b200lu_factor, libb200lu[]), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Int64, Ptr{BlasInt}, Ref{BlasInt}),
        c.handle, n, A, max(1, stride(A, 2)), c.ipiv, c.info)
    rc == 0 || _b200lu_error(c, rc)
    c.n = n
    if alg isa B200LUFactorization && alg.host_register
        c.registered = A        # the registration moved to this buffer (or stayed on it)
    end
    return c.info[]
end

# getrs: same call shape as openblas_getrs! (src/openblas.jl:247-278); u may alias b
@inline function _direct_lu_solve!(c::B200LUCache, u::StridedVecOrMat{T}, b::StridedVecOrMat{T}, alg;
        trans::Char = 'N') where {T}
    chkstride1(u, b)
    size(b, 1) == c.n || throw(DimensionMismatch("b has leading dimension $(size(b, 1)), but needs $(c.n)"))
    rc = ccall((:b200lu_solve, libb200lu[]), Cint,
        (Ptr{Cvoid}, UInt8, Int64, Ptr{T}, Int64, Ptr{T}, Int64),
        c.handle, UInt8(trans), size(b, 2), b, max(1, stride(b, 2)), u, max(1, stride(u, 2)))
    rc == 0 || _b200lu_error(c, rc)
    return u
end

# ||b - A u|| <= abstol + reltol ||b|| with both norms from b200lu_residual_norms (Frobenius over the
# columns of a matrix right-hand side, like `norm` in the reference's check)
function _b200lu_residual_ok(c::B200LUCache, u::StridedVecOrMat{T}, b::StridedVecOrMat{T}, abstol, reltol) where {T}
    nrhs = size(b, 2)
    resid = Vector{Float64}(undef, nrhs)
    bnorm = Vector{Float64}(undef, nrhs)
    rc = ccall((:b200lu_residual_norms, libb200lu[]), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Int64, Ptr{T}, Int64, Ptr{Float64}, Ptr{Float64}),
        c.handle, nrhs, b, max(1, stride(b, 2)), u, max(1, stride(u, 2)), resid, bnorm)
    rc == 0 || _b200lu_error(c, rc)
    return sqrt(sum(abs2, resid)) <= abstol + reltol * sqrt(sum(abs2, bnorm))
end

# `solve!(cache; adjoint = true)` (src/common.jl:1012-1027) reuses the cached factorization:
# the hooks of src/adjoint_factorization.jl:149-153.  Real element types: adjoint == transpose.
# The mixed-precision handle refines the TRANSPOSED system (FP32 transposed sweeps + FP64 residual
# b - A'x), so both algorithms reuse their factors.
const _B200LUAlgs = Union{B200LUFactorization, B200LU32MixedLUFactorization}
_custom_can_reuse_adjoint_factorization(::B200LU32MixedLUFactorization, c::B200LUCache) = c.handle != C_NULL
# (the multi-GPU handle solves with trans = 'N' only: its adjoint is refactorized by the generic path)
_custom_can_reuse_adjoint_factorization(alg::B200LUFactorization, c::B200LUCache) =
    c.handle != C_NULL && length(alg.devices) == 1
# `defaultalg_adjoint_eval` (src/default.jl:1262-1348): dy is overwritten with adjoint(A) \ dy
function _b200lu_solve_trans!(c::B200LUCache, dy::StridedVecOrMat{T}) where {T}
    rc = ccall((:b200lu_solve, libb200lu[]), Cint,
        (Ptr{Cvoid}, UInt8, Int64, Ptr{T}, Int64, Ptr{T}, Int64),
        c.handle, UInt8('T'), size(dy, 2), dy, max(1, stride(dy, 2)), dy, max(1, stride(dy, 2)))
    rc == 0 || _b200lu_error(c, rc)
    return dy
end
function _custom_adjoint_factorization_solve(alg::_B200LUAlgs, c::B200LUCache, A, b)
    u = similar(b)
    return _direct_lu_solve!(c, u, b, alg; trans = 'T')
end

function SciMLBase.solve!(
        cache::LinearCache, alg::Union{B200LUFactorization, B200LU32MixedLUFactorization};
        kwargs...
    )
    A_work = convert(AbstractMatrix, cache.A)
    check_safety = alg isa B200LUFactorization && alg.residualsafety && cache.isfresh
    # host A is never overwritten by the device path, so no A_backup copy is needed
    # for the default solver's QR rescue (contrast src/openblas.jl:369-372)
    cacheval = alg isa B200LUFactorization ?
        @get_cacheval(cache, :B200LUFactorization) :
        @get_cacheval(cache, :B200LU32MixedLUFactorization)
    if cache.isfresh
        info_value = _direct_lu_factorize!(cacheval, A_work, alg)
        if info_value != 0
            @SciMLMessage("Solver failed", cache.verbose, :solver_failure)
            return SciMLBase.build_linear_solution(
                alg, cache.u, nothing, nothing; retcode = ReturnCode.Failure
            )                                   # isfresh stays true (src/factorization.jl:714-722)
        end
        cache.isfresh = false
    end
    require_one_based_indexing(cache.u, cache.b)
    _direct_lu_solve!(cacheval, cache.u, cache.b, alg)
    if check_safety
        # the reference's a-posteriori check (src/factorization.jl:127-156) with the two norms taken on
        # the device from the copy of A the handle kept (B200LU_OPT_KEEP_A, set in
        # _b200lu_ensure_handle! when alg.residualsafety): 0.18 ms at n = 8192 instead of a host GEMV
        if !_b200lu_residual_ok(cacheval, cache.u, cache.b, cache.abstol, cache.reltol)
            @SciMLMessage("Residual safety check failed", cache.verbose, :residual_safety)
            return SciMLBase.build_linear_solution(
                alg, cache.u, nothing, nothing; retcode = ReturnCode.APosterioriSafetyFailure
            )                                   # as `_check_residual_safety`, src/factorization.jl:150-153
        end
    end
    return SciMLBase.build_linear_solution(
        alg, cache.u, nothing, nothing; retcode = ReturnCode.Success
    )
end

# BlockDiagonal surface (ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205).  Blocks of up to 160
# rows go through ONE batched call per size class (<= 16, 32, 64, 96, 128, 160 rows) however ragged their
# sizes are: block B is embedded in the class's largest member size m as diag(B, I).  Partial
# pivoting never looks at the padding (exact zeros below B, unit pivots after it), so ipiv[1:n],
# info and the leading n x n factors are those of B itself (tests/test_host_logic.py::
# test_blockdiag_padding_is_exact checks this bit for bit against the per-block lu!).  Larger
# blocks take `_direct_lu_factorize!`, one cache each.  Python twin: interface.py::plan_blockdiag,
# pad_blocks, _factor_blockdiag, _solve_blockdiag.
function _b200lu_factor_blocks!(c::B200LUCache, blocks::Vector{Matrix{T}}, alg) where {T}
    m = maximum(B -> size(B, 1), blocks)              # all blocks of one kernel class
    batch = length(blocks)
    _b200lu_ensure_handle!(c, alg)
    packed = zeros(T, m, m, batch)                    # column-major systems back to back
    for (s, B) in enumerate(blocks)
        n = size(B, 1)
        copyto!(view(packed, 1:n, 1:n, s), B)
        for d in (n + 1):m
            packed[d, d, s] = one(T)
        end
    end
    ipiv = Vector{BlasInt}(undef, m * batch)
    info = Vector{BlasInt}(undef, batch)
    rc = ccall((:b200lu_factor_batched, libb200lu[]), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{T}, Int64, Int64, Ptr{BlasInt}, Ptr{BlasInt}),
        c.handle, batch, m, packed, m, m * m, ipiv, info)
    rc == 0 || _b200lu_error(c, rc)
    c.n = m
    return all(iszero, info)                           # success = all(issuccess), :121-124
end

# `solve!` on a FRESH cache with a vector right-hand side, blocks of up to 64 rows: getrf of every block and
# getrs of its segment of b in ONE kernel launch (A read once, factors written once and kept for later
# solves) — per-block `lu!` followed by per-block `ldiv!`, ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205
function _b200lu_factor_solve_blocks!(c::B200LUCache, x::Matrix{T}, packed::Array{T, 3}, rhs::Matrix{T}, alg) where {T}
    m, _, batch = size(packed)
    _b200lu_ensure_handle!(c, alg)
    ipiv = Vector{BlasInt}(undef, m * batch)
    info = Vector{BlasInt}(undef, batch)
    rc = ccall((:b200lu_factor_solve_batched, libb200lu[]), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{T}, Int64, Int64, Ptr{T}, Int64, Ptr{T}, Int64, Ptr{BlasInt}, Ptr{BlasInt}),
        c.handle, batch, m, packed, m, m * m, rhs, m, x, m, ipiv, info)
    rc == 0 || _b200lu_error(c, rc)
    c.n = m
    return all(iszero, info)
end

# per-block ldiv! on views of b (:183-205); rhs[:, :, s] is the zero-padded m x nrhs block of b that
# belongs to system s.  trans = 'T': `solve!(cache; adjoint = true)` with the same factors.
function _b200lu_solve_blocks!(c::B200LUCache, x::Array{T, 3}, rhs::Array{T, 3}; trans::Char = 'N') where {T}
    m, nrhs, batch = size(rhs)
    rc = ccall((:b200lu_solve_batched_trans, libb200lu[]), Cint,
        (Ptr{Cvoid}, UInt8, Int64, Ptr{T}, Int64, Int64, Ptr{T}, Int64, Int64),
        c.handle, UInt8(trans), nrhs, rhs, m, m * nrhs, x, m, m * nrhs)
    rc == 0 || _b200lu_error(c, rc)
    return x
end
