#!/usr/bin/env python
"""Generate the unified diffs that wire `B200LUFactorization` into SciML/LinearSolve.jl v5.12.0.

    python linearsolve.jl_b200/julia/patches/make_patches.py [/path/to/LinearSolve.jl]   (default /root/reference)

Every edit site is the footprint of the newest `DefaultAlgorithmChoice` member (`LHLFactorization`)
plus the sites of the GPU-offload sibling (`CudaOffloadLUFactorization`); INTEGRATION.md §2 lists
them with their file:line.  The diffs are plain `diff -u` output relative to the repository root
(apply with `patch -p1`); `tests/test_julia_patches.py` applies them to a scratch copy and checks
that no site of the footprint is missed.  Nothing here reads reference source at product run
time: this is maintainer tooling and test infrastructure.
"""
import difflib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def sub1(text, old, new, what):
    if text.count(old) != 1:
        raise SystemExit(f"edit site not found exactly once ({text.count(old)}x): {what}")
    return text.replace(old, new)


def edit_linearsolve_jl(t):
    t = sub1(t, "    SparseColumnPivotedQRFactorization\n    LHLFactorization\nend",
             "    SparseColumnPivotedQRFactorization\n    LHLFactorization\n    B200LUFactorization\nend",
             "DefaultAlgorithmChoice enum (src/LinearSolve.jl:330-357)")
    t = sub1(t, "    elseif alg === DefaultAlgorithmChoice.LHLFactorization\n"
                "        return true  # LHLFactorization.jl is a hard dependency, always available\n",
             "    elseif alg === DefaultAlgorithmChoice.LHLFactorization\n"
             "        return true  # LHLFactorization.jl is a hard dependency, always available\n"
             "    elseif alg === DefaultAlgorithmChoice.B200LUFactorization\n"
             "        return useb200()  # Available if libb200lu.so opens and finds an sm_100 device\n",
             "is_algorithm_available (src/LinearSolve.jl:367-395)")
    t = sub1(t, 'include("openblas.jl")\n', 'include("openblas.jl")\ninclude("b200lu.jl")\n',
             "include list (src/LinearSolve.jl:445-465)")
    t = sub1(t, "        :STRUMPACKFactorization,\n    )\n    @eval needs_square_A(::$(alg)) = true",
             "        :STRUMPACKFactorization,\n        :B200LUFactorization, :B200LU32MixedLUFactorization,\n    )\n"
             "    @eval needs_square_A(::$(alg)) = true", "needs_square_A list (src/LinearSolve.jl:720-731)")
    t = sub1(t, "export LHLFactorization, update_gamma!\n",
             "export LHLFactorization, update_gamma!\nexport B200LUFactorization, B200LU32MixedLUFactorization\n",
             "exports (src/LinearSolve.jl:789)")
    return t


def edit_default_jl(t):
    t = sub1(t, "        T13, T14, T15, T16, T17, T18, T19, T20, T21, T22, T23, T24, T25, T26,\n        TA, Tb, TR,",
             "        T13, T14, T15, T16, T17, T18, T19, T20, T21, T22, T23, T24, T25, T26, T27,\n        TA, Tb, TR,",
             "DefaultLinearSolverInit type parameters (src/default.jl:8-12)")
    t = sub1(t, "    LHLFactorization::T26\n", "    LHLFactorization::T26\n    B200LUFactorization::T27\n",
             "DefaultLinearSolverInit field in enum order (src/default.jl:38)")
    # heuristic arm: above every CPU band (n >= 1024), gated on availability, so machines without the
    # library select exactly what they select today (test/Core/default_algs.jl:4-66)
    t = sub1(t, "                    if tuned_alg !== nothing\n                        tuned_alg\n",
             "                    if tuned_alg !== nothing\n                        tuned_alg\n"
             "                    elseif matrix_size >= B200LU_DEFAULT_MIN_N && b isa DenseArray &&\n"
             "                            !(b isa GPUArraysCore.AnyGPUArray) &&\n"
             "                            (\n"
             "                            A === nothing ? eltype(b) <: Union{Float32, Float64} :\n"
             "                                eltype(A) <: Union{Float32, Float64}\n"
             "                        ) && useb200()\n"
             "                        DefaultAlgorithmChoice.B200LUFactorization\n",
             "defaultalg heuristic arm (src/default.jl:444-474)")
    t = sub1(t, "    elseif alg === :LHLFactorization\n        LHLFactorization()\n",
             "    elseif alg === :LHLFactorization\n        LHLFactorization()\n"
             "    elseif alg === :B200LUFactorization\n        B200LUFactorization(throwerror = false)\n",
             "algchoice_to_alg (src/default.jl:522-581)")
    t = sub1(t, "    elseif alg === :MetalLUFactorization\n"
                "        :(MetalLUFactorization(throwerror = false, residualsafety = alg.residualsafety))\n",
             "    elseif alg === :MetalLUFactorization\n"
             "        :(MetalLUFactorization(throwerror = false, residualsafety = alg.residualsafety))\n"
             "    elseif alg === :B200LUFactorization\n"
             "        :(B200LUFactorization(throwerror = false, residualsafety = alg.residualsafety))\n",
             "_algchoice_to_alg_with_safety (src/default.jl:1024-1044)")
    t = sub1(t, "        elseif alg == Symbol(DefaultAlgorithmChoice.MetalLUFactorization)\n"
                "            inner_alg_expr = _algchoice_to_alg_with_safety(alg)\n",
             "        elseif alg == Symbol(DefaultAlgorithmChoice.B200LUFactorization)\n"
             "            inner_alg_expr = _algchoice_to_alg_with_safety(alg)\n"
             "            newex = quote\n"
             "                if !useb200()\n"
             "                    error(\"Default algorithm calling solve on B200LUFactorization without libb200lu.so and a B200 being available. This shouldn't happen.\")\n"
             "                end\n"
             "                sol = SciMLBase.solve!(cache, $inner_alg_expr)\n"
             "                _default_lu_solve_with_fallback(cache, alg, sol)\n"
             "            end\n"
             "        elseif alg == Symbol(DefaultAlgorithmChoice.MetalLUFactorization)\n"
             "            inner_alg_expr = _algchoice_to_alg_with_safety(alg)\n",
             "@generated solve! branch with the LU -> QR safety fallback (src/default.jl:1096-1113)")
    t = sub1(t, "        elseif alg == Symbol(DefaultAlgorithmChoice.AppleAccelerateLUFactorization)\n"
                "            quote\n"
                "                A = getproperty(cache.cacheval, $(Meta.quot(alg)))\n"
                "                aa_getrs!('T', A.factors, A.ipiv, dy, A.info)\n"
                "            end\n",
             "        elseif alg == Symbol(DefaultAlgorithmChoice.AppleAccelerateLUFactorization)\n"
             "            quote\n"
             "                A = getproperty(cache.cacheval, $(Meta.quot(alg)))\n"
             "                aa_getrs!('T', A.factors, A.ipiv, dy, A.info)\n"
             "            end\n"
             "        elseif alg == Symbol(DefaultAlgorithmChoice.B200LUFactorization)\n"
             "            quote\n"
             "                # getrs with trans = 'T' on the factors cached on the device (real element types)\n"
             "                _b200lu_solve_trans!(getproperty(cache.cacheval, $(Meta.quot(alg))), copy(dy))\n"
             "            end\n",
             "defaultalg_adjoint_eval branch (src/default.jl:1262-1348)")
    return t


def edit_adjoint_factorization_jl(t):
    t = sub1(t, "        HSLMA57Factorization,\n        HSLMA97Factorization,\n    )\n"
                "    @eval _adjoint_factorization_reuse(::Type{<:$Alg}) =\n        _CustomAdjointFactorizationReuse()",
             "        HSLMA57Factorization,\n        HSLMA97Factorization,\n"
             "        # the factors live on the device; getrs with trans = 'T' / 'C' reuses them (src/b200lu.jl)\n"
             "        B200LUFactorization,\n        B200LU32MixedLUFactorization,\n    )\n"
             "    @eval _adjoint_factorization_reuse(::Type{<:$Alg}) =\n        _CustomAdjointFactorizationReuse()",
             "_adjoint_factorization_reuse policy table (src/adjoint_factorization.jl:60-70; enforced by test/Core/adjoint.jl:60-66)")
    return t


def edit_preferences_jl(t):
    t = sub1(t, "    elseif algorithm_name == \"MetalLUFactorization\"\n"
                "        return DefaultAlgorithmChoice.MetalLUFactorization  # Now supported as a separate choice\n",
             "    elseif algorithm_name == \"MetalLUFactorization\"\n"
             "        return DefaultAlgorithmChoice.MetalLUFactorization  # Now supported as a separate choice\n"
             "    elseif algorithm_name == \"B200LUFactorization\"\n"
             "        return DefaultAlgorithmChoice.B200LUFactorization\n",
             "_string_to_algorithm_choice (src/preferences.jl:6-37)")
    return t


def edit_resolve_test(t):
    t = sub1(t, "                CudaOffloadLUFactorization,\n                CudaOffloadQRFactorization,\n",
             "                CudaOffloadLUFactorization,\n                CudaOffloadQRFactorization,\n"
             "                B200LUFactorization,\n                B200LU32MixedLUFactorization,\n",
             "hardware-dependent skip set (test/Core/resolve.jl:28-44)")
    return t


def edit_autotune_algorithms(t):
    t = sub1(t, "    # Metal algorithms for Apple Silicon\n    if is_metal_available()\n",
             "    # B200 dense LU: no package to load, the library either opens and finds a device or it does not\n"
             "    if LinearSolve.useb200()\n"
             "        push!(gpu_algs, B200LUFactorization())\n"
             "        push!(gpu_names, \"B200LUFactorization\")\n"
             "    end\n\n"
             "    # Metal algorithms for Apple Silicon\n    if is_metal_available()\n",
             "get_gpu_algorithms (lib/LinearSolveAutotune/src/algorithms.jl:98-139)")
    return t


def edit_autotune_module(t):
    t = sub1(t, "using LinearSolve: LinearSolve, AppleAccelerateLUFactorization,\n",
             "using LinearSolve: LinearSolve, AppleAccelerateLUFactorization, B200LUFactorization,\n",
             "imports (lib/LinearSolveAutotune/src/LinearSolveAutotune.jl:21-25)")
    return t


def edit_autotune_benchmarking(t):
    t = sub1(t, "            \"CudaOffloadLUFactorization\", \"CudaOffloadQRFactorization\", \"CudaOffloadFactorization\",\n",
             "            \"CudaOffloadLUFactorization\", \"CudaOffloadQRFactorization\", \"CudaOffloadFactorization\",\n"
             "            \"B200LUFactorization\",\n",
             "Float16 exclusion list of the GPU offload algorithms (lib/LinearSolveAutotune/src/benchmarking.jl:46-52)")
    return t


EDITS = [
    ("src/LinearSolve.jl", edit_linearsolve_jl),
    ("src/default.jl", edit_default_jl),
    ("src/adjoint_factorization.jl", edit_adjoint_factorization_jl),
    ("src/preferences.jl", edit_preferences_jl),
    ("test/Core/resolve.jl", edit_resolve_test),
    ("lib/LinearSolveAutotune/src/algorithms.jl", edit_autotune_algorithms),
    ("lib/LinearSolveAutotune/src/LinearSolveAutotune.jl", edit_autotune_module),
    ("lib/LinearSolveAutotune/src/benchmarking.jl", edit_autotune_benchmarking),
]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    for i, (rel, fn) in enumerate(EDITS, 1):
        old = open(os.path.join(ref, rel)).read()
        new = fn(old)
        diff = "".join(difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/" + rel, "b/" + rel, n=3))
        name = "%02d-%s.diff" % (i, rel.replace("/", "_"))
        open(os.path.join(HERE, name), "w").write(diff)
        print(name, diff.count("\n+") - 1, "lines added")
    # the new source file itself: src/b200lu.jl is julia/B200LUFactorization.jl
    src = open(os.path.join(HERE, "..", "B200LUFactorization.jl")).read()
    diff = "".join(difflib.unified_diff([], src.splitlines(True), "/dev/null", "b/src/b200lu.jl", n=0))
    open(os.path.join(HERE, "00-src_b200lu.jl.diff"), "w").write(diff)
    print("00-src_b200lu.jl.diff", len(src.splitlines()), "lines (new file)")


if __name__ == "__main__":
    main()
