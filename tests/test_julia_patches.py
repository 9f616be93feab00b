"""The Julia-side wiring of `B200LUFactorization` as real diffs against the reference tree (v5.12.0):
linearsolve.jl_b200/julia/patches/*.diff.  CPU-only; needs the read-only reference checkout at
/root/reference (present where the driver runs the `-m "not gpu"` suite; skipped elsewhere, e.g. on
the GPU box).  Checks:
  * every diff applies cleanly (`patch -p1 --dry-run`) to a scratch copy of the reference;
  * the committed diffs are what make_patches.py generates today (no stale patch);
  * after applying, every place the newest `DefaultAlgorithmChoice` member (`LHLFactorization`) and the
    GPU-offload sibling (`CudaOffloadLUFactorization`) touch in the selection machinery is also touched
    for `B200LUFactorization` (SURVEY §8b selection checklist);
  * the Python `defaultalg` twin selects what src/default.jl:427-475 selects on the bands of
    test/Core/default_algs.jl:4-66 when the library is unavailable.
"""
import glob
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = os.path.join(ROOT, "linearsolve.jl_b200", "julia", "patches")
FILES = ["src/LinearSolve.jl", "src/default.jl", "src/adjoint_factorization.jl", "src/preferences.jl",
         "test/Core/resolve.jl", "lib/LinearSolveAutotune/src/algorithms.jl",
         "lib/LinearSolveAutotune/src/LinearSolveAutotune.jl", "lib/LinearSolveAutotune/src/benchmarking.jl"]

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present")


def _scratch(tmp_path):
    for rel in FILES:
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, rel), dst)
    return tmp_path


@needs_ref
def test_patches_apply_and_cover_the_enum_footprint(tmp_path):
    work = _scratch(tmp_path)
    diffs = sorted(glob.glob(os.path.join(PATCHES, "*.diff")))
    assert len(diffs) == 9
    for d in diffs:
        for extra in (["--dry-run"], []):
            out = subprocess.run(["patch", "-p1", "--batch", "--forward", "-i", d] + extra, cwd=work,
                                 capture_output=True, text=True)
            assert out.returncode == 0, (d, out.stdout, out.stderr)
            assert "fuzz" not in out.stdout and "offset" not in out.stdout, (d, out.stdout)
    assert (work / "src" / "b200lu.jl").read_text() == open(
        os.path.join(ROOT, "linearsolve.jl_b200", "julia", "B200LUFactorization.jl")).read()
    ls = (work / "src" / "LinearSolve.jl").read_text()
    de = (work / "src" / "default.jl").read_text()
    # enum member LAST (fields of DefaultLinearSolverInit are positional, in enum order: src/default.jl:689-693)
    enum = re.search(r"@enumx DefaultAlgorithmChoice begin\n(.*?)\nend", ls, re.S).group(1).split()
    assert enum[-2:] == ["LHLFactorization", "B200LUFactorization"]
    fields = re.findall(r"^    (\w+!?)::T(\d+)$", de, re.M)
    assert [f for f, _ in fields] == enum and [int(i) for _, i in fields] == list(range(1, len(enum) + 1))
    assert "T26, T27,\n        TA, Tb, TR," in de
    # every selection site of the footprint
    assert "alg === DefaultAlgorithmChoice.B200LUFactorization\n        return useb200()" in ls
    assert 'include("b200lu.jl")' in ls and ls.index('include("openblas.jl")') < ls.index('include("b200lu.jl")') < ls.index('include("default.jl")')
    assert ":B200LUFactorization, :B200LU32MixedLUFactorization," in ls          # needs_square_A
    assert "export B200LUFactorization, B200LU32MixedLUFactorization" in ls
    assert "elseif alg === :B200LUFactorization\n        B200LUFactorization(throwerror = false)" in de     # algchoice_to_alg
    assert ":(B200LUFactorization(throwerror = false, residualsafety = alg.residualsafety))" in de     # _algchoice_to_alg_with_safety
    assert "alg == Symbol(DefaultAlgorithmChoice.B200LUFactorization)\n            inner_alg_expr" in de  # generated solve! + LU->QR fallback
    assert de.count("_default_lu_solve_with_fallback(cache, alg, sol)") == open(os.path.join(REF, "src", "default.jl")).read().count("_default_lu_solve_with_fallback(cache, alg, sol)") + 1
    assert "alg == Symbol(DefaultAlgorithmChoice.B200LUFactorization)\n            quote\n                # getrs with trans" in de  # adjoint eval
    arm = de.index("DefaultAlgorithmChoice.B200LUFactorization\n                    elseif appleaccelerate_isavailable()")
    assert de.index("if tuned_alg !== nothing") < arm                                  # autotune preference still wins
    assert "matrix_size >= B200LU_DEFAULT_MIN_N" in de and "&& useb200()" in de            # gated on availability
    adj = (work / "src" / "adjoint_factorization.jl").read_text()
    assert re.search(r"B200LUFactorization,\n        B200LU32MixedLUFactorization,\n    \)\n    @eval _adjoint_factorization_reuse\(::Type\{<:\$Alg\}\) =\n        _CustomAdjointFactorizationReuse\(\)", adj)
    assert 'algorithm_name == "B200LUFactorization"' in (work / "src" / "preferences.jl").read_text()
    assert "B200LUFactorization,\n                B200LU32MixedLUFactorization," in (work / "test" / "Core" / "resolve.jl").read_text()
    at = (work / "lib" / "LinearSolveAutotune" / "src" / "algorithms.jl").read_text()
    assert 'push!(gpu_names, "B200LUFactorization")' in at and "LinearSolve.useb200()" in at
    # the same sites the two reference members touch are touched for the new one (file by file)
    for rel, member in (("src/LinearSolve.jl", "LHLFactorization"), ("src/default.jl", "CudaOffloadLUFactorization"),
                        ("src/preferences.jl", "CudaOffloadLUFactorization"), ("src/adjoint_factorization.jl", "LHLFactorization")):
        txt = (work / rel).read_text()
        code = [ln for ln in txt.splitlines() if not ln.lstrip().startswith("#")]
        n_ref = sum(member in ln and "using " not in ln and '"' + member not in ln.replace("algorithm_name ==", "") for ln in code)
        n_new = sum("B200LUFactorization" in ln for ln in code)
        assert n_new >= 1 and n_new >= min(n_ref, 4) - 1, (rel, member, n_ref, n_new)
    # what the glue must define for those sites
    glue = (work / "src" / "b200lu.jl").read_text()
    for sym in ("useb200()", "const B200LU_DEFAULT_MIN_N", "function _b200lu_solve_trans!", "_get_residualsafety(alg::B200LUFactorization)",
                "_custom_can_reuse_adjoint_factorization", "_custom_adjoint_factorization_solve", "function init_cacheval(",
                "ReturnCode.APosterioriSafetyFailure", "@get_cacheval(cache, :B200LUFactorization)"):
        assert sym in glue, sym


@needs_ref
def test_committed_patches_are_current(tmp_path):
    out = tmp_path / "patches"
    shutil.copytree(PATCHES, out)
    for d in glob.glob(str(out / "*.diff")):
        os.remove(d)
    shutil.copytree(os.path.join(ROOT, "linearsolve.jl_b200", "julia"), tmp_path / "julia", dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns("patches"))
    shutil.copytree(out, tmp_path / "julia" / "patches", dirs_exist_ok=True)
    r = subprocess.run([sys.executable, str(tmp_path / "julia" / "patches" / "make_patches.py"), REF], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for d in sorted(glob.glob(os.path.join(PATCHES, "*.diff"))):
        assert open(d).read() == open(tmp_path / "julia" / "patches" / os.path.basename(d)).read(), os.path.basename(d)


def test_python_defaultalg_matches_reference_bands(ls):
    """src/default.jl:427-475 on the sizes of test/Core/default_algs.jl:4-66, library unavailable: the new arm must
    be invisible; with the library it takes over from n = 1024 for Float32/Float64 only."""
    C = ls.DefaultAlgorithmChoice

    def pick(n, dtype=np.float64, **kw):
        A = np.zeros((1, 1), dtype=dtype)
        b = np.zeros(n, dtype=dtype)
        return ls.defaultalg(A, b, b200_available=kw.pop("avail", False), **kw).alg

    # OpenBLAS host with RecursiveFactorization (the reference CI's configuration)
    assert pick(1) == pick(10) == C.GenericLUFactorization
    assert pick(11) == pick(100) == pick(500) == C.RFLUFactorization
    assert pick(501) == pick(600) == pick(1000) == pick(5000) == C.LUFactorization
    # without RecursiveFactorization: blocked generic kernel through 256, then LAPACK
    assert pick(50, userecursivefactorization=False) == pick(256, userecursivefactorization=False) == C.GenericLUFactorization
    assert pick(257, userecursivefactorization=False) == C.LUFactorization
    # MKL host
    assert pick(150, isopenblas=False, usemkl=True) == C.RFLUFactorization
    assert pick(201, isopenblas=False, usemkl=True) == C.MKLLUFactorization
    assert pick(33, isopenblas=False, usemkl=True, userecursivefactorization=False) == C.MKLLUFactorization
    assert pick(32, isopenblas=False, usemkl=True, userecursivefactorization=False) == C.GenericLUFactorization
    # with the library: unchanged below 1024, the GPU arm from 1024, real floating point only
    for n in (1, 10, 11, 100, 500, 600, 1000, 1023):
        assert pick(n, avail=True) == pick(n, avail=False)
    assert pick(1024, avail=True) == pick(32768, avail=True) == pick(2000, np.float32, avail=True) == C.B200LUFactorization
    assert pick(4096, np.complex128, avail=True) == C.LUFactorization
    cond = ls.OperatorAssumptions(issq=True, condition="VeryIllConditioned")
    assert ls.defaultalg(np.zeros((1, 1)), np.zeros(4096), cond, b200_available=True).alg == C.QRFactorization
