"""Pin the CPU oracle (oracle/lu_oracle.{c,py}) against every known-answer case the
reference's own tests hold for this path, and against LAPACK
(scipy OpenBLAS dgetrf/sgetrf = the arithmetic behind the reference's
LUFactorization).  Runs on CPU (`-m "not gpu"`)."""
import json
import os

import numpy as np
import pytest

from conftest import decisive_matrix

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_residual_square_default_params(oracle):
    """test/Core/blocked_lufact.jl:44-54 (F64 and F32)"""
    rng = np.random.default_rng(1234)
    for dtype in (np.float64, np.float32):
        for n in (1, 2, 3, 5, 7, 8, 9, 13, 16, 17, 31, 32, 33, 40, 41, 63, 64, 65, 100, 127, 128, 129,
                  200, 250, 256, 257, 300):
            A = rng.standard_normal((n, n)).astype(dtype)
            F, ipiv, info = oracle.ref_lufact(A)
            assert info == 0
            assert oracle.scaled_residual(A, F, ipiv) < 20, (dtype, n)


def test_forced_blocked_driver(oracle):
    """test/Core/blocked_lufact.jl:56-70"""
    rng = np.random.default_rng(5)
    for n in (11, 41, 67, 70, 97, 130, 190, 257):
        for nb in (4, 5, 8, 13, 16, 32):
            if n <= nb:
                continue
            A = rng.standard_normal((n, n))
            F, ipiv, info = oracle.ref_lufact(A, "blocked", nb)
            assert oracle.scaled_residual(A, F, ipiv) < 20
            # blocked and unblocked orders agree on the pivots of a generic matrix
            _, ipiv_u, _ = oracle.ref_lufact(A, "unblocked")
            assert np.array_equal(ipiv, ipiv_u)


def test_rectangular(oracle):
    """test/Core/blocked_lufact.jl:119-130"""
    rng = np.random.default_rng(6)
    for m, n in ((100, 3), (3, 100), (128, 40), (40, 128), (257, 130), (130, 257), (65, 64), (64, 65)):
        A = rng.standard_normal((m, n))
        F, ipiv, info = oracle.ref_lufact(A)
        assert oracle.scaled_residual(A, F, ipiv) < 20


def test_pivots_match_lapack_decisive(oracle):
    """test/Core/blocked_lufact.jl:148-161"""
    rng = np.random.default_rng(42)
    for n in (17, 64, 129, 300):
        for _ in range(3):
            A = decisive_matrix(rng, n)
            _, ipiv_l, _ = oracle.lapack_getrf(A)
            F, ipiv, info = oracle.ref_lufact(A)
            assert np.array_equal(ipiv, ipiv_l)
            _, ipiv_g, _ = oracle.ref_lufact(A, "generic")
            assert np.array_equal(ipiv_g, ipiv_l)
            assert oracle.scaled_residual(A, F, ipiv) < 20


def test_permutation_matrix_exact(oracle):
    """test/Core/blocked_lufact.jl:163-173"""
    rng = np.random.default_rng(3)
    for n in (16, 65, 200):
        A = np.eye(n)[rng.permutation(n), :]
        lu_l, ipiv_l, info_l = oracle.lapack_getrf(A)
        F, ipiv, info = oracle.ref_lufact(A)
        assert info == 0 == info_l
        assert np.array_equal(ipiv, ipiv_l)
        assert np.array_equal(F, lu_l)


def test_wilkinson_growth(oracle):
    """test/Core/blocked_lufact.jl:175-184"""
    for n in (24, 53):
        A = np.eye(n) - np.tril(np.ones((n, n)), -1)
        A[:, n - 1] = 1.0
        F, ipiv, info = oracle.ref_lufact(A)
        assert np.array_equal(ipiv, np.arange(1, n + 1))
        assert F[n - 1, n - 1] == 2.0 ** (n - 1)
        assert oracle.scaled_residual(A, F, ipiv) < 20


def test_singularity(oracle):
    """test/Core/blocked_lufact.jl:186-214"""
    rng = np.random.default_rng(11)
    for n in (10, 50, 130):
        for zc in (0, 3, n - 1):
            A = rng.standard_normal((n, n))
            A[:, zc] = 0.0
            _, _, info_l = oracle.lapack_getrf(A)
            for variant in ("reference", "generic", "unblocked"):
                _, _, info = oracle.ref_lufact(A, variant)
                assert info > 0 and info == info_l, (n, zc, variant)
    assert oracle.ref_lufact(np.zeros((50, 50)))[2] == 1
    An = rng.standard_normal((30, 30))
    An[1, 1] = np.nan
    F, _, _ = oracle.ref_lufact(An)
    assert np.isnan(F).any()


def test_known_answers_2x2(oracle):
    """test/Core/retcodes.jl:17-18,41-42; test/Core/resolve.jl:84-96; test/Trim/runtests.jl:7"""
    def solve(A, b):
        F, ipiv, info = oracle.ref_lufact(np.array(A, dtype=np.float64))
        return info, (oracle.ref_ldiv(F, ipiv, np.array(b, dtype=np.float64)) if info == 0 else None)
    info, x = solve([[2, 1], [-1, 1]], [-1, 1])
    assert info == 0 and np.allclose(x, np.linalg.solve([[2, 1], [-1, 1]], [-1, 1]))
    assert solve([[1, 1], [1, 1]], [1, 1])[0] > 0
    A = np.array([[1.0, 2.0], [3.0, 4.0]])
    info, x = solve(A.T @ A, [1, 2])
    assert np.allclose(x, [-2.0, 1.5], rtol=1e-12)
    info, x = solve([[4, 1], [1, 3]], [1, 2])
    assert np.allclose(x, [0.09090909090909091, 0.6363636363636364], rtol=1e-14)


def test_ldiv_matches_getrs(oracle):
    """test/Core/genericlu_naive_ldiv.jl:28-46: back-solve vs getrs! to rtol 1e-14-ish"""
    rng = np.random.default_rng(8)
    for n in (2, 4, 8, 9, 16, 51, 100, 257):
        A = rng.random((n, n)) + n * np.eye(n)
        b = rng.random(n)
        B = rng.random((n, 4))
        lu_l, ipiv_l, _ = oracle.lapack_getrf(A)
        np.testing.assert_allclose(oracle.ref_ldiv(lu_l, ipiv_l, b), oracle.lapack_getrs(lu_l, ipiv_l, b), rtol=1e-12)
        np.testing.assert_allclose(oracle.ref_ldiv(lu_l, ipiv_l, B), oracle.lapack_getrs(lu_l, ipiv_l, B), rtol=1e-12)
        F, ipiv, _ = oracle.ref_lufact(A)
        x = oracle.ref_ldiv(F, ipiv, b)
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=100 * np.finfo(float).eps * n)


def test_mixed_precision_reference_bar(oracle):
    """test/Core/test_mixed_precision.jl:9-31: 100x100 rand+5I, FP32 LU => rel err < 1e-5"""
    rng = np.random.default_rng(123)
    n = 100
    A = rng.random((n, n)) + 5 * np.eye(n)
    b = rng.random(n)
    F, ipiv, info = oracle.ref_lufact(A.astype(np.float32))
    x = oracle.ref_ldiv(F, ipiv, b.astype(np.float32)).astype(np.float64)
    xr = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-5
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-5


def test_refactor_sequence_n51(oracle):
    """test/Core/direct_blas_refactorization.jl:15-46 on the oracle"""
    rng = np.random.default_rng(42)
    n = 51
    A1 = rng.random((n, n)) + n * np.eye(n)
    A2 = rng.random((n, n)) + n * np.eye(n)
    b = rng.random(n)
    for Ak in (A1, A2, A1, A2):
        F, ipiv, info = oracle.ref_lufact(Ak)
        assert info == 0
        np.testing.assert_allclose(oracle.ref_ldiv(F, ipiv, b), np.linalg.solve(Ak, b), rtol=1e-10)
    As = A1.copy(); As[:, 0] = 0
    assert oracle.ref_lufact(As)[2] == 1


def test_batched_blocks(oracle):
    """BlockDiagonal per-block LU: blocks [3,3,3,3]-style equal blocks and a singular block"""
    rng = np.random.default_rng(1)
    A = rng.random((4, 3, 3)) + 3 * np.eye(3)
    b = rng.random((4, 3))
    F, ipiv, info, x = oracle.ref_batched(A, b)
    assert not info.any()
    for s in range(4):
        np.testing.assert_allclose(x[s], np.linalg.solve(A[s].T, b[s]), rtol=1e-12)
    A[2] = 1.0
    assert oracle.ref_batched(A, b)[2][2] > 0


def test_compare_ipiv_tie_logic(oracle):
    A = np.array([[1.0, 2.0], [1.0, 3.0]], order="F")      # exact tie in column 1
    assert oracle.compare_ipiv(A, [1, 2], [1, 2])[1] == "exact"
    assert oracle.compare_ipiv(A, [2, 2], [1, 2])[1] == "tie"
    B = np.array([[1.0, 2.0], [5.0, 3.0]], order="F")
    assert oracle.compare_ipiv(B, [1, 2], [2, 2])[1] == "mismatch"


def test_golden_fixtures(oracle):
    """committed golden vectors (tests/golden/make_golden.py): inputs are regenerated
    from their seeds; expected ipiv / info / checksums were produced by LAPACK here."""
    meta = json.load(open(os.path.join(GOLD, "golden.json")))
    data = np.load(os.path.join(GOLD, "golden.npz"))
    from golden.make_golden import make_case
    for case in meta["cases"]:
        A, b = make_case(case)
        lu_l, ipiv_l, info_l = oracle.lapack_getrf(A)
        assert np.array_equal(ipiv_l, data[case["name"] + "_ipiv"])
        assert info_l == case["info"]
        F, ipiv, info = oracle.ref_lufact(A)
        assert info == case["info"]
        if case["decisive"]:
            assert np.array_equal(ipiv, data[case["name"] + "_ipiv"])
        if info_l == 0:
            x = oracle.ref_ldiv(F, ipiv, b)
            np.testing.assert_allclose(x, data[case["name"] + "_x"], rtol=1e-9)
