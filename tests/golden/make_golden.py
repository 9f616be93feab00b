"""Generates tests/golden/golden.{json,npz}: small seeded inputs with the pivot
vector / info / solution produced HERE by LAPACK (scipy OpenBLAS dgetrf/dgetrs —
the arithmetic behind the reference's LUFactorization).  The reference itself is
Julia and cannot be imported; the case list restates the matrices of its tests
(test/Core/blocked_lufact.jl:148-198, test/Core/direct_blas_refactorization.jl:15-23).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(case):
    rng = np.random.default_rng(case["seed"])
    n = case["n"]
    kind = case["kind"]
    if kind == "decisive":
        mags = 2.0 ** rng.permutation(np.arange(1, n + 1)).astype(np.float64)
        A = np.diag(mags) @ (np.eye(n) + 0.01 * rng.standard_normal((n, n)))
        A = A[rng.permutation(n), :]
    elif kind == "permutation":
        A = np.eye(n)[rng.permutation(n), :]
    elif kind == "wilkinson":
        A = np.eye(n) - np.tril(np.ones((n, n)), -1)
        A[:, n - 1] = 1.0
    elif kind == "zero_column":
        A = rng.standard_normal((n, n))
        A[:, case["zc"]] = 0.0
    elif kind == "shifted_uniform":
        A = rng.random((n, n)) + n * np.eye(n)
    elif kind == "uniform":
        A = rng.random((n, n))
    else:
        raise ValueError(kind)
    b = rng.random(n)
    return np.asfortranarray(A), b


CASES = (
    [{"name": f"decisive_{n}", "kind": "decisive", "n": n, "seed": 100 + n, "decisive": True} for n in (17, 64, 129, 300)]
    + [{"name": f"perm_{n}", "kind": "permutation", "n": n, "seed": 200 + n, "decisive": True} for n in (16, 65, 200)]
    + [{"name": f"wilk_{n}", "kind": "wilkinson", "n": n, "seed": 1, "decisive": True} for n in (24, 53)]
    + [{"name": f"zerocol_{n}_{zc}", "kind": "zero_column", "n": n, "zc": zc, "seed": 300 + n + zc, "decisive": False}
       for n in (10, 50, 130) for zc in (0, 3)]
    + [{"name": "shifted_51", "kind": "shifted_uniform", "n": 51, "seed": 42, "decisive": True},
       {"name": "uniform_200", "kind": "uniform", "n": 200, "seed": 123, "decisive": False}]
)


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import lu_oracle
    arrays, meta = {}, []
    for case in CASES:
        A, b = make_case(case)
        lu, ipiv, info = lu_oracle.lapack_getrf(A)
        case = dict(case, info=info)
        arrays[case["name"] + "_ipiv"] = ipiv
        if info == 0:
            arrays[case["name"] + "_x"] = lu_oracle.lapack_getrs(lu, ipiv, b)
        meta.append(case)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    json.dump({"generator": "tests/golden/make_golden.py", "lapack": "scipy OpenBLAS dgetrf/dgetrs", "cases": meta},
              open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print(f"wrote {len(meta)} cases")


if __name__ == "__main__":
    main()
