/*
 * lu_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into or called by the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it).
 *
 * Plain-C CPU restatement of the reference's in-tree dense LU path
 * (SciML/LinearSolve.jl v5.12.0).  Each function cites the reference lines it
 * follows.  The reference is pure Julia and cannot run in this environment
 * (no Julia binary), so this restatement is pinned instead against
 *   (a) the known-answer cases of the reference's own tests
 *       (test/Core/blocked_lufact.jl, test/Core/retcodes.jl, test/Trim/runtests.jl,
 *        test/Core/resolve.jl) — see tests/test_oracle.py, and
 *   (b) LAPACK dgetrf/dgetrs/sgetrf/sgetrs from scipy's OpenBLAS, which IS the
 *       arithmetic behind the reference's LUFactorization/OpenBLASLUFactorization
 *       (src/factorization.jl:632-637, src/openblas.jl:144-151).
 *
 * Conventions: column-major, leading dimension lda (elements), ipiv 1-based
 * int64 (Julia BlasInt), info = first zero pivot (1-based) or 0.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; FMAs are explicit so
 * the result does not depend on compiler contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define A_(i, j) A[(size_t)(j) * lda + (i)]

/* ------------------------------------------------------------------ double */

/* Row-maximum pivot search, two passes: `>`-select max from amax = 0 (NaN
 * compares false and is ignored), then first index equal to the max.
 * reference: src/blocked_lufact.jl:38-54 (_blocked_lu_find_pivot).
 * k, m are 0-based here; returns the 0-based pivot row. */
static int64_t find_pivot_d(const double* A, int64_t lda, int64_t k, int64_t m, double* amax_out) {
    double amax = 0.0;
    for (int64_t i = k; i < m; ++i) {
        double a = fabs(A_(i, k));
        amax = (a > amax) ? a : amax;
    }
    int64_t kp = k;
    if (amax != 0.0) {
        for (int64_t i = k; i < m; ++i) {
            if (fabs(A_(i, k)) == amax) { kp = i; break; }
        }
    }
    if (amax_out) *amax_out = amax;
    return kp;
}

/* Scalar right-looking LU with the stdlib pivot rule: amax starts at
 * abs(A[k,k]), strict `>`; zero pivot => info = k once, continue.
 * reference: src/generic_lufact.jl:86-131 (RowMaximum branch).
 * Note `A[i,j] -= A[i,k]*A[k,j]` is NOT a fused multiply-add there. */
int64_t oracle_generic_lufact_d(double* A, int64_t lda, int64_t m, int64_t n, int64_t* ipiv) {
    int64_t minmn = m < n ? m : n, info = 0;
    for (int64_t k = 0; k < minmn; ++k) {
        int64_t kp = k;
        if (k < m - 1) {
            double amax = fabs(A_(k, k));
            for (int64_t i = k + 1; i < m; ++i) {
                double a = fabs(A_(i, k));
                if (a > amax) { kp = i; amax = a; }
            }
        }
        ipiv[k] = kp + 1;
        if (A_(kp, k) != 0.0) {
            if (k != kp)
                for (int64_t j = 0; j < n; ++j) { double t = A_(k, j); A_(k, j) = A_(kp, j); A_(kp, j) = t; }
            double inv = 1.0 / A_(k, k);
            for (int64_t i = k + 1; i < m; ++i) A_(i, k) *= inv;
        } else if (info == 0) {
            info = k + 1;
        }
        for (int64_t j = k + 1; j < n; ++j)
            for (int64_t i = k + 1; i < m; ++i) {
                volatile double p = A_(i, k) * A_(k, j); /* forbid contraction */
                A_(i, j) -= p;
            }
    }
    return info;
}

/* Unblocked kernel of the blocked file (used for min(m,n) <= 8).
 * reference: src/blocked_lufact.jl:58-90 (_blocked_lu_unblocked!), muladd => fma. */
int64_t oracle_unblocked_lufact_d(double* A, int64_t lda, int64_t m, int64_t n, int64_t* ipiv) {
    int64_t minmn = m < n ? m : n, info = 0;
    for (int64_t k = 0; k < minmn; ++k) {
        int64_t kp = find_pivot_d(A, lda, k, m, NULL);
        ipiv[k] = kp + 1;
        if (A_(kp, k) != 0.0) {
            if (k != kp)
                for (int64_t j = 0; j < n; ++j) { double t = A_(k, j); A_(k, j) = A_(kp, j); A_(kp, j) = t; }
            double inv = 1.0 / A_(k, k);
            for (int64_t i = k + 1; i < m; ++i) A_(i, k) *= inv;
        } else if (info == 0) {
            info = k + 1;
        }
        for (int64_t j = k + 1; j < n; ++j) {
            double akj = A_(k, j);
            for (int64_t i = k + 1; i < m; ++i) A_(i, j) = fma(-A_(i, k), akj, A_(i, j));
        }
    }
    return info;
}

/* Panel factorization: unblocked LU of A[j0:m, j0:j1] (0-based, j1 exclusive),
 * swaps applied inside the panel only.  reference: src/blocked_lufact.jl:93-122. */
static int64_t panel_d(double* A, int64_t lda, int64_t* ipiv, int64_t m, int64_t j0, int64_t j1,
                       int64_t info) {
    for (int64_t k = j0; k < j1; ++k) {
        int64_t kp = find_pivot_d(A, lda, k, m, NULL);
        ipiv[k] = kp + 1;
        if (A_(kp, k) != 0.0) {
            if (k != kp)
                for (int64_t j = j0; j < j1; ++j) { double t = A_(k, j); A_(k, j) = A_(kp, j); A_(kp, j) = t; }
            double inv = 1.0 / A_(k, k);
            for (int64_t i = k + 1; i < m; ++i) A_(i, k) *= inv;
        } else if (info == 0) {
            info = k + 1;
        }
        for (int64_t j = k + 1; j < j1; ++j) {
            double akj = A_(k, j);
            for (int64_t i = k + 1; i < m; ++i) A_(i, j) = fma(-A_(i, k), akj, A_(i, j));
        }
    }
    return info;
}

/* reference: src/blocked_lufact.jl:126-141 (_blocked_lu_swap_rows!) */
static void swap_rows_d(double* A, int64_t lda, const int64_t* ipiv, int64_t k0, int64_t k1,
                        int64_t c0, int64_t c1) {
    for (int64_t j = c0; j < c1; ++j)
        for (int64_t k = k0; k < k1; ++k) {
            int64_t kp = ipiv[k] - 1;
            if (kp != k) { double t = A_(k, j); A_(k, j) = A_(kp, j); A_(kp, j) = t; }
        }
}

/* U12 := L11 \ A12, unit lower, forward substitution with FMAs.
 * reference: src/blocked_lufact.jl:146-178 (_blocked_lu_trsm_unit_lower!) */
static void trsm_unit_lower_d(double* A, int64_t lda, int64_t j0, int64_t j1, int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1; ++c)
        for (int64_t k = j0; k < j1; ++k) {
            double bk = A_(k, c);
            for (int64_t i = k + 1; i < j1; ++i) A_(i, c) = fma(-A_(i, k), bk, A_(i, c));
        }
}

/* C -= L21 * U12, k-ordered FMA chain per element (the order every variant of
 * the reference's Schur kernels uses: pivot columns ascending).
 * reference: src/blocked_lufact.jl:186-620 (_blocked_lu_schur!) */
static void schur_d(double* A, int64_t lda, int64_t m, int64_t j0, int64_t j1, int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1; ++c)
        for (int64_t k = j0; k < j1; ++k) {
            double ukc = A_(k, c);
            for (int64_t i = j1; i < m; ++i) A_(i, c) = fma(-A_(i, k), ukc, A_(i, c));
        }
}

/* Blocked right-looking driver. reference: src/blocked_lufact.jl:658-679 */
int64_t oracle_blocked_lufact_d(double* A, int64_t lda, int64_t m, int64_t n, int64_t nb, int64_t* ipiv) {
    int64_t minmn = m < n ? m : n, info = 0;
    for (int64_t j0 = 0; j0 < minmn; j0 += nb) {
        int64_t jb = (nb < minmn - j0) ? nb : (minmn - j0);
        int64_t j1 = j0 + jb;
        info = panel_d(A, lda, ipiv, m, j0, j1, info);
        swap_rows_d(A, lda, ipiv, j0, j1, 0, j0);
        if (j1 < n) {
            swap_rows_d(A, lda, ipiv, j0, j1, j1, n);
            trsm_unit_lower_d(A, lda, j0, j1, j1, n);
            if (j1 < m) schur_d(A, lda, m, j0, j1, j1, n);
        }
    }
    return info;
}

/* Dispatch of `generic_lufact!(A::StridedMatrix{Float64}, RowMaximum(), ipiv)`:
 * unblocked for min(m,n) <= 8, else blocked with nb = 8 (minmn <= 160) or 16.
 * reference: src/blocked_lufact.jl:10-15,711-743 */
int64_t oracle_reference_lufact_d(double* A, int64_t lda, int64_t m, int64_t n, int64_t* ipiv) {
    int64_t minmn = m < n ? m : n;
    if (minmn <= 8) return oracle_unblocked_lufact_d(A, lda, m, n, ipiv);
    return oracle_blocked_lufact_d(A, lda, m, n, minmn <= 160 ? 8 : 16, ipiv);
}

/* getrs, vector and matrix right-hand sides.  Vector form divides by the
 * diagonal; matrix form multiplies by its inverse.
 * reference: src/factorization.jl:433-491 (_naive_lu_ldiv!) */
void oracle_lu_ldiv_d(const double* A, int64_t lda, int64_t n, const int64_t* ipiv, double* B,
                      int64_t ldb, int64_t nrhs, int matrix_form) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t p = ipiv[i] - 1;
        if (p != i)
            for (int64_t c = 0; c < nrhs; ++c) {
                double t = B[c * ldb + i]; B[c * ldb + i] = B[c * ldb + p]; B[c * ldb + p] = t;
            }
    }
    for (int64_t j = 0; j < n; ++j)
        for (int64_t c = 0; c < nrhs; ++c) {
            double bj = B[c * ldb + j];
            for (int64_t i = j + 1; i < n; ++i) B[c * ldb + i] = fma(-A_(i, j), bj, B[c * ldb + i]);
        }
    for (int64_t j = n - 1; j >= 0; --j) {
        double invd = 1.0 / A_(j, j);
        for (int64_t c = 0; c < nrhs; ++c) {
            if (matrix_form) B[c * ldb + j] *= invd; else B[c * ldb + j] /= A_(j, j);
            double bj = B[c * ldb + j];
            for (int64_t i = 0; i < j; ++i) B[c * ldb + i] = fma(-A_(i, j), bj, B[c * ldb + i]);
        }
    }
}

/* Per-block LU + solve of a block-diagonal system with equal block size
 * (BlockDiagonal surface).  reference: ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205
 * (each block: lu!(B; check=false) => LAPACK getrf; here the in-tree kernel). */
int64_t oracle_batched_lufact_solve_d(double* A, int64_t n, int64_t batch, int64_t* ipiv, int64_t* info,
                                      double* B, int64_t nrhs) {
    int64_t bad = 0;
    for (int64_t s = 0; s < batch; ++s) {
        double* As = A + (size_t)s * n * n;
        info[s] = oracle_reference_lufact_d(As, n, n, n, ipiv + s * n);
        if (info[s]) { ++bad; continue; }
        if (B) oracle_lu_ldiv_d(As, n, n, ipiv + s * n, B + (size_t)s * n * nrhs, n, nrhs, nrhs > 1);
    }
    return bad;
}

/* ------------------------------------------------------------------- float */
#undef A_
#define A_(i, j) A[(size_t)(j) * lda + (i)]

static int64_t find_pivot_s(const float* A, int64_t lda, int64_t k, int64_t m) {
    float amax = 0.0f;
    for (int64_t i = k; i < m; ++i) {
        float a = fabsf(A_(i, k));
        amax = (a > amax) ? a : amax;
    }
    int64_t kp = k;
    if (amax != 0.0f)
        for (int64_t i = k; i < m; ++i)
            if (fabsf(A_(i, k)) == amax) { kp = i; break; }
    return kp;
}

/* Same algorithm as oracle_blocked_lufact_d in Float32 (the reference kernel is
 * generic over Float32/Float64, src/blocked_lufact.jl:695-743). */
int64_t oracle_blocked_lufact_s(float* A, int64_t lda, int64_t m, int64_t n, int64_t nb, int64_t* ipiv) {
    int64_t minmn = m < n ? m : n, info = 0;
    for (int64_t j0 = 0; j0 < minmn; j0 += nb) {
        int64_t jb = (nb < minmn - j0) ? nb : (minmn - j0);
        int64_t j1 = j0 + jb;
        for (int64_t k = j0; k < j1; ++k) {
            int64_t kp = find_pivot_s(A, lda, k, m);
            ipiv[k] = kp + 1;
            if (A_(kp, k) != 0.0f) {
                if (k != kp)
                    for (int64_t j = j0; j < j1; ++j) { float t = A_(k, j); A_(k, j) = A_(kp, j); A_(kp, j) = t; }
                float inv = 1.0f / A_(k, k);
                for (int64_t i = k + 1; i < m; ++i) A_(i, k) *= inv;
            } else if (info == 0) {
                info = k + 1;
            }
            for (int64_t j = k + 1; j < j1; ++j) {
                float akj = A_(k, j);
                for (int64_t i = k + 1; i < m; ++i) A_(i, j) = fmaf(-A_(i, k), akj, A_(i, j));
            }
        }
        for (int pass = 0; pass < 2; ++pass) {
            int64_t c0 = pass == 0 ? 0 : j1, c1 = pass == 0 ? j0 : n;
            for (int64_t j = c0; j < c1; ++j)
                for (int64_t k = j0; k < j1; ++k) {
                    int64_t kp = ipiv[k] - 1;
                    if (kp != k) { float t = A_(k, j); A_(k, j) = A_(kp, j); A_(kp, j) = t; }
                }
        }
        if (j1 < n) {
            for (int64_t c = j1; c < n; ++c)
                for (int64_t k = j0; k < j1; ++k) {
                    float bk = A_(k, c);
                    for (int64_t i = k + 1; i < j1; ++i) A_(i, c) = fmaf(-A_(i, k), bk, A_(i, c));
                }
            if (j1 < m)
                for (int64_t c = j1; c < n; ++c)
                    for (int64_t k = j0; k < j1; ++k) {
                        float ukc = A_(k, c);
                        for (int64_t i = j1; i < m; ++i) A_(i, c) = fmaf(-A_(i, k), ukc, A_(i, c));
                    }
        }
    }
    return info;
}

int64_t oracle_reference_lufact_s(float* A, int64_t lda, int64_t m, int64_t n, int64_t* ipiv) {
    int64_t minmn = m < n ? m : n;
    if (minmn <= 8) return oracle_blocked_lufact_s(A, lda, m, n, minmn > 0 ? minmn : 1, ipiv);
    return oracle_blocked_lufact_s(A, lda, m, n, minmn <= 160 ? 8 : 16, ipiv);
}

void oracle_lu_ldiv_s(const float* A, int64_t lda, int64_t n, const int64_t* ipiv, float* B,
                      int64_t ldb, int64_t nrhs, int matrix_form) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t p = ipiv[i] - 1;
        if (p != i)
            for (int64_t c = 0; c < nrhs; ++c) {
                float t = B[c * ldb + i]; B[c * ldb + i] = B[c * ldb + p]; B[c * ldb + p] = t;
            }
    }
    for (int64_t j = 0; j < n; ++j)
        for (int64_t c = 0; c < nrhs; ++c) {
            float bj = B[c * ldb + j];
            for (int64_t i = j + 1; i < n; ++i) B[c * ldb + i] = fmaf(-A_(i, j), bj, B[c * ldb + i]);
        }
    for (int64_t j = n - 1; j >= 0; --j) {
        float invd = 1.0f / A_(j, j);
        for (int64_t c = 0; c < nrhs; ++c) {
            if (matrix_form) B[c * ldb + j] *= invd; else B[c * ldb + j] /= A_(j, j);
            float bj = B[c * ldb + j];
            for (int64_t i = 0; i < j; ++i) B[c * ldb + i] = fmaf(-A_(i, j), bj, B[c * ldb + i]);
        }
    }
}
