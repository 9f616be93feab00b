// batched.cuh — many independent small systems (n <= 64 in registers, 65 ... 160 in shared memory,
// see getrf_batched_smem_kernel at the end), the reference's
// BlockDiagonal surface: per-block `lu!(B; check=false)` + per-block `ldiv!`
// (ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205).  Pivot/info semantics
// as src/blocked_lufact.jl:38-54,58-90.
//
// HBM-bound in the limit (read A once, write LU once: 66,816 bytes per 64x64 FP64
// system).  B200 mapping of getrf: one CTA of max(32, NMAX) threads per system, thread t
// holds ROW t in registers; rows never move (each thread tracks the logical position of
// its row, the LAPACK interchange sequence acts on positions).  Same instruction economy
// as the cluster panel kernel (panel_cluster.cuh): LEFT-SHIFTING LIVE WINDOW — the rank-1
// update writes element c to register c - 1, so the current column is always register 0 in
// a ROLLED loop (no 64-fold unrolled body that overflows the instruction cache, no
// per-element predicates); finished entries (multipliers, frozen pivot rows) go to a
// shared-memory tile that is written out once, permuted, fully coalesced; one barrier
// per column (each warp publishes its candidate row, every thread picks the winner).
// getrs: one warp per system, the packed LU staged in shared memory with cp.async, the
// substitution chains run on warp shuffles (no barriers); the row permutation is the
// `perm` vector the factorization leaves behind.
#pragma once
#include "common.cuh"
#include "panel_cluster.cuh"  // pcl_warp_argmax
#include <limits.h>

namespace b200lu {

__host__ __device__ constexpr int batched_threads(int nmax) { return nmax < 32 ? 32 : nmax; }

template <typename T, int NMAX>
struct BatchedShared {
    static constexpr int NW = (NMAX + 31) / 32;
    T tile[NMAX * NMAX];          // finished entry (row t, column c) at c * NMAX + ((t + c) % NMAX):
                                  // the skew makes row-wise AND column-wise warp accesses conflict-free
    T cand_row[2][NW][NMAX];      // per-warp candidate rows, double-buffered by column parity
    T cand_val[2][NW];
    T cand_rinv[2][NW];           // 1 / candidate value, computed by the candidate lane
    int cand_pos[2][NW];
    int cand_thr[2][NW];
    int red[NW];
};

// A: batch systems, column-major n x n (lda, strideA).  LU: packed n x n (ldlu = n).
// ipiv: 0-based LAPACK interchange sequence per system; perm[sys*n + p] = original row that
// ends at position p (for getrs); info: 0 or 1-based first zero pivot.
template <typename T, int NMAX>
__global__ void __launch_bounds__(batched_threads(NMAX)) getrf_batched_kernel(
    const T* __restrict__ A, long long lda, long long strideA, T* __restrict__ LU,
    long long ldlu, long long strideLU, int* __restrict__ ipiv, int* __restrict__ perm,
    int* __restrict__ info, int n) {
    constexpr int NW = (NMAX + 31) / 32;
    __shared__ __align__(16) BatchedShared<T, NMAX> sh;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long sys = blockIdx.x;
    const T* Ab = A + sys * strideA;

    T a[NMAX];   // a[c]: column (k + c) of my row while column k is being eliminated
#pragma unroll
    for (int c = 0; c < NMAX; ++c) a[c] = (t < n && c < n) ? Ab[(long long)c * lda + t] : T(0);
    int pos = t < n ? t : INT_MAX;   // logical position of my row; padding rows never compete
    bool done = t >= n;              // row already used as a pivot row (frozen)
    int myinfo = 0;

#define B200_BATCHED_COLUMN(LIVE)                                                                  \
    {                                                                                              \
        const int par = k & 1;                                                                     \
        /* pivot search: |a| max over the rows not yet used, lowest position on ties */            \
        /* zeros AND NaNs are no candidates (key 0): the redux works on bit patterns, where a NaN */ \
        /* would be the largest key and then fail the v > 0 test for the whole warp              */ \
        const T av_ = tabs(a[0]);                                                                  \
        const T v = (done || !(av_ > T(0))) ? T(0) : av_;                                          \
        const int wl = pcl_warp_argmax(v, pos);                                                    \
        if (lane == wl) {                                                                          \
            if (NW == 1) {                                                                         \
                _Pragma("unroll") for (int c = 0; c < (LIVE); ++c) sh.cand_row[par][0][c] = a[c];  \
            }                                                                                      \
            sh.cand_val[par][warp] = v;                                                            \
            sh.cand_rinv[par][warp] = T(1) / a[0];   /* off the other threads' critical path */    \
            sh.cand_pos[par][warp] = pos;                                                          \
            sh.cand_thr[par][warp] = t;                                                            \
        }                                                                                          \
        if (wl < 0 && lane == 0) { sh.cand_val[par][warp] = T(0); sh.cand_pos[par][warp] = INT_MAX; } \
        if (NW > 1) __syncthreads(); else __syncwarp();                                            \
        int bw = 0;                                                                                \
        T bv = sh.cand_val[par][0];                                                                \
        int bp = sh.cand_pos[par][0];                                                              \
        _Pragma("unroll") for (int w = 1; w < NW; ++w) {                                           \
            const T ov = sh.cand_val[par][w];                                                      \
            const int op = sh.cand_pos[par][w];                                                    \
            if (ov > bv || (ov == bv && op < bp)) { bv = ov; bp = op; bw = w; }                    \
        }                                                                                          \
        if (NW > 1) {                                                                              \
            /* only the system's winner stages its row (half the shared-memory stores of staging */ \
            /* every warp's candidate), at the price of a second barrier                         */ \
            if (bv > T(0) && t == sh.cand_thr[par][bw]) {                                          \
                _Pragma("unroll") for (int c = 0; c < (LIVE); ++c) sh.cand_row[par][0][c] = a[c];  \
            }                                                                                      \
            __syncthreads();                                                                       \
        }                                                                                          \
        const bool none = !(bv > T(0));   /* all-zero / all-NaN subcolumn: kp = k */               \
        if (none) {                                                                                \
            /* the "pivot" row is the row at position k: a second, rare exchange */                \
            if (NW > 1) __syncthreads(); else __syncwarp();                                        \
            if (!done && pos == k) {                                                               \
                _Pragma("unroll") for (int c = 0; c < (LIVE); ++c) sh.cand_row[par][0][c] = a[c];  \
                sh.cand_thr[par][0] = t;                                                           \
                sh.cand_rinv[par][0] = T(1) / a[0];                                                \
            }                                                                                      \
            if (NW > 1) __syncthreads(); else __syncwarp();                                        \
            bw = 0;                                                                                \
            bp = k;                                                                                \
        }                                                                                          \
        const T* prow = &sh.cand_row[par][0][0];                                                   \
        const int pthr = sh.cand_thr[par][bw];                                                     \
        const T pv = prow[0];                                                                      \
        /* the pivot row is frozen: all threads copy its staged entries into the tile */           \
        if (t < n - k) sh.tile[(k + t) * NMAX + ((pthr + k + t) & (NMAX - 1))] = prow[t];                                   \
        if (t == pthr) {                                                                           \
            done = true;                                                                           \
            pos = k;                                                                               \
            ipiv[sys * n + k] = bp;                                                                \
            if (pv == T(0) && myinfo == 0) myinfo = k + 1;                                         \
        } else if (!done && pos == k) {                                                            \
            pos = bp;                                                                              \
        }                                                                                          \
        /* multiplier and rank-1 update, shifted one register to the left (frozen and padding */   \
        /* rows compute garbage that is never read)                                           */   \
        T l = a[0];                                                                                \
        if (pv != T(0)) l *= sh.cand_rinv[par][bw];                                                        \
        if (!done) sh.tile[k * NMAX + ((t + k) & (NMAX - 1))] = l;                                                      \
        const T nl = -l;                                                                           \
        _Pragma("unroll") for (int c = 1; c < (LIVE); ++c) a[c - 1] = tfma(nl, prow[c], a[c]);     \
    }

    int k = 0;
    // copies of the rolled column loop with 64/56/.../8 live elements bound the dead elements
    // a column still computes on to < 8
#define B200_BATCHED_PHASE(LIVE)                                        \
    if constexpr (NMAX >= (LIVE)) {                                     \
        _Pragma("unroll 1") for (; k < n && k <= NMAX - ((LIVE) - 7); ++k) B200_BATCHED_COLUMN(LIVE) \
    }
    B200_BATCHED_PHASE(64)
    B200_BATCHED_PHASE(56)
    B200_BATCHED_PHASE(48)
    B200_BATCHED_PHASE(40)
    B200_BATCHED_PHASE(32)
    B200_BATCHED_PHASE(24)
    B200_BATCHED_PHASE(16)
    B200_BATCHED_PHASE(8)
#undef B200_BATCHED_PHASE
#undef B200_BATCHED_COLUMN

    // first zero pivot over the system (each pivot thread saw at most one)
    {
        int v = myinfo ? myinfo : INT_MAX;
        v = __reduce_min_sync(0xffffffffu, v);
        if (lane == 0) sh.red[warp] = v;
        __syncthreads();
        v = sh.red[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) v = min(v, sh.red[w]);
        if (t == 0) info[sys] = (v == INT_MAX) ? 0 : v;
    }
    // every entry of row t is in the tile now: write it at its final position
    if (t < n) {
        T* Lb = LU + sys * strideLU + pos;
        perm[sys * n + pos] = t;
#pragma unroll 8
        for (int c = 0; c < n; ++c) Lb[(long long)c * ldlu] = sh.tile[c * NMAX + ((t + c) & (NMAX - 1))];
    }
}

// getrs on the cached batched factors: X = U \ (L \ (P B)), nrhs columns.
// One warp per system (lane l holds rows l and l + 32); the packed LU is staged in shared
// memory; the substitution chains run on shuffles.  WPC systems per CTA.
template <typename T, int NMAX, int WPC>
__global__ void __launch_bounds__(32 * WPC) getrs_batched_kernel(
    const T* __restrict__ LU, long long ldlu, long long strideLU, const int* __restrict__ perm,
    const T* __restrict__ B, long long ldb, long long strideB, T* __restrict__ X, long long ldx,
    long long strideX, int n, int nrhs, long long batch) {
    constexpr int RPL = (NMAX + 31) / 32;
    extern __shared__ __align__(16) unsigned char getrs_b_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long sys = (long long)blockIdx.x * WPC + warp;
    if (sys >= batch) return;
    T* s_lu = reinterpret_cast<T*>(getrs_b_smem) + (size_t)warp * NMAX * NMAX;   // packed, ld = n
    const T* Lb = LU + sys * strideLU;
    {
        // packed n x n block: 16-byte chunks when the block is 16-byte sized and aligned
        const int total = n * n;
        constexpr int EPV = 16 / (int)sizeof(T);
        if (ldlu == n && (total % EPV) == 0 && ((strideLU * (long long)sizeof(T)) % 16) == 0) {
            for (int i = lane * EPV; i < total; i += 32 * EPV) cp_async16(s_lu + i, Lb + i, true);
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            for (int i = lane; i < total; i += 32) {
                const int c = i / n, r = i - c * n;
                s_lu[i] = Lb[(long long)c * ldlu + r];
            }
        }
        __syncwarp();
    }
    int src[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int row = r * 32 + lane;
        src[r] = row < n ? perm[sys * n + row] : 0;
    }
    for (int rhs = 0; rhs < nrhs; ++rhs) {
        T b[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            b[r] = (r * 32 + lane < n) ? B[sys * strideB + (long long)rhs * ldb + src[r]] : T(0);
        // forward: unit lower, column oriented
#pragma unroll
        for (int kr = 0; kr < RPL; ++kr) {
            const int kend = min(32, n - kr * 32);
#pragma unroll 4
            for (int kk = 0; kk < kend; ++kk) {
                const int k = kr * 32 + kk;
                const T xk = __shfl_sync(0xffffffffu, b[kr], kk);
#pragma unroll
                for (int r = kr; r < RPL; ++r) {
                    const int row = r * 32 + lane;
                    if (row > k && row < n) b[r] = tfma(-s_lu[k * n + row], xk, b[r]);
                }
            }
        }
        // backward: upper, divide by the diagonal (vector right-hand side form)
#pragma unroll
        for (int kr = RPL - 1; kr >= 0; --kr) {
            const int kend = min(32, n - kr * 32);
#pragma unroll 4
            for (int kk = kend - 1; kk >= 0; --kk) {
                const int k = kr * 32 + kk;
                if (lane == kk) b[kr] = b[kr] / s_lu[k * n + k];
                const T xk = __shfl_sync(0xffffffffu, b[kr], kk);
#pragma unroll
                for (int r = 0; r <= kr; ++r) {
                    const int row = r * 32 + lane;
                    if (row < k) b[r] = tfma(-s_lu[k * n + row], xk, b[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (r * 32 + lane < n) X[sys * strideX + (long long)rhs * ldx + r * 32 + lane] = b[r];
    }
}

// getrs with op(A) = A^T on the cached batched factors (the reference's `solve!(cache; adjoint = true)`,
// src/common.jl:1012-1027, for BlockDiagonal problems): P A = L U  =>  A^T = U^T L^T P, so
// U^T y = b (lower, divide by the diagonal), L^T z = y (unit upper), x = P^T z (x[perm[p]] = z[p]).
// Same mapping as getrs_batched_kernel (one warp per system, lane l holds rows l and l + 32,
// chains on shuffles); the factors are staged ROW-major with a padded leading dimension
// (s_t[r * (NMAX + 1) + c] = LU[r, c]) so that the column-oriented sweeps of the transposed
// triangles read consecutive shared-memory words.
template <typename T, int NMAX, int WPC>
__global__ void __launch_bounds__(32 * WPC) getrs_batched_trans_kernel(
    const T* __restrict__ LU, long long ldlu, long long strideLU, const int* __restrict__ perm,
    const T* __restrict__ B, long long ldb, long long strideB, T* __restrict__ X, long long ldx,
    long long strideX, int n, int nrhs, long long batch) {
    constexpr int RPL = (NMAX + 31) / 32;
    constexpr int LDS = NMAX + 1;
    extern __shared__ __align__(16) unsigned char getrs_bt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long sys = (long long)blockIdx.x * WPC + warp;
    if (sys >= batch) return;
    T* s_t = reinterpret_cast<T*>(getrs_bt_smem) + (size_t)warp * NMAX * LDS;
    const T* Lb = LU + sys * strideLU;
    for (int i = lane; i < n * n; i += 32) {
        const int c = i / n, r = i - c * n;
        s_t[r * LDS + c] = Lb[(long long)c * ldlu + r];
    }
    __syncwarp();
    int dst[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int row = r * 32 + lane;
        dst[r] = row < n ? perm[sys * n + row] : 0;
    }
    for (int rhs = 0; rhs < nrhs; ++rhs) {
        T b[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            b[r] = (r * 32 + lane < n) ? B[sys * strideB + (long long)rhs * ldb + r * 32 + lane] : T(0);
        // forward: U^T y = b; (U^T)[row, k] = U[k, row] = s_t[k * LDS + row] for row > k
#pragma unroll
        for (int kr = 0; kr < RPL; ++kr) {
            const int kend = min(32, n - kr * 32);
#pragma unroll 4
            for (int kk = 0; kk < kend; ++kk) {
                const int k = kr * 32 + kk;
                if (lane == kk) b[kr] = b[kr] / s_t[k * LDS + k];
                const T xk = __shfl_sync(0xffffffffu, b[kr], kk);
#pragma unroll
                for (int r = kr; r < RPL; ++r) {
                    const int row = r * 32 + lane;
                    if (row > k && row < n) b[r] = tfma(-s_t[k * LDS + row], xk, b[r]);
                }
            }
        }
        // backward: L^T z = y, unit diagonal; (L^T)[row, k] = L[k, row] = s_t[k * LDS + row] for row < k
#pragma unroll
        for (int kr = RPL - 1; kr >= 0; --kr) {
            const int kend = min(32, n - kr * 32);
#pragma unroll 4
            for (int kk = kend - 1; kk >= 0; --kk) {
                const int k = kr * 32 + kk;
                const T xk = __shfl_sync(0xffffffffu, b[kr], kk);
#pragma unroll
                for (int r = 0; r <= kr; ++r) {
                    const int row = r * 32 + lane;
                    if (row < k) b[r] = tfma(-s_t[k * LDS + row], xk, b[r]);
                }
            }
        }
        // x = P^T z: the entry at position p belongs to original row perm[p] (all loads of this
        // right-hand side are done, so X may alias B)
        __syncwarp();
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (r * 32 + lane < n) X[sys * strideX + (long long)rhs * ldx + dst[r]] = b[r];
    }
}

// ---------------------------------------------------------------------------------------------
// Blocks of 65 ... BATCHED_SMEM_NMAX rows (variable-size BlockDiagonal / supernode blocks, SURVEY
// 8(f)3): one CTA per system, the whole system in shared memory (column-major, odd leading
// dimension), unblocked right-looking getrf with the same arithmetic contract as the register
// kernel above — amax from 0 with strict '>', lowest row on ties (NaN never wins), kp = k and
// info = k on an all-zero subcolumn with the rank-1 update still run, multipliers scaled by the
// reciprocal of the pivot, FMA updates.  Three CTA barriers per column: after the pivot search,
// after the row interchange, after the rank-1 update (which reads the UNSCALED column k; the scaled
// multipliers are written behind that barrier by warp 0, and nothing reads them before the next
// interchange, which sits behind the next search barrier).
constexpr int BATCHED_SMEM_NMAX = 160;   // 160 x 161 FP64 words = 206,080 bytes of the 227 KB
constexpr int BATCHED_SMEM_NT = 256;

template <typename T>
__global__ void __launch_bounds__(BATCHED_SMEM_NT) getrf_batched_smem_kernel(
    const T* __restrict__ A, long long lda, long long strideA, T* __restrict__ LU, long long ldlu,
    long long strideLU, int* __restrict__ ipiv, int* __restrict__ perm, int* __restrict__ info, int n) {
    constexpr int NT = BATCHED_SMEM_NT, NW = NT / 32, QMAX = (BATCHED_SMEM_NMAX + 31) / 32;
    extern __shared__ __align__(16) unsigned char bsm_dyn[];
    T* s = reinterpret_cast<T*>(bsm_dyn);
    __shared__ T c_val[NW];
    __shared__ int c_pos[NW];
    __shared__ int s_perm[BATCHED_SMEM_NMAX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long sys = blockIdx.x;
    const int ld = n | 1;
    const T* Ab = A + sys * strideA;
    for (int i = tid; i < n * n; i += NT) {
        const int c = i / n, r = i - c * n;
        s[r + c * ld] = Ab[(long long)c * lda + r];
    }
    for (int i = tid; i < n; i += NT) s_perm[i] = i;
    int myinfo = 0;   // thread 0
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        // 1. pivot search over rows k .. n-1 of column k (n - k <= 160 < NT rows: one per thread)
        T v = T(0);
        int pos = INT_MAX;
        if (k + tid < n) {
            v = tabs(s[k + tid + k * ld]);
            pos = k + tid;
            if (!(v > T(0))) { v = T(0); pos = INT_MAX; }   // zeros and NaNs are no candidates
        }
        const int wl = pcl_warp_argmax(v, pos);
        if (lane == wl) { c_val[warp] = v; c_pos[warp] = pos; }
        if (wl < 0 && lane == 0) { c_val[warp] = T(0); c_pos[warp] = INT_MAX; }
        __syncthreads();
        T bv = c_val[0];
        int bp = c_pos[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const T ov = c_val[w];
            const int op = c_pos[w];
            if (ov > bv || (ov == bv && op < bp)) { bv = ov; bp = op; }
        }
        const int piv = (bv > T(0)) ? bp : k;   // all-zero / all-NaN subcolumn: kp = k
        // 2. interchange of rows k and piv over ALL columns (LAPACK: left and right of the panel)
        if (piv != k) {
            for (int c = tid; c < n; c += NT) {
                const T x = s[k + c * ld];
                s[k + c * ld] = s[piv + c * ld];
                s[piv + c * ld] = x;
            }
        }
        if (tid == 0) {
            ipiv[sys * n + k] = piv;
            if (piv != k) { const int x = s_perm[k]; s_perm[k] = s_perm[piv]; s_perm[piv] = x; }
        }
        __syncthreads();
        const T pv = s[k + k * ld];
        const bool scale = pv != T(0);
        if (tid == 0 && !scale && myinfo == 0) myinfo = k + 1;
        const T rinv = T(1) / pv;
        // 3. rank-1 update of the trailing block: lane <-> row (mod 32), warp <-> column (mod NW)
        T l[QMAX];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int r = k + 1 + lane + 32 * q;
            l[q] = T(0);
            if (r < n) {
                l[q] = s[r + k * ld];
                if (scale) l[q] *= rinv;
            }
        }
        for (int c = k + 1 + warp; c < n; c += NW) {
            const T u = s[k + c * ld];
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                const int r = k + 1 + lane + 32 * q;
                if (r < n) s[r + c * ld] = tfma(-l[q], u, s[r + c * ld]);
            }
        }
        __syncthreads();
        // 4. the scaled multipliers of column k
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                const int r = k + 1 + lane + 32 * q;
                if (r < n) s[r + k * ld] = l[q];
            }
        }
    }
    __syncthreads();
    T* Lb = LU + sys * strideLU;
    for (int i = tid; i < n * n; i += NT) {
        const int c = i / n, r = i - c * n;
        Lb[(long long)c * ldlu + r] = s[r + c * ld];
    }
    for (int i = tid; i < n; i += NT) perm[sys * n + i] = s_perm[i];
    if (tid == 0) info[sys] = myinfo;
}

// ---------------------------------------------------------------------------------------------
// getrf (+ optionally the first getrs) of one small system per WARP, the system resident in shared
// memory in LAPACK layout, blocked right-looking with panels of 8 columns:
//   * the panel (rows kb.. x 8 columns) lives in registers (lane l holds rows l and l + 32): per
//     column one redux arg-max, the winning lane stages its 8-wide row, one __syncwarp, scale +
//     update in registers; rows do not move inside a panel (positions are tracked as in the
//     register kernel above) and are written back at their final positions when the panel is done;
//   * the panel's 8 interchanges are then applied to all other columns (lane <-> column), U12 is
//     solved with the 8 x 8 unit-lower block (lane <-> column) and A22 -= L21 U12 runs lane <-> rows
//     with the 2 x 8 multipliers in registers and U12 broadcast from shared memory.
// Every entry sees the FMA sequence of the unblocked right-looking algorithm (pivot 0, 1, 2, ... in
// order), so factors and pivots are those of the reference's `generic_lufact!`
// (src/generic_lufact.jl:86-131) bit for bit; pivot / tie / zero-pivot / NaN rules as above.
// Why: the register kernel above keeps 12 warps per SM and spends ~210 instructions per column and
// warp staging and re-reading 64-wide rows (13 % of the HBM bound for 64 x 64 FP64).  Here a column
// step moves 8 values, not 64; the bulk of the flops is a straight FMA stream; one warp needs no
// CTA barrier; a system costs ~34 KB of shared memory (6 systems in flight per SM for FP64/64).
// SOLVE: the right-hand side is solved while the factors are still in shared memory and the
// factors are written out as well ("factors kept"): HBM traffic per 64 x 64 FP64 system is A in,
// LU out, pivots, b in, x out = 66,816 bytes — the bound of SURVEY §8(d).
template <typename T, int NMAX>
__host__ __device__ constexpr int bw_ld() { return NMAX + 16 / (int)sizeof(T); }   // 16-byte aligned columns, rows staggered over the banks
template <typename T, int NMAX>
__host__ __device__ constexpr size_t bw_smem_bytes() { return (size_t)NMAX * bw_ld<T, NMAX>() * sizeof(T) + 2 * NMAX * sizeof(int) + 2 * 8 * sizeof(T); }

template <typename T>
struct BwArgs {
    const T* A; long long lda, strideA;
    T* LU; long long ldlu, strideLU;
    int* ipiv; int* perm; int* info;
    int n;
    const T* B; long long strideB;   // SOLVE: one right-hand side per system
    T* X; long long strideX;
};

template <typename T, int NMAX, bool SOLVE>
__global__ void __launch_bounds__(32) getrf_batched_warp_kernel(BwArgs<T> p) {
    constexpr int RPL = NMAX / 32, LD = bw_ld<T, NMAX>(), EPV = 16 / (int)sizeof(T), NB = 8;
    static_assert(NMAX == 32 || NMAX == 64, "one or two rows per lane");
    extern __shared__ __align__(16) unsigned char bw_smem[];
    T* S = reinterpret_cast<T*>(bw_smem);                       // S[c * LD + r]: rows stay where they were loaded
    int* s_rowof = reinterpret_cast<int*>(S + NMAX * LD);       // physical row at each position (for the solve)
    int* s_ipiv = s_rowof + NMAX;
    T* s_stage = reinterpret_cast<T*>(s_ipiv + NMAX);           // [2][8]: the pivot row of the current column
    const int lane = threadIdx.x;
    const long long sys = blockIdx.x;
    const int n = p.n;
    const T* Ab = p.A + sys * p.strideA;
    {
        const bool vec = (n % EPV == 0) && (p.lda % EPV == 0) && (p.strideA % EPV == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
        if (vec) {
            const int cpc = n / EPV, total = cpc * n;
            for (int i = lane; i < total; i += 32) {
                const int c = i / cpc, r = (i - c * cpc) * EPV;
                cp_async16(S + c * LD + r, Ab + (long long)c * p.lda + r, true);
            }
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            for (int i = lane; i < n * n; i += 32) {
                const int c = i / n, r = i - c * n;
                S[c * LD + r] = Ab[(long long)c * p.lda + r];
            }
        }
    }
    __syncwarp();
    // Rows never move: pos[q] is the current position of physical row lane + 32 q in the LAPACK row order
    // (the interchange sequence acts on positions), act[q] says it has not been a pivot row yet.
    int pos[RPL];
    bool act[RPL];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int row = lane + 32 * q;
        act[q] = row < n;
        pos[q] = act[q] ? row : INT_MAX;
    }
    int myinfo = 0;   // uniform

    for (int kb = 0; kb < n; kb += NB) {
        const int jb = min(NB, n - kb);
        // ---- panel in registers (rows that are still active; finished rows are final already)
        T a[RPL][NB];
        bool inpanel[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            inpanel[q] = act[q];
#pragma unroll
            for (int c = 0; c < NB; ++c) a[q][c] = (act[q] && c < jb) ? S[(kb + c) * LD + lane + 32 * q] : T(0);
        }
        int pv[NB], pr[NB];   // uniform: position and physical row of the pivot of column kb + j
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            pv[j] = kb + j;
            pr[j] = 0;
            if (j < jb) {
                const int k = kb + j;
                // local candidate: strict '>' from 0 (zeros and NaNs are no candidates), lowest position on ties
                T best = T(0), cand = T(1);
                int bpos = INT_MAX;
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    if (act[q]) {
                        const T v = tabs(a[q][j]);
                        if (v > best || (v == best && v > T(0) && pos[q] < bpos)) { best = v; bpos = pos[q]; cand = a[q][j]; }
                    }
                }
                T rinv = T(1) / cand;                       // under the redux latency
                const int wl = pcl_warp_argmax(best, bpos);
                const bool none = wl < 0;                   // all-zero / all-NaN subcolumn: kp = k
                const int piv = none ? k : __shfl_sync(0xffffffffu, bpos, wl);
                // the pivot row (at position piv) stages its 8-wide row and tells its physical index
                T* stg = s_stage + (j & 1) * NB;
                int ownq = -1;
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    if (act[q] && pos[q] == piv) {
                        ownq = q;
#pragma unroll
                        for (int c = 0; c < NB; ++c) stg[c] = a[q][c];
                    }
                }
                const unsigned ob = __ballot_sync(0xffffffffu, ownq >= 0);
                const int ol = __ffs(ob) - 1;
                rinv = __shfl_sync(0xffffffffu, rinv, ol);
                pr[j] = ol + 32 * __shfl_sync(0xffffffffu, ownq, ol);
                __syncwarp();
                T prow[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) prow[c] = stg[c];
                bool scale = true;
                if (none) {
                    const T pvv = prow[j];
                    scale = (pvv != T(0));
                    rinv = T(1) / pvv;
                    if (!scale && myinfo == 0) myinfo = k + 1;
                }
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    if (act[q]) {
                        if (pos[q] == piv) { pos[q] = k; act[q] = false; }   // the pivot row: frozen
                        else if (pos[q] == k) pos[q] = piv;                  // the displaced top row stays active
                    }
                    if (act[q]) {
                        T l = a[q][j];
                        if (scale) l *= rinv;
                        a[q][j] = l;
#pragma unroll
                        for (int c = j + 1; c < NB; ++c) a[q][c] = tfma(-l, prow[c], a[q][c]);
                    }
                }
                pv[j] = piv;
            }
        }
        // ---- panel back to shared memory (rows in place)
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            if (inpanel[q]) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < jb) S[(kb + c) * LD + lane + 32 * q] = a[q][c];
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j)
            if (j < jb && lane == j) s_ipiv[kb + j] = pv[j];
        __syncwarp();
        if (kb + jb < n) {   // (then jb == NB)
            // ---- U12 = L11^{-1} A12 on the 8 pivot rows: lane <-> trailing column
            T L[NB][NB];
#pragma unroll
            for (int i = 1; i < NB; ++i)
#pragma unroll
                for (int i2 = 0; i2 < i; ++i2) L[i][i2] = S[(kb + i2) * LD + pr[i]];
            for (int c = kb + NB + lane; c < n; c += 32) {
                T* col = S + c * LD;
                T x[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) x[i] = col[pr[i]];
#pragma unroll
                for (int i = 1; i < NB; ++i)
#pragma unroll
                    for (int i2 = 0; i2 < i; ++i2) x[i] = tfma(-L[i][i2], x[i2], x[i]);
#pragma unroll
                for (int i = 1; i < NB; ++i) col[pr[i]] = x[i];
            }
            __syncwarp();
            // ---- A22 -= L21 U12 on the rows that are still active: lane <-> rows
            T l21[RPL][NB];
#pragma unroll
            for (int q = 0; q < RPL; ++q)
#pragma unroll
                for (int i = 0; i < NB; ++i) l21[q][i] = act[q] ? S[(kb + i) * LD + lane + 32 * q] : T(0);
#pragma unroll 4
            for (int c = kb + NB; c < n; ++c) {
                T* col = S + c * LD;
                T u[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) u[i] = col[pr[i]];
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    if (act[q]) {
                        T v = col[lane + 32 * q];
#pragma unroll
                        for (int i = 0; i < NB; ++i) v = tfma(-l21[q][i], u[i], v);
                        col[lane + 32 * q] = v;
                    }
                }
            }
            __syncwarp();
        }
    }

    // ---- outputs: info, pivots, permutation, the packed factors (every row stored at its final position)
    if (lane == 0) p.info[sys] = myinfo;
    for (int i = lane; i < n; i += 32) p.ipiv[sys * n + i] = s_ipiv[i];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int row = lane + 32 * q;
        if (row < n) {
            p.perm[sys * n + pos[q]] = row;
            s_rowof[pos[q]] = row;
        }
    }
    {
        T* Lb = p.LU + sys * p.strideLU;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int row = lane + 32 * q;
            if (row < n) {
                T* dst = Lb + pos[q];
                const T* src = S + row;
#pragma unroll 8
                for (int c = 0; c < n; ++c) dst[(long long)c * p.ldlu] = src[c * LD];
            }
        }
    }
    if constexpr (SOLVE) {
        // x = U \ (L \ (P b)) with the factors still in shared memory; b stays with its physical row, the sweeps
        // run in position order (s_rowof); the diagonal reciprocals are taken up front, off the chain
        __syncwarp();
        T b[RPL], rd[RPL];
        int ro[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int row = lane + 32 * q;
            b[q] = row < n ? p.B[sys * p.strideB + row] : T(0);
            rd[q] = row < n ? T(1) / S[pos[q] * LD + row] : T(1);
            ro[q] = row < n ? s_rowof[row] : 0;       // lane l, slot q: physical row at POSITION l + 32 q
        }
#pragma unroll
        for (int kr = 0; kr < RPL; ++kr) {
            const int kend = min(32, n - kr * 32);
#pragma unroll 4
            for (int kk = 0; kk < kend; ++kk) {
                const int k = kr * 32 + kk;
                const int r = __shfl_sync(0xffffffffu, ro[kr], kk);          // physical row at position k
                const T mine = (RPL > 1 && (r >> 5)) ? b[RPL - 1] : b[0];
                const T xk = __shfl_sync(0xffffffffu, mine, r & 31);
                const T* col = S + k * LD;
#pragma unroll
                for (int q = 0; q < RPL; ++q)
                    if (pos[q] > k && pos[q] != INT_MAX) b[q] = tfma(-col[lane + 32 * q], xk, b[q]);
            }
        }
#pragma unroll
        for (int kr = RPL - 1; kr >= 0; --kr) {
            const int kend = min(32, n - kr * 32);
#pragma unroll 4
            for (int kk = kend - 1; kk >= 0; --kk) {
                const int k = kr * 32 + kk;
                const int r = __shfl_sync(0xffffffffu, ro[kr], kk);
#pragma unroll
                for (int q = 0; q < RPL; ++q)
                    if (pos[q] == k) b[q] *= rd[q];
                const T mine = (RPL > 1 && (r >> 5)) ? b[RPL - 1] : b[0];
                const T xk = __shfl_sync(0xffffffffu, mine, r & 31);
                const T* col = S + k * LD;
#pragma unroll
                for (int q = 0; q < RPL; ++q)
                    if (pos[q] < k) b[q] = tfma(-col[lane + 32 * q], xk, b[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < RPL; ++q)
            if (lane + 32 * q < n) p.X[sys * p.strideX + pos[q]] = b[q];
    }
}

}  // namespace b200lu
