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

}  // namespace b200lu
