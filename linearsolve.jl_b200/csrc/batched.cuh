// batched.cuh — many independent small systems (n <= 64), the reference's
// BlockDiagonal surface: per-block `lu!(B; check=false)` + per-block `ldiv!`
// (ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205).  Pivot/info semantics
// as src/blocked_lufact.jl:38-54,58-90.
//
// HBM-bound (read A once, write LU once).  B200 mapping: one CTA of NMAX threads
// per system, thread t holds ROW t in registers for the whole factorization;
// rows never move — each thread tracks the logical position of its row
// (implicit permutation), the pivot row tail is broadcast through shared
// memory, the pivot search is a warp-shuffle arg-max (lowest position on ties).
#pragma once
#include "common.cuh"
#include "panel.cuh"  // warp_argmax
#include <limits.h>

namespace b200lu {

__host__ __device__ constexpr int batched_threads(int nmax) { return nmax < 32 ? 32 : nmax; }

template <typename T, int NMAX>
__global__ void __launch_bounds__(batched_threads(NMAX)) getrf_batched_kernel(
    const T* __restrict__ A, long long lda, long long strideA, T* __restrict__ LU,
    long long ldlu, long long strideLU, int* __restrict__ ipiv, int* __restrict__ info, int n) {
    constexpr int NW = (NMAX + 31) / 32;  // blockDim.x == max(32, NMAX): lanes >= NMAX are padding rows
    constexpr int NPAD = NMAX < 2 ? 2 : NMAX;
    // double-buffered by step parity so one barrier per exchange suffices
    __shared__ __align__(16) T s_row[2][NPAD];
    __shared__ T s_val[2][NW];
    __shared__ int s_pos[2][NW];
    __shared__ int s_thr[2][NW];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long sys = blockIdx.x;
    const T* Ab = A + sys * strideA;

    T a[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c)
        a[c] = (t < n && c < n) ? Ab[(long long)c * lda + t] : (t == c ? T(1) : T(0));
    int pos = t;        // logical row position of this thread's row (t >= n: never a candidate)
    bool done = false;  // row already used as a pivot row
    int myinfo = 0;

#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
        if (k < n) {
            const int par = k & 1;
            // pivot search: |a| max over the rows not yet used, lowest position on ties
            T best = T(0);
            int bp = INT_MAX;
            if (!done && t < n) {
                const T v = tabs(a[k]);
                if (v > best) { best = v; bp = pos; }
            }
            int wl;
            warp_argmax(best, bp, wl);
            int bt = warp * 32 + wl;   // thread holding the warp's winner (unused when none)
            if (NW > 1) {
                if (lane == 0) { s_val[par][warp] = best; s_pos[par][warp] = bp; s_thr[par][warp] = bt; }
                __syncthreads();
                best = s_val[par][0]; bp = s_pos[par][0]; bt = s_thr[par][0];
#pragma unroll
                for (int w = 1; w < NW; ++w) {
                    const T ob = s_val[par][w];
                    const int op = s_pos[par][w];
                    if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; bt = s_thr[par][w]; }
                }
            }
            // all-zero / all-NaN subcolumn: kp = k, the row at position k is the "pivot" row
            const bool none = !(best > T(0));
            const bool i_am_piv = none ? (pos == k && !done) : (t == bt);
            if (i_am_piv) {
#pragma unroll
                for (int c = k; c < NMAX; ++c) s_row[par][c] = a[c];
            }
            __syncthreads();
            const T pv = s_row[par][k];
            const int ppos = none ? k : bp;
            if (i_am_piv) {
                // this row lands at position k; whoever sat at k takes my old position
                done = true;
                pos = k;
                ipiv[sys * n + k] = ppos;
                if (pv == T(0) && myinfo == 0) myinfo = k + 1;
            } else if (!done && pos == k) {
                pos = ppos;
            }
            if (!done && t < n) {
                T l = a[k];
                if (pv != T(0)) l *= (T(1) / pv);
                a[k] = l;
#pragma unroll
                for (int c = k + 1; c < NMAX; ++c) a[c] = tfma(-l, s_row[par][c], a[c]);
            }
        }
    }
    // first zero pivot over the system (each pivot thread saw at most one)
    {
        int v = myinfo ? myinfo : INT_MAX;
        v = __reduce_min_sync(0xffffffffu, v);
        __syncthreads();
        if (NW > 1) {
            if (lane == 0) s_pos[0][warp] = v;
            __syncthreads();
            v = s_pos[0][0];
#pragma unroll
            for (int w = 1; w < NW; ++w) v = min(v, s_pos[0][w]);
        }
        if (t == 0) info[sys] = (v == INT_MAX) ? 0 : v;
    }
    if (t < n) {
        T* Lb = LU + sys * strideLU;
#pragma unroll
        for (int c = 0; c < NMAX; ++c)
            if (c < n) Lb[(long long)c * ldlu + pos] = a[c];
    }
}

// getrs on the cached batched factors: X = U \ (L \ (P B)), nrhs columns.
// One CTA of NMAX threads per system; thread t holds row t of the packed LU.
template <typename T, int NMAX>
__global__ void __launch_bounds__(batched_threads(NMAX)) getrs_batched_kernel(
    const T* __restrict__ LU, long long ldlu, long long strideLU, const int* __restrict__ ipiv,
    const T* __restrict__ B, long long ldb, long long strideB, T* __restrict__ X, long long ldx,
    long long strideX, int n, int nrhs) {
    __shared__ T s_b[batched_threads(NMAX)];
    __shared__ int s_perm[batched_threads(NMAX)];
    const int t = threadIdx.x;
    const long long sys = blockIdx.x;
    const T* Lb = LU + sys * strideLU;
    T a[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) a[c] = (t < n && c < n) ? Lb[(long long)c * ldlu + t] : T(0);
    s_perm[t] = t;
    __syncthreads();
    if (t == 0) {
        for (int k = 0; k < n; ++k) {
            const int p = ipiv[sys * n + k];
            if (p != k) { int tmp = s_perm[k]; s_perm[k] = s_perm[p]; s_perm[p] = tmp; }
        }
    }
    __syncthreads();
    const int src = s_perm[t];
    for (int r = 0; r < nrhs; ++r) {
        T b = (t < n) ? B[sys * strideB + (long long)r * ldb + src] : T(0);
        // forward: unit lower, column oriented
#pragma unroll
        for (int k = 0; k < NMAX; ++k) {
            if (k < n) {
                if (t == k) s_b[k] = b;
                __syncthreads();
                if (t > k) b = tfma(-a[k], s_b[k], b);
            }
        }
        // backward: upper, divide by the diagonal (vector right-hand side form)
#pragma unroll
        for (int k = NMAX - 1; k >= 0; --k) {
            if (k < n) {
                if (t == k) { b = b / a[k]; s_b[k] = b; }
                __syncthreads();
                if (t < k) b = tfma(-a[k], s_b[k], b);
            }
        }
        if (t < n) X[sys * strideX + (long long)r * ldx + t] = b;
        __syncthreads();
    }
}

}  // namespace b200lu
