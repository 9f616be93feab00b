// laswp.cuh — row interchanges outside the panel (reference
// `_blocked_lu_swap_rows!`, src/blocked_lufact.jl:126-141) as a two-kernel
// gather: (1) one warp folds the jb sequential interchanges into a net
// "which original row ends up where" map touching <= 2*jb rows; (2) a
// bandwidth kernel reads every affected element once and writes it once
// (all reads of a column chunk happen before its writes).
// Also: pivots -> permutation vector for getrs (reference `_naive_lu_ldiv!`
// pivot loop, src/factorization.jl:437-443).
#pragma once
#include "common.cuh"

namespace b200lu {

constexpr int LASWP_MAXSW = 256;  // max interchanges folded by one plan

struct LaswpPlan {
    int n_tot;                     // entries used in dst/src
    int dst[2 * LASWP_MAXSW];      // destination row
    int src[2 * LASWP_MAXSW];      // original row whose content lands there
};

// ipiv: global 0-based rows; interchanges k0 .. k0+nsw-1 (nsw <= LASWP_MAXSW).
// One CTA of 2*LASWP_MAXSW threads.  Thread t < nsw owns top position k0+t;
// thread LASWP_MAXSW+t owns the outside pivot row ipiv[k0+t] (first occurrence
// only).  Each thread finds where the content that ends up at ITS position
// came from by walking the interchange sequence backwards — positions are
// independent, so the jb sequential swaps cost O(jb) steps in parallel instead
// of a serial simulation.
// body shared by laswp_plan_kernel and the distributed engine's receive kernel: s_piv[0 .. nsw) (shared memory)
// holds the pivots; all 2 * LASWP_MAXSW threads of the CTA take part
__device__ __forceinline__ void laswp_plan_body(const int* s_piv, int* s_cnt, int k0, int nsw, LaswpPlan* __restrict__ plan) {
    const int t = threadIdx.x;
    // Both loops have a uniform trip count and no early exit: an early `break` here left
    // the warp diverged for the rest of the kernel and serialised the 256-step walks
    // thread by thread (110 us instead of a few).
    const bool is_top = t < LASWP_MAXSW;
    const int k_me = is_top ? t : t - LASWP_MAXSW;
    const int pv_me = (k_me < nsw) ? s_piv[k_me] : -1;
    bool dup = false;
    for (int kk = 0; kk < nsw; ++kk) dup |= (kk < k_me) & (s_piv[kk] == pv_me);
    int pos = -1;
    if (k_me < nsw) {
        if (is_top) pos = k0 + t;
        else if (pv_me >= k0 + nsw && !dup) pos = pv_me;   // outside pivot row, first occurrence
    }
    int y = pos;
    for (int k = nsw - 1; k >= 0; --k) {
        const int a = k0 + k, b = s_piv[k];
        y = (y == a) ? b : ((y == b) ? a : y);
    }
    if (pos >= 0 && y != pos) {
        const int slot = atomicAdd(s_cnt, 1);
        plan->dst[slot] = pos;
        plan->src[slot] = y;
    }
    __syncthreads();
    if (t == 0) plan->n_tot = *s_cnt;
}

__global__ void __launch_bounds__(2 * LASWP_MAXSW)
    laswp_plan_kernel(const int* __restrict__ ipiv, int k0, int nsw, LaswpPlan* __restrict__ plan) {
    __shared__ int s_piv[LASWP_MAXSW];
    __shared__ int s_cnt;
    const int t = threadIdx.x;
    if (t < nsw) s_piv[t] = ipiv[k0 + t];
    if (t == 0) s_cnt = 0;
    __syncthreads();
    laswp_plan_body(s_piv, &s_cnt, k0, nsw, plan);
}

// Columns [c0, c1) of A get the planned row gather. Each CTA walks column
// chunks of CW; thread e owns plan entries e, e+NT.
template <typename T, int CW, int NT>
__global__ void __launch_bounds__(NT) laswp_apply_kernel(T* __restrict__ A, long long lda, int c0,
                                                         int c1,
                                                         const LaswpPlan* __restrict__ plan) {
    const int n_tot = plan->n_tot;
    if (n_tot == 0) return;
    constexpr int EPT = (2 * LASWP_MAXSW + NT - 1) / NT;
    int d[EPT], s[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const int e = threadIdx.x + r * NT;
        d[r] = (e < n_tot) ? plan->dst[e] : -1;
        s[r] = (e < n_tot) ? plan->src[e] : -1;
    }
    for (int cb = c0 + blockIdx.x * CW; cb < c1; cb += gridDim.x * CW) {
        T v[EPT][CW];
#pragma unroll
        for (int r = 0; r < EPT; ++r)
#pragma unroll
            for (int c = 0; c < CW; ++c)
                if (d[r] >= 0 && cb + c < c1) v[r][c] = A[(long long)(cb + c) * lda + s[r]];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < EPT; ++r)
#pragma unroll
            for (int c = 0; c < CW; ++c)
                if (d[r] >= 0 && cb + c < c1) A[(long long)(cb + c) * lda + d[r]] = v[r][c];
        // chunks are disjoint columns: no barrier needed before the next one
    }
}

// perm[i] = original row of b that lands at position i after all n interchanges.
// Single CTA; the index vector lives in shared memory when it fits (n <= 48K),
// otherwise in global scratch (`perm` itself).
__global__ void ipiv_to_perm_kernel(const int* __restrict__ ipiv, int n, int* __restrict__ perm,
                                    int use_smem) {
    extern __shared__ int sp[];
    int* w = use_smem ? sp : perm;
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < n; ++k) {
            const int pv = ipiv[k];
            if (pv != k) {
                int t = w[k];
                w[k] = w[pv];
                w[pv] = t;
            }
        }
    }
    __syncthreads();
    if (use_smem)
        for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = sp[i];
}

// X[i, r] = B[perm[i], r]  (or the inverse scatter for the transposed solve)
template <typename T>
__global__ void permute_rows_kernel(const T* __restrict__ B, long long ldb, T* __restrict__ X,
                                    long long ldx, const int* __restrict__ perm, int n, int nrhs,
                                    int inverse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= n || r >= nrhs) return;
    const int pi = perm[i];
    if (!inverse) X[(long long)r * ldx + i] = B[(long long)r * ldb + pi];
    else X[(long long)r * ldx + pi] = B[(long long)r * ldb + i];
}

__global__ void ipiv_to_i64_kernel(const int* __restrict__ ipiv, long long* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (long long)ipiv[i] + 1;  // 0-based device -> 1-based LAPACK
}

}  // namespace b200lu
