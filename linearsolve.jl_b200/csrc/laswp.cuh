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

// One warp. ipiv: global 0-based rows; interchanges k0 .. k0+nsw-1.
__global__ void laswp_plan_kernel(const int* __restrict__ ipiv, int k0, int nsw,
                                  LaswpPlan* __restrict__ plan) {
    __shared__ int top[LASWP_MAXSW];
    __shared__ int ext_row[LASWP_MAXSW];
    __shared__ int ext_cont[LASWP_MAXSW];
    __shared__ int s_piv[LASWP_MAXSW];
    const int lane = threadIdx.x;
    for (int i = lane; i < nsw; i += 32) {
        top[i] = k0 + i;
        s_piv[i] = ipiv[k0 + i];
    }
    __syncwarp();
    int ne = 0;
    for (int k = 0; k < nsw; ++k) {
        const int pv = s_piv[k];
        if (pv == k0 + k) continue;
        if (pv < k0 + nsw) {
            if (lane == 0) {
                int t = top[k];
                top[k] = top[pv - k0];
                top[pv - k0] = t;
            }
        } else {
            int found = -1;
            for (int e = lane; e < ne; e += 32)
                if (ext_row[e] == pv) found = e;
            unsigned msk = __ballot_sync(0xffffffffu, found >= 0);
            int e;
            if (msk) {
                e = __shfl_sync(0xffffffffu, found, __ffs(msk) - 1);
            } else {
                e = ne++;
                if (lane == 0) { ext_row[e] = pv; ext_cont[e] = pv; }
            }
            __syncwarp();
            if (lane == 0) {
                int t = top[k];
                top[k] = ext_cont[e];
                ext_cont[e] = t;
            }
        }
        __syncwarp();
    }
    // compact: only entries that actually move
    int cnt = 0;
    for (int base = 0; base < nsw + ne; base += 32) {
        const int i = base + lane;
        int d = -1, s = -1;
        if (i < nsw) { d = k0 + i; s = top[i]; }
        else if (i < nsw + ne) { d = ext_row[i - nsw]; s = ext_cont[i - nsw]; }
        const bool mv = (i < nsw + ne) && (d != s);
        unsigned msk = __ballot_sync(0xffffffffu, mv);
        if (mv) {
            int pos = cnt + __popc(msk & ((1u << lane) - 1));
            plan->dst[pos] = d;
            plan->src[pos] = s;
        }
        cnt += __popc(msk);
    }
    if (lane == 0) plan->n_tot = cnt;
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
