// trsm.cuh — U12 := L11 \ A12 with L11 unit lower (reference
// `_blocked_lu_trsm_unit_lower!`, src/blocked_lufact.jl:146-178: forward
// substitution, column-oriented FMA updates, no division).
//
// B200 mapping: each warp owns CC right-hand-side columns entirely in
// registers (lane l holds rows l, l+32, ...); L11 streams through shared
// memory in 32-column chunks shared by the CTA's warps; the solved value of
// row k is broadcast with a warp shuffle.  In place: a column is read and
// written only by its own warp.
#pragma once
#include "common.cuh"

namespace b200lu {

// L: w x w unit-lower block at Lp (leading dim ldl).  B: w x ncols at Bp (ldb).
// RPL = ceil(w / 32) rows per lane (compile-time), CC columns per warp.
template <typename T, int RPL, int CC, int NWARP>
__global__ void __launch_bounds__(NWARP * 32) trsm_lunit_kernel(const T* __restrict__ Lp,
                                                                long long ldl, T* __restrict__ Bp,
                                                                long long ldb, int w, int ncols) {
    constexpr int WMAXR = RPL * 32;
    extern __shared__ unsigned char smem_raw[];
    T* Ls = reinterpret_cast<T*>(smem_raw);  // [32][WMAXR]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int col0 = (blockIdx.x * NWARP + warp) * CC;

    T b[RPL][CC];
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            const int row = r * 32 + lane;
            b[r][c] = (row < w && col0 + c < ncols) ? Bp[(long long)(col0 + c) * ldb + row] : T(0);
        }

    // L11 chunk r (32 columns) is prefetched into registers while chunk r-1 is being
    // consumed, so the global-load latency of the staging never sits on the k-chain.
    constexpr int NT = NWARP * 32;
    constexpr int PER = 32 * WMAXR / NT;  // elements of one chunk per thread
    T pf[PER];
    auto prefetch = [&](int r) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int idx = threadIdx.x + u * NT;
            const int kk = idx / WMAXR;
            const int row = idx - kk * WMAXR;
            const int k = r * 32 + kk;
            pf[u] = (row < w && k < w && row > k) ? Lp[(long long)k * ldl + row] : T(0);
        }
    };
    prefetch(0);
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        if (r * 32 < w) {
            __syncthreads();
#pragma unroll
            for (int u = 0; u < PER; ++u) Ls[threadIdx.x + u * NT] = pf[u];
            __syncthreads();
            if ((r + 1) * 32 < w) prefetch(r + 1);
#pragma unroll 4
            for (int kk = 0; kk < 32; ++kk) {
                T xk[CC];
#pragma unroll
                for (int c = 0; c < CC; ++c) xk[c] = shfl(b[r][c], kk);
#pragma unroll
                for (int rr = r; rr < RPL; ++rr) {
                    const T l = Ls[kk * WMAXR + rr * 32 + lane];  // 0 for rows <= k
#pragma unroll
                    for (int c = 0; c < CC; ++c) b[rr][c] = tfma(-l, xk[c], b[rr][c]);
                }
            }
        }
    }

#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            const int row = r * 32 + lane;
            if (row < w && col0 + c < ncols) Bp[(long long)(col0 + c) * ldb + row] = b[r][c];
        }
}

}  // namespace b200lu
