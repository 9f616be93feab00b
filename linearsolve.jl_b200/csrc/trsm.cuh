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

#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        if (r * 32 < w) {
            // stage L[:, 32r .. 32r+31] (rows >= 32r) into shared memory
            __syncthreads();
            for (int idx = threadIdx.x; idx < 32 * WMAXR; idx += NWARP * 32) {
                const int kk = idx / WMAXR;
                const int row = idx - kk * WMAXR;
                const int k = r * 32 + kk;
                Ls[idx] = (row < w && k < w && row > k) ? Lp[(long long)k * ldl + row] : T(0);
            }
            __syncthreads();
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
