// trsv.cuh — getrs: forward (unit lower) and backward (upper) substitution
// (reference `_naive_lu_ldiv!`, src/factorization.jl:433-491, and LAPACK
// getrs as called at src/openblas.jl:247-278).
//
// Bandwidth-bound (8 n^2 bytes per right-hand side).  B200 mapping: ONE launch
// per triangular sweep; CTA = one 64-row block row, taken in dependency order
// from an atomic ticket.  A CTA streams its off-diagonal 64x64 blocks from HBM
// (next block prefetched in registers while it waits), multiplies each by the
// already-solved segment of x as soon as that segment's release flag shows up,
// and finishes with a 64x64 mat-vec against the pre-inverted diagonal block.
// No host round trips, no kernel-per-block launches.
#pragma once
#include "common.cuh"

namespace b200lu {

constexpr int TRSV_TB = 64;

// Inverse of every 64x64 diagonal block: unit-lower part of L (upper==0) or the
// upper triangle U (upper==1).  dinv: [nblk][64*64] column-major; the ragged
// last block is padded with the identity.  One CTA of 64 threads per block.
template <typename T>
__global__ void __launch_bounds__(TRSV_TB) trtri_diag_kernel(const T* __restrict__ A, long long lda,
                                                             int n, T* __restrict__ dinv,
                                                             int upper) {
    extern __shared__ __align__(16) unsigned char trtri_smem[];
    T(*Ls)[TRSV_TB + 1] = reinterpret_cast<T(*)[TRSV_TB + 1]>(trtri_smem);  // Ls[col][row]
    T(*Ys)[TRSV_TB] = reinterpret_cast<T(*)[TRSV_TB]>(trtri_smem + sizeof(T) * TRSV_TB * (TRSV_TB + 1));  // Ys[row][col j]
    const int blk = blockIdx.x;
    const int j = threadIdx.x;
    const int base = blk * TRSV_TB;
    for (int c = 0; c < TRSV_TB; ++c) {
        const int gr = base + j, gc = base + c;
        T v = (gr < n && gc < n) ? A[(long long)gc * lda + gr] : (j == c ? T(1) : T(0));
        if (!upper) v = (j > c) ? v : (j == c ? T(1) : T(0));
        else v = (j <= c) ? v : T(0);
        Ls[c][j] = v;
    }
    for (int i = 0; i < TRSV_TB; ++i) Ys[i][j] = (i == j) ? T(1) : T(0);
    __syncthreads();
    if (!upper) {
        for (int k = 0; k < TRSV_TB; ++k) {
            const T yk = Ys[k][j];
            for (int i = k + 1; i < TRSV_TB; ++i) Ys[i][j] = tfma(-Ls[k][i], yk, Ys[i][j]);
        }
    } else {
        for (int k = TRSV_TB - 1; k >= 0; --k) {
            const T yk = Ys[k][j] / Ls[k][k];
            Ys[k][j] = yk;
            for (int i = 0; i < k; ++i) Ys[i][j] = tfma(-Ls[k][i], yk, Ys[i][j]);
        }
    }
    __syncthreads();
    // thread j now reads ROW j of the inverse so the global store is coalesced
    T* out = dinv + (long long)blk * TRSV_TB * TRSV_TB;
    for (int c = 0; c < TRSV_TB; ++c) out[c * TRSV_TB + j] = Ys[j][c];
}

struct TrsvSync {
    int* flags;    // [nblk] : == epoch when the block's x segment is final
    int* ticket;   // monotonically increasing
    int* deverr;
};

// x (n x nrhs, ldx) holds the right-hand side on entry and the solution on exit.
// NR right-hand sides per CTA pass (gridDim.y walks the column groups).
template <typename T, int NR, bool UPPER>
__global__ void __launch_bounds__(256) trsv_block_kernel(const T* __restrict__ A, long long lda,
                                                         int n, const T* __restrict__ dinv,
                                                         T* __restrict__ x, long long ldx, int nrhs,
                                                         TrsvSync sy, int epoch, int ticket_base,
                                                         int nblk) {
    constexpr int TB = TRSV_TB;
    __shared__ T s_x[8][NR][16];      // per-warp staging of the x segment it multiplies by
    __shared__ T s_part[4][NR][TB];   // partial sums per column-quarter
    __shared__ T s_rhs[NR][TB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & (TB - 1);
    const int q = tid >> 6;  // column quarter 0..3 (warps 2q, 2q+1)
    const int grp = blockIdx.y;
    const int r0 = grp * NR;
    int* flags = sy.flags + (long long)grp * nblk;

    if (tid == 0) s_ticket = atomicAdd(sy.ticket + grp, 1) - ticket_base;
    __syncthreads();
    const int t = s_ticket;
    const int r = UPPER ? (nblk - 1 - t) : t;
    const int grow = r * TB + row;
    const bool rok = grow < n;

    // diagonal-inverse slice for the final mat-vec: row `row`, columns q*16..q*16+15
    T dv[16];
    {
        const T* dp = dinv + (long long)r * TB * TB + (q * 16) * TB + row;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) dv[jj] = dp[jj * TB];
    }

    T acc[NR];
#pragma unroll
    for (int v = 0; v < NR; ++v) acc[v] = T(0);

    const int ndep = UPPER ? (nblk - 1 - r) : r;
    T an[16];
    auto load_blk = [&](int d, T* dst) {
        const int c = UPPER ? (nblk - 1 - d) : d;
        const T* ap = A + (long long)(c * TB + q * 16) * lda + grow;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int gc = c * TB + q * 16 + jj;
            dst[jj] = (rok && gc < n) ? ap[(long long)jj * lda] : T(0);
        }
    };
    if (ndep > 0) load_blk(0, an);
    bool dead = false;
    for (int d = 0; d < ndep; ++d) {
        T ac[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) ac[jj] = an[jj];
        if (d + 1 < ndep) load_blk(d + 1, an);
        const int c = UPPER ? (nblk - 1 - d) : d;
        if (lane == 0) {
            long long t0 = clock64();
            while (ld_acquire(flags + c) != epoch) {
                if (clock64() - t0 > kSpinTimeoutCycles) { dead = true; break; }
            }
        }
        __syncwarp();
        if (lane < 16) {
#pragma unroll
            for (int v = 0; v < NR; ++v) {
                const int gc = c * TB + q * 16 + lane;
                s_x[warp][v][lane] =
                    (gc < n && r0 + v < nrhs) ? ld_cg(x + (long long)(r0 + v) * ldx + gc) : T(0);
            }
        }
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
#pragma unroll
            for (int v = 0; v < NR; ++v) acc[v] = tfma(ac[jj], s_x[warp][v][jj], acc[v]);
        __syncwarp();
    }
    if (__any_sync(0xffffffffu, dead) && lane == 0) atomicExch(sy.deverr, DEV_ERR_TRSV_TIMEOUT);

#pragma unroll
    for (int v = 0; v < NR; ++v) s_part[q][v][row] = acc[v];
    __syncthreads();
    if (tid < TB) {
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            T b = (rok && r0 + v < nrhs) ? x[(long long)(r0 + v) * ldx + grow] : T(0);
            b -= (s_part[0][v][row] + s_part[1][v][row]) + (s_part[2][v][row] + s_part[3][v][row]);
            s_rhs[v][row] = b;
        }
    }
    __syncthreads();
    // x_r = dinv_r * rhs : thread (row, q) does 16 columns, then reduce over q
#pragma unroll
    for (int v = 0; v < NR; ++v) {
        T s = T(0);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) s = tfma(dv[jj], s_rhs[v][q * 16 + jj], s);
        s_part[q][v][row] = s;
    }
    __syncthreads();
    if (tid < TB) {
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            if (rok && r0 + v < nrhs)
                x[(long long)(r0 + v) * ldx + grow] =
                    (s_part[0][v][row] + s_part[1][v][row]) + (s_part[2][v][row] + s_part[3][v][row]);
        }
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) st_release(flags + r, epoch);
}

// y = b - A x  (FP64 residual for the refinement loop; A n x n column-major).
// CTA: 256 rows x `cchunk` columns; one atomicAdd per row per CTA.
template <typename TA>
__global__ void __launch_bounds__(256) residual_gemv_kernel(const TA* __restrict__ A, long long lda,
                                                            int n, const double* __restrict__ x,
                                                            double* __restrict__ y, int cchunk) {
    const int row = blockIdx.x * 256 + threadIdx.x;
    const int c0 = blockIdx.y * cchunk;
    const int c1 = min(n, c0 + cchunk);
    if (row >= n) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const TA* ap = A + (long long)c0 * lda + row;
    int c = c0;
    for (; c + 3 < c1; c += 4) {
        s0 = fma((double)ap[0], x[c], s0);
        s1 = fma((double)ap[lda], x[c + 1], s1);
        s2 = fma((double)ap[2 * lda], x[c + 2], s2);
        s3 = fma((double)ap[3 * lda], x[c + 3], s3);
        ap += 4 * lda;
    }
    for (; c < c1; ++c) {
        s0 = fma((double)ap[0], x[c], s0);
        ap += lda;
    }
    atomicAdd(y + row, -((s0 + s1) + (s2 + s3)));
}

}  // namespace b200lu
