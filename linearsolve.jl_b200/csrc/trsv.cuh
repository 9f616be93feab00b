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
#include "panel.cuh"  // LL packet helpers

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

// Coupling blocks W^m_r = dinv_r * A[r, r-m] (lower) or dinv_r * A[r, r+m] (upper), m = 1 +
// blockIdx.y, 64x64 each, column-major like dinv; plane m-1 of `wmat` holds the nblk blocks of
// offset m.  With them the only work a block row has to do AFTER its last dependency arrives is
// one 64x64 mat-vec:
//     x_r = dinv_r (b_r - sum_{c not adjacent} A_rc x_c)  -  W_r x_adjacent .
// (trsv_block_kernel / trsv2_kernel use plane 0; the cluster chain of trsv_cluster.cuh planes 0..3)
template <typename T>
__global__ void __launch_bounds__(256) trsv_coupling_kernel(const T* __restrict__ A, long long lda,
                                                            int n, const T* __restrict__ dinv,
                                                            T* __restrict__ wmat, int upper, int nblk,
                                                            int trans) {
    constexpr int TB = TRSV_TB;
    __shared__ T s_a[TB][TB + 1];   // block [k][col]
    const int r = blockIdx.x;
    const int off = 1 + (int)blockIdx.y;
    const int c = upper ? r + off : r - off;
    T* out = wmat + ((long long)blockIdx.y * nblk + r) * TB * TB;
    if (c < 0 || c >= nblk) {
        for (int i = threadIdx.x; i < TB * TB; i += 256) out[i] = T(0);
        return;
    }
    // trans: the block of the TRANSPOSED triangular factor, T(r, c)(i, j) = A[c*TB + j, r*TB + i]
    for (int i = threadIdx.x; i < TB * TB; i += 256) {
        if (!trans) {
            const int row = i % TB, col = i / TB;
            const int gr = r * TB + row, gc = c * TB + col;
            s_a[row][col] = (gr < n && gc < n) ? A[(long long)gc * lda + gr] : T(0);
        } else {
            const int col = i % TB, row = i / TB;   // col is contiguous in memory
            const int gi = r * TB + row, gj = c * TB + col;
            s_a[row][col] = (gi < n && gj < n) ? A[(long long)gi * lda + gj] : T(0);
        }
    }
    __syncthreads();
    const T* dp = dinv + (long long)r * TB * TB;   // column-major: element (row, k) at k*TB + row
    for (int i = threadIdx.x; i < TB * TB; i += 256) {
        const int row = i % TB, col = i / TB;
        T acc = T(0);
#pragma unroll 8
        for (int k = 0; k < TB; ++k) acc = tfma(trans ? dp[row * TB + k] : dp[k * TB + row], s_a[k][col], acc);
        out[i] = acc;
    }
}

// X[i, c] = B[perm[i], c]: the row interchanges of getrs applied to a block of right-hand sides
template <typename T>
__global__ void perm_gather_kernel(const T* __restrict__ B, long long ldb, const int* __restrict__ perm,
                                   T* __restrict__ X, long long ldx, int n, int nrhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int src = perm[i];
    for (long long c = blockIdx.y; c < nrhs; c += gridDim.y) X[c * ldx + i] = B[c * ldb + src];
}

// x[perm[i]] = z[i]: the row interchanges of a transposed solve (A^T = U^T L^T P)
template <typename T>
__global__ void perm_scatter_kernel(const T* __restrict__ z, const int* __restrict__ perm, T* __restrict__ x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[perm[i]] = z[i];
}

struct TrsvSync {
    unsigned long long* xll;  // LL packets of solved x segments: [grp][npad][NR][words]
    int* ticket;              // [groups], self-resetting
    int* deverr;
};

// One triangular sweep.  Lower (UPPER=false): x = L \ rhs with unit L, where
// rhs = B[perm[i]] when perm != nullptr (row interchanges fused into the load,
// reference `_naive_lu_ldiv!` pivot loop src/factorization.jl:437-443) else X.
// Upper: x = U \ X in place.  The solution lands in X (n x nrhs, ldx).
// NR right-hand sides per CTA; the 1-D grid has nblk * groups CTAs.
// Solved 64-row segments travel between CTAs as 64-bit {data32, flag32} LL
// packets: no fence, no separate flag, one L2 round trip per dependency step.
// Critical path per block row (after the adjacent segment arrives): one 64x64
// mat-vec with the precomputed coupling block, a 4-way shared-memory reduce, publish.
template <typename T, int NR, bool UPPER>
__global__ void __launch_bounds__(256) trsv_block_kernel(const T* __restrict__ A, long long lda,
                                                         int n, const T* __restrict__ dinv,
                                                         const T* __restrict__ wmat,
                                                         const T* __restrict__ B, long long ldb,
                                                         const int* __restrict__ perm,
                                                         T* __restrict__ X, long long ldx, int nrhs,
                                                         TrsvSync sy, unsigned epoch, int nblk) {
    constexpr int TB = TRSV_TB;
    constexpr int WN = sizeof(T) / 4;
    __shared__ __align__(16) T s_x[8][NR][16];   // per-warp staging of the x segment it multiplies by
    __shared__ T s_part[4][NR][TB];              // partial sums per column-quarter
    __shared__ T s_rhs[NR][TB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & (TB - 1);
    const int q = tid >> 6;  // column quarter 0..3 (warps 2q, 2q+1)
    // One ticket counter for the whole launch; tickets are dealt round-robin over the
    // right-hand-side groups so that all groups advance through the block rows together:
    // the CTAs of one block row (one per group) are co-resident and share the A blocks in L2,
    // and every group's dependency chain is pipelined instead of the groups running one
    // after another.
    const int groups = gridDim.x / nblk;
    if (tid == 0) {
        const int t = atomicAdd(sy.ticket, 1);
        if (t == (int)gridDim.x - 1) sy.ticket[0] = 0;   // everyone has drawn: re-arm for the next sweep
        s_ticket = t;
    }
    __syncthreads();
    const int grp = s_ticket % groups;
    const int t = s_ticket / groups;
    const int r0 = grp * NR;
    unsigned long long* xll = sy.xll + (size_t)grp * nblk * TB * NR * WN;
    const int r = UPPER ? (nblk - 1 - t) : t;
    const int grow = r * TB + row;
    const bool rok = grow < n;

    // right-hand side of my rows (prefetched; only tid < TB uses it)
    T myb[NR];
    if (tid < TB) {
        const int srow = (!UPPER && perm != nullptr && rok) ? perm[grow] : grow;
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            myb[v] = T(0);
            if (rok && r0 + v < nrhs)
                myb[v] = (!UPPER && perm != nullptr) ? B[(long long)(r0 + v) * ldb + srow]
                                                    : X[(long long)(r0 + v) * ldx + grow];
        }
    }
    // slices (row `row`, columns q*16..q*16+15) of the diagonal inverse and the coupling block
    T dv[16], wv[16];
    {
        const T* dp = dinv + (long long)r * TB * TB + (q * 16) * TB + row;
        const T* wp = wmat + (long long)r * TB * TB + (q * 16) * TB + row;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) { dv[jj] = dp[jj * TB]; wv[jj] = wp[jj * TB]; }
    }

    // gather 16 x-values (x NR x WN words) of block c, quarter q, into this warp's staging
    bool dead = false;
    auto gather_x = [&](int c) {
        const unsigned long long* src = xll + (size_t)(c * TB + q * 16) * NR * WN;
        unsigned* dstw = reinterpret_cast<unsigned*>(&s_x[warp][0][0]);
        for (int idx = lane; idx < 16 * NR * WN; idx += 32) {
            // packet order in xll: [row jj][v][word]; staging order: [v][jj][word]
            const int jj = idx / (NR * WN);
            const int rem = idx - jj * (NR * WN);
            const int v = rem / WN, wd = rem - v * WN;
            unsigned data = 0;
            if (!ll_wait(src + idx, epoch, data)) dead = true;
            dstw[(v * 16 + jj) * WN + wd] = data;
        }
        __syncwarp();
    };

    // ---- phase 1 (off the critical path): all dependencies except the adjacent block ----
    constexpr int NP = NR == 1 ? 4 : (NR <= 4 ? 2 : 1);  // independent FMA chains per right-hand side
    T acc[NR][NP];
#pragma unroll
    for (int v = 0; v < NR; ++v)
#pragma unroll
        for (int u = 0; u < NP; ++u) acc[v][u] = T(0);
    const int ndep = UPPER ? (nblk - 1 - r) : r;   // blocks this row depends on
    const int nfar = ndep > 0 ? ndep - 1 : 0;      // ... all but the adjacent one
    // The far blocks stream through a ring of PF prefetched blocks per thread: with one block in
    // flight the stream ran at one HBM latency per 32 KB block and the LAST block row (n/64 - 2
    // blocks) took as long as the whole dependency chain.
    constexpr int PF = NR == 1 ? ((sizeof(T) == 8) ? 4 : 6) : (NR <= 4 ? 2 : 1);   // register budget: 16 values per block
    T an[PF][16];
    auto load_blk = [&](int d, T* dst) {
        const int c = UPPER ? (nblk - 1 - d) : d;
        const T* ap = A + (long long)(c * TB + q * 16) * lda + grow;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int gc = c * TB + q * 16 + jj;
            dst[jj] = (rok && gc < n) ? ap[(long long)jj * lda] : T(0);
        }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u)
        if (u < nfar) load_blk(u, an[u]);
    for (int d0 = 0; d0 < nfar; d0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int d = d0 + u;
            if (d < nfar) {
                T ac[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) ac[jj] = an[u][jj];
                if (d + PF < nfar) load_blk(d + PF, an[u]);
                gather_x(UPPER ? (nblk - 1 - d) : d);
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
#pragma unroll
                    for (int v = 0; v < NR; ++v) acc[v][jj % NP] = tfma(ac[jj], s_x[warp][v][jj], acc[v][jj % NP]);
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NR; ++v) {
        T sacc = acc[v][0];
#pragma unroll
        for (int u = 1; u < NP; ++u) sacc += acc[v][u];
        s_part[q][v][row] = sacc;
    }
    __syncthreads();
    if (tid < TB) {
#pragma unroll
        for (int v = 0; v < NR; ++v)
            s_rhs[v][row] = myb[v] - ((s_part[0][v][row] + s_part[1][v][row]) +
                                      (s_part[2][v][row] + s_part[3][v][row]));
    }
    __syncthreads();
    // t = dinv_r * rhs  (partial over my 16 columns; summed over q at the very end)
    T tq[NR][NP];
#pragma unroll
    for (int v = 0; v < NR; ++v) {
#pragma unroll
        for (int u = 0; u < NP; ++u) tq[v][u] = T(0);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) tq[v][jj % NP] = tfma(dv[jj], s_rhs[v][q * 16 + jj], tq[v][jj % NP]);
    }
    // ---- phase 2 (critical path): x_r = t - W_r * x_adjacent ----
    if (ndep > 0) {
        gather_x(UPPER ? r + 1 : r - 1);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
#pragma unroll
            for (int v = 0; v < NR; ++v) tq[v][jj % NP] = tfma(-wv[jj], s_x[warp][v][jj], tq[v][jj % NP]);
    }
    if (__any_sync(0xffffffffu, dead) && lane == 0) atomicExch(sy.deverr, DEV_ERR_TRSV_TIMEOUT);
#pragma unroll
    for (int v = 0; v < NR; ++v) {
        T st = tq[v][0];
#pragma unroll
        for (int u = 1; u < NP; ++u) st += tq[v][u];
        s_part[q][v][row] = st;
    }
    __syncthreads();
    if (tid < TB) {
        unsigned long long* dst = xll + (size_t)grow * NR * WN;
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            const T xv = (s_part[0][v][row] + s_part[1][v][row]) + (s_part[2][v][row] + s_part[3][v][row]);
            unsigned w[WN];
            Words<T>::split(xv, w);
#pragma unroll
            for (int x = 0; x < WN; ++x) ll_store(dst + v * WN + x, w[x], epoch);
            if (rok && r0 + v < nrhs) X[(long long)(r0 + v) * ldx + grow] = xv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Single right-hand side, version 2: 2-D work items.  One CTA per block row (above) leaves the
// LAST block row streaming n/64 - 2 blocks alone — twice the average — and serialises each row's
// stream behind one thread block.  Here a block row t (logical order: t = 0 is solved first) is
// split into
//   * partial items (t, k): the far dependencies d in [8k, 8k + 8) ∩ [0, t - 2): eight 64x64
//     blocks times already-solved segments, result = 64 partial sums written as LL packets;
//   * one final item (t): rhs - sum of the partial items (fixed order: deterministic), minus the
//     block at d = t - 2, times the inverted diagonal block, minus the coupling block times the
//     adjacent segment x_{t-1}; publishes x_t.
// Items are drawn from one ticket counter in (t, k) order by a persistent grid that is fully
// resident, so an item only ever waits for items with smaller tickets: no deadlock.  The
// dependency chain per block row is the same single 64x64 mat-vec as before; the streaming work
// is spread over every SM.
struct Trsv2Item { int t, k; };   // k < 0: final item

struct Trsv2Sync {
    unsigned long long* xll;     // [nblk*64][WN]   solved x segments (LL packets)
    unsigned long long* pll;     // [nblk][kmax][64][WN] partial sums (LL packets)
    int* ticket;
    int* deverr;
    const Trsv2Item* items;
    int nitems, kmax;
};

constexpr int TRSV2_CH = 8;   // far blocks per partial item

template <typename T, bool UPPER, bool TRANS>
__global__ void __launch_bounds__(256, 1) trsv2_kernel(const T* __restrict__ A, long long lda, int n,
                                                       const T* __restrict__ dinv, const T* __restrict__ wmat,
                                                       const T* __restrict__ B, const int* __restrict__ perm,
                                                       T* __restrict__ X, Trsv2Sync sy, unsigned epoch, int nblk) {
    constexpr int TB = TRSV_TB;
    constexpr int WN = sizeof(T) / 4;
    __shared__ __align__(16) T s_x[8][16];       // per-warp staging of the 16 x values it multiplies by
    __shared__ T s_part[4][TB];
    __shared__ T s_rhs[TB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & (TB - 1);
    const int q = tid >> 6;   // column quarter 0..3
    bool dead = false;

    // gather the 16 x values (quarter q of logical block d) into this warp's staging
    auto gather_x = [&](int d) {
        const int c = UPPER ? (nblk - 1 - d) : d;
        const unsigned long long* src = sy.xll + (size_t)(c * TB + q * 16) * WN;
        unsigned* dstw = reinterpret_cast<unsigned*>(&s_x[warp][0]);
        for (int idx = lane; idx < 16 * WN; idx += 32) {
            unsigned data = 0;
            if (!ll_wait(src + idx, epoch, data)) dead = true;
            dstw[idx] = data;
        }
        __syncwarp();
    };
    // block (r, c) of the triangular factor being solved with; TRANS: of its transpose,
    // element (i, j) = A[c*TB + j, r*TB + i] — 16 contiguous values per thread
    auto load_blk = [&](int d, int grow, bool rok, T* dst) {
        const int c = UPPER ? (nblk - 1 - d) : d;
        if constexpr (!TRANS) {
            const T* ap = A + (long long)(c * TB + q * 16) * lda + grow;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int gc = c * TB + q * 16 + jj;
                dst[jj] = (rok && gc < n) ? ap[(long long)jj * lda] : T(0);
            }
        } else {
            const T* ap = A + (long long)grow * lda + (c * TB + q * 16);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int gc = c * TB + q * 16 + jj;
                dst[jj] = (rok && gc < n) ? ap[jj] : T(0);
            }
        }
    };

    for (;;) {
        __syncthreads();   // s_ticket / staging reuse
        if (tid == 0) {
            const int tk = atomicAdd(sy.ticket, 1);
            if (tk == sy.nitems + (int)gridDim.x - 1) sy.ticket[0] = 0;   // last draw of the sweep: re-arm
            s_ticket = tk;
        }
        __syncthreads();
        const int tk = s_ticket;
        if (tk >= sy.nitems) break;
        const Trsv2Item it = sy.items[tk];
        const int t = it.t;
        const int r = UPPER ? (nblk - 1 - t) : t;
        const int grow = r * TB + row;
        const bool rok = grow < n;
        const int nfar = t > 2 ? t - 2 : 0;
        if (it.k >= 0) {
            // ---------------- partial item ----------------
            const int d0 = it.k * TRSV2_CH, d1 = min(d0 + TRSV2_CH, nfar);
            constexpr int PF = 4;
            T an[PF][16];
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (d0 + u < d1) load_blk(d0 + u, grow, rok, an[u]);
            T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
            for (int db = d0; db < d1; db += PF) {
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int d = db + u;
                    if (d < d1) {
                        T ac[16];
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) ac[jj] = an[u][jj];
                        if (d + PF < d1) load_blk(d + PF, grow, rok, an[u]);
                        gather_x(d);
#pragma unroll
                        for (int jj = 0; jj < 16; jj += 4) {
                            acc0 = tfma(ac[jj], s_x[warp][jj], acc0);
                            acc1 = tfma(ac[jj + 1], s_x[warp][jj + 1], acc1);
                            acc2 = tfma(ac[jj + 2], s_x[warp][jj + 2], acc2);
                            acc3 = tfma(ac[jj + 3], s_x[warp][jj + 3], acc3);
                        }
                        __syncwarp();
                    }
                }
            }
            s_part[q][row] = (acc0 + acc1) + (acc2 + acc3);
            __syncthreads();
            if (tid < TB) {
                const T sum = (s_part[0][row] + s_part[1][row]) + (s_part[2][row] + s_part[3][row]);
                unsigned w[WN];
                Words<T>::split(sum, w);
                unsigned long long* dst = sy.pll + ((size_t)(t * sy.kmax + it.k) * TB + row) * WN;
#pragma unroll
                for (int x = 0; x < WN; ++x) ll_store(dst + x, w[x], epoch);
            }
        } else {
            // ---------------- final item ----------------
            T myb = T(0);
            if (tid < TB && rok) {
                if (!UPPER && perm != nullptr) myb = B[perm[grow]];
                else if (!UPPER && TRANS) myb = B[grow];
                else myb = X[grow];
            }
            T dv[16], wv[16], av[16];
            {
                // TRANS: the inverse of the transposed diagonal block is the transpose of the inverse
                const T* dp = TRANS ? dinv + (long long)r * TB * TB + row * TB + q * 16
                                    : dinv + (long long)r * TB * TB + (q * 16) * TB + row;
                const T* wp = wmat + (long long)r * TB * TB + (q * 16) * TB + row;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) { dv[jj] = TRANS ? dp[jj] : dp[jj * TB]; wv[jj] = wp[jj * TB]; }
            }
            if (t >= 2) load_blk(t - 2, grow, rok, av);
            // sum of the partial items, fixed order
            const int nch = (nfar + TRSV2_CH - 1) / TRSV2_CH;
            if (tid < TB) {
                T sum = T(0);
                const unsigned long long* src = sy.pll + ((size_t)(t * sy.kmax) * TB + row) * WN;
                for (int k = 0; k < nch; ++k) {
                    unsigned w[WN];
#pragma unroll
                    for (int x = 0; x < WN; ++x) {
                        unsigned data = 0;
                        if (!ll_wait(src + (size_t)k * TB * WN + x, epoch, data)) dead = true;
                        w[x] = data;
                    }
                    sum += Words<T>::join(w);
                }
                s_rhs[row] = myb - sum;
            }
            // the block at d = t - 2
            if (t >= 2) {
                gather_x(t - 2);
                T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
#pragma unroll
                for (int jj = 0; jj < 16; jj += 4) {
                    a0 = tfma(av[jj], s_x[warp][jj], a0);
                    a1 = tfma(av[jj + 1], s_x[warp][jj + 1], a1);
                    a2 = tfma(av[jj + 2], s_x[warp][jj + 2], a2);
                    a3 = tfma(av[jj + 3], s_x[warp][jj + 3], a3);
                }
                __syncwarp();
                s_part[q][row] = (a0 + a1) + (a2 + a3);
            }
            __syncthreads();
            if (t >= 2 && tid < TB)
                s_rhs[row] -= (s_part[0][row] + s_part[1][row]) + (s_part[2][row] + s_part[3][row]);
            __syncthreads();
            // tq = dinv_r * rhs (my 16 columns)
            T t0 = T(0), t1 = T(0), t2 = T(0), t3 = T(0);
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
                t0 = tfma(dv[jj], s_rhs[q * 16 + jj], t0);
                t1 = tfma(dv[jj + 1], s_rhs[q * 16 + jj + 1], t1);
                t2 = tfma(dv[jj + 2], s_rhs[q * 16 + jj + 2], t2);
                t3 = tfma(dv[jj + 3], s_rhs[q * 16 + jj + 3], t3);
            }
            // critical path: x_t = tq - W_t * x_{t-1}
            if (t >= 1) {
                gather_x(t - 1);
#pragma unroll
                for (int jj = 0; jj < 16; jj += 4) {
                    t0 = tfma(-wv[jj], s_x[warp][jj], t0);
                    t1 = tfma(-wv[jj + 1], s_x[warp][jj + 1], t1);
                    t2 = tfma(-wv[jj + 2], s_x[warp][jj + 2], t2);
                    t3 = tfma(-wv[jj + 3], s_x[warp][jj + 3], t3);
                }
                __syncwarp();
            }
            __syncthreads();   // all reads of s_part (d = t-2 sums) done before it is reused
            s_part[q][row] = (t0 + t1) + (t2 + t3);
            __syncthreads();
            if (tid < TB) {
                const T xv = (s_part[0][row] + s_part[1][row]) + (s_part[2][row] + s_part[3][row]);
                unsigned w[WN];
                Words<T>::split(xv, w);
                unsigned long long* dst = sy.xll + (size_t)grow * WN;
#pragma unroll
                for (int x = 0; x < WN; ++x) ll_store(dst + x, w[x], epoch);
                if (rok) X[grow] = xv;
            }
        }
        if (__any_sync(0xffffffffu, dead) && lane == 0) atomicExch(sy.deverr, DEV_ERR_TRSV_TIMEOUT);
        if (__syncthreads_or(dead ? 1 : 0)) break;
    }
}

// y = b - A x  (FP64 residual for the refinement loop; A n x n column-major).
// CTA (bx, by): 256 rows x `cchunk` columns, partial sums to part[by * n + row]; residual_finish_kernel then
// subtracts the chunks' partial sums from y in ascending chunk order — one writer per entry, no atomics: the
// residual, and with it the refined solution, is the same bits on every run.
template <typename TA>
__global__ void __launch_bounds__(256) residual_gemv_kernel(const TA* __restrict__ A, long long lda,
                                                            int n, const double* __restrict__ x,
                                                            double* __restrict__ part, int cchunk) {
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
    part[(long long)blockIdx.y * n + row] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256) residual_finish_kernel(const double* __restrict__ part, int n, int nchunks,
                                                              double* __restrict__ y) {
    const int row = blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    double s = 0.0;
    for (int k = 0; k < nchunks; ++k) s += part[(long long)k * n + row];
    y[row] -= s;
}

// y -= A^T x (the residual of the transposed system, MIXED refinement with trans = 'T'): one warp per
// COLUMN of the column-major A — a contiguous dot product, 256 bytes per warp request, four
// independent accumulators, shuffle tree, ONE writer per y[c] (no atomics: deterministic).
template <typename TA>
__global__ void __launch_bounds__(256) residual_gemvT_kernel(const TA* __restrict__ A, long long lda,
                                                             int n, const double* __restrict__ x,
                                                             double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n) return;   // whole warps leave together
    const TA* ap = A + (long long)c * lda;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = lane;
    for (; r + 96 < n; r += 128) {
        s0 = fma((double)ap[r], x[r], s0);
        s1 = fma((double)ap[r + 32], x[r + 32], s1);
        s2 = fma((double)ap[r + 64], x[r + 64], s2);
        s3 = fma((double)ap[r + 96], x[r + 96], s3);
    }
    for (; r < n; r += 32) s0 = fma((double)ap[r], x[r], s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[c] -= s;
}

}  // namespace b200lu
