// trsv_cluster.cuh — getrs, one right-hand side, version 3: the dependency chain of the
// substitution runs inside ONE thread-block cluster and hands the solved segments from CTA to
// CTA through distributed shared memory; every other CTA streams the far blocks from HBM.
// (reference `_naive_lu_ldiv!`, src/factorization.jl:433-491; LAPACK getrs, src/openblas.jl:247-278)
//
// Why: version 2 (trsv.cuh) publishes x_t as LL packets in global memory, so a chain step costs an
// L2 store + an L2 poll (~0.9 us measured per 64-row block row: n/64 steps bound the sweep).  A
// DSMEM store lands directly in the consumer's shared memory and the consumer polls its OWN
// shared memory.
//
// Roles (grid = clusters of CS CTAs, all co-resident):
//   * cluster 0 — the chain.  CTA `rank` owns block rows t = rank, rank + CS, ...  Per row:
//       x_t = dinv_t (b_t - far_t) - sum_{m=1..NEAR} W^m_t x_{t-m},   W^m_t = dinv_t A[t, t-m]
//     (coupling blocks precomputed once per factorization).  dinv_t, W^1_t wait in registers and
//     W^2.._t in shared memory (cp.async) before they are needed; x_{t-NEAR..t-1} arrive in the
//     CTA's own ring of the last 16 segments as {data32, tag32} packets; far_t = the fixed-order
//     sum of the partial items below.  After x_{t-1} lands the step is 16 FMAs per thread, one
//     shared-memory reduction and the remote stores of x_t into the rings of the next NEAR rows' CTAs.
//   * every other cluster — workers.  They draw partial items (t, k) from a ticket counter in
//     t-major order: 8 far blocks A[t, d], d in [8k, 8k + 8) ∩ [0, t - NEAR), times x_d read as LL
//     packets from global memory (the chain publishes x there as well; all eight packets of an item
//     are requested at once, the next ticket is drawn while the item runs), result = 64 partial
//     sums written as LL packets.
// Every CTA is resident, items wait only on earlier rows: no deadlock; every wait has a watchdog.
//
// Measured on B200 (in-kernel clock64 stamps, B200LU_TRSV_DBG=1; n = 8192, FP64, NEAR 4, CS 8):
// one chain step = ~1270 cycles with the far sums taken out of the loop (570 from the arrival of
// x_{t-1} to the remote stores of x_t: mat-vec, CTA barrier, reduction; ~700 for the DSMEM
// transit and its detection) and ~1500 cycles with them (the far sums of row t need x_{t-NEAR-1},
// two L2 round trips and a worker's pass: a second feedback loop).  Version 2 needs ~1770.
// getrs: 232 us vs 240 us (n = 8192), 620 vs 685 us (n = 16384); below n = 6144 version 2 is as
// fast or faster and stays the default.  A chain CTA also has to pull (NEAR + 1) x 32 KiB of
// operands per row (~3200 cycles to issue), which is why NEAR = 6 and 16-CTA clusters (7 instead
// of 15 co-resident clusters: fewer workers) measured no better; taking the near terms before the
// far sums measured worse (232 -> 260 us), as did a single polling lane per warp (232 -> 268 us).
#pragma once
#include "common.cuh"
#include "panel.cuh"           // LL packets
#include "panel_cluster.cuh"   // cluster / DSMEM helpers
#include "trsv.cuh"

namespace b200lu {

constexpr int TRSV3_NEAR_MAX = 4; // coupling planes kept per factorization (block columns t-1 .. t-4)
constexpr int TRSV3_RING = 16;   // solved segments kept in every chain CTA's shared memory
constexpr int TRSV3_CH = 8;      // far blocks per partial item

struct Trsv3Sync {
    unsigned long long* xll;     // [nblk*64][WN]       solved x (LL packets) for the workers
    unsigned long long* pll;     // [nblk][kmax][64][WN] partial sums (LL packets)
    int* ticket;
    int* deverr;
    const Trsv2Item* items;      // partial items only, t-major
    int nitems, kmax;
    long long* dbg;              // B200LU_TRSV_DBG: per chain CTA, cycles spent in each wait
};

template <typename T, int NEAR>
constexpr int trsv3_smem_bytes() {
    return (NEAR - 1) * TRSV_TB * TRSV_TB * (int)sizeof(T) + TRSV3_RING * TRSV_TB * (int)(sizeof(T) / 4) * 8;
}

__device__ __forceinline__ void t3_st_cluster_u64(unsigned raddr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(raddr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long t3_ld_shared_u64(unsigned addr) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

// wmat: planes of nblk coupling blocks, plane m-1 = dinv_r * A[r, r -+ m]; NEAR of them are used.
// CS = cluster size (8 portable, 16 opt-in): the chain CTA of a block row has CS chain steps to
// fetch its operands.
template <typename T, bool UPPER, int NEAR, int CS>
__global__ void __launch_bounds__(256, 1) trsv3_kernel(const T* __restrict__ A, long long lda, int n,
                                                       const T* __restrict__ dinv, const T* __restrict__ wmat,
                                                       const T* __restrict__ B, const int* __restrict__ perm,
                                                       T* __restrict__ X, Trsv3Sync sy, unsigned epoch, int nblk) {
    constexpr int TB = TRSV_TB;
    constexpr int WN = sizeof(T) / 4;
    extern __shared__ __align__(16) unsigned char t3_smem[];
    T* wbuf = reinterpret_cast<T*>(t3_smem);                                        // [NEAR-1][TB*TB]
    unsigned long long* ring = reinterpret_cast<unsigned long long*>(t3_smem + (NEAR - 1) * TB * TB * sizeof(T));   // [RING][TB*WN]
    __shared__ __align__(16) T s_x[8][16];       // per-warp staging of the 16 x values it multiplies by
    __shared__ T s_part[4][TB];
    __shared__ T s_far[4][TB];
    __shared__ T s_rhs[TB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & (TB - 1);
    const int q = tid >> 6;   // column quarter 0..3
    const unsigned rank = pcl_cluster_rank();
    const bool chain = blockIdx.x < (unsigned)CS;
    bool dead = false;

    if (chain)
        for (int i = tid; i < TRSV3_RING * TB * WN; i += 256) ring[i] = 0ull;
    // every CTA of the cluster is running and the rings are clear before any remote store
    pcl_cluster_sync();

    if (!chain) {
        // ------------------------------ workers: far partial items ------------------------------
        const int nworkers = (int)gridDim.x - CS;
        auto load_blk = [&](int d, int grow, bool rok, T* dst) {
            const int c = UPPER ? (nblk - 1 - d) : d;
            const T* ap = A + (long long)(c * TB + q * 16) * lda + grow;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int gc = c * TB + q * 16 + jj;
                dst[jj] = (rok && gc < n) ? ap[(long long)jj * lda] : T(0);
            }
        };
        // packet `lane` of the 16 x values (quarter q of logical block d) this warp multiplies by
        auto x_src = [&](int d) {
            const int c = UPPER ? (nblk - 1 - d) : d;
            return sy.xll + (size_t)(c * TB + q * 16) * WN + lane;
        };
        auto draw = [&]() {
            const int tk = atomicAdd(sy.ticket, 1);
            if (tk == sy.nitems + nworkers - 1) sy.ticket[0] = 0;   // last draw of the sweep: re-arm
            return tk;
        };
        if (tid == 0) s_ticket = draw();
        __syncthreads();
        int tk = s_ticket;
        while (tk < sy.nitems) {
            const Trsv2Item it = sy.items[tk];
            __syncthreads();                       // everyone has read s_ticket
            if (tid == 0) s_ticket = draw();       // the next ticket's round trip hides behind this item
            const int t = it.t;
            const int r = UPPER ? (nblk - 1 - t) : t;
            const int grow = r * TB + row;
            const bool rok = grow < n;
            const int nfar = t - NEAR;
            const int d0 = it.k * TRSV3_CH, d1 = min(d0 + TRSV3_CH, nfar);
            constexpr int PF = 4;
            T an[PF][16];
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (d0 + u < d1) load_blk(d0 + u, grow, rok, an[u]);
            // all x packets of the item in flight at once: one L2 round trip instead of eight
            const bool xl = lane < 16 * WN;
            unsigned long long xv[TRSV3_CH];
#pragma unroll
            for (int u = 0; u < TRSV3_CH; ++u) xv[u] = (xl && d0 + u < d1) ? ll_load(x_src(d0 + u)) : 0ull;
            T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
            unsigned* dstw = reinterpret_cast<unsigned*>(&s_x[warp][0]);
#pragma unroll
            for (int u = 0; u < TRSV3_CH; ++u) {
                const int d = d0 + u;
                if (d < d1) {
                    if (xl) {
                        unsigned data = (unsigned)xv[u];
                        if ((unsigned)(xv[u] >> 32) != epoch)
                            if (!ll_wait(x_src(d), epoch, data)) dead = true;
                        dstw[lane] = data;
                    }
                    __syncwarp();
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4) {
                        acc0 = tfma(an[u % PF][jj], s_x[warp][jj], acc0);
                        acc1 = tfma(an[u % PF][jj + 1], s_x[warp][jj + 1], acc1);
                        acc2 = tfma(an[u % PF][jj + 2], s_x[warp][jj + 2], acc2);
                        acc3 = tfma(an[u % PF][jj + 3], s_x[warp][jj + 3], acc3);
                    }
                    __syncwarp();
                    if (d + PF < d1) load_blk(d + PF, grow, rok, an[u % PF]);   // refill the slot just consumed
                }
            }
            s_part[q][row] = (acc0 + acc1) + (acc2 + acc3);
            const int any_dead = __syncthreads_or(dead ? 1 : 0);
            if (tid < TB) {
                const T sum = (s_part[0][row] + s_part[1][row]) + (s_part[2][row] + s_part[3][row]);
                unsigned w[WN];
                Words<T>::split(sum, w);
                unsigned long long* dst = sy.pll + ((size_t)(t * sy.kmax + it.k) * TB + row) * WN;
#pragma unroll
                for (int x = 0; x < WN; ++x) ll_store(dst + x, w[x], epoch);
            }
            if (any_dead) {
                if (tid == 0) atomicExch(sy.deverr, DEV_ERR_TRSV_TIMEOUT);
                break;
            }
            tk = s_ticket;   // written before the barrier above
        }
    } else {
        // ------------------------------------ the chain ------------------------------------
        const long long bstride = (long long)TB * TB;
        const long long plane = (long long)nblk * bstride;
        const unsigned ring_u32 = pcl_smem_u32(ring);
        T dv[16], w1[16];
        T myb = T(0);
        int pnext = 0;   // perm index of the row after next: B[perm[.]] never waits on a dependent load
        // operands of block row t: dinv and W^1 into registers, W^2..4 into shared memory, b_t
        auto issue_loads = [&](int t) {
            const int r = UPPER ? (nblk - 1 - t) : t;
            const T* dp = dinv + (long long)r * bstride + (q * 16) * TB + row;
            const T* wp = wmat + (long long)r * bstride + (q * 16) * TB + row;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) { dv[jj] = dp[jj * TB]; w1[jj] = wp[jj * TB]; }
#pragma unroll
            for (int m = 0; m < NEAR - 1; ++m) {
                const char* src = reinterpret_cast<const char*>(wmat + (long long)(m + 1) * plane + (long long)r * bstride);
                char* dst = reinterpret_cast<char*>(wbuf + m * TB * TB);
                for (int i = tid; i < TB * TB * (int)sizeof(T) / 16; i += 256) cp_async16(dst + i * 16, src + i * 16, true);
            }
            cp_async_commit();
            const int grow = r * TB + row;
            myb = T(0);
            if (tid < TB && grow < n) {
                if (!UPPER) myb = B[pnext];
                else myb = X[grow];
            }
            if (!UPPER && tid < TB && t + CS < nblk) {
                const int g2 = (t + CS) * TB + row;
                pnext = g2 < n ? perm[g2] : 0;
            }
        };
        // the 16 values of x_d this thread's quarter multiplies by, from the local ring
        auto gather_local = [&](int d) {
            const unsigned tag = (unsigned)d + 1u;
            const unsigned src = ring_u32 + (unsigned)(((d & (TRSV3_RING - 1)) * TB + q * 16) * WN) * 8u;
            unsigned* dstw = reinterpret_cast<unsigned*>(&s_x[warp][0]);
            // every lane polls its own packet (measured: a single polling lane per warp followed
            // by a check of the others is ~15 % slower end to end)
            for (int idx = lane; idx < 16 * WN; idx += 32) {
                unsigned long long v = t3_ld_shared_u64(src + idx * 8);
                if ((unsigned)(v >> 32) != tag) {
                    const long long c0 = clock64();
                    do {
                        v = t3_ld_shared_u64(src + idx * 8);
                        if (clock64() - c0 > kSpinTimeoutCycles) { dead = true; break; }
                    } while ((unsigned)(v >> 32) != tag);
                }
                dstw[idx] = (unsigned)v;
            }
            __syncwarp();
        };

        long long dg0 = 0, dg1 = 0, dg2 = 0, dg3 = 0, dg4 = 0, dg5 = 0, c_prev = 0, g_start = 0, k_start = 0;
        if (sy.dbg && tid == 0) {
            k_start = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
        }
        if (!UPPER && tid < TB && (int)rank * TB + row < n) pnext = perm[(int)rank * TB + row];
        if ((int)rank < nblk) issue_loads((int)rank);
        for (int t = (int)rank; t < nblk; t += CS) {
            const int r = UPPER ? (nblk - 1 - t) : t;
            const int grow = r * TB + row;
            const bool rok = grow < n;
            const int nfar = t > NEAR ? t - NEAR : 0;
            const int nch = (nfar + TRSV3_CH - 1) / TRSV3_CH;
            long long c_a = 0, c_b = 0, c_c = 0, c_d = 0;
            if (sy.dbg && tid == 0) {
                c_a = clock64();
                if (c_prev) dg5 += c_a - c_prev;   // end of the previous row -> here: operand loads issued
            }
            {
                // far partial sums: quarter q takes chunks q, q+4, ... — up to four packets (and
                // both of their words) in flight per thread; fixed assignment and order: deterministic
                T sum = T(0);
                const unsigned long long* src = sy.pll + ((size_t)(t * sy.kmax) * TB + row) * WN;
                for (int k0 = q; k0 < nch; k0 += 16) {
                    unsigned long long v[4][WN];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + 4 * u;
#pragma unroll
                        for (int x = 0; x < WN; ++x) v[u][x] = k < nch ? ll_load(src + (size_t)k * TB * WN + x) : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + 4 * u;
                        if (k < nch) {
                            unsigned w[WN];
#pragma unroll
                            for (int x = 0; x < WN; ++x) {
                                unsigned data = (unsigned)v[u][x];
                                if ((unsigned)(v[u][x] >> 32) != epoch)
                                    if (!ll_wait(src + (size_t)k * TB * WN + x, epoch, data)) dead = true;
                                w[x] = data;
                            }
                            sum += Words<T>::join(w);
                        }
                    }
                }
                s_far[q][row] = sum;
            }
            cp_async_wait<0>();
            __syncthreads();
            if (tid < TB) s_rhs[row] = myb - ((s_far[0][row] + s_far[1][row]) + (s_far[2][row] + s_far[3][row]));
            __syncthreads();
            if (sy.dbg && tid == 0) c_b = clock64();
            T t0 = T(0), t1 = T(0), t2 = T(0), t3 = T(0);
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
                t0 = tfma(dv[jj], s_rhs[q * 16 + jj], t0);
                t1 = tfma(dv[jj + 1], s_rhs[q * 16 + jj + 1], t1);
                t2 = tfma(dv[jj + 2], s_rhs[q * 16 + jj + 2], t2);
                t3 = tfma(dv[jj + 3], s_rhs[q * 16 + jj + 3], t3);
            }
            // near block columns, oldest first; the last one (x_{t-1}) is the critical path
#pragma unroll
            for (int m = NEAR; m >= 2; --m) {
                if (t >= m) {
                    gather_local(t - m);
                    const T* wb = wbuf + (m - 2) * TB * TB + (q * 16) * TB + row;
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4) {
                        t0 = tfma(-wb[jj * TB], s_x[warp][jj], t0);
                        t1 = tfma(-wb[(jj + 1) * TB], s_x[warp][jj + 1], t1);
                        t2 = tfma(-wb[(jj + 2) * TB], s_x[warp][jj + 2], t2);
                        t3 = tfma(-wb[(jj + 3) * TB], s_x[warp][jj + 3], t3);
                    }
                    __syncwarp();
                }
            }
            if (sy.dbg && tid == 0) c_c = clock64();
            if (t >= 1) {
                gather_local(t - 1);
                if (sy.dbg && tid == 0) c_d = clock64();
#pragma unroll
                for (int jj = 0; jj < 16; jj += 4) {
                    t0 = tfma(-w1[jj], s_x[warp][jj], t0);
                    t1 = tfma(-w1[jj + 1], s_x[warp][jj + 1], t1);
                    t2 = tfma(-w1[jj + 2], s_x[warp][jj + 2], t2);
                    t3 = tfma(-w1[jj + 3], s_x[warp][jj + 3], t3);
                }
                __syncwarp();
            }
            s_part[q][row] = (t0 + t1) + (t2 + t3);
            __syncthreads();
            {
                // every quarter forms the same x_t and pushes it into the rings of the CTAs that own
                // the next NEAR block rows, the next one first
                const T xv = (s_part[0][row] + s_part[1][row]) + (s_part[2][row] + s_part[3][row]);
                unsigned w[WN];
                Words<T>::split(xv, w);
                const unsigned long long tag = (unsigned long long)((unsigned)t + 1u) << 32;
                const unsigned slot = ring_u32 + (unsigned)(((t & (TRSV3_RING - 1)) * TB + row) * WN) * 8u;
#pragma unroll
                for (int u = 0; u < (NEAR + 3) / 4; ++u) {
                    const int m = 1 + q + 4 * u;
                    if (m <= NEAR && t + m < nblk) {
                        const unsigned ra = pcl_mapa(slot, (rank + (unsigned)m) % CS);
#pragma unroll
                        for (int x = 0; x < WN; ++x) t3_st_cluster_u64(ra + 8u * x, tag | w[x]);
                    }
                }
                if (q == 3) {
                    unsigned long long* dst = sy.xll + (size_t)grow * WN;
#pragma unroll
                    for (int x = 0; x < WN; ++x) ll_store(dst + x, w[x], epoch);
                    if (rok) X[grow] = xv;
                }
            }
            if (sy.dbg && tid == 0) {
                const long long c_e = clock64();
                dg0 += c_b - c_a;                       // far partial sums + operands
                dg1 += c_c - c_b;                       // dinv mat-vec + x_{t-NEAR} .. x_{t-2}
                dg2 += t >= 1 ? c_d - c_c : 0;          // wait for x_{t-1}
                dg3 += c_e - (t >= 1 ? c_d : c_c);      // step: mat-vec, reduce, push
                dg4 += 1;
                c_prev = c_e;
            }
            if (__syncthreads_or(dead ? 1 : 0)) {
                if (tid == 0) atomicExch(sy.deverr, DEV_ERR_TRSV_TIMEOUT);
                break;
            }
            if (t + CS < nblk) issue_loads(t + CS);
        }
        cp_async_wait<0>();
        if (sy.dbg && tid == 0) {
            long long g_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
            unsigned long long* o = reinterpret_cast<unsigned long long*>(sy.dbg) + rank * 8;
            atomicAdd(o + 0, (unsigned long long)dg0); atomicAdd(o + 1, (unsigned long long)dg1);
            atomicAdd(o + 2, (unsigned long long)dg2); atomicAdd(o + 3, (unsigned long long)dg3);
            atomicAdd(o + 4, (unsigned long long)dg4); atomicAdd(o + 5, (unsigned long long)dg5);
            atomicAdd(o + 6, (unsigned long long)(clock64() - k_start));
            atomicAdd(o + 7, (unsigned long long)(g_end - g_start));
        }
    }
    // no CTA leaves while a peer can still store into its ring
    pcl_cluster_sync();
}

}  // namespace b200lu
