// panel_cluster.cuh — partial-pivoting base panel (getf2) on ONE thread-block cluster.
//
// Same arithmetic contract as panel.cuh (reference src/blocked_lufact.jl:38-54 pivot rule,
// :93-122 column loop; src/generic_lufact.jl:117-123 zero-pivot rule): amax starts at 0,
// strict `>` (NaN never wins), lowest row on ties; zero pivot => info = k once, no scaling,
// the rank-1 update still runs; scaling multiplies by inv(pivot); updates are FMAs.
//
// What changed is the exchange.  Measured on B200 (scripts/microbench/exchange_latency*.cu):
// the L2-mailbox protocol of panel.cuh costs 3000-4700 cycles per column at 32 CTAs; a
// cluster-wide all-to-all with `st.async` + mbarrier transaction counts costs 870 (8 CTAs) to
// 1100 cycles (16 CTAs) including the CTA barrier, and never touches L2, so the concurrent
// trailing-update GEMM cannot slow it down.  Per column:
//   1. every warp: redux arg-max over its rows (the reciprocal of each lane's own candidate is
//      computed under the redux latency); the winning lane stages its row in shared memory;
//   2. one __syncthreads; the sender threads reduce the warp candidates and push the CTA
//      candidate {1/pivot, position, row} into EVERY CTA's inbox with st.async (complete_tx on
//      the receiver's mbarrier) — only the vectors that still hold unfactored columns travel;
//   3. every thread waits on its own CTA's mbarrier (hardware sleep, no polling traffic), every
//      warp reduces the <= 16 candidates redundantly and reads the winning row from its own
//      shared memory;
//   4. scale + rank-1 update in registers.
// Rows never move between threads: each thread tracks the current POSITION of its rows (the
// LAPACK interchange sequence acts on positions), candidates are compared by position, and the
// rows are scattered to their final positions when the block is written back.  That removes the
// "top row" message of panel.cuh.  The interchanges of the other columns of the outer panel are
// applied at the end by the same cluster as one gather/scatter of the <= 2W rows that moved.
//
// Instruction economy.  In-kernel clock64 stamps and ncu stall sampling of the first versions
// showed the exchange wait at only ~400 of ~4300 cycles per column: the rest was instruction
// issue (register rotation, per-element predicates, PHI moves around a switch) and instruction
// cache misses of an unrolled column loop, in warps that all sit in the same phase.  So:
//  * LEFT-SHIFTING LIVE WINDOW: a finished column leaves the registers at once (its multipliers
//    go to a shared-memory tile), and the rank-1 update writes element c to index c - 1, so the
//    current column is always register 0 with static indices in a ROLLED loop: no rotation, no
//    per-element predicate (dead elements compute garbage that is never read);
//  * four copies of the loop body with 32/24/16/8 live elements bound the wasted FMAs;
//  * a pivot row is frozen when chosen: its warp copies the staged row into the tile;
//  * arg-max = ONE redux on the top 32 bits + a ballot (ties on the top word take the exact slow
//    path); the next column's arg-max is in flight under the rest of the rank-1 update;
//  * 1/pivot is computed by every candidate lane while the redux is in flight.
#pragma once
#include "common.cuh"
#include "panel.cuh"

namespace b200lu {

#ifdef PCL_TIMING
#define PCL_T(i) do { if (dbg) { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } } while (0)
#else
#define PCL_T(i) do {} while (0)
#endif

constexpr int PCL_GMAX = 16;   // non-portable cluster size limit on sm_100
constexpr int PCL_NT = 256;    // threads per CTA (up to 255 registers per thread)
constexpr int PCL_U = 8;       // columns unrolled per loop iteration

__device__ __forceinline__ unsigned pcl_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned pcl_mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void pcl_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned pcl_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned pcl_cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pcl_mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void pcl_mbar_expect_tx(unsigned addr, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// The inbox lives in this CTA's shared memory (never cached in L1), so observing the phase flip
// orders the st.async payload; the CTA-scope form avoids the L1 invalidate (CCTL.IVALL) that the
// .acquire.cluster form costs every column.
__device__ __forceinline__ bool pcl_mbar_try_wait(unsigned addr, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool pcl_mbar_wait(unsigned addr, unsigned parity) {
    if (pcl_mbar_try_wait(addr, parity)) return true;
    const long long t0 = clock64();
    while (!pcl_mbar_try_wait(addr, parity))
        if (clock64() - t0 > kSpinTimeoutCycles) return false;
    return true;
}
__device__ __forceinline__ void pcl_st_async_v2(unsigned raddr, unsigned long long a, unsigned long long b, unsigned rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "l"(a), "l"(b), "r"(rmbar) : "memory");
}
__device__ __forceinline__ unsigned long long pcl_bits(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ unsigned long long pcl_bits(float x) { return (unsigned long long)__float_as_uint(x); }
__device__ __forceinline__ void pcl_from_bits(unsigned long long b, double& x) { x = __longlong_as_double((long long)b); }
__device__ __forceinline__ void pcl_from_bits(unsigned long long b, float& x) { x = __uint_as_float((unsigned)b); }


// Exact arg-max over the lanes of a warp: largest v (v >= 0, candidates have v > 0), lowest pos
// among equals.  Returns the winning lane or -1.  Fast path: one redux on the top 32 bits of the
// value and one ballot; only if several lanes share the top word do the low word / position
// reductions run.  Split in two so that independent work can be placed under the redux latency.
__device__ __forceinline__ unsigned pcl_argmax_key(double v) { return (unsigned)((unsigned long long)__double_as_longlong(v) >> 32); }
__device__ __forceinline__ unsigned pcl_argmax_key(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned pcl_argmax_issue(unsigned key) { return __reduce_max_sync(0xffffffffu, key); }
__device__ __forceinline__ int pcl_argmax_finish(double v, int pos, unsigned key, unsigned mkey) {
    const bool c1 = (key == mkey) && (v > 0.0);
    const unsigned m1 = __ballot_sync(0xffffffffu, c1);
    if (m1 == 0u) return -1;
    if ((m1 & (m1 - 1u)) == 0u) return __ffs(m1) - 1;
    const unsigned lo = (unsigned)(unsigned long long)__double_as_longlong(v);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
    const bool c2 = c1 && lo == mlo;
    const int mpos = __reduce_min_sync(0xffffffffu, c2 ? pos : INT_MAX);
    return __ffs(__ballot_sync(0xffffffffu, c2 && pos == mpos)) - 1;
}
__device__ __forceinline__ int pcl_argmax_finish(float v, int pos, unsigned key, unsigned mkey) {
    const bool c1 = (key == mkey) && (v > 0.0f);
    const unsigned m1 = __ballot_sync(0xffffffffu, c1);
    if (m1 == 0u) return -1;
    if ((m1 & (m1 - 1u)) == 0u) return __ffs(m1) - 1;
    const int mpos = __reduce_min_sync(0xffffffffu, c1 ? pos : INT_MAX);
    return __ffs(__ballot_sync(0xffffffffu, c1 && pos == mpos)) - 1;
}
template <typename T>
__device__ __forceinline__ int pcl_warp_argmax(T v, int pos) {
    const unsigned key = pcl_argmax_key(v);
    return pcl_argmax_finish(v, pos, key, pcl_argmax_issue(key));
}

template <typename T, int W, int RPT, int NT>
struct PclShared {
    static constexpr int NW = NT / 32;
    static constexpr int EPV = 16 / (int)sizeof(T);
    static constexpr int NV = W / EPV;
    ulonglong2 box[2][PCL_GMAX][1 + NV];   // inbox: one message {header, row} per source CTA
    // per-warp candidate rows (coordinates of their column) and candidates, double-buffered by column parity like the
    // inbox: the readers of column j (after that column's CTA barrier) and the writers of column j + 1 (before the next
    // barrier) are not ordered by any barrier — a warp that stalls right behind the barrier must not find its
    // candidates overwritten by a warp that is a whole column ahead (racecheck: profiles/r02_sanitizer_racecheck_*)
    T stage[2][NW][W];
    T s_top[W];                            // zero/NaN-pivot path only
    T s_val[2][NW];
    int s_pos[2][NW];
    unsigned long long mbar[2];
    unsigned long long mbar_bulk;          // TMA bulk copies of the multipliers (fused kernel)
    int s_mv_dst[2 * W], s_mv_src[2 * W];
    int s_nmv;
    int s_has_top;
};
template <typename T, int W, int RPT, int NT>
constexpr size_t pcl_tile_bytes() { return (size_t)W * NT * RPT * sizeof(T); }
// NSUB > 1 (left-looking fused kernel): behind the tile, U of the current sub-block's columns on the rows of the
// earlier sub-blocks (MAXTOP x W) and the strictly lower triangle of their L11 (MAXTOP x MAXTOP)
template <typename T, int W, int RPT, int NT, int NSUB>
constexpr size_t pcl_smem_bytes() {
    return pcl_tile_bytes<T, W, RPT, NT>() + (size_t)(NSUB - 1) * W * W * sizeof(T) +
           (size_t)(NSUB - 1) * W * (NSUB - 1) * W * sizeof(T);
}

// local candidate of the current column (register index 0): strict '>' from amax = 0 (NaN never
// wins), lowest position on ties; rows at positions < j are finished
template <typename T, int W, int RPT>
__device__ __forceinline__ void pcl_local_cand(const T (&a)[RPT][W], const int (&pos)[RPT], int j, T& best, int& bpos) {
    best = T(0);
    bpos = INT_MAX;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        if (pos[q] >= j && pos[q] != INT_MAX) {
            const T v = tabs(a[q][0]);
            if (v > best || (v == best && v > T(0) && pos[q] < bpos)) { best = v; bpos = pos[q]; }
        }
    }
}

// NSUB > 1: ONE launch factors up to NSUB consecutive W-wide sub-blocks of the outer panel (p.wc = their total
// width), LEFT-LOOKING: before sub-block s is eliminated, the cluster itself applies the updates of the sub-blocks
// 0 .. s-1 of this launch to it — U = L11^{-1} A12 on the s W rows already pivoted (every CTA redundantly, warp-local
// forward substitution in registers, L11 staged in shared memory), then a[c] -= sum_k L[row, k] U[k, c] for the
// thread's own rows with its multipliers streamed from global memory (L2) and U broadcast from shared memory.
// That replaces the unit-lower TRSM + Schur GEMM launch pairs of the recursive panel between these sub-blocks and
// NSUB - 1 base-kernel launches (their load / exit / launch gaps); the arithmetic of an entry is the FMA sequence
// pivot 0, 1, 2, ... of the unblocked algorithm.  Sub-blocks are separated by a cluster barrier (global-memory
// visibility: barrier.cluster release / acquire + __threadfence, loads of data other CTAs wrote go through L2).
template <typename T, int W, int RPT, int NT, int NSUB = 1>
__global__ void __launch_bounds__(NT, 1) panel_cluster_kernel(PanelArgs<T> p) {
    constexpr int NW = NT / 32;
    constexpr int EPV = 16 / (int)sizeof(T);       // elements per 16-byte vector
    constexpr int NV = W / EPV;                    // vectors per row
    constexpr int ROWS = NT * RPT;                 // rows of one CTA
    static_assert(W % 8 == 0 && W <= 32, "live-window steps of 8");
    static_assert(NT / PCL_GMAX >= NV, "one sender thread per (destination, row vector)");
    static_assert(NW <= 32 && 2 * W <= NT, "warp-level bookkeeping");
    __shared__ __align__(16) PclShared<T, W, RPT, NT> sh;
    extern __shared__ __align__(16) unsigned char pcl_dyn[];
    T* tile = reinterpret_cast<T*>(pcl_dyn);       // tile[c * ROWS + local row]: finished entries

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int me = (int)pcl_cluster_rank();
    const int G = (int)pcl_cluster_size();   // power of two
    const int lgG = 31 - __clz(G);
    constexpr int MAXTOP = (NSUB - 1) * W;
    T* Us = tile + (size_t)W * ROWS;           // [MAXTOP][W]   (NSUB > 1)
    T* Ls = Us + (size_t)MAXTOP * W;           // [MAXTOP][MAXTOP]: Ls[k * MAXTOP + r] = L11[r, k], r > k
    const int J0 = p.j0;                       // first row / column of the launch
    const int wtot = p.wc;
    const int nsub = NSUB > 1 ? (wtot + W - 1) / W : 1;
    if (NSUB > 1) p.wc = min(W, wtot);

    if (tid == 0) {
        pcl_mbar_init(pcl_smem_u32(&sh.mbar[0]), 1);
        pcl_mbar_init(pcl_smem_u32(&sh.mbar[1]), 1);
        pcl_mbar_init(pcl_smem_u32(&sh.mbar_bulk), 1);
        sh.s_has_top = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool dbg = p.dbg != nullptr && me == 0 && tid == 0;
    if (dbg) p.dbg[0] = clock64();
    const long long t_launch = dbg ? clock64() : 0;

    // sender role / inbox addresses do not depend on the sub-block
    const int s_dst = tid & (G - 1), s_k = tid >> lgG;
    const unsigned raddr0 = pcl_mapa(pcl_smem_u32(&sh.box[0][me][1 + (s_k < NV ? s_k : 0)]), (unsigned)s_dst);
    const unsigned rhdr0 = pcl_mapa(pcl_smem_u32(&sh.box[0][me][0]), (unsigned)s_dst);
    const unsigned rbar0 = pcl_mapa(pcl_smem_u32(&sh.mbar[0]), (unsigned)s_dst);
    constexpr unsigned BOXB = (unsigned)sizeof(sh.box[0]);   // parity stride of the inbox

    unsigned bulk_parity = 0;   // phase of sh.mbar_bulk
#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
    const int wc = p.wc;
    long long t_pro = dbg ? clock64() : 0;
    T a[RPT][W];   // a[q][c]: column (j + c) of row q while column j is being eliminated
    int ri[RPT];   // panel-local ORIGINAL row of each owned row (where it is loaded from)
    int pos[RPT];  // its current position in the LAPACK row order
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        ri[q] = me * ROWS + q * NT + tid;
        pos[q] = ri[q] < p.m ? ri[q] : INT_MAX;   // padding rows are never candidates
        const T* src = p.A + (long long)p.j0 * p.lda + (p.j0 + (ri[q] < p.m ? ri[q] : 0));
#pragma unroll
        for (int c = 0; c < W; ++c) {
            // (NSUB > 1: other CTAs of the cluster moved these rows in the previous sub-block: read through L2)
            a[q][c] = (ri[q] < p.m && c < wc) ? (NSUB > 1 ? __ldcg(src + (long long)c * p.lda) : src[(long long)c * p.lda]) : T(0);
        }
    }
    if (dbg) { const long long t_ = clock64(); p.dbg[16] = t_ - t_pro; t_pro = t_; p.dbg[17] = 0; p.dbg[18] = 0; }
    if constexpr (NSUB > 1) {
        if (sub > 0) {
            const int sW = p.j0 - J0;   // rows / columns already factored in this launch
            // ---- A: U = L11^{-1} A[J0 .. J0 + sW, sub-block] in every CTA: L11 to shared memory ...
            for (int idx = tid; idx < sW * sW; idx += NT) {
                const int k = idx / sW, r = idx - k * sW;
                Ls[k * MAXTOP + r] = (r > k) ? __ldcg(p.A + (long long)(J0 + k) * p.lda + J0 + r) : T(0);
            }
            __syncthreads();
            // ... each warp solves its CC columns in registers (lane l holds rows l, l + 32, ...)
            constexpr int CC = (W + NW - 1) / NW, RR = (MAXTOP + 31) / 32;
            T bt[RR][CC];
#pragma unroll
            for (int rr = 0; rr < RR; ++rr)
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    const int row = rr * 32 + lane, col = warp * CC + c;
                    bt[rr][c] = (row < sW && col < wc) ? __ldcg(p.A + (long long)(p.j0 + col) * p.lda + J0 + row) : T(0);
                }
#pragma unroll
            for (int kr = 0; kr < RR; ++kr) {
                if (kr * 32 < sW) {
                    const int kend = min(32, sW - kr * 32);
                    for (int kk = 0; kk < kend; ++kk) {
                        const int k = kr * 32 + kk;
                        T xk[CC];
#pragma unroll
                        for (int c = 0; c < CC; ++c) xk[c] = __shfl_sync(0xffffffffu, bt[kr][c], kk);
#pragma unroll
                        for (int rr = kr; rr < RR; ++rr) {
                            const T l = (rr * 32 + lane < MAXTOP) ? Ls[k * MAXTOP + rr * 32 + lane] : T(0);   // 0 for rows <= k
#pragma unroll
                            for (int c = 0; c < CC; ++c) bt[rr][c] = tfma(-l, xk[c], bt[rr][c]);
                        }
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < RR; ++rr)
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    const int row = rr * 32 + lane, col = warp * CC + c;
                    // (global memory still holds the unsolved rows: every CTA reads them above, so the final U
                    // entries are written out only after the column loop, when all CTAs are past this point)
                    if (row < sW && col < W) Us[row * W + col] = bt[rr][c];
                }
            __syncthreads();
            if (dbg) { const long long t_ = clock64(); p.dbg[17] = t_ - t_pro; t_pro = t_; }
            // ---- B: the thread's own rows: a[c] -= sum_k L[row, k] U[k, c], k ascending.  The multipliers L[row, k]
            // (written to global memory by this launch's earlier sub-blocks) are staged W columns at a time into the
            // tile, which is idle until the column loop, with asynchronous 8-byte copies (cp.async: every copy of
            // the CTA in flight at once; loading them into registers a few at a time is bound by L2 latency with two
            // warps per scheduler: measured 30 k cycles per 8-column sub-block at 8 rows per thread)
            for (int k0 = 0; k0 < sW; k0 += W) {
                {
                    // one TMA bulk copy per multiplier column: this CTA's rows are contiguous in global memory and in
                    // the tile (8-byte cp.async copies took ~8 k cycles per round: 16 B per clock and SM)
                    const long long rbase = (long long)p.j0 + (long long)me * ROWS;
                    const int nrow = max(0, min(ROWS, p.m - me * ROWS));   // rows of this CTA inside the panel
                    constexpr int EPB = 16 / (int)sizeof(T);
                    const int nbulk = nrow & ~(EPB - 1);                    // 16-byte multiple
                    const unsigned mb = pcl_smem_u32(&sh.mbar_bulk);
                    if (tid == 0 && nbulk > 0) {
                        // the tile was last touched through the generic proxy (all of it behind a CTA barrier)
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        pcl_mbar_expect_tx(mb, (unsigned)(W * nbulk * (int)sizeof(T)));
                        for (int c = 0; c < W; ++c) {
                            const T* gcol = p.A + (long long)(J0 + k0 + c) * p.lda + rbase;
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(pcl_smem_u32(tile + c * ROWS)), "l"(gcol), "r"((unsigned)(nbulk * (int)sizeof(T))), "r"(mb)
                                         : "memory");
                        }
                    }
                    for (int i = tid; i < (nrow - nbulk) * W; i += NT) {   // the odd tail rows
                        const int c = i / (nrow - nbulk), r = nbulk + (i - c * (nrow - nbulk));
                        tile[c * ROWS + r] = __ldcg(p.A + (long long)(J0 + k0 + c) * p.lda + rbase + r);
                    }
                    if (nbulk > 0) {
                        if (!pcl_mbar_wait(mb, bulk_parity)) {
                            atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT);
                            return;
                        }
                        bulk_parity ^= 1u;
                    }
                }
                __syncthreads();
                if constexpr (W <= 16 && RPT > 1) {
                    // narrow blocks, several rows per thread: one row of U in registers serves all the thread's rows
                    // (measured with the loops the other way round: 9.5 k cycles per round of 8 multiplier columns
                    // at 8 rows per thread, five shared-memory loads per eight FMAs)
#pragma unroll 2
                    for (int c = 0; c < W; ++c) {
                        const T* urow = Us + (k0 + c) * W;
                        T u[W];
#pragma unroll
                        for (int cc = 0; cc < W; ++cc) u[cc] = urow[cc];
#pragma unroll
                        for (int q = 0; q < RPT; ++q) {
                            if (ri[q] < p.m) {
                                const T nl = -tile[c * ROWS + q * NT + tid];
#pragma unroll
                                for (int cc = 0; cc < W; ++cc) a[q][cc] = tfma(nl, u[cc], a[q][cc]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < RPT; ++q) {
                        if (ri[q] < p.m) {
                            const T* lq = tile + q * NT + tid;
#pragma unroll 4
                            for (int c = 0; c < W; ++c) {
                                const T nl = -lq[c * ROWS];
                                const T* urow = Us + (k0 + c) * W;
#pragma unroll
                                for (int cc = 0; cc < W; ++cc) a[q][cc] = tfma(nl, urow[cc], a[q][cc]);
                            }
                        }
                    }
                }
                __syncthreads();   // the tile is overwritten by the next round / the column loop
            }
            if (dbg) { const long long t_ = clock64(); p.dbg[18] = t_ - t_pro; t_pro = t_; }
        }
    }
    // sender role (set up before the sub-block loop): thread (dst, k) pushes row vector k of this CTA's
    // message to CTA dst; the threads with k == 0 also push the header
    // warp 0 of every CTA tracks which original row sits at each touched position
    int top_src = lane;      // lane l < W: original row now at position l
    int ext_row = -1;        // lane e: e-th position >= W that took part in an interchange
    int ext_src = -1;        //         and the original row now sitting there
    int ext_n = 0;

    if (sub == 0) pcl_cluster_sync();  // mbarriers initialised cluster-wide before any st.async
    if (dbg) p.dbg[1] = clock64();
#ifdef PCL_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#endif

    T best;
    int bpos, wl;
    pcl_local_cand<T, W, RPT>(a, pos, 0, best, bpos);
    wl = pcl_warp_argmax(best, bpos);
    if (lane == wl) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            if (pos[q] == bpos) {
#pragma unroll
                for (int c = 0; c < W; ++c) sh.stage[0][warp][c] = a[q][c];
            }
        }
    }

    // publish the candidates of column jn (row staged by the caller in that column's coordinates)
#define PCL_PUBLISH(jn)                                                                              \
    {                                                                                                \
        const int par_ = (jn) & 1;                                                                   \
        if (lane == wl) { sh.s_val[par_][warp] = best; sh.s_pos[par_][warp] = bpos; }                \
        if (wl < 0 && lane == 0) { sh.s_val[par_][warp] = T(0); sh.s_pos[par_][warp] = INT_MAX; }    \
        const int nv_ = (W - (jn) + EPV - 1) / EPV;   /* live row vectors of that column */          \
        if (tid == 0) pcl_mbar_expect_tx(pcl_smem_u32(&sh.mbar[par_]), (unsigned)(G * (1 + nv_)) * 16u); \
        PCL_T(0);                                                                                    \
        __syncthreads();                                                                             \
        const T cv_ = lane < NW ? sh.s_val[par_][lane] : T(0);                                       \
        const int cpos_ = lane < NW ? sh.s_pos[par_][lane] : INT_MAX;                                \
        int cw_ = pcl_warp_argmax(cv_, cpos_);                                                       \
        const int cp_ = cw_ < 0 ? INT_MAX : __shfl_sync(0xffffffffu, cpos_, cw_);                    \
        if (cw_ < 0) cw_ = 0;                                                                        \
        const unsigned rb_ = rbar0 + (unsigned)par_ * 8u;                                            \
        if (s_k < nv_) {                                                                             \
            const ulonglong2 v_ = reinterpret_cast<const ulonglong2*>(&sh.stage[par_][cw_][0])[s_k]; \
            pcl_st_async_v2(raddr0 + (unsigned)par_ * BOXB, v_.x, v_.y, rb_);                        \
        }                                                                                            \
        if (s_k == 0) pcl_st_async_v2(rhdr0 + (unsigned)par_ * BOXB, (unsigned long long)(unsigned)cp_, 0ull, rb_); \
        PCL_T(1);                                                                                    \
    }

    // one column with LIVE live register elements (a[q][0 .. LIVE-1])
#define PCL_COLUMN(LIVE)                                                                             \
    {                                                                                                \
        const int par = j & 1;                                                                       \
        PCL_T(5);                                                                                    \
        if (!pcl_mbar_wait(pcl_smem_u32(&sh.mbar[par]), (unsigned)(j >> 1) & 1u)) {                  \
            atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT);                                             \
            return;                                                                                  \
        }                                                                                            \
        PCL_T(2);                                                                                    \
        /* every warp picks the winner among the G candidates; each candidate lane also computes */  \
        /* the reciprocal of its own value under the redux latency                               */  \
        T gval = T(1), gv = T(0);                                                                    \
        int gpos = INT_MAX;                                                                          \
        if (lane < G) {                                                                              \
            gpos = (int)(unsigned)sh.box[par][lane][0].x;                                            \
            const T v0_ = reinterpret_cast<const T*>(&sh.box[par][lane][1])[0];                      \
            if (gpos != INT_MAX) { gval = v0_; gv = tabs(v0_); }                                     \
        }                                                                                            \
        const unsigned gkey = pcl_argmax_key(gv);                                                    \
        const unsigned gmax = pcl_argmax_issue(gkey);                                                \
        T rinv = T(1) / gval;                                                                        \
        const int gl = pcl_argmax_finish(gv, gpos, gkey, gmax);                                      \
        const bool none = gl < 0;  /* all-zero (or all-NaN) subcolumn: kp = k */                     \
        const int gsel = none ? 0 : gl;                                                              \
        rinv = __shfl_sync(0xffffffffu, rinv, gsel);                                                 \
        const int piv = none ? j : __shfl_sync(0xffffffffu, gpos, gsel);                             \
        const T* prow = reinterpret_cast<const T*>(&sh.box[par][gsel][1]);                           \
        bool scale = true;                                                                           \
        if (none) PCL_NONE_PATH()                                                                    \
        PCL_T(3);                                                                                    \
        if (me == 0 && tid == 0) p.ipiv[p.j0 + j] = p.j0 + piv;                                      \
        /* the pivot row is frozen: its warp copies the staged row (pivot, U entries) to the tile */ \
        {                                                                                            \
            int own = -1;                                                                            \
            _Pragma("unroll") for (int q = 0; q < RPT; ++q) if (pos[q] == piv) own = q * NT + tid;   \
            const unsigned ob = __ballot_sync(0xffffffffu, own >= 0);                                \
            if (ob) {                                                                                \
                const int orow = __shfl_sync(0xffffffffu, own, __ffs(ob) - 1);                       \
                const T* srow = none ? sh.s_top : &sh.stage[par][warp][0];                           \
                if (lane < W - j && lane < W) tile[(j + lane) * ROWS + orow] = srow[lane];           \
            }                                                                                        \
        }                                                                                            \
        /* positions, multipliers */                                                                 \
        _Pragma("unroll") for (int q = 0; q < RPT; ++q) {                                            \
            if (pos[q] == piv) pos[q] = j;            /* the pivot row: finished */                  \
            else if (pos[q] == j) pos[q] = piv;       /* the displaced top row stays active */       \
            const bool upd = pos[q] > j && pos[q] != INT_MAX;                                        \
            if (scale) a[q][0] *= rinv;               /* garbage in frozen rows is never read */     \
            if (upd) tile[j * ROWS + q * NT + tid] = a[q][0];                                        \
        }                                                                                            \
        PCL_T(6);                                                                                    \
        /* next column first, so that its arg-max is in flight under the rest of the update */       \
        {                                                                                            \
            const T pr1 = prow[1];                                                                   \
            _Pragma("unroll") for (int q = 0; q < RPT; ++q) a[q][1] = tfma(-a[q][0], pr1, a[q][1]);  \
        }                                                                                            \
        best = T(0);                                                                                 \
        bpos = INT_MAX;                                                                              \
        _Pragma("unroll") for (int q = 0; q < RPT; ++q) {                                            \
            if (pos[q] >= j + 1 && pos[q] != INT_MAX) {                                              \
                const T v = tabs(a[q][1]);                                                           \
                if (v > best || (v == best && v > T(0) && pos[q] < bpos)) { best = v; bpos = pos[q]; } \
            }                                                                                        \
        }                                                                                            \
        const unsigned akey = pcl_argmax_key(best);                                                  \
        const unsigned amax = pcl_argmax_issue(akey);          /* in flight under the update */      \
        _Pragma("unroll") for (int q = 0; q < RPT; ++q) {                                            \
            const T nl = -a[q][0];                                                                   \
            a[q][0] = a[q][1];                                                                       \
            _Pragma("unroll") for (int c = 2; c < (LIVE); ++c) a[q][c - 1] = tfma(nl, prow[c], a[q][c]); \
        }                                                                                            \
        PCL_T(7);                                                                                    \
        if (warp == 0 && piv != j) PCL_BOOKKEEP()                                                    \
        wl = pcl_argmax_finish(best, bpos, akey, amax);                                              \
        if (lane == wl) {   /* stage the row in the coordinates of column j + 1 */                   \
            _Pragma("unroll") for (int q = 0; q < RPT; ++q) {                                        \
                if (pos[q] == bpos) {                                                                \
                    _Pragma("unroll") for (int c = 0; c < (LIVE) - 1; ++c) sh.stage[par ^ 1][warp][c] = a[q][c]; \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
        PCL_T(4);                                                                                    \
        if (j + 1 < wc) PCL_PUBLISH(j + 1)                                                           \
    }

#define PCL_NONE_PATH()                                                                              \
    {   /* the pivot row is the row at position j: its owner hands it to every CTA (rare path) */    \
        _Pragma("unroll") for (int q = 0; q < RPT; ++q) {                                            \
            if (pos[q] == j) {                                                                       \
                _Pragma("unroll") for (int c = 0; c < W; ++c) sh.s_top[c] = a[q][c];                 \
                sh.s_has_top = 1;                                                                    \
            }                                                                                        \
        }                                                                                            \
        __syncthreads();                                                                             \
        if (sh.s_has_top) {                                                                          \
            for (int i = tid; i < G * W; i += NT) {                                                  \
                const int d = i / W, c = i - d * W;                                                  \
                if (d == me) continue;                                                               \
                const unsigned ra = pcl_mapa(pcl_smem_u32(&sh.s_top[c]), (unsigned)d);               \
                if constexpr (sizeof(T) == 8)                                                        \
                    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(sh.s_top[c]) : "memory"); \
                else                                                                                 \
                    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(sh.s_top[c]) : "memory"); \
            }                                                                                        \
        }                                                                                            \
        pcl_cluster_sync();                                                                          \
        if (tid == 0) sh.s_has_top = 0;                                                              \
        prow = sh.s_top;                                                                             \
        const T pv = prow[0];                                                                        \
        scale = (pv != T(0));                                                                        \
        rinv = T(1) / pv;                                                                            \
        if (me == 0 && tid == 0 && pv == T(0) && *p.info == 0) *p.info = p.j0 + j + 1;               \
    }

#define PCL_BOOKKEEP()                                                                               \
    {   /* interchange of positions j and piv */                                                     \
        const int sj = __shfl_sync(0xffffffffu, top_src, j);                                         \
        if (piv < W) {                                                                               \
            const int sp = __shfl_sync(0xffffffffu, top_src, piv);                                   \
            if (lane == j) top_src = sp;                                                             \
            if (lane == piv) top_src = sj;                                                           \
        } else {                                                                                     \
            const unsigned hit = __ballot_sync(0xffffffffu, ext_row == piv);                         \
            int e;                                                                                   \
            if (hit) e = __ffs(hit) - 1;                                                             \
            else {                                                                                   \
                e = ext_n++;                                                                         \
                if (lane == e) { ext_row = piv; ext_src = piv; }                                     \
            }                                                                                        \
            const int sp = __shfl_sync(0xffffffffu, ext_src, e);                                     \
            if (lane == j) top_src = sp;                                                             \
            if (lane == e) ext_src = sj;                                                             \
        }                                                                                            \
    }

    PCL_PUBLISH(0)
    // Four copies of the rolled column loop: the live window shrinks by one element per column,
    // the copies bound the dead elements a column still computes on to < 8.
    int j = 0;
    if constexpr (W >= 32) {
#pragma unroll 1
        for (; j < wc && j < W - 24; ++j) PCL_COLUMN(W)
    }
    if constexpr (W >= 24) {
#pragma unroll 1
        for (; j < wc && j < W - 16; ++j) PCL_COLUMN(W >= 32 ? 24 : W)
    }
    if constexpr (W >= 16) {
#pragma unroll 1
        for (; j < wc && j < W - 8; ++j) PCL_COLUMN(W >= 24 ? 16 : W)
    }
#pragma unroll 1
    for (; j < wc; ++j) PCL_COLUMN(W >= 16 ? 8 : W)
#undef PCL_COLUMN
#undef PCL_PUBLISH
#undef PCL_NONE_PATH
#undef PCL_BOOKKEEP
    if (dbg) p.dbg[2] = clock64();
#ifdef PCL_TIMING
    if (dbg) for (int i = 0; i < 8; ++i) p.dbg[8 + i] = tacc[i];
#endif
    // every finished entry of this CTA's rows now sits in the tile: write the block back, every
    // row at its final position
    __syncthreads();
    if constexpr (NSUB > 1) {
        if (sub > 0 && me == 0) {   // U of this sub-block's columns on the rows of the earlier sub-blocks
            const int sW = p.j0 - J0;
            for (int idx = tid; idx < sW * wc; idx += NT) {
                const int c = idx / sW, r = idx - c * sW;
                p.A[(long long)(p.j0 + c) * p.lda + J0 + r] = Us[r * W + c];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        if (ri[q] < p.m) {
            T* dst = p.A + (long long)p.j0 * p.lda + (p.j0 + pos[q]);
            const T* srcT = tile + q * NT + tid;
#pragma unroll 8
            for (int c = 0; c < wc; ++c) dst[(long long)c * p.lda] = srcT[c * ROWS];
        }
    }
    if (dbg) p.dbg[3] = clock64();
    // the same interchanges on the other columns of the outer panel [pc0, pc1)
    const int nleft = p.j0 - p.pc0;
    const int ncols = nleft + (p.pc1 - (p.j0 + wc));
    if (ncols > 0) {
        if (warp == 0) {
            const bool m1 = lane < W && top_src != lane;
            const bool m2 = lane < ext_n && ext_src != ext_row;
            const unsigned b1 = __ballot_sync(0xffffffffu, m1), b2 = __ballot_sync(0xffffffffu, m2);
            const unsigned below = (1u << lane) - 1u;
            if (m1) { const int s = __popc(b1 & below); sh.s_mv_dst[s] = lane; sh.s_mv_src[s] = top_src; }
            if (m2) { const int s = __popc(b1) + __popc(b2 & below); sh.s_mv_dst[s] = ext_row; sh.s_mv_src[s] = ext_src; }
            if (lane == 0) sh.s_nmv = __popc(b1) + __popc(b2);
        }
        __syncthreads();
        const int nmv = sh.s_nmv;
        if (nmv > 0) {
            // one group of 2W threads per column (all loads of a column before its stores); every
            // thread keeps up to 4 columns in flight so the pass costs one memory latency, not four
            constexpr int GRP = 2 * W, CPP = NT / GRP, MLP = 4;
            const int slot = tid % GRP, sub = tid / GRP;
            const int src = slot < nmv ? sh.s_mv_src[slot] : 0, dst = slot < nmv ? sh.s_mv_dst[slot] : 0;
            for (int c0 = me * CPP; c0 < ncols; c0 += G * CPP * MLP) {
                T v[MLP];
                T* base[MLP];
                bool on[MLP];
#pragma unroll
                for (int u = 0; u < MLP; ++u) {
                    const int c = c0 + u * G * CPP + sub;
                    on[u] = c < ncols && slot < nmv;
                    const int col = (c < nleft) ? (p.pc0 + c) : (p.j0 + wc + (c - nleft));
                    base[u] = p.A + (long long)col * p.lda + p.j0;
                    v[u] = T(0);
                    if (on[u]) v[u] = NSUB > 1 ? __ldcg(base[u] + src) : base[u][src];
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < MLP; ++u)
                    if (on[u]) base[u][dst] = v[u];
            }
        }
    }
    if (dbg) p.dbg[4] = clock64();
    if (NSUB > 1) __threadfence();   // this sub-block's global writes before the barrier's release
    pcl_cluster_sync();  // no CTA leaves (or starts the next sub-block) while a peer could still address its shared memory
    if (dbg) { p.dbg[5] = clock64(); p.dbg[6] = p.m; p.dbg[7] = G; p.dbg[19] = clock64() - t_launch; }
    p.j0 += W;
    p.m -= W;
    p.wc = min(W, wtot - (sub + 1) * W);
    }   // sub-blocks
}

}  // namespace b200lu
