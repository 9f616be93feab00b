// panel.cuh — partial-pivoting base panel (getf2) as ONE persistent multi-CTA kernel.
//
// Restates the column loop of `_blocked_lu_panel!` (reference
// src/blocked_lufact.jl:93-122) and the pivot rule of `_blocked_lu_find_pivot`
// (src/blocked_lufact.jl:38-54): amax starts at 0, strict `>` (NaN never wins),
// first index of the maximum; zero pivot => info = k once, no swap, no scaling,
// the rank-1 update still runs; scaling multiplies by inv(pivot); updates are FMAs.
//
// B200 mapping: the (m x W) block lives in REGISTERS for the whole kernel: thread
// t of CTA c owns RPT rows (all W columns of each).  Per column:
//   1. redux.sync arg-max (lowest index on ties) -> CTA candidate;
//   2. the candidate (|value|, row index, the whole candidate row) is written to an
//      L2-resident mailbox as 64-bit {data32, flag32} packets ("LL" style: the
//      flag travels with the data, so there is NO fence and NO separate flag);
//   3. every warp of every CTA reads the G candidate headers and reduces them on
//      its own (no barrier), then one thread per packet fetches the winning row
//      and the current top row into shared memory;
//   4. swap (pure register moves) + scale + rank-1 update, fused with a register
//      rotation that keeps the column loop compact (not unrolled).
// Mailboxes are double-buffered by column parity; flags are a monotone epoch, so
// nothing is ever reset between columns or launches.
// An extra "swapper" CTA follows the published pivots and applies each row
// interchange eagerly to the remaining columns of the OUTER panel
// [pc0, pc1) \ [j0, j0+wc), so the recursive panel needs no laswp kernels.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace b200lu {

constexpr int PANEL_GMAX = 128;  // max CTAs cooperating on one base panel
constexpr int PANEL_WMAX = 32;   // max base width
// Mailbox (all 64-bit {data32, flag32} packets, double-buffered by column parity):
//   hdr[par][cta][0..3] : candidate |value| (1 or 2 words) and row index
//   row[par][cta][..]   : the candidate row (W values)
//   top[par][..]        : the current top row (from CTA 0)
//   piv[j]              : the chosen pivot row of column j (for the swapper CTA)
struct PanelMail {
    unsigned long long hdr[2][PANEL_GMAX][32];   // one 256-byte slot per CTA: spreads the polled lines over L2 slices
    unsigned long long row[2][PANEL_GMAX][2 * PANEL_WMAX];
    unsigned long long top[2][2 * PANEL_WMAX];
    unsigned long long piv[PANEL_WMAX];
};

template <typename T>
struct PanelArgs {
    T* A;                // full matrix (device), column-major
    long long lda;
    int j0;              // global row == column index of the block's top-left
    int m;               // rows in the block: global rows j0 .. j0+m-1
    int wc;              // columns in the block (<= W)
    int pc0, pc1;        // outer panel column range (for the swapper)
    int* ipiv;           // global, 0-based row indices
    int* info;           // 0 or 1-based first zero pivot
    int G;               // panel CTAs (grid = G + has_swapper)
    unsigned epoch;      // packets carry epoch + j + 1 at column j
    PanelMail* mail;
    int* deverr;
    long long* dbg;      // optional: 8 clock64 stamps per launch (B200LU_PANEL_DBG=1)
};

__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned data, unsigned flag) {
    const unsigned long long v = ((unsigned long long)flag << 32) | data;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until the packet carries `want`; returns false on watchdog timeout
__device__ __forceinline__ bool ll_wait(const unsigned long long* p, unsigned want, unsigned& data) {
    unsigned long long v = ll_load(p);
    if ((unsigned)(v >> 32) != want) {
        const long long t0 = clock64();
        do {
            v = ll_load(p);
            if (clock64() - t0 > kSpinTimeoutCycles) return false;
        } while ((unsigned)(v >> 32) != want);
    }
    data = (unsigned)v;
    return true;
}

template <typename T> struct Words;
template <> struct Words<double> {
    static constexpr int N = 2;
    __device__ static void split(double x, unsigned* w) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(x);
        w[0] = (unsigned)b; w[1] = (unsigned)(b >> 32);
    }
    __device__ static double join(const unsigned* w) {
        return __longlong_as_double((long long)(((unsigned long long)w[1] << 32) | w[0]));
    }
};
template <> struct Words<float> {
    static constexpr int N = 1;
    __device__ static void split(float x, unsigned* w) { w[0] = __float_as_uint(x); }
    __device__ static float join(const unsigned* w) { return __uint_as_float(w[0]); }
};

// Warp arg-max of non-negative values with lowest-index tie-break, on the redux unit
// (3 REDUX for double, 2 for float, instead of 5 dependent shuffle rounds).
// Lanes without a candidate pass v = 0, idx = INT_MAX.  Returns the winning
// (v, idx) in every lane and the winning lane in `wl` (undefined value if none).
__device__ __forceinline__ void warp_argmax(double& v, int& idx, int& wl) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(v);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool ismax = (hi == mhi) && (lo == mlo);
    const int midx = __reduce_min_sync(0xffffffffu, ismax ? idx : INT_MAX);
    wl = __ffs(__ballot_sync(0xffffffffu, ismax && idx == midx)) - 1;
    v = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    idx = midx;
}
__device__ __forceinline__ void warp_argmax(float& v, int& idx, int& wl) {
    const unsigned key = __float_as_uint(v);
    const unsigned mk = __reduce_max_sync(0xffffffffu, key);
    const bool ismax = key == mk;
    const int midx = __reduce_min_sync(0xffffffffu, ismax ? idx : INT_MAX);
    wl = __ffs(__ballot_sync(0xffffffffu, ismax && idx == midx)) - 1;
    v = __uint_as_float(mk);
    idx = midx;
}

template <typename T, int W, int RPT, int NT>
__global__ void __launch_bounds__(NT, 1) panel_base_kernel(PanelArgs<T> p) {
    constexpr int NW = NT / 32;
    constexpr int WN = Words<T>::N;
    constexpr int ROWW = W * WN;               // 32-bit words of one row
    static_assert(2 * ROWW <= NT && 32 + ROWW <= NT, "one publishing / fetching thread per packet");
    __shared__ T s_val[NW];
    __shared__ int s_idx[NW];
    __shared__ __align__(16) unsigned s_stage[4 + 2 * ROWW];  // [hdr 4][row ROWW][top ROWW]
    __shared__ __align__(16) unsigned s_rows[2 * ROWW];       // fetched [pivot row][top row]
    __shared__ int s_piv;
    __shared__ int s_win;
    __shared__ int s_abort;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int cta = blockIdx.x;
    const int G = p.G;
    const int wc = p.wc;
    PanelMail* mail = p.mail;

    // ------------------------------------------------------------------ swapper
    if (cta >= G) {
        const int nleft = p.j0 - p.pc0;
        const int nright = p.pc1 - (p.j0 + wc);
        const int ncols = nleft + nright;
        for (int j = 0; j < wc; ++j) {
            if (tid == 0) {
                unsigned d = 0;
                const bool ok = ll_wait(&mail->piv[j], p.epoch + j + 1, d);
                if (!ok) atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT);
                s_piv = ok ? (int)d : -1;
            }
            __syncthreads();
            const int piv = s_piv;
            __syncthreads();
            if (piv < 0) return;
            const int k = p.j0 + j;
            if (piv != k) {
                for (int c = tid; c < ncols; c += NT) {
                    const int col = (c < nleft) ? (p.pc0 + c) : (p.j0 + wc + (c - nleft));
                    T* pk = p.A + (long long)col * p.lda + k;
                    T* pp = p.A + (long long)col * p.lda + piv;
                    const T vk = *pk, vp = *pp;
                    *pk = vp;
                    *pp = vk;
                }
            }
        }
        return;
    }

    // ------------------------------------------------------------- panel CTAs
    T a[RPT][W];
    int ri[RPT];  // panel-local row index of each owned row
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        ri[q] = cta * (NT * RPT) + q * NT + tid;
#pragma unroll
        for (int c = 0; c < W; ++c) {
            a[q][c] = (ri[q] < p.m && c < wc)
                          ? p.A[(long long)(p.j0 + c) * p.lda + (p.j0 + ri[q])]
                          : T(0);
        }
    }
    if (tid == 0) s_abort = 0;
    T* stage_row = reinterpret_cast<T*>(&s_stage[4]);
    T* stage_top = reinterpret_cast<T*>(&s_stage[4 + ROWW]);
    const T* prow = reinterpret_cast<const T*>(&s_rows[0]);
    const T* trow = reinterpret_cast<const T*>(&s_rows[ROWW]);

    // The column loop is NOT unrolled (an unrolled body is ~400 KB of SASS and every
    // column then runs from a cold instruction cache).  To keep all register indices
    // static, the row registers are ROTATED left by one each column: the current
    // column always sits at index 0, the finished multiplier re-enters at index W-1,
    // and after W steps the row is back in natural order.  Every CTA rotates in
    // lock-step, so staged rows line up across CTAs.
#pragma unroll 1
    for (int j = 0; j < wc; ++j) {
        const int par = j & 1;
        const unsigned want = p.epoch + j + 1;
        const int nact = W - j;  // rotated indices [0, nact) are not yet factored columns
        // 1. local candidate: strict '>' from amax = 0 (NaN never wins), lowest row on ties
        T best = T(0);
        int bi = INT_MAX;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            if (ri[q] >= j && ri[q] < p.m) {
                const T v = tabs(a[q][0]);
                if (v > best) { best = v; bi = ri[q]; }
            }
        }
        int wl;
        warp_argmax(best, bi, wl);
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        T cb = s_val[0];
        int ci = s_idx[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const T ob = s_val[w];
            const int oi = s_idx[w];
            if (ob > cb || (ob == cb && oi < ci)) { cb = ob; ci = oi; }
        }
        // 2. stage the candidate message in shared memory, then one thread per packet
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            if (ri[q] == ci) {
#pragma unroll
                for (int c = 0; c < W; ++c) stage_row[c] = a[q][c];
            }
            if (ri[q] == j) {  // current top row (CTA 0 only)
#pragma unroll
                for (int c = 0; c < W; ++c) stage_top[c] = a[q][c];
            }
        }
        if (tid == 0) {
            *reinterpret_cast<T*>(&s_stage[0]) = cb;
            s_stage[WN] = (unsigned)ci;
        }
        __syncthreads();
        if (tid <= WN) ll_store(&mail->hdr[par][cta][tid], s_stage[tid], want);
        else if (tid >= 32 && tid < 32 + ROWW) ll_store(&mail->row[par][cta][tid - 32], s_stage[4 + tid - 32], want);
        if (cta == 0 && tid >= NT - ROWW) ll_store(&mail->top[par][tid - (NT - ROWW)], s_stage[4 + ROWW + tid - (NT - ROWW)], want);
        // 3a. warp 0 polls the G candidate headers (one lane per CTA; a single polling warp per
        //     CTA keeps the request pressure on the polled L2 lines low) and picks the winner
        bool dead = false;
        if (warp == 0) {
            T gv = T(0);
            int gi = INT_MAX, gc = 0;
            for (int c = lane; c < G; c += 32) {
                unsigned w[WN + 1];
                unsigned long long v[WN + 1];
#pragma unroll
                for (int x = 0; x <= WN; ++x) v[x] = ll_load(&mail->hdr[par][c][x]);
#pragma unroll
                for (int x = 0; x <= WN; ++x) {
                    if ((unsigned)(v[x] >> 32) != want) {
                        const long long t0 = clock64();
                        do {
                            __nanosleep(20);
                            v[x] = ll_load(&mail->hdr[par][c][x]);
                            if (clock64() - t0 > kSpinTimeoutCycles) { dead = true; break; }
                        } while ((unsigned)(v[x] >> 32) != want);
                    }
                    w[x] = (unsigned)v[x];
                }
                const T cv = Words<T>::join(w);
                const int cidx = (int)w[WN];
                if (cv > gv || (cv == gv && cidx < gi)) { gv = cv; gi = cidx; gc = c; }
            }
            int gl;
            if (!(gv > T(0))) gi = INT_MAX;
            warp_argmax(gv, gi, gl);
            gc = __shfl_sync(0xffffffffu, gc, gl < 0 ? 0 : gl);
            if (lane == 0) {
                const bool none0 = !(gv > T(0));
                s_piv = none0 ? j : gi;
                s_win = none0 ? -1 : gc;
            }
            if (__any_sync(0xffffffffu, dead) && lane == 0) { atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT); s_abort = 1; }
        }
        __syncthreads();
        if (s_abort) return;
        const int piv = s_piv;
        const int gc = s_win;
        const bool none = gc < 0;  // all-zero (or all-NaN) subcolumn: kp = k
        // 3b. fetch the winning row and the top row: one packet per thread
        if (tid < 2 * ROWW) {
            const unsigned long long* src = (tid < ROWW)
                ? (none ? &mail->top[par][tid] : &mail->row[par][gc][tid])
                : &mail->top[par][tid - ROWW];
            unsigned d = 0;
            if (!ll_wait(src, want, d)) dead = true;
            s_rows[tid] = d;
        }
        if (dead) { atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT); s_abort = 1; }
        __syncthreads();
        if (s_abort) return;
        // 4. swap + scale + rank-1 update fused with the rotation, all in registers
        const T pv = prow[0];
        if (cta == 0 && tid == 0) {
            p.ipiv[p.j0 + j] = p.j0 + piv;
            ll_store(&mail->piv[j], (unsigned)(p.j0 + piv), want);
            if (pv == T(0) && *p.info == 0) *p.info = p.j0 + j + 1;
        }
        const bool scale = (pv != T(0));
        const T rinv = T(1) / pv;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            if (ri[q] == j) {
#pragma unroll
                for (int c = 0; c < W; ++c) a[q][c] = prow[c];
            } else if (ri[q] == piv) {  // piv != j here
#pragma unroll
                for (int c = 0; c < W; ++c) a[q][c] = trow[c];
            }
            const bool upd = ri[q] > j && ri[q] < p.m;
            T l = a[q][0];
            if (upd && scale) l *= rinv;
#pragma unroll
            for (int c = 1; c < W; ++c) {
                T nv = a[q][c];
                if (upd && c < nact) nv = tfma(-l, prow[c], nv);
                a[q][c - 1] = nv;
            }
            a[q][W - 1] = l;
        }
    }
    // ragged last block: finish the W-step rotation so registers are in natural order
#pragma unroll 1
    for (int j = wc; j < W; ++j) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const T f = a[q][0];
#pragma unroll
            for (int c = 1; c < W; ++c) a[q][c - 1] = a[q][c];
            a[q][W - 1] = f;
        }
    }

#pragma unroll
    for (int q = 0; q < RPT; ++q) {
#pragma unroll
        for (int c = 0; c < W; ++c) {
            if (ri[q] < p.m && c < wc)
                p.A[(long long)(p.j0 + c) * p.lda + (p.j0 + ri[q])] = a[q][c];
        }
    }
}

}  // namespace b200lu
