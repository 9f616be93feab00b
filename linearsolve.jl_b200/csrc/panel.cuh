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
//   1. warp-shuffle arg-max (lowest index on ties) -> CTA candidate;
//   2. the candidate (|value|, row index, the whole candidate row) is written to an
//      L2-resident mailbox as 64-bit {data32, flag32} packets ("LL" style: the
//      flag travels with the data, so there is NO fence and NO separate flag);
//   3. every CTA gathers all G candidate messages (one L2 round trip, all
//      threads polling their own packets), warp 0 reduces them and stages the
//      winning row + the current top row in shared memory;
//   4. swap (pure register moves) + scale + rank-1 update.
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
// message words (32-bit): |val| (2) + idx (1) + row (2*W)  [float uses half of the value words]
constexpr int PANEL_MSG_WORDS = 3 + 2 * PANEL_WMAX;  // 67
constexpr int PANEL_TOP_WORDS = 2 * PANEL_WMAX;

struct PanelMail {
    unsigned long long msg[2][PANEL_GMAX][PANEL_MSG_WORDS + 1];
    unsigned long long top[2][PANEL_TOP_WORDS];
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
};

__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned data, unsigned flag) {
    const unsigned long long v = ((unsigned long long)flag << 32) | data;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
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

template <typename T, int W, int RPT, int NT>
__global__ void __launch_bounds__(NT, 1) panel_base_kernel(PanelArgs<T> p) {
    constexpr int NW = NT / 32;
    constexpr int WN = Words<T>::N;
    constexpr int OFF_IDX = WN;              // message layout: [val WN][idx 1][row W*WN]
    constexpr int OFF_ROW = WN + 1;
    constexpr int MSG = OFF_ROW + W * WN;    // words actually used
    extern __shared__ unsigned s_tab[];      // [G*MSG + W*WN] gathered message words
    __shared__ T s_val[NW];
    __shared__ int s_idx[NW];
    __shared__ __align__(16) T s_prow[PANEL_WMAX];
    __shared__ __align__(16) T s_trow[PANEL_WMAX];
    __shared__ int s_piv;
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
    const int total_words = G * MSG + W * WN;

#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (j < wc) {
            const int par = j & 1;
            const unsigned want = p.epoch + j + 1;
            // 1. local candidate: strict '>' from amax = 0, rows ascending
            T best = T(0);
            int bi = INT_MAX;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (ri[q] >= j && ri[q] < p.m) {
                    const T v = tabs(a[q][j]);
                    if (v > best) { best = v; bi = ri[q]; }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T ob = shfl_xor(best, off);
                const int oi = shfl_xor(bi, off);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
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
            // 2. publish: the candidate's owner writes the whole message as LL packets
            {
                unsigned long long* mb = mail->msg[par][cta];
                const bool nocand = (ci == INT_MAX);
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const bool writer = nocand ? (q == 0 && tid == 0) : (ri[q] == ci);
                    if (writer) {
                        unsigned w[WN];
                        Words<T>::split(cb, w);
#pragma unroll
                        for (int x = 0; x < WN; ++x) ll_store(mb + x, w[x], want);
                        ll_store(mb + OFF_IDX, (unsigned)ci, want);
#pragma unroll
                        for (int c = 0; c < W; ++c) {
                            Words<T>::split(nocand ? T(0) : a[q][c], w);
#pragma unroll
                            for (int x = 0; x < WN; ++x) ll_store(mb + OFF_ROW + c * WN + x, w[x], want);
                        }
                    }
                    if (ri[q] == j) {  // current top row (CTA 0 only)
                        unsigned w[WN];
#pragma unroll
                        for (int c = 0; c < W; ++c) {
                            Words<T>::split(a[q][c], w);
#pragma unroll
                            for (int x = 0; x < WN; ++x) ll_store(&mail->top[par][c * WN + x], w[x], want);
                        }
                    }
                }
            }
            // 3. gather all messages (every thread polls its own packets)
            {
                bool dead = false;
                for (int idx = tid; idx < total_words; idx += NT) {
                    const unsigned long long* src;
                    if (idx < G * MSG) {
                        const int c = idx / MSG;
                        src = &mail->msg[par][c][idx - c * MSG];
                    } else {
                        src = &mail->top[par][idx - G * MSG];
                    }
                    unsigned d = 0;
                    if (!ll_wait(src, want, d)) dead = true;
                    s_tab[idx] = d;
                }
                if (dead) { atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT); s_abort = 1; }
            }
            __syncthreads();
            if (warp == 0) {
                T gv = T(0);
                int gi = INT_MAX, gc = -1;
                for (int c = lane; c < G; c += 32) {
                    const T v = Words<T>::join(&s_tab[c * MSG]);
                    const int idx = (int)s_tab[c * MSG + OFF_IDX];
                    if (v > gv || (v == gv && idx < gi)) { gv = v; gi = idx; gc = c; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const T ob = shfl_xor(gv, off);
                    const int oi = shfl_xor(gi, off);
                    const int oc = shfl_xor(gc, off);
                    if (ob > gv || (ob == gv && oi < gi)) { gv = ob; gi = oi; gc = oc; }
                }
                const bool none = !(gv > T(0));  // all-zero (or all-NaN) subcolumn: kp = k
                if (lane < W) {
                    const T tr = Words<T>::join(&s_tab[G * MSG + lane * WN]);
                    const T pr = none ? tr : Words<T>::join(&s_tab[gc * MSG + OFF_ROW + lane * WN]);
                    s_trow[lane] = tr;
                    s_prow[lane] = pr;
                }
                if (lane == 0) s_piv = none ? j : gi;
            }
            __syncthreads();
            if (s_abort) return;
            // 4. swap + scale + rank-1 update, all in registers
            const int piv = s_piv;
            const T pv = s_prow[j];
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
                    for (int c = 0; c < W; ++c) a[q][c] = s_prow[c];
                } else if (ri[q] == piv) {  // piv != j here
#pragma unroll
                    for (int c = 0; c < W; ++c) a[q][c] = s_trow[c];
                }
                if (ri[q] > j && ri[q] < p.m) {
                    T l = a[q][j];
                    if (scale) l *= rinv;
                    a[q][j] = l;
#pragma unroll
                    for (int c = j + 1; c < W; ++c) a[q][c] = tfma(-l, s_prow[c], a[q][c]);
                }
            }
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
