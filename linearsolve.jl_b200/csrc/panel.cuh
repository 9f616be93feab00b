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
//   2. the candidate row and (val, idx) are published to a small L2-resident
//      mailbox with a release flag (double-buffered by column parity);
//   3. warp 0 of every CTA acquires all G flags, reduces the candidates, and
//      stages the winning row + the current top row in shared memory;
//   4. swap (pure register moves) + scale + rank-1 update.
// One L2 round trip per column; no grid.sync, no kernel relaunch.
// An extra "swapper" CTA follows the published pivots and applies each row
// interchange eagerly to the remaining columns of the OUTER panel
// [pc0, pc1) \ [j0, j0+wc), so the recursive panel needs no laswp kernels.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace b200lu {

constexpr int PANEL_GMAX = 128;  // max CTAs cooperating on one base panel
constexpr int PANEL_WMAX = 32;   // max base width

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
    int epoch;           // flags carry epoch + j + 1 at column j
    int* flags;          // [2][PANEL_GMAX]
    T* cand_val;         // [2][PANEL_GMAX]
    int* cand_idx;       // [2][PANEL_GMAX]
    T* rowbuf;           // [2][PANEL_GMAX][PANEL_WMAX]
    T* toprow;           // [2][PANEL_WMAX]
    int* progress;       // last finished epoch (for the swapper)
    int* deverr;
};

template <typename T, int W, int RPT, int NT>
__global__ void __launch_bounds__(NT, 1) panel_base_kernel(PanelArgs<T> p) {
    constexpr int NW = NT / 32;
    __shared__ T s_val[NW];
    __shared__ int s_idx[NW];
    __shared__ T s_prow[PANEL_WMAX];
    __shared__ T s_trow[PANEL_WMAX];
    __shared__ int s_piv;
    __shared__ int s_abort;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int cta = blockIdx.x;
    const int G = p.G;
    const int wc = p.wc;

    // ------------------------------------------------------------------ swapper
    if (cta >= G) {
        const int nleft = p.j0 - p.pc0;
        const int nright = p.pc1 - (p.j0 + wc);
        const int ncols = nleft + nright;
        if (tid == 0) s_abort = 0;
        __syncthreads();
        for (int j = 0; j < wc; ++j) {
            if (tid == 0) {
                const int want = p.epoch + j + 1;
                long long t0 = clock64();
                while (ld_acquire(p.progress) - want < 0) {
                    if (clock64() - t0 > kSpinTimeoutCycles) {
                        atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT);
                        s_abort = 1;
                        break;
                    }
                }
            }
            __syncthreads();
            if (s_abort) return;
            const int k = p.j0 + j;
            const int piv = ld_cg(p.ipiv + k);
            if (piv != k) {
                for (int c = tid; c < ncols; c += NT) {
                    const int col = (c < nleft) ? (p.pc0 + c) : (p.j0 + wc + (c - nleft));
                    T* pk = p.A + (long long)col * p.lda + k;
                    T* pp = p.A + (long long)col * p.lda + piv;
                    T vk = *pk, vp = *pp;
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

#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (j < wc) {
            const int par = j & 1;
            const int want = p.epoch + j + 1;
            // 1. local candidate: strict '>' from amax = 0, rows ascending
            T best = T(0);
            int bi = INT_MAX;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (ri[q] >= j && ri[q] < p.m) {
                    T v = tabs(a[q][j]);
                    if (v > best) { best = v; bi = ri[q]; }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                T ob = shfl_xor(best, off);
                int oi = shfl_xor(bi, off);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
            __syncthreads();
            T cb = s_val[0];
            int ci = s_idx[0];
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                T ob = s_val[w];
                int oi = s_idx[w];
                if (ob > cb || (ob == cb && oi < ci)) { cb = ob; ci = oi; }
            }
            // 2. publish candidate row (+ the current top row from its owner)
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (ri[q] == ci) {
                    T* dst = p.rowbuf + ((long long)par * PANEL_GMAX + cta) * PANEL_WMAX;
#pragma unroll
                    for (int c = 0; c < W; ++c) dst[c] = a[q][c];
                    __threadfence();
                }
                if (ri[q] == j) {
                    T* dst = p.toprow + par * PANEL_WMAX;
#pragma unroll
                    for (int c = 0; c < W; ++c) dst[c] = a[q][c];
                    __threadfence();
                }
            }
            __syncthreads();
            if (tid == 0) {
                p.cand_val[par * PANEL_GMAX + cta] = cb;
                p.cand_idx[par * PANEL_GMAX + cta] = ci;
                __threadfence();
                st_release(p.flags + par * PANEL_GMAX + cta, want);
            }
            // 3. warp 0: acquire all candidates, pick the pivot, stage rows
            if (warp == 0) {
                T gv = T(0);
                int gi = INT_MAX;
                int gc = -1;
                bool dead = false;
                for (int c = lane; c < G; c += 32) {
                    long long t0 = clock64();
                    while (ld_acquire(p.flags + par * PANEL_GMAX + c) != want) {
                        if (clock64() - t0 > kSpinTimeoutCycles) { dead = true; break; }
                    }
                    T v = ld_cg(p.cand_val + par * PANEL_GMAX + c);
                    int idx = ld_cg(p.cand_idx + par * PANEL_GMAX + c);
                    if (v > gv || (v == gv && idx < gi)) { gv = v; gi = idx; gc = c; }
                }
                if (__any_sync(0xffffffffu, dead)) {
                    if (lane == 0) { atomicExch(p.deverr, DEV_ERR_PANEL_TIMEOUT); s_abort = 1; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    T ob = shfl_xor(gv, off);
                    int oi = shfl_xor(gi, off);
                    int oc = shfl_xor(gc, off);
                    if (ob > gv || (ob == gv && oi < gi)) { gv = ob; gi = oi; gc = oc; }
                }
                const bool none = !(gv > T(0));  // all-zero (or all-NaN) subcolumn: kp = k
                if (lane < W) {
                    T tr = ld_cg(p.toprow + par * PANEL_WMAX + lane);
                    T pr = none ? tr
                                : ld_cg(p.rowbuf + ((long long)par * PANEL_GMAX + gc) * PANEL_WMAX + lane);
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
                if (pv == T(0) && *p.info == 0) *p.info = p.j0 + j + 1;
                __threadfence();
                st_release(p.progress, want);
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
