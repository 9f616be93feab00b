// gemm_tc.cuh — FP32 trailing-matrix update C -= A * B on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32 issued by one thread, operands staged in shared memory by TMA
// (cp.async.bulk.tensor, 128-byte swizzle), accumulator in TMEM, read back with tcgen05.ld.
// (reference `_blocked_lu_schur!`, src/blocked_lufact.jl:186-620 — the FP32 instantiation that
// the *32MixedLUFactorization algorithms factor with, src/openblas.jl:470-543.)
//
// FP32 accuracy from TF32 tensor cores: error-compensated 3xTF32.  Every operand x is split
// once, in a bandwidth-trivial pre-pass over the panel, into hi = rn_tf32(x) and
// lo = rn_tf32(x - hi) (both exactly representable, so the tensor core's truncation to TF32 is
// exact), and the product is accumulated as  A_hi*B_lo + A_lo*B_hi + A_hi*B_hi  in the FP32
// TMEM accumulator (the dropped lo*lo term is 2^-22 relative).
//
// Operand layouts: both operands are fed K-major (K contiguous), the canonical UMMA layout.
//   B = U12 (K x N, K contiguous in the column-major factor matrix) is K-major as it lies;
//   A = L21 (M x K, M contiguous) is TRANSPOSED by the split pre-pass (measured: an MN-major
//       32-bit operand under the plain 128-byte swizzle yields zeros — it needs the 32-byte-atom
//       swizzle variant — so the pre-pass that splits anyway also transposes).
//   One TMA box {32 (k), rows} per operand lands as `rows` rows of 128 bytes (SBO = 1024 bytes
//   between groups of 8 rows); a UMMA K-step (8 TF32) advances the start address by 32 bytes.
// CTA tile 128 x 256 (UMMA M = 128, N = 256, K = 8 per instruction), BK = 32, two stages of
// 96 KB; persistent CTAs (one per SM) with TWO TMEM accumulators so the HBM-bound epilogue of a
// tile overlaps the main loop of the next; warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer, warps 2.. = epilogue (TMEM -> registers, C += D with D = (-A) B by the instruction
// descriptor's negate bit).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encode entry point is fetched at run time)
#include "common.cuh"

namespace b200lu {

// BK = 16 (64-byte rows, 64-byte swizzle) x 4 stages of 48 KB: the producer runs three stages
// (~1.2 us of tensor work) ahead of the MMA issuer — with BK = 32 x 2 stages of 96 KB the ring was one
// stage deep and the tensor pipe idled on TMA latency (ncu: 35 % active).
#ifndef TC_BK_CFG
#define TC_BK_CFG 16
#endif
constexpr int TC_BM = 128, TC_BN = 256, TC_BK = TC_BK_CFG, TC_STAGES = (TC_BK_CFG == 16 ? 4 : 2);
constexpr unsigned TC_SWIZZLE_BYTES = TC_BK * 4;                 // 64 or 128: one K-row of a stage
constexpr unsigned TC_SBO = 8 * TC_SWIZZLE_BYTES;                // 8 rows of the swizzle atom
constexpr unsigned long long TC_LAYOUT = (TC_BK_CFG == 16 ? 4ull : 2ull);   // UMMA layout type: SWIZZLE_64B / SWIZZLE_128B
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;              // 16 KB: one of A_hi / A_lo
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;              // 32 KB: one of B_hi / B_lo
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 96 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;   // 197,888 B: one CTA per SM
constexpr int TC_EPI_WARPS = 4;                 // one warp per TMEM lane quarter (8 warps measured: no gain once C is prefetched into L2)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

// hi = rn_tf32(x), lo = rn_tf32(x - hi) for a rows x cols column-major block
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, long long lds,
                                                         float* __restrict__ hi, float* __restrict__ lo,
                                                         long long ldd, int rows, int cols) {
    const int r4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (r4 >= rows) return;
    // columns are strided over gridDim.y: the launcher clamps it to the 65535 limit of the y dimension
    for (int c = blockIdx.y; c < cols; c += gridDim.y) {
    const float* s = src + (long long)c * lds + r4;
    float x[4];
    if (r4 + 3 < rows && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
        const float4 v = *reinterpret_cast<const float4*>(s);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = (r4 + i < rows) ? s[i] : 0.f;
    }
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[i]));
        h[i] = __uint_as_float(hb);
        const float d = x[i] - h[i];
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d));
        l[i] = __uint_as_float(lb);
    }
    float* ph = hi + (long long)c * ldd + r4;
    float* pl = lo + (long long)c * ldd + r4;
    if (r4 + 3 < rows) {   // ldd is a multiple of 4 and the scratch base is 16-byte aligned
        *reinterpret_cast<float4*>(ph) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(pl) = make_float4(l[0], l[1], l[2], l[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r4 + i < rows) { ph[i] = h[i]; pl[i] = l[i]; }
    }
    }
}

// transposing variant for A: src is rows x cols column-major (rows = M contiguous); hiT/loT are
// [rows][ldt] with the `cols` (K) values of one row contiguous
__global__ void __launch_bounds__(256) split_tf32_transpose_kernel(const float* __restrict__ src, long long lds,
                                                                   float* __restrict__ hiT, float* __restrict__ loT,
                                                                   long long ldt, int rows, int cols) {
    __shared__ float th[32][33], tl[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        const float x = (r < rows && c < cols) ? src[(long long)c * lds + r] : 0.f;
        unsigned hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        const float h = __uint_as_float(hb);
        const float d = x - h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d));
        th[ty + 8 * i][tx] = h;
        tl[ty + 8 * i][tx] = __uint_as_float(lb);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < cols) {
            hiT[(long long)r * ldt + c] = th[tx][ty + 8 * i];
            loT[(long long)r * ldt + c] = tl[tx][ty + 8 * i];
        }
    }
}

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ unsigned tc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// returns false on watchdog timeout (a lost TMA / MMA completion would otherwise hang the GPU)
__device__ __forceinline__ bool tc_mbar_wait(unsigned bar, unsigned parity) {
    if (tc_mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!tc_mbar_try_wait(bar, parity))
        if (clock64() - t0 > kSpinTimeoutCycles) return false;
    return true;
}
__device__ __forceinline__ void tc_tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle, Blackwell version field = 1
__device__ __forceinline__ unsigned long long tc_smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
           ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46) | (TC_LAYOUT << 61);
}

struct TcGemmParams {
    float* C;
    long long ldc;
    int M, N, K;
    int tiles_per_cta;   // consecutive tiles per CTA (see the kernel comment)
    int* deverr;
};

// Semi-persistent: CTA b works on `tiles_per_cta` consecutive tiles (M-fastest order, so neighbours
// share the B strip in L2) and then EXITS — a fully persistent grid would hold every SM for the whole
// update and starve the look-ahead panel kernels of the high-priority stream (measured).  Two TMEM accumulators (2 x 256 columns):
// the epilogue of tile i (TMEM -> registers -> C, HBM-bound: 128 KB read + 128 KB written) runs
// under the main loop of tile i + 1.  The shared-memory ring runs continuously across tiles.
__global__ void __launch_bounds__(TC_THREADS, 1)
sgemm3x_tc_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                  TcGemmParams p) {
    extern __shared__ unsigned char tc_smem_raw[];
    const unsigned raw = tc_smem_u32(tc_smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;            // 128-byte swizzle atoms need 1024-byte alignment
    unsigned char* gen = tc_smem_raw + (base - raw);
    // [stage][A_hi | A_lo | B_hi | B_lo] ... then the barriers
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(gen + TC_STAGES * TC_STAGE_BYTES);
    const unsigned bar_full = tc_smem_u32(&bars[0]);              // [TC_STAGES] TMA -> MMA
    const unsigned bar_empty = tc_smem_u32(&bars[TC_STAGES]);     // [TC_STAGES] MMA -> TMA
    const unsigned bar_accf = tc_smem_u32(&bars[2 * TC_STAGES]);  // [2] MMA -> epilogue (accumulator full)
    const unsigned bar_acce = tc_smem_u32(&bars[2 * TC_STAGES + 2]);  // [2] epilogue -> MMA (accumulator drained)
    unsigned* s_tmem = reinterpret_cast<unsigned*>(&bars[2 * TC_STAGES + 4]);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = (p.K + TC_BK - 1) / TC_BK;
    const int tiles_m = (p.M + TC_BM - 1) / TC_BM, tiles_n = (p.N + TC_BN - 1) / TC_BN;
    const int ntiles = tiles_m * tiles_n;
    const int tile0 = blockIdx.x * p.tiles_per_cta, tile1 = min(tile0 + p.tiles_per_cta, ntiles);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc_mbar_init(bar_full + 8 * s, 1);
            tc_mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc_mbar_init(bar_accf + 8 * a, 1);
            tc_mbar_init(bar_acce + 8 * a, TC_EPI_WARPS);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tc_prefetch_tensormap(&tmAhi);
        tc_prefetch_tensormap(&tmAlo);
        tc_prefetch_tensormap(&tmBhi);
        tc_prefetch_tensormap(&tmBlo);
    }
    if (warp == 1) {   // TMEM: all 512 columns = two 128 x 256 FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *s_tmem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;   // running k-block index across tiles
            bool ok = true;
            for (int tile = tile0; tile < tile1 && ok; ++tile) {
                const int m0 = (tile % tiles_m) * TC_BM, n0 = (tile / tiles_m) * TC_BN;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const unsigned ph = (unsigned)(it / TC_STAGES) & 1u;
                    ok = tc_mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    if (!ok) { atomicExch(p.deverr, DEV_ERR_GEMM_TIMEOUT); break; }
                    const unsigned st = base + s * TC_STAGE_BYTES;
                    tc_mbar_expect_tx(bar_full + 8 * s, (unsigned)TC_STAGE_BYTES);
                    const int k0 = kb * TC_BK;
                    tc_tma_load_2d(st, &tmAhi, k0, m0, bar_full + 8 * s);
                    tc_tma_load_2d(st + TC_A_BYTES, &tmAlo, k0, m0, bar_full + 8 * s);
                    tc_tma_load_2d(st + 2 * TC_A_BYTES, &tmBhi, k0, n0, bar_full + 8 * s);
                    tc_tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tmBlo, k0, n0, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, A negated, both K-major, N = 256, M = 128
            constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) |
                                       ((unsigned)(TC_BN >> 3) << 17) | ((unsigned)(TC_BM >> 4) << 24);
            int it = 0, j = 0;
            bool ok = true;
            for (int tile = tile0; tile < tile1 && ok; ++tile, ++j) {
                const int acc = j & 1;
                const unsigned aph = (unsigned)(j >> 1) & 1u;
                // the epilogue must have drained this accumulator (first use: passes immediately)
                ok = tc_mbar_wait(bar_acce + 8 * acc, aph ^ 1u);
                if (!ok) { atomicExch(p.deverr, DEV_ERR_GEMM_TIMEOUT); break; }
                tc_fence_after();
                const unsigned tacc = tmem + (unsigned)(acc * TC_BN);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const unsigned ph = (unsigned)(it / TC_STAGES) & 1u;
                    ok = tc_mbar_wait(bar_full + 8 * s, ph);
                    if (!ok) { atomicExch(p.deverr, DEV_ERR_GEMM_TIMEOUT); break; }
                    tc_fence_after();
                    const unsigned st = base + s * TC_STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        const unsigned long long a_hi = tc_smem_desc(st + ks * 32, 16, TC_SBO);
                        const unsigned long long a_lo = tc_smem_desc(st + TC_A_BYTES + ks * 32, 16, TC_SBO);
                        const unsigned long long b_hi = tc_smem_desc(st + 2 * TC_A_BYTES + ks * 32, 16, TC_SBO);
                        const unsigned long long b_lo = tc_smem_desc(st + 2 * TC_A_BYTES + TC_B_BYTES + ks * 32, 16, TC_SBO);
                        tc_mma_tf32(tacc, a_hi, b_lo, IDESC, (kb | ks) != 0 ? 1u : 0u);
                        tc_mma_tf32(tacc, a_lo, b_hi, IDESC, 1u);
                        tc_mma_tf32(tacc, a_hi, b_hi, IDESC, 1u);
                    }
                    tc_commit(bar_empty + 8 * s);   // frees the stage when these MMAs have read it
                }
                if (ok) tc_commit(bar_accf + 8 * acc);   // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2.., TMEM lane quarter = warp % 4, column slice = (warp - 2) / 4 =====
        const int q = warp & 3;
        constexpr int CPW = TC_BN / (TC_EPI_WARPS / 4);   // columns per epilogue warp
        const int cbase = ((warp - 2) >> 2) * CPW;
        int j = 0;
        for (int tile = tile0; tile < tile1; ++tile, ++j) {
            const int m0 = (tile % tiles_m) * TC_BM, n0 = (tile / tiles_m) * TC_BN;
            const int acc = j & 1;
            const unsigned aph = (unsigned)(j >> 1) & 1u;
            const int row = m0 + 32 * q + lane;
            // The epilogue is latency-bound (32 loads in flight per thread, 8 dependent chunks per tile:
            // ~12 us against a 9 us main loop).  While the main loop of THIS tile runs, pull the warp's
            // 256 lines of C (one 128-byte line per column) into L2.
            if (m0 + 32 * q < p.M) {
#pragma unroll
                for (int jj = 0; jj < CPW / 32; ++jj) {
                    const int col = n0 + cbase + lane + 32 * jj;
                    if (col < p.N) {
                        const float* pf = p.C + (long long)col * p.ldc + (m0 + 32 * q);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                    }
                }
            }
            const bool ok = tc_mbar_wait(bar_accf + 8 * acc, aph);
            if (!ok) { if (lane == 0) atomicExch(p.deverr, DEV_ERR_GEMM_TIMEOUT); break; }
            tc_fence_after();
#pragma unroll 1
            for (int c = cbase / 32; c < (cbase + CPW) / 32; ++c) {
                float cv[32];
                float* cp = p.C + (long long)(n0 + c * 32) * p.ldc + row;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    cv[jj] = (row < p.M && n0 + c * 32 + jj < p.N) ? cp[(long long)jj * p.ldc] : 0.f;
                unsigned v[32];
                const unsigned taddr = tmem + ((unsigned)(32 * q) << 16) + (unsigned)(acc * TC_BN + c * 32);
                // load + wait in ONE asm statement: nothing may read v[] before tcgen05.wait::ld
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    "tcgen05.wait::ld.sync.aligned;"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    if (row < p.M && n0 + c * 32 + jj < p.N) cp[(long long)jj * p.ldc] = cv[jj] + __uint_as_float(v[jj]);
            }
            // this warp has read its lanes of the accumulator: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acce + 8 * acc) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace b200lu
