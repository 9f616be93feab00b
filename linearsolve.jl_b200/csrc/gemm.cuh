// gemm.cuh — trailing-matrix (Schur) update C -= A * B (reference
// `_blocked_lu_schur!`, src/blocked_lufact.jl:186-620,658-679: "~all the flops
// at mid/large N live here").
//
// FP64: sm_100 has no tcgen05 kind for f64; the FP64 tensor path is the
// warp-level DMMA (SASS DMMA.8x8x4).  The kernel is a multi-stage cp.async
// (LDGSTS, 16-byte, zero-fill predicated) shared-memory pipeline feeding
// m8n8k4 DMMAs with register accumulators; C is read once into the
// accumulators and written once.  Operands are column-major as they lie in the
// factor matrix: A = L21 (m x k, m contiguous), B = U12 (k x n, k contiguous).
//
// FP32: register-tiled FFMA kernel (first version of the FP32-factor mode).
#pragma once
#include "common.cuh"

namespace b200lu {

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
struct DgemmCfg {
    static constexpr int NT = WARPS_M * WARPS_N * 32;
    static constexpr int WTM = BM / WARPS_M;
    static constexpr int WTN = BN / WARPS_N;
    static constexpr int MI = WTM / 8;
    static constexpr int NI = WTN / 8;
    static constexpr int LDAS = BM + 4;  // (LDAS mod 16) == 4: conflict-free a-fragment loads
    static constexpr int LDBS = BK + 4;  // (LDBS mod 16) == 4: conflict-free b-fragment loads
    static constexpr int A_STAGE = BK * LDAS;
    static constexpr int B_STAGE = BN * LDBS;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    static_assert(BM % 16 == 0 && BK % 16 == 0, "tile shape");
};

// C[M x N] -= A[M x K] * B[K x N].  All pointers 16-byte aligned, lda/ldb even.
// K may be any value (zero-filled to a multiple of BK; K must be even).
template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, int MINB>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB)
    dgemm_sub_kernel(int M, int N, int K, const double* __restrict__ A, long long lda,
                     const double* __restrict__ B, long long ldb, double* __restrict__ C,
                     long long ldc, int tiles_m, int tiles_n, int group_n) {
    using Cfg = DgemmCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    constexpr int NT = Cfg::NT, MI = Cfg::MI, NI = Cfg::NI;
    constexpr int LDAS = Cfg::LDAS, LDBS = Cfg::LDBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + (size_t)STAGES * Cfg::A_STAGE;

    // tile rasterisation: groups of `group_n` tile-columns, tile-rows inside a
    // group vary slowest so concurrently resident CTAs share A and B strips in L2.
    int tm, tn;
    {
        const int t = blockIdx.x;
        const int per_group = tiles_m * group_n;
        const int g = t / per_group;
        const int r = t - g * per_group;
        const int gw = min(group_n, tiles_n - g * group_n);
        tm = r / gw;
        tn = g * group_n + (r - tm * gw);
    }
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lr = lane >> 2, lc = lane & 3;  // T/4, T%4

    const int KT = (K + BK - 1) / BK;

    auto load_tile = [&](int kt, int stage) {
        const int k0 = kt * BK;
        double* as = As + (size_t)stage * Cfg::A_STAGE;
        double* bs = Bs + (size_t)stage * Cfg::B_STAGE;
        // A tile: BK columns of BM rows; 16-byte chunks along m
        constexpr int A_CH = BK * (BM / 2);
#pragma unroll
        for (int i = 0; i < (A_CH + NT - 1) / NT; ++i) {
            const int ch = tid + i * NT;
            if (A_CH % NT == 0 || ch < A_CH) {
                const int kk = ch / (BM / 2);
                const int mm = (ch - kk * (BM / 2)) * 2;
                const bool ok = (m0 + mm < M) && (k0 + kk < K);
                const double* g = A + (long long)(k0 + kk) * lda + (m0 + mm);
                cp_async16(as + kk * LDAS + mm, ok ? g : A, ok);
            }
        }
        // B tile: BN columns of BK rows; 16-byte chunks along k
        constexpr int B_CH = BN * (BK / 2);
#pragma unroll
        for (int i = 0; i < (B_CH + NT - 1) / NT; ++i) {
            const int ch = tid + i * NT;
            if (B_CH % NT == 0 || ch < B_CH) {
                const int nn = ch / (BK / 2);
                const int kk = (ch - nn * (BK / 2)) * 2;
                const bool ok = (n0 + nn < N) && (k0 + kk < K);
                const double* g = B + (long long)(n0 + nn) * ldb + (k0 + kk);
                cp_async16(bs + nn * LDBS + kk, ok ? g : B, ok);
            }
        }
    };

    // prologue: start the pipeline, then pull C into the accumulators
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int row = m0 + wm * Cfg::WTM + mi * 8 + lr;
            const int col = n0 + wn * Cfg::WTN + ni * 8 + 2 * lc;
            const double* cp = C + (long long)col * ldc + row;
            acc[mi][ni][0] = (row < M && col < N) ? cp[0] : 0.0;
            acc[mi][ni][1] = (row < M && col + 1 < N) ? cp[ldc] : 0.0;
        }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_tile(nk, nk % STAGES);
            cp_async_commit();
        }
        const double* as = As + (size_t)(kt % STAGES) * Cfg::A_STAGE + wm * Cfg::WTM + lr;
        const double* bs = Bs + (size_t)(kt % STAGES) * Cfg::B_STAGE + (wn * Cfg::WTN + lr) * LDBS;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) a[mi] = as[(ks * 4 + lc) * LDAS + mi * 8];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) b[ni] = neg_bits(bs[ni * 8 * LDBS + ks * 4 + lc]);
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int row = m0 + wm * Cfg::WTM + mi * 8 + lr;
            const int col = n0 + wn * Cfg::WTN + ni * 8 + 2 * lc;
            double* cp = C + (long long)col * ldc + row;
            if (row < M && col < N) cp[0] = acc[mi][ni][0];
            if (row < M && col + 1 < N) cp[ldc] = acc[mi][ni][1];
        }
}

// ---------------------------------------------------------------- FP32 ------
// C -= A*B, 128x128x8 tiles, 256 threads, 8x8 register tile per thread.
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) sgemm_sub_kernel(int M, int N, int K,
                                                        const float* __restrict__ A, long long lda,
                                                        const float* __restrict__ B, long long ldb,
                                                        float* __restrict__ C, long long ldc) {
    static_assert(BM == 128 && BN == 128 && BK == 8, "fixed tile");
    __shared__ float As[2][BK][BM];
    __shared__ float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; rows tx*4 (+64), cols ty*4 (+64)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loader mapping: A tile 8 x 128 -> thread loads 4 consecutive rows of one k
    const int a_k = tid >> 5, a_m = (tid & 31) * 4;
    // B tile 8 x 128 (k contiguous in memory): thread loads k = tid&7, 4 columns
    const int b_k = tid & 7, b_n = (tid >> 3) * 4;

    const int KT = (K + BK - 1) / BK;
    float ra[4], rb[4];
    auto gload = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + a_m + i, k = k0 + a_k;
            ra[i] = (m < M && k < K) ? A[(long long)k * lda + m] : 0.f;
            const int n = n0 + b_n + i, kb = k0 + b_k;
            rb[i] = (n < N && kb < K) ? B[(long long)n * ldb + kb] : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[buf][a_k][a_m + i] = ra[i];
            Bs[buf][b_k][b_n + i] = rb[i];
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = As[buf][k][tx * 4 + i];
                a[4 + i] = As[buf][k][64 + tx * 4 + i];
                b[i] = Bs[buf][k][ty * 4 + i];
                b[4 + i] = Bs[buf][k][64 + ty * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (col >= N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
            if (row < M) C[(long long)col * ldc + row] -= acc[i][j];
        }
    }
}

}  // namespace b200lu
