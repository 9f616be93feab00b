// b200lu.cu — host driver and C ABI of libb200lu.so (see include/b200lu.h).
//
// getrf driver: right-looking blocked LU with a recursive panel and one-panel
// look-ahead on a high-priority stream — the GPU restatement of the
// reference's `_blocked_lufact!` (src/blocked_lufact.jl:658-679): panel ->
// laswp (left + right) -> unit-lower TRSM -> Schur GEMM.  Everything is
// stream-ordered; the host never synchronises inside a factorization.
#include "../../include/b200lu.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "batched.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"
#include "laswp.cuh"
#include "panel.cuh"
#include "panel_cluster.cuh"
#include "trsm.cuh"
#include "trsv.cuh"
#include "trsv_cluster.cuh"
#include "util_kernels.cuh"

namespace b200lu {
std::atomic<unsigned long long> g_launch_count{0};
}
using namespace b200lu;

#define B200LU_VERSION 300   // 0.3.0: ngpus > 1 in b200lu_create, fused batched factor+solve, dist transport query, HOST_REGISTER

// ------------------------------------------------------------------ handle --
struct b200lu_handle {
    int dtype = 0;
    int dev = 0;
    int sms = 148;             // multiprocessors of `dev`
    cudaStream_t s_main = nullptr, s_panel = nullptr, s_copy = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_next = nullptr, ev_h2d = nullptr;
    std::vector<cudaEvent_t> ev_panel;
    // streamed upload of A (B200LU_OPT_STREAM_H2D): column chunks in flight on s_copy
    std::vector<cudaEvent_t> ev_chunk;   // chunk c has landed
    std::vector<int> chunk_end;          // one past its last column (multiples of nb; last = n)
    void* reg_ptr = nullptr;             // B200LU_OPT_HOST_REGISTER: the caller's matrix this handle page-locked
    size_t reg_bytes = 0;
    char err[512] = {0};
    double timing[B200LU_T_COUNT] = {0};
    int64_t opt[B200LU_OPT_COUNT];
    bool nb_user = false;                // B200LU_OPT_NB was set by the caller (no automatic choice by GPU count)

    // single large system
    int64_t n = 0, ldd = 0, cap_n = 0, cap_meta = 0;
    void* dA = nullptr;        // factors (double for F64, float for F32/MIXED)
    double* dA64 = nullptr;    // MIXED: FP64 copy of A for residuals
    void* dA_keep = nullptr;   // B200LU_OPT_KEEP_A: copy of A in the interface type, for b200lu_residual_norms
    int64_t cap_keep = 0;      // bytes
    bool keep_valid = false;
    int* d_ipiv = nullptr;
    int* d_perm = nullptr;
    int* d_info = nullptr;
    int* d_deverr = nullptr;
    LaswpPlan* d_plans = nullptr;
    int cap_plans = 0;
    void* d_panelsync = nullptr;
    float* d_split = nullptr;      // tcgen05 SGEMM: hi/lo TF32 splits of the current L21 and U12
    int64_t cap_split_n = 0;
    long long* d_pdbg = nullptr;   // B200LU_PANEL_DBG=1: clock64 stamps of the cluster panel launches
    int pdbg_n = 0;
    unsigned panel_epoch = 0;
    bool factored = false;
    int64_t info = 0;
    // solve state
    bool solve_ready = false;
    void* d_dinvL = nullptr;
    void* d_dinvU = nullptr;
    void* d_wL = nullptr;   // coupling blocks dinv_r * A[r, r-1]
    void* d_wU = nullptr;   // coupling blocks dinv_r * A[r, r+1]
    int* d_tflags = nullptr;
    int* d_tticket = nullptr;
    int cap_tgroups = 0;
    size_t cap_tflag_bytes = 0;
    unsigned trsv_epoch = 0;
    void* d_wLt = nullptr;  // transposed solves: coupling blocks of U^T (first sweep) ...
    void* d_wUt = nullptr;  // ... and of L^T (second sweep)
    bool solve_ready_t = false;
    // single-RHS TRSV v2 (2-D work items)
    Trsv2Item* d_t2items = nullptr;
    unsigned long long* d_t2x = nullptr;
    unsigned long long* d_t2p = nullptr;
    int* d_t2ticket = nullptr;
    int t2_nblk = 0, t2_nitems = 0, t2_kmax = 0, t2_grid = 0;
    unsigned t2_epoch = 0;
    // single-RHS TRSV v3 (chain in one cluster over DSMEM, workers on the far blocks)
    Trsv2Item* d_t3items = nullptr;
    unsigned long long* d_t3x = nullptr;
    unsigned long long* d_t3p = nullptr;
    int* d_t3ticket = nullptr;
    int t3_nblk = 0, t3_nitems = 0, t3_kmax = 0, t3_clusters = 0;
    long long* d_t3dbg = nullptr;
    int t3_ok = -1;            // -1 not probed, 0 cluster launch not possible (v2 is used), 1 ok
    unsigned t3_epoch = 0;
    void* d_B = nullptr;       // staging for host solves / permuted rhs
    void* d_X = nullptr;
    int64_t cap_rhs = 0;
    // refinement scratch (MIXED)
    double* d_r = nullptr;     // residual (n)
    float* d_r32 = nullptr;    // residual cast / correction (n)
    double* d_scal = nullptr;  // [4] norms
    double* d_rpart = nullptr; // partial sums of the deterministic residual / norm reductions (ensure_rpart)
    int64_t cap_rpart = 0;
    double* d_cscal = nullptr; // [2 * cap_cscal] per-column norms of the matrix-RHS refinement
    double* h_cscal = nullptr;
    int cap_cscal = 0;
    double normA_F = 0.0;
    int last_refine_iters = 0;
    // page-locked, device-mapped staging of one right-hand side and its solution (2 * cap_hrhs * 8 bytes)
    char* h_rhs = nullptr;
    char* d_rhs_map = nullptr;
    int64_t cap_hrhs = 0;
    // pinned host staging
    long long* h_ipiv = nullptr;
    int64_t cap_hipiv = 0;
    int* h_small = nullptr;    // [16] info/deverr/...
    double* h_scal = nullptr;  // [4]

    // batched
    int64_t b_batch = 0, b_n = 0, b_cap_bytes = 0, b_cap_batch_n = 0;
    void* dB_LU = nullptr;
    int* dB_ipiv = nullptr;
    int* dB_info = nullptr;
    int* dB_perm = nullptr;    // batched: original row that ends at each position (for getrs)
    void* dB_in = nullptr;     // host-path staging for A
    int64_t b_cap_in = 0;
    void* dB_rhs = nullptr;
    void* dB_x = nullptr;
    int64_t b_cap_rhs = 0;
    int* hB_stage = nullptr;   // pinned staging of batched pivots / info on their way to the host
    int64_t hb_cap = 0;        // ints
    bool b_factored = false;

    // profiling (B200LU_OPT_PROFILE)
    std::vector<cudaEvent_t> prof_ev;
    int prof_used = 0;
    // chain profile of the distributed getrf (B200LU_OPT_PROFILE): event pairs on the panel stream, by phase
    std::vector<cudaEvent_t> chain_ev;
    std::vector<int> chain_kind;
    double prof_flops = 0.0;
    double counters[B200LU_C_COUNT] = {0};

    // distributed (filled by dist.inc)
    void* comm = nullptr;  // DistState*: this handle is one rank of a multi-GPU factorization
    int rank = 0, nranks = 1;
    bool factored_dist = false;
    void* team = nullptr;  // Team*: this handle was created with ngpus > 1 and drives one sub-handle per GPU
};

// multi-GPU (dist.inc, included at the end of this file)
static int dist_destroy(b200lu_handle* h);
static void team_destroy(b200lu_handle* h);
static int team_create(b200lu_handle** out, int dtype, int ngpus, const int* devices);
static int team_set_option(b200lu_handle* h, int option, int64_t value);
static int team_factor(b200lu_handle* h, int64_t n, const void* A_host, int64_t lda, int64_t* ipiv_out, int64_t* info);
static int team_solve(b200lu_handle* h, char trans, int64_t nrhs, const void* B_host, int64_t ldb, void* X_host, int64_t ldx);
static int team_get_factors(b200lu_handle* h, void* LU_host, int64_t ldlu);
static int team_factor_batched(b200lu_handle* h, int64_t batch, int64_t n, const void* A_host, int64_t lda,
                               int64_t strideA, int64_t* ipiv_out, int64_t* info_out);
static int team_solve_batched(b200lu_handle* h, bool is_trans_call, char trans, int64_t nrhs, const void* B_host, int64_t ldb,
                              int64_t strideB, void* X_host, int64_t ldx, int64_t strideX);
static int team_get_factors_batched(b200lu_handle* h, void* LU_host, int64_t lda, int64_t strideA, int64_t* ipiv_out,
                                    int64_t* info_out);
static int team_factor_solve_batched(b200lu_handle* h, int64_t batch, int64_t n, const void* A_host, int64_t lda, int64_t strideA,
                                     const void* B_host, int64_t strideB, void* X_host, int64_t strideX, int64_t* ipiv_out,
                                     int64_t* info_out);
static b200lu_handle* team_first(b200lu_handle* h);
#define TEAM_ONLY_HOST(h)                                                                                     \
    if ((h) && (h)->team)                                                                                     \
        return set_err((h), -1, "a multi-GPU handle (ngpus > 1) takes HOST matrices: use the entry points without _device")


static int set_err(b200lu_handle* h, int status, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return status;
}

#define CU_TRY(h, expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return set_err((h), 1, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                           __LINE__, cudaGetErrorString(_e));                                  \
    } while (0)

#define LAUNCH_CHECK(h)                                                                        \
    do {                                                                                       \
        B200LU_COUNT_LAUNCH();                                                                 \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return set_err((h), 1, "kernel launch failed %s at %s:%d: %s", cudaGetErrorName(_e), \
                           __FILE__, __LINE__, cudaGetErrorString(_e));                        \
    } while (0)

static size_t elem_size(const b200lu_handle* h) { return h->dtype == B200LU_F64 ? 8 : 4; }
static size_t iface_size(const b200lu_handle* h) { return h->dtype == B200LU_F32 ? 4 : 8; }



template <typename P>
static void free_dev(P*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// columns per CTA of residual_gemv_kernel: at most 32 chunks of partial sums per row
static int residual_chunk(int64_t n) { return (int)std::max<int64_t>(512, (cdiv(n, 32) + 3) / 4 * 4); }
// scratch of the fixed-order reductions: 32 partial sums per row (residual), one per CTA (Frobenius norm)
static int ensure_rpart(b200lu_handle* h, int64_t n) {
    const int64_t need = std::max<int64_t>(32 * n, (int64_t)cdiv(n, 256) * 256);
    if (need <= h->cap_rpart) return 0;
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    free_dev(h->d_rpart);
    h->cap_rpart = 0;
    CU_TRY(h, cudaMalloc((void**)&h->d_rpart, (size_t)need * sizeof(double)));
    h->cap_rpart = need;
    return 0;
}

// ------------------------------------------------------------ kernel launch --
constexpr int BASE_W = 32;    // base panel width (columns held in registers)
constexpr int BASE_NT = 256;  // threads per panel CTA

template <typename T>
static int launch_panel_base(b200lu_handle* h, cudaStream_t st, T* A, int64_t lda, int nrows,
                             int j0, int wc, int pc0, int pc1) {
    PanelArgs<T> p;
    p.A = A;
    p.lda = lda;
    p.j0 = j0;
    p.m = nrows - j0;
    p.wc = wc;
    p.pc0 = pc0;
    p.pc1 = pc1;
    p.ipiv = h->d_ipiv;
    p.info = h->d_info;
    // rows per thread: 1 keeps the block spill-free and halves the per-column work of a
    // thread; 2 halves the number of cooperating CTAs (and mailbox traffic) for tall panels
    int rpt = (int)h->opt[B200LU_OPT_PANEL_RPT];
    if (rpt == 0) rpt = (p.m <= 16384) ? 1 : 2;
    const int gmax = (int)std::min<int64_t>(PANEL_GMAX, h->opt[B200LU_OPT_PANEL_CTAS]);
    if (rpt == 1 && cdiv(p.m, BASE_NT) > gmax) rpt = 2;
    const int G = cdiv(p.m, BASE_NT * rpt);
    if (G > gmax)
        return set_err(h, 2, "panel of %d rows needs %d CTAs > limit %d", p.m, G, gmax);
    p.G = G;
    if (h->panel_epoch > (1u << 30)) {
        CU_TRY(h, cudaMemsetAsync(h->d_panelsync, 0, sizeof(PanelMail), st));
        h->panel_epoch = 0;
    }
    p.epoch = h->panel_epoch;
    h->panel_epoch += BASE_W;
    p.mail = (PanelMail*)h->d_panelsync;
    p.deverr = h->d_deverr;
    p.dbg = nullptr;
    const int has_swapper = (pc1 - pc0) > wc ? 1 : 0;
    if (rpt == 1)
        panel_base_kernel<T, BASE_W, 1, BASE_NT><<<G + has_swapper, BASE_NT, 0, st>>>(p);
    else
        panel_base_kernel<T, BASE_W, 2, BASE_NT><<<G + has_swapper, BASE_NT, 0, st>>>(p);
    LAUNCH_CHECK(h);
    return 0;
}

template <typename T, int CC>
static int launch_trsm_cc(b200lu_handle* h, cudaStream_t st, const T* Lp, int64_t ldl, T* Bp,
                          int64_t ldb, int w, int ncols) {
    constexpr int NWARP = 8;
    const int grid = cdiv(ncols, CC * NWARP);
    const int rpl = cdiv(w, 32);
#define TRSM_CASE(R)                                                                             \
    {                                                                                            \
        const size_t smem = (size_t)32 * (R * 32) * sizeof(T);                                   \
        static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */                                                            \
        if (!attr_set) {                                                                         \
            CU_TRY(h, cudaFuncSetAttribute(trsm_lunit_kernel<T, R, CC, NWARP>,                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set = true;                                                                     \
        }                                                                                        \
        trsm_lunit_kernel<T, R, CC, NWARP><<<grid, NWARP * 32, smem, st>>>(Lp, ldl, Bp, ldb, w, ncols); \
    }
    if (rpl <= 1) TRSM_CASE(1)
    else if (rpl <= 2) TRSM_CASE(2)
    else if (rpl <= 4) TRSM_CASE(4)
    else if (rpl <= 8) TRSM_CASE(8)
    else return set_err(h, 2, "trsm block %d too wide", w);
#undef TRSM_CASE
    LAUNCH_CHECK(h);
    return 0;
}
// Few right-hand-side columns (the panel recursion, the look-ahead block): the k-chain latency
// is the cost, so one column per warp spreads it over 4x more CTAs; many columns: 4 per warp.
template <typename T>
static int launch_trsm(b200lu_handle* h, cudaStream_t st, const T* Lp, int64_t ldl, T* Bp,
                       int64_t ldb, int w, int ncols) {
    if (w <= 0 || ncols <= 0) return 0;
    if (ncols <= 1024) return launch_trsm_cc<T, 1>(h, st, Lp, ldl, Bp, ldb, w, ncols);
    return launch_trsm_cc<T, 4>(h, st, Lp, ldl, Bp, ldb, w, ncols);
}

// FP64 Schur update on DMMA.  Tile configurations (B200LU_OPT_GEMM_CFG):
//   0: 128x64 CTA, 4 warps of 64x32, 2 CTAs/SM   (fewest shared-memory reads per DMMA)
//   1: 128x64 CTA, 8 warps of 32x32, 2 CTAs/SM   (4 warps per scheduler: better latency hiding)
//   2: 128x128 CTA, 8 warps of 64x32, 1 CTA/SM   (half the L2->smem traffic per flop)
template <int BM, int BN, int WM, int WN, int STAGES, int MINB>
static int launch_dgemm_cfg(b200lu_handle* h, cudaStream_t st, int M, int N, int K, const double* A,
                            int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc) {
    using Cfg = DgemmCfg<BM, BN, 16, WM, WN, STAGES>;
    auto kern = dgemm_sub_kernel<BM, BN, 16, WM, WN, STAGES, MINB>;
    static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
    if (!attr_set) {
        CU_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    const int tm = cdiv(M, BM), tn = cdiv(N, BN);
    kern<<<tm * tn, Cfg::NT, Cfg::SMEM, st>>>(M, N, K, A, lda, B, ldb, C, ldc, tm, tn, 16);
    LAUNCH_CHECK(h);
    return 0;
}
static int launch_gemm(b200lu_handle* h, cudaStream_t st, int M, int N, int K, const double* A,
                       int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int cfg = (int)h->opt[B200LU_OPT_GEMM_CFG];
    if (cfg == 3) cfg = (h->n >= 12288) ? 1 : 0;   // auto: 8 warps of 32x32 win by ~1.3 % once the update dominates
    switch (cfg) {
        case 1: return launch_dgemm_cfg<128, 64, 4, 2, 3, 2>(h, st, M, N, K, A, lda, B, ldb, C, ldc);
        case 2: return launch_dgemm_cfg<128, 128, 2, 4, 4, 1>(h, st, M, N, K, A, lda, B, ldb, C, ldc);
        default: return launch_dgemm_cfg<128, 64, 2, 2, 3, 2>(h, st, M, N, K, A, lda, B, ldb, C, ldc);
    }
}
// ---- FP32: tcgen05 (TMA + TMEM) 3xTF32 kernel for the large updates, FFMA kernel otherwise ----
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_tmap_encode() {
    // function-local static: initialised once, thread-safe (handles may live on different host threads)
    static const PFN_tmapEncodeTiled fn = []() -> PFN_tmapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (PFN_tmapEncodeTiled)p;
        return nullptr;
    }();
    return fn;
}
// 2-D FP32 tensor map: dim0 = rows (contiguous), dim1 = cols (stride ld), 128-byte swizzle
static int make_tmap_2d(b200lu_handle* h, CUtensorMap* m, const float* base, int64_t rows, int64_t cols,
                        int64_t ld, int box0, int box1) {
    PFN_tmapEncodeTiled enc = get_tmap_encode();
    if (!enc) return set_err(h, 1, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, TC_BK == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(h, 1, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}
static int ensure_split(b200lu_handle* h, int64_t n) {
    if (n <= h->cap_split_n) return 0;
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    free_dev(h->d_split);
    const int64_t n32 = ((n + 31) / 32) * 32;
    // A_hi, A_lo: n32 x 256 each; B_hi, B_lo: 256 x n each
    CU_TRY(h, cudaMalloc((void**)&h->d_split, (size_t)(2 * n32 * 256 + 2 * 256 * n32) * sizeof(float)));
    h->cap_split_n = n;
    return 0;
}
static int launch_sgemm_tc(b200lu_handle* h, cudaStream_t st, int M, int N, int K, const float* A,
                           int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc) {
    const int64_t need = std::max(M, N);
    int rc = ensure_split(h, std::max<int64_t>(need, h->n));
    if (rc) return rc;
    const int64_t n32 = ((h->cap_split_n + 31) / 32) * 32;
    float* Ahi = h->d_split;
    float* Alo = Ahi + n32 * 256;
    float* Bhi = Alo + n32 * 256;
    float* Blo = Bhi + 256 * n32;
    const int64_t lst = 256;   // both splits are stored K-major with a fixed leading dimension of 256
    split_tf32_transpose_kernel<<<dim3(cdiv(M, 32), cdiv(K, 32)), 256, 0, st>>>(A, lda, Ahi, Alo, lst, M, K);
    LAUNCH_CHECK(h);
    split_tf32_kernel<<<dim3(cdiv(K, 1024), grid_y(N)), 256, 0, st>>>(B, ldb, Bhi, Blo, lst, K, N);
    LAUNCH_CHECK(h);
    CUtensorMap tAhi, tAlo, tBhi, tBlo;
    if ((rc = make_tmap_2d(h, &tAhi, Ahi, K, M, lst, TC_BK, TC_BM))) return rc;
    if ((rc = make_tmap_2d(h, &tAlo, Alo, K, M, lst, TC_BK, TC_BM))) return rc;
    if ((rc = make_tmap_2d(h, &tBhi, Bhi, K, N, lst, TC_BK, TC_BN))) return rc;
    if ((rc = make_tmap_2d(h, &tBlo, Blo, K, N, lst, TC_BK, TC_BN))) return rc;
    static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
    if (!attr_set) {
        CU_TRY(h, cudaFuncSetAttribute(sgemm3x_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set = true;
    }
    const int sms = h->sms;
    const int ntiles = cdiv(M, TC_BM) * cdiv(N, TC_BN);
    const int tpc = std::max(1, std::min(4, ntiles / (2 * sms)));
    TcGemmParams p{C, ldc, M, N, K, tpc, h->d_deverr};
    sgemm3x_tc_kernel<<<cdiv(ntiles, tpc), TC_THREADS, TC_SMEM_BYTES, st>>>(tAhi, tAlo, tBhi, tBlo, p);
    LAUNCH_CHECK(h);
    return 0;
}
static int launch_sgemm_ffma(b200lu_handle* h, cudaStream_t st, int M, int N, int K, const float* A,
                             int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc) {
    dim3 grid(cdiv(M, 128), cdiv(N, 128));
    sgemm_sub_kernel<128, 128, 8><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc);
    LAUNCH_CHECK(h);
    return 0;
}
static int launch_gemm(b200lu_handle* h, cudaStream_t st, int M, int N, int K, const float* A,
                       int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    // The tcgen05 path owns one scratch set: main-stream updates only (the panel recursion's small
    // GEMMs run concurrently on the panel stream); K <= 256 by construction (K = panel width).
    const int64_t mode = h->opt[B200LU_OPT_SGEMM_MODE];   // 0 auto, 1 FFMA, 2 tcgen05 whenever legal (tests)
    const bool tc = mode != 1 && st == h->s_main && (ldc % 4) == 0 &&
                    (mode == 2 || (int64_t)M * N >= (int64_t)512 * 512);
    if (tc) {
        // the split scratch holds 256 k-columns: deeper updates (the blocked TRSM of getrs) go in slabs
        for (int k0 = 0; k0 < K; k0 += 256) {
            const int rc = launch_sgemm_tc(h, st, M, N, std::min(256, K - k0), A + (int64_t)k0 * lda, lda, B + k0, ldb, C, ldc);
            if (rc) return rc;
        }
        return 0;
    }
    return launch_sgemm_ffma(h, st, M, N, K, A, lda, B, ldb, C, ldc);
}

// trailing-update GEMM, optionally bracketed by events for the in-situ roofline
template <typename T>
static int launch_gemm_prof(b200lu_handle* h, cudaStream_t st, int M, int N, int K, const T* A,
                            int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const bool prof = h->opt[B200LU_OPT_PROFILE] != 0;
    if (prof) {
        while ((int)h->prof_ev.size() < h->prof_used + 2) {
            cudaEvent_t e;
            CU_TRY(h, cudaEventCreate(&e));
            h->prof_ev.push_back(e);
        }
        CU_TRY(h, cudaEventRecord(h->prof_ev[h->prof_used], st));
    }
    int rc = launch_gemm(h, st, M, N, K, A, lda, B, ldb, C, ldc);
    if (rc) return rc;
    if (prof) {
        CU_TRY(h, cudaEventRecord(h->prof_ev[h->prof_used + 1], st));
        h->prof_used += 2;
        h->prof_flops += 2.0 * M * (double)N * K;
    }
    return 0;
}

template <typename T>
static int launch_laswp(b200lu_handle* h, cudaStream_t st, T* A, int64_t lda, int c0, int c1,
                        const LaswpPlan* plan) {
    if (c1 <= c0) return 0;
    constexpr int CW = 8, NT = 256;
    const int grid = std::min(cdiv(c1 - c0, CW), 148 * 8);
    laswp_apply_kernel<T, CW, NT><<<grid, NT, 0, st>>>(A, lda, c0, c1, plan);
    LAUNCH_CHECK(h);
    return 0;
}

// ------------------------------------------------------ cluster base panel --
template <typename T, int W, int RPT, int NSUB = 1, int NTV = PCL_NT>
static int launch_panel_cluster_cfg(b200lu_handle* h, cudaStream_t st, PanelArgs<T> p) {
    constexpr int PCL_NT = NTV;   // threads per CTA of this instantiation
    auto kern = panel_cluster_kernel<T, W, RPT, PCL_NT, NSUB>;
    constexpr size_t smem = pcl_smem_bytes<T, W, RPT, PCL_NT, NSUB>();
    static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
    if (!attr_set) {
        CU_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        CU_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int G = 1;
    while (G * PCL_NT * RPT < p.m) G *= 2;  // power-of-two cluster sizes only
    if (G > PCL_GMAX) return set_err(h, 2, "cluster panel of %d rows needs %d CTAs > %d", p.m, G, PCL_GMAX);
    p.G = G;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(PCL_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) {
        // a cluster of this size cannot be placed on this device / partition (MIG, MPS limits):
        // fall back to the L2-mailbox kernel for the rest of the handle's life
        cudaGetLastError();
        h->opt[B200LU_OPT_PANEL_MODE] = 1;
        return -1000;
    }
    LAUNCH_CHECK(h);
    return 0;
}

// Panel classes of the cluster kernel (rows m of the panel -> base width W, rows per thread, sub-blocks fused
// into one launch; B200LU_OPT_PANEL_MODE 0 = fused, 2 = the round-1 kernels: one W-wide block per launch):
//   FP64:  m <=  4096: 32 x 1, 4 sub-blocks (128 columns per launch)     FP32:  m <=  4096: 32 x 1, 4
//          m <=  8192: 32 x 2, 2 sub-blocks ( 64)                                m <=  8192: 32 x 2, 4
//          m <= 16384: 16 x 4, 4 sub-blocks ( 64)                                m <= 16384: 32 x 4, 2
//          m <= 32768:  8 x 8, 4 sub-blocks ( 32)   [fused only; round 1: L2 mailbox; 8 sub-blocks measured: no gain]
// 256 threads hold up to 128 data registers per thread: 32 doubles x 2 rows, 16 x 4, 8 x 8.
constexpr int PCL_ROWS1 = PCL_GMAX * PCL_NT;       // 4096 rows at one row per thread
constexpr int PCL_ROWS2 = PCL_GMAX * PCL_NT * 2;
static bool panel_fused(const b200lu_handle* h) { return h->opt[B200LU_OPT_PANEL_MODE] == 0; }

// Base width of the cluster kernel for an OUTER panel of m rows (0 = does not fit: L2-mailbox kernel of panel.cuh).
template <typename T>
static int cluster_base_width(const b200lu_handle* h, int m) {
    if (h->opt[B200LU_OPT_PANEL_MODE] == 1) return 0;
    if (m <= PCL_ROWS2) return 32;
    if (m <= 2 * PCL_ROWS2) return sizeof(T) == 8 ? 16 : 32;
    if (m <= 4 * PCL_ROWS2 && sizeof(T) == 8 && panel_fused(h)) return 8;
    return 0;
}
// Columns one launch factors: the base width times the sub-blocks its class fuses
template <typename T>
static int cluster_group_width(const b200lu_handle* h, int m, int bw) {
    if (!panel_fused(h) || bw == 0) return bw;
    if (sizeof(T) == 8) {
        if (bw == 32) return m <= PCL_ROWS1 ? 128 : 64;
        if (bw == 16) return 64;
        return 32;   // bw == 8
    }
    return m <= PCL_ROWS2 ? 128 : 64;
}

// can a panel of m rows with base width bw run on the cluster kernel?
template <typename T>
static bool cluster_usable(const b200lu_handle* h, int m, int bw) {
    if (h->opt[B200LU_OPT_PANEL_MODE] == 1) return false;
    if (m <= PCL_ROWS2) return true;
    if (m <= 2 * PCL_ROWS2) return sizeof(T) == 4 || bw <= 16;
    return m <= 4 * PCL_ROWS2 && sizeof(T) == 8 && panel_fused(h) && bw <= 8;
}

template <typename T>
static int launch_panel_any(b200lu_handle* h, cudaStream_t st, T* A, int64_t lda, int nrows, int j0,
                            int wc, int pc0, int pc1, int bw) {
    const int m = nrows - j0;
    const bool fused = panel_fused(h);
    const bool cluster_ok = cluster_usable<T>(h, m, bw);
    if (!cluster_ok) {
        if (wc > BASE_W) return -1000;   // the caller splits down to the mailbox kernel's width
        return launch_panel_base<T>(h, st, A, lda, nrows, j0, wc, pc0, pc1);
    }
    PanelArgs<T> p;
    p.A = A;
    p.lda = lda;
    p.j0 = j0;
    p.m = m;
    p.wc = wc;
    p.pc0 = pc0;
    p.pc1 = pc1;
    p.ipiv = h->d_ipiv;
    p.info = h->d_info;
    p.G = 0;
    p.epoch = 0;
    p.mail = nullptr;
    p.deverr = h->d_deverr;
    p.dbg = h->d_pdbg ? h->d_pdbg + 24 * (h->pdbg_n++ % 4096) : nullptr;
    int rc;
    if constexpr (sizeof(T) == 8) {
        if (fused) {
            if (bw > 16) rc = (m <= PCL_ROWS1) ? launch_panel_cluster_cfg<T, 32, 1, 4>(h, st, p)
                                               : launch_panel_cluster_cfg<T, 32, 2, 2>(h, st, p);
            else if (bw > 8) rc = launch_panel_cluster_cfg<T, 16, 4, 4>(h, st, p);
            else rc = launch_panel_cluster_cfg<T, 8, 8, 4>(h, st, p);
        } else {
            // one row per thread while 16 CTAs x 256 threads cover the panel (measured: n = 4096 10.3 -> 9.6 ms)
            if (bw > 16) rc = (m <= PCL_ROWS1) ? launch_panel_cluster_cfg<T, 32, 1>(h, st, p)
                                               : launch_panel_cluster_cfg<T, 32, 2>(h, st, p);
            else if (m <= PCL_ROWS2) rc = launch_panel_cluster_cfg<T, 16, 2>(h, st, p);
            else rc = launch_panel_cluster_cfg<T, 16, 4>(h, st, p);
        }
    } else {
        if (fused) {
            if (m <= PCL_ROWS1) rc = launch_panel_cluster_cfg<T, 32, 1, 4>(h, st, p);
            else if (m <= PCL_ROWS2) rc = launch_panel_cluster_cfg<T, 32, 2, 4>(h, st, p);
            else rc = launch_panel_cluster_cfg<T, 32, 4, 2>(h, st, p);
        } else {
            if (m <= PCL_ROWS1) rc = launch_panel_cluster_cfg<T, 32, 1>(h, st, p);
            else if (m <= PCL_ROWS2) rc = launch_panel_cluster_cfg<T, 32, 2>(h, st, p);
            else rc = launch_panel_cluster_cfg<T, 32, 4>(h, st, p);
        }
    }
    if (rc == -1000 && wc <= BASE_W) return launch_panel_base<T>(h, st, A, lda, nrows, j0, wc, pc0, pc1);
    return rc;
}

// ---------------------------------------------------------- recursive panel --
// Factor columns [j0, j0+w) of the nrows x * matrix (rows j0..nrows-1) inside the
// outer panel [pc0, pc1), recursing down to groups the base kernels take in one launch (bw columns
// for the round-1 kernels and the L2 mailbox, bw x fused sub-blocks for the left-looking cluster kernel).
// Interchanges are applied to the whole outer panel by the base kernels, so no laswp appears here.
template <typename T>
static int panel_recursive(b200lu_handle* h, cudaStream_t st, T* A, int64_t lda, int nrows, int j0,
                           int w, int pc0, int pc1, int bw) {
    const int gw = cluster_usable<T>(h, nrows - j0, bw) ? std::max(bw, cluster_group_width<T>(h, nrows - j0, bw)) : bw;
    if (w <= gw) {
        const int rc0 = launch_panel_any<T>(h, st, A, lda, nrows, j0, w, pc0, pc1, bw);
        if (rc0 != -1000) return rc0;
        // the cluster launch was refused (the handle has switched to the L2 mailbox): split down to its width
        if (w <= BASE_W) return launch_panel_base<T>(h, st, A, lda, nrows, j0, w, pc0, pc1);
        bw = BASE_W;
    }
    const int w1 = ((w / 2 + bw - 1) / bw) * bw;
    const int w2 = w - w1;
    int rc = panel_recursive<T>(h, st, A, lda, nrows, j0, w1, pc0, pc1, bw);
    if (rc) return rc;
    T* L11 = A + (int64_t)j0 * lda + j0;
    T* A12 = A + (int64_t)(j0 + w1) * lda + j0;
    rc = launch_trsm<T>(h, st, L11, lda, A12, lda, w1, w2);
    if (rc) return rc;
    const int mrest = nrows - (j0 + w1);
    rc = launch_gemm(h, st, mrest, w2, w1, A + (int64_t)j0 * lda + (j0 + w1), lda, A12, lda,
                     A + (int64_t)(j0 + w1) * lda + (j0 + w1), lda);
    if (rc) return rc;
    return panel_recursive<T>(h, st, A, lda, nrows, j0 + w1, w2, pc0, pc1, bw);
}
// outer-panel entry: pick the base width from the panel height
template <typename T>
static int panel_factor(b200lu_handle* h, cudaStream_t st, T* A, int64_t lda, int nrows, int j0,
                        int w, int pc0, int pc1) {
    int bw = cluster_base_width<T>(h, nrows - j0);
    if (bw == 0) bw = BASE_W;
    return panel_recursive<T>(h, st, A, lda, nrows, j0, w, pc0, pc1, bw);
}

static int ensure_events(b200lu_handle* h, int count) {
    while ((int)h->ev_panel.size() < count) {
        cudaEvent_t e;
        CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_panel.push_back(e);
    }
    return 0;
}

// getrf of the n x n matrix in `A` (device, leading dim lda).
// nchunks > 0: A is still being uploaded in column chunks (h->chunk_end / h->ev_chunk, recorded
// on the copy stream).  A chunk is admitted when it has landed (host-side event query, the host
// is paced one panel behind the device while chunks are outstanding) or at the latest when the
// next panel needs its columns; on admission the chunk is caught up left-looking: interchanges
// of every finished panel, the U block rows by TRSM, one deep GEMM below them.  Each element
// sees the same arithmetic in the same order as in the resident factorization.
template <typename T>
static int getrf_device(b200lu_handle* h, T* A, int64_t lda, int n, int nchunks = 0) {
    const int nb = (int)h->opt[B200LU_OPT_NB];
    const bool la = h->opt[B200LU_OPT_LOOKAHEAD] != 0;
    const int nblk = cdiv(n, nb);
    int rc = ensure_events(h, nblk + 1);
    if (rc) return rc;
    cudaStream_t sm = h->s_main;
    cudaStream_t sp = la ? h->s_panel : h->s_main;
    int adm = nchunks > 0 ? 0 : n;   // columns [0, adm) are on the device
    static const bool sdbg = getenv("B200LU_STREAM_DBG") != nullptr;
    int next_chunk = 0;
    // admit the next chunk; panels [0, kdone) are finished and applied to everything admitted so far
    auto admit = [&](int kdone) -> int {
        const int c0 = adm, c1 = h->chunk_end[next_chunk];
        CU_TRY(h, cudaStreamWaitEvent(sm, h->ev_chunk[next_chunk], 0));
        ++next_chunk;
        adm = c1;
        if (kdone == 0) return 0;
        int r;
        for (int j = 0; j < kdone; ++j) {
            r = launch_laswp<T>(h, sm, A, lda, c0, c1, h->d_plans + j);
            if (r) return r;
        }
        const int kd = kdone * nb;
        for (int j = 0; j < kdone; ++j) {
            const int r0 = j * nb, r1 = r0 + nb;
            r = launch_trsm<T>(h, sm, A + (int64_t)r0 * lda + r0, lda, A + (int64_t)c0 * lda + r0, lda, nb, c1 - c0);
            if (r) return r;
            if (r1 < kd) {
                r = launch_gemm(h, sm, kd - r1, c1 - c0, nb, A + (int64_t)r0 * lda + r1, lda,
                                A + (int64_t)c0 * lda + r0, lda, A + (int64_t)c0 * lda + r1, lda);
                if (r) return r;
            }
        }
        if (kd < n) {
            r = launch_gemm_prof<T>(h, sm, n - kd, c1 - c0, kd, A + kd, lda, A + (int64_t)c0 * lda, lda,
                                    A + (int64_t)c0 * lda + kd, lda);
            if (r) return r;
        }
        return 0;
    };

    CU_TRY(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), sm));
    iota_kernel<<<cdiv(n, 256), 256, 0, sm>>>(h->d_perm, n);
    LAUNCH_CHECK(h);
    while (next_chunk < nchunks && adm < std::min(nb, n)) {
        rc = admit(0);
        if (rc) return rc;
    }
    if (la) {
        CU_TRY(h, cudaEventRecord(h->ev_fork, sm));
        CU_TRY(h, cudaStreamWaitEvent(sp, h->ev_fork, 0));
    }
    {
        const int jb = std::min(nb, n);
        rc = panel_factor<T>(h, sp, A, lda, n, 0, jb, 0, jb);
        if (rc) return rc;
        laswp_plan_kernel<<<1, 2 * LASWP_MAXSW, 0, sp>>>(h->d_ipiv, 0, jb, h->d_plans + 0);
        LAUNCH_CHECK(h);
        if (la) CU_TRY(h, cudaEventRecord(h->ev_panel[0], sp));
    }
    for (int k = 0; k < nblk; ++k) {
        const int j0 = k * nb;
        const int jb = std::min(nb, n - j0);
        const int j1 = j0 + jb;
        const LaswpPlan* plan = h->d_plans + k;
        if (la) CU_TRY(h, cudaStreamWaitEvent(sm, h->ev_panel[k], 0));
        const int jb2 = std::min(nb, n - j1);
        // a chunk the next panel needs is admitted (and waited for) now; chunks that have merely
        // landed are admitted below, after the next panel has been handed to the panel stream,
        // so that their catch-up never delays the panel chain
        while (next_chunk < nchunks && j1 + jb2 > adm) {
            if (sdbg) fprintf(stderr, "[stream] step %d admits chunk %d (needed)\n", k, next_chunk);
            rc = admit(k);
            if (rc) return rc;
        }
        // interchanges: the next panel's columns first (they gate the look-ahead), the rest
        // of the matrix after the next panel has been handed to the panel stream
        if (j1 < n) {
            rc = launch_laswp<T>(h, sm, A, lda, j1, j1 + jb2, plan);
            if (rc) return rc;
        }
        if (j1 >= n) {
            rc = launch_laswp<T>(h, sm, A, lda, 0, j0, plan);
            if (rc) return rc;
            rc = launch_laswp<int>(h, sm, h->d_perm, n, 0, 1, plan);
            if (rc) return rc;
            break;
        }
        T* L11 = A + (int64_t)j0 * lda + j0;
        T* L21 = A + (int64_t)j0 * lda + j1;
        // next panel's columns first (look-ahead)
        rc = launch_trsm<T>(h, sm, L11, lda, A + (int64_t)j1 * lda + j0, lda, jb, jb2);
        if (rc) return rc;
        rc = launch_gemm(h, sm, n - j1, jb2, jb, L21, lda, A + (int64_t)j1 * lda + j0, lda,
                         A + (int64_t)j1 * lda + j1, lda);
        if (rc) return rc;
        if (la) {
            CU_TRY(h, cudaEventRecord(h->ev_next, sm));
            CU_TRY(h, cudaStreamWaitEvent(sp, h->ev_next, 0));
        }
        rc = panel_factor<T>(h, sp, A, lda, n, j1, jb2, j1, j1 + jb2);
        if (rc) return rc;
        laswp_plan_kernel<<<1, 2 * LASWP_MAXSW, 0, sp>>>(h->d_ipiv, j1, jb2, h->d_plans + (k + 1));
        LAUNCH_CHECK(h);
        if (la) CU_TRY(h, cudaEventRecord(h->ev_panel[k + 1], sp));
        if (next_chunk < nchunks) {
            // the host stays one panel behind the device while chunks are outstanding, so that
            // the arrival query is fresh
            CU_TRY(h, cudaEventSynchronize(h->ev_panel[k]));
            while (next_chunk < nchunks) {
                if (cudaEventQuery(h->ev_chunk[next_chunk]) != cudaSuccess) {
                    (void)cudaGetLastError();   // cudaErrorNotReady
                    break;
                }
                if (sdbg) fprintf(stderr, "[stream] step %d admits chunk %d (landed)\n", k, next_chunk);
                rc = admit(k);
                if (rc) return rc;
            }
        }
        // the remaining interchanges of panel k, then the rest of the trailing matrix
        rc = launch_laswp<T>(h, sm, A, lda, 0, j0, plan);
        if (rc) return rc;
        rc = launch_laswp<int>(h, sm, h->d_perm, n, 0, 1, plan);
        if (rc) return rc;
        const int c2 = j1 + jb2;
        if (c2 < adm) {
            rc = launch_laswp<T>(h, sm, A, lda, c2, adm, plan);
            if (rc) return rc;
            rc = launch_trsm<T>(h, sm, L11, lda, A + (int64_t)c2 * lda + j0, lda, jb, adm - c2);
            if (rc) return rc;
            rc = launch_gemm_prof<T>(h, sm, n - j1, adm - c2, jb, L21, lda, A + (int64_t)c2 * lda + j0, lda,
                                     A + (int64_t)c2 * lda + j1, lda);
            if (rc) return rc;
        }
    }
    return 0;
}

// ----------------------------------------------------------------- capacity --
// pivots / permutation / laswp plans / pinned staging: needed by every mode
static int ensure_meta(b200lu_handle* h, int64_t n) {
    if (n <= h->cap_meta) return 0;
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    int* old_ipiv = h->d_ipiv;
    int* old_perm = h->d_perm;
    const int64_t old_cap = h->cap_meta;
    h->d_ipiv = nullptr;
    h->d_perm = nullptr;
    h->cap_meta = 0;   // committed at the end; a failed allocation leaves an empty, consistent handle
    free_dev(h->d_plans);
    h->cap_plans = 0;
    bool ok = cudaMalloc((void**)&h->d_ipiv, (size_t)n * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&h->d_perm, (size_t)n * sizeof(int)) == cudaSuccess;
    if (ok && old_ipiv && old_perm && old_cap > 0) {  // keep a cached factorization's pivots alive
        ok = cudaMemcpy(h->d_ipiv, old_ipiv, (size_t)old_cap * sizeof(int), cudaMemcpyDeviceToDevice) == cudaSuccess &&
             cudaMemcpy(h->d_perm, old_perm, (size_t)old_cap * sizeof(int), cudaMemcpyDeviceToDevice) == cudaSuccess;
    }
    free_dev(old_ipiv);
    free_dev(old_perm);
    const int nplans = cdiv(n, 16) + 1;
    ok = ok && cudaMalloc((void**)&h->d_plans, (size_t)nplans * sizeof(LaswpPlan)) == cudaSuccess;
    if (ok && n > h->cap_hipiv) {
        if (h->h_ipiv) cudaFreeHost(h->h_ipiv);
        h->h_ipiv = nullptr;
        h->cap_hipiv = 0;
        ok = cudaMallocHost((void**)&h->h_ipiv, (size_t)n * sizeof(long long)) == cudaSuccess;
        if (ok) h->cap_hipiv = n;
    }
    if (!ok) {
        const cudaError_t e = cudaGetLastError();
        h->factored = false;
        h->factored_dist = false;
        free_dev(h->d_ipiv);
        free_dev(h->d_perm);
        free_dev(h->d_plans);
        return set_err(h, 1, "CUDA error %s while growing the pivot buffers to n = %lld: %s", cudaGetErrorName(e),
                       (long long)n, cudaGetErrorString(e));
    }
    h->cap_plans = nplans;
    h->cap_meta = n;
    return 0;
}

// drop every per-size buffer of the dense path and return the handle to the empty state
static void reset_dense_state(b200lu_handle* h) {
    h->factored = false;
    h->solve_ready = false;
    h->solve_ready_t = false;
    h->keep_valid = false;
    h->n = 0;
    h->ldd = 0;
    h->cap_n = 0;
    h->cap_rhs = 0;
    h->cap_tgroups = 0;
    h->cap_tflag_bytes = 0;
    h->t2_nblk = 0;
    h->t3_nblk = 0;
    free_dev(h->dA);
    free_dev(h->dA64);
    free_dev(h->d_dinvL);
    free_dev(h->d_dinvU);
    free_dev(h->d_wL);
    free_dev(h->d_wU);
    free_dev(h->d_wLt);
    free_dev(h->d_wUt);
    free_dev(h->d_r);
    free_dev(h->d_r32);
    free_dev(h->d_tflags);
    free_dev(h->d_tticket);
    free_dev(h->d_B);
    free_dev(h->d_X);
}

static int ensure_capacity(b200lu_handle* h, int64_t n) {
    int rc = ensure_meta(h, n);
    if (rc) return rc;
    if (n <= h->cap_n) {
        if (n != h->n || h->ldd != ((n + 15) / 16) * 16) {
            h->factored = false;   // the cached factors are about to be overwritten
            h->solve_ready = false;
            h->solve_ready_t = false;
            h->n = n;
            h->ldd = ((n + 15) / 16) * 16;
            // keep padding rows finite for the 16-byte chunk loads
            CU_TRY(h, cudaMemsetAsync(h->dA, 0, (size_t)h->ldd * n * elem_size(h), h->s_main));
        }
        return 0;
    }
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    // growth: from here until every allocation has succeeded the handle is EMPTY (no cached
    // factorization, zero capacity), so a failed cudaMalloc cannot leave stale state behind
    // that a later solve / smaller factor call would run on freed or null buffers
    reset_dense_state(h);
    const int64_t ldd = ((n + 15) / 16) * 16;
    const size_t es = elem_size(h);
    const int nblk = cdiv(n, TRSV_TB);
    bool ok = cudaMalloc(&h->dA, (size_t)ldd * n * es) == cudaSuccess;
    if (ok && h->dtype == B200LU_MIXED) {
        ok = cudaMalloc((void**)&h->dA64, (size_t)ldd * n * 8) == cudaSuccess &&
             cudaMalloc((void**)&h->d_r, (size_t)n * 8) == cudaSuccess &&
             cudaMalloc((void**)&h->d_r32, (size_t)n * 4) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&h->d_dinvL, (size_t)nblk * TRSV_TB * TRSV_TB * es) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_dinvU, (size_t)nblk * TRSV_TB * TRSV_TB * es) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_wL, (size_t)TRSV3_NEAR_MAX * nblk * TRSV_TB * TRSV_TB * es) == cudaSuccess;   // coupling planes
    ok = ok && cudaMalloc(&h->d_wU, (size_t)TRSV3_NEAR_MAX * nblk * TRSV_TB * TRSV_TB * es) == cudaSuccess;
    ok = ok && cudaMemsetAsync(h->dA, 0, (size_t)ldd * n * es, h->s_main) == cudaSuccess;
    if (!ok) {
        const cudaError_t e = cudaGetLastError();
        reset_dense_state(h);
        return set_err(h, 1, "CUDA error %s while growing the factor buffers to n = %lld: %s", cudaGetErrorName(e),
                       (long long)n, cudaGetErrorString(e));
    }
    h->n = n;
    h->ldd = ldd;
    h->cap_n = n;
    return 0;
}

static int ensure_rhs(b200lu_handle* h, int64_t nrhs) {
    if (nrhs <= h->cap_rhs) return 0;
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    free_dev(h->d_B);
    free_dev(h->d_X);
    h->cap_rhs = 0;   // committed only when both allocations have succeeded
    if (cudaMalloc(&h->d_B, (size_t)h->cap_n * nrhs * 8) != cudaSuccess ||
        cudaMalloc(&h->d_X, (size_t)h->cap_n * nrhs * 8) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        free_dev(h->d_B);
        free_dev(h->d_X);
        return set_err(h, 1, "CUDA error %s allocating %lld right-hand sides: %s", cudaGetErrorName(e), (long long)nrhs,
                       cudaGetErrorString(e));
    }
    h->cap_rhs = nrhs;
    return 0;
}

static int ensure_trsv_groups(b200lu_handle* h, int groups, int NR) {
    // LL packet buffer: [groups][nblk*64 rows][NR][2 words] x 8 bytes (sized for FP64)
    const int nblk = cdiv(h->cap_n, TRSV_TB);
    const size_t need = (size_t)groups * nblk * TRSV_TB * NR * 2 * sizeof(unsigned long long);
    if (need <= h->cap_tflag_bytes && groups <= h->cap_tgroups) return 0;
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    free_dev(h->d_tflags);
    free_dev(h->d_tticket);
    const size_t bytes = std::max(need, h->cap_tflag_bytes);
    const int g = std::max(groups, h->cap_tgroups);
    CU_TRY(h, cudaMalloc((void**)&h->d_tflags, bytes));
    CU_TRY(h, cudaMemset(h->d_tflags, 0, bytes));
    CU_TRY(h, cudaMalloc((void**)&h->d_tticket, (size_t)g * sizeof(int)));
    CU_TRY(h, cudaMemset(h->d_tticket, 0, (size_t)g * sizeof(int)));
    h->cap_tflag_bytes = bytes;
    h->cap_tgroups = g;
    h->trsv_epoch = 0;
    return 0;
}

// getrs, one right-hand side, version 3 (trsv_cluster.cuh): both sweeps.  Returns -1000 when the
// cluster launch is not possible on this device (the caller falls back to version 2).
template <typename T, int NEAR, int CS>
static int trsv3_sweeps(b200lu_handle* h, const T* A, int64_t lda, int n, const T* B, T* X, int nblk) {
    cudaStream_t st = h->s_main;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = trsv3_smem_bytes<T, NEAR>();
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    auto kl = trsv3_kernel<T, false, NEAR, CS>;
    auto ku = trsv3_kernel<T, true, NEAR, CS>;
    if (h->t3_ok < 0 || h->t3_nblk != nblk) {
        static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
        if (!attr_set) {
            CU_TRY(h, cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, trsv3_smem_bytes<T, NEAR>()));
            CU_TRY(h, cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, trsv3_smem_bytes<T, NEAR>()));
            if (CS > 8) {
                CU_TRY(h, cudaFuncSetAttribute(kl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                CU_TRY(h, cudaFuncSetAttribute(ku, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            }
            attr_set = true;
        }
        int ncl_l = 0, ncl_u = 0;
        cfg.gridDim = dim3(CS * 2);
        cudaError_t e1 = cudaOccupancyMaxActiveClusters(&ncl_l, kl, &cfg);
        cudaError_t e2 = cudaOccupancyMaxActiveClusters(&ncl_u, ku, &cfg);
        const int ncl = std::min(ncl_l, ncl_u);
        if (getenv("B200LU_TRSV_DBG")) fprintf(stderr, "[trsv3] NEAR=%d CS=%d: %d co-resident clusters\n", NEAR, CS, ncl);
        if (e1 != cudaSuccess || e2 != cudaSuccess || ncl < 2) {
            (void)cudaGetLastError();
            h->t3_ok = 0;
            return -1000;
        }
        CU_TRY(h, cudaStreamSynchronize(st));
        free_dev(h->d_t3items); free_dev(h->d_t3x); free_dev(h->d_t3p);
        std::vector<Trsv2Item> items;
        int kmax = 1;
        for (int t = NEAR + 1; t < nblk; ++t) {
            const int nch = cdiv(t - NEAR, TRSV3_CH);
            kmax = std::max(kmax, nch);
            for (int k = 0; k < nch; ++k) items.push_back(Trsv2Item{t, k});
        }
        CU_TRY(h, cudaMalloc((void**)&h->d_t3items, items.size() * sizeof(Trsv2Item)));
        CU_TRY(h, cudaMemcpy(h->d_t3items, items.data(), items.size() * sizeof(Trsv2Item), cudaMemcpyHostToDevice));
        const size_t xb = (size_t)nblk * TRSV_TB * 2 * sizeof(unsigned long long);
        const size_t pb = (size_t)nblk * kmax * TRSV_TB * 2 * sizeof(unsigned long long);
        CU_TRY(h, cudaMalloc((void**)&h->d_t3x, xb));
        CU_TRY(h, cudaMemset(h->d_t3x, 0, xb));
        CU_TRY(h, cudaMalloc((void**)&h->d_t3p, pb));
        CU_TRY(h, cudaMemset(h->d_t3p, 0, pb));
        if (!h->d_t3ticket) {
            CU_TRY(h, cudaMalloc((void**)&h->d_t3ticket, 64));
            CU_TRY(h, cudaMemset(h->d_t3ticket, 0, 64));
        }
        if (getenv("B200LU_TRSV_DBG") && !h->d_t3dbg) {
            CU_TRY(h, cudaMalloc((void**)&h->d_t3dbg, 16 * 8 * sizeof(long long)));
            CU_TRY(h, cudaMemset(h->d_t3dbg, 0, 16 * 8 * sizeof(long long)));
        }
        h->t3_clusters = ncl;   // every cluster resident: the chain and the tickets cannot deadlock
        h->t3_nblk = nblk;
        h->t3_nitems = (int)items.size();
        h->t3_kmax = kmax;
        h->t3_epoch = 0;
        h->t3_ok = 1;
    }
    // one cluster for the chain, workers up to one item each
    const int clusters = std::max(2, std::min(h->t3_clusters, 1 + cdiv(h->t3_nitems, CS)));
    cfg.gridDim = dim3(clusters * CS);
    for (int upper = 0; upper < 2; ++upper) {
        if (h->t3_epoch > (1u << 30)) {
            CU_TRY(h, cudaMemsetAsync(h->d_t3x, 0, (size_t)nblk * TRSV_TB * 2 * sizeof(unsigned long long), st));
            CU_TRY(h, cudaMemsetAsync(h->d_t3p, 0, (size_t)nblk * h->t3_kmax * TRSV_TB * 2 * sizeof(unsigned long long), st));
            h->t3_epoch = 0;
        }
        const unsigned epoch = ++h->t3_epoch;
        Trsv3Sync sy{h->d_t3x, h->d_t3p, h->d_t3ticket, h->d_deverr, h->d_t3items, h->t3_nitems, h->t3_kmax, h->d_t3dbg};
        const T* nullT = nullptr;
        const int* nullI = nullptr;
        if (upper)
            CU_TRY(h, cudaLaunchKernelEx(&cfg, ku, A, (long long)lda, n, (const T*)h->d_dinvU, (const T*)h->d_wU,
                                         nullT, nullI, X, sy, epoch, nblk));
        else
            CU_TRY(h, cudaLaunchKernelEx(&cfg, kl, A, (long long)lda, n, (const T*)h->d_dinvL, (const T*)h->d_wL,
                                         B, (const int*)h->d_perm, X, sy, epoch, nblk));
        B200LU_COUNT_LAUNCH();
    }
    return 0;
}

// ------------------------------------------------------------------- getrs --
// X = U \ (L \ (P B)).  B must not alias X here (the lower sweep gathers B rows
// through the permutation); getrs_device stages an aliased right-hand side.
template <typename T>
static int trsv_sweeps(b200lu_handle* h, const T* A, int64_t lda, int n, const T* B, int64_t ldb,
                       T* X, int64_t ldx, int nrhs, bool trans = false) {
    cudaStream_t st = h->s_main;
    const int nblk = cdiv(n, TRSV_TB);
    if (!h->solve_ready) {
        const size_t tsm = sizeof(T) * TRSV_TB * (2 * TRSV_TB + 1);
        static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
        if (!attr_set) {
            CU_TRY(h, cudaFuncSetAttribute(trtri_diag_kernel<T>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
            attr_set = true;
        }
        trtri_diag_kernel<T><<<nblk, TRSV_TB, tsm, st>>>(A, lda, n, (T*)h->d_dinvL, 0);
        LAUNCH_CHECK(h);
        trtri_diag_kernel<T><<<nblk, TRSV_TB, tsm, st>>>(A, lda, n, (T*)h->d_dinvU, 1);
        LAUNCH_CHECK(h);
        trsv_coupling_kernel<T><<<dim3(nblk, TRSV3_NEAR_MAX), 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvL, (T*)h->d_wL, 0, nblk, 0);
        LAUNCH_CHECK(h);
        trsv_coupling_kernel<T><<<dim3(nblk, TRSV3_NEAR_MAX), 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvU, (T*)h->d_wU, 1, nblk, 0);
        LAUNCH_CHECK(h);
        h->solve_ready = true;
    }
    if (trans && !h->solve_ready_t) {
        // A^T = U^T L^T P: first sweep with U^T (lower order, coupling with block r-1), second with L^T
        const size_t wb = (size_t)nblk * TRSV_TB * TRSV_TB * sizeof(T);
        if (!h->d_wLt) CU_TRY(h, cudaMalloc(&h->d_wLt, (size_t)cdiv(h->cap_n, TRSV_TB) * TRSV_TB * TRSV_TB * sizeof(T)));
        if (!h->d_wUt) CU_TRY(h, cudaMalloc(&h->d_wUt, (size_t)cdiv(h->cap_n, TRSV_TB) * TRSV_TB * TRSV_TB * sizeof(T)));
        (void)wb;
        trsv_coupling_kernel<T><<<nblk, 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvU, (T*)h->d_wUt, 0, nblk, 1);
        LAUNCH_CHECK(h);
        trsv_coupling_kernel<T><<<nblk, 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvL, (T*)h->d_wLt, 1, nblk, 1);
        LAUNCH_CHECK(h);
        h->solve_ready_t = true;
    }
    if (!trans && nrhs == 1 && h->t3_ok != 0 &&
        (h->opt[B200LU_OPT_TRSV_MODE] == 3 || (h->opt[B200LU_OPT_TRSV_MODE] == 0 && nblk >= 96)) && nblk >= 16) {
        // ---- version 3: the dependency chain in one cluster over DSMEM, workers on the far blocks ----
        // measured and dropped: NEAR = 6 and 16-CTA clusters (see trsv_cluster.cuh)
        const int rc3 = trsv3_sweeps<T, 4, 8>(h, A, lda, n, B, X, nblk);
        if (rc3 != -1000) return rc3;   // -1000: cluster launch not possible here, version 2 takes over
    }
    if (trans || (nrhs == 1 && h->opt[B200LU_OPT_TRSV_MODE] != 1 && nblk >= 4)) {
        // ---- version 2: 2-D work items on a persistent, fully resident grid ----
        if (h->t2_nblk != nblk) {
            CU_TRY(h, cudaStreamSynchronize(st));
            free_dev(h->d_t2items); free_dev(h->d_t2x); free_dev(h->d_t2p);
            std::vector<Trsv2Item> items;
            int kmax = 1;
            for (int t = 0; t < nblk; ++t) {
                const int nfar = t > 2 ? t - 2 : 0;
                const int nch = (nfar + TRSV2_CH - 1) / TRSV2_CH;
                kmax = std::max(kmax, nch);
                for (int k = 0; k < nch; ++k) items.push_back(Trsv2Item{t, k});
                items.push_back(Trsv2Item{t, -1});
            }
            CU_TRY(h, cudaMalloc((void**)&h->d_t2items, items.size() * sizeof(Trsv2Item)));
            CU_TRY(h, cudaMemcpy(h->d_t2items, items.data(), items.size() * sizeof(Trsv2Item), cudaMemcpyHostToDevice));
            const size_t xb = (size_t)nblk * TRSV_TB * 2 * sizeof(unsigned long long);
            const size_t pb = (size_t)nblk * kmax * TRSV_TB * 2 * sizeof(unsigned long long);
            CU_TRY(h, cudaMalloc((void**)&h->d_t2x, xb));
            CU_TRY(h, cudaMemset(h->d_t2x, 0, xb));
            CU_TRY(h, cudaMalloc((void**)&h->d_t2p, pb));
            CU_TRY(h, cudaMemset(h->d_t2p, 0, pb));
            if (!h->d_t2ticket) {
                CU_TRY(h, cudaMalloc((void**)&h->d_t2ticket, 64));
                CU_TRY(h, cudaMemset(h->d_t2ticket, 0, 64));
            }
            int occ = 0, sms = 0;
            CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trsv2_kernel<T, false, false>, 256, 0));
            int occ_u = 0;
            CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, trsv2_kernel<T, true, true>, 256, 0));
            CU_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->dev));
            h->t2_grid = std::max(1, std::min(occ, occ_u)) * sms;   // every CTA resident: tickets cannot deadlock
            h->t2_nblk = nblk;
            h->t2_nitems = (int)items.size();
            h->t2_kmax = kmax;
            h->t2_epoch = 0;
        }
        const int grid = std::min(h->t2_grid, h->t2_nitems);
        if (trans) {
            int rc = ensure_rhs(h, 1);
            if (rc) return rc;
        }
        for (int col = 0; col < nrhs; ++col) {
            const T* Bc = B + (int64_t)col * ldb;
            T* Xc = X + (int64_t)col * ldx;
            for (int upper = 0; upper < 2; ++upper) {
                if (h->t2_epoch > (1u << 30)) {
                    CU_TRY(h, cudaMemsetAsync(h->d_t2x, 0, (size_t)nblk * TRSV_TB * 2 * sizeof(unsigned long long), st));
                    CU_TRY(h, cudaMemsetAsync(h->d_t2p, 0, (size_t)nblk * h->t2_kmax * TRSV_TB * 2 * sizeof(unsigned long long), st));
                    h->t2_epoch = 0;
                }
                const unsigned epoch = ++h->t2_epoch;
                Trsv2Sync sy{h->d_t2x, h->d_t2p, h->d_t2ticket, h->d_deverr, h->d_t2items, h->t2_nitems, h->t2_kmax};
                if (!trans) {
                    if (upper)
                        trsv2_kernel<T, true, false><<<grid, 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvU, (const T*)h->d_wU, nullptr, nullptr, Xc, sy, epoch, nblk);
                    else
                        trsv2_kernel<T, false, false><<<grid, 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvL, (const T*)h->d_wL, Bc, h->d_perm, Xc, sy, epoch, nblk);
                } else {
                    if (upper)   // L^T z = y: block rows in descending order
                        trsv2_kernel<T, true, true><<<grid, 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvL, (const T*)h->d_wLt, nullptr, nullptr, Xc, sy, epoch, nblk);
                    else         // U^T y = b: block rows in ascending order, no row gather
                        trsv2_kernel<T, false, true><<<grid, 256, 0, st>>>(A, lda, n, (const T*)h->d_dinvU, (const T*)h->d_wUt, Bc, nullptr, Xc, sy, epoch, nblk);
                }
                LAUNCH_CHECK(h);
            }
            if (trans) {   // x = P^T z
                T* tmp = (T*)h->d_B;
                CU_TRY(h, cudaMemcpyAsync(tmp, Xc, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
                perm_scatter_kernel<T><<<cdiv(n, 256), 256, 0, st>>>(tmp, h->d_perm, Xc, n);
                LAUNCH_CHECK(h);
            }
        }
        return 0;
    }
    const int tile = (int)h->opt[B200LU_OPT_SOLVE_NRHS_TILE];
    const int NR = nrhs == 1 ? 1 : (tile >= 8 ? 8 : (tile >= 4 ? 4 : 1));
    const int groups = cdiv(nrhs, NR);
    int rc = ensure_trsv_groups(h, groups, NR);
    if (rc) return rc;
    TrsvSync sy{(unsigned long long*)h->d_tflags, h->d_tticket, h->d_deverr};
    // ---- many right-hand sides: blocked TRSM.  Diagonal blocks of `db` rows are solved by the
    // block-row kernel, everything off the diagonal is one tensor-core GEMM per block column
    // (X_rest -= L_panel X_j / U_panel X_j): 2 n^2 nrhs flops on the trailing-update kernels
    // instead of FMA mat-vecs.  (reference: getrs with a matrix right-hand side,
    // src/openblas.jl:247-278; `_naive_lu_ldiv!` matrix branch, src/factorization.jl:480-488)
    static const int db_env = getenv("B200LU_TRSM_DB") ? atoi(getenv("B200LU_TRSM_DB")) : 0;
    const int db = db_env >= 64 ? (db_env / 64) * 64 : 1024;   // measured: 128/256/512/1024 -> 4.06/3.52/3.22/3.02 ms at n = 8192, 100 rhs
    if (nrhs >= 16 && h->opt[B200LU_OPT_TRSV_MODE] == 0 && n > db && (ldx % 4) == 0 && (n % 4) == 0 &&
        (reinterpret_cast<uintptr_t>(X) % 16) == 0) {   // 16-byte cp.async chunks of the GEMM operands
        // P B -> X (B does not alias X here)
        perm_gather_kernel<T><<<dim3(cdiv(n, 256), grid_y(nrhs)), 256, 0, st>>>(B, ldb, h->d_perm, X, ldx, n, nrhs);
        LAUNCH_CHECK(h);
        const int nd = cdiv(n, db);
        auto diag_solve = [&](int j, bool upper) -> int {
            const int j0 = j * db, nj = std::min(db, n - j0), nb4 = cdiv(nj, TRSV_TB);
            if (h->trsv_epoch > (1u << 30)) {
                CU_TRY(h, cudaMemsetAsync(h->d_tflags, 0, h->cap_tflag_bytes, st));
                h->trsv_epoch = 0;
            }
            const unsigned epoch = ++h->trsv_epoch;
            const size_t boff = (size_t)(j0 / TRSV_TB) * TRSV_TB * TRSV_TB;
            const T* Ad = A + (int64_t)j0 * lda + j0;
            T* Xd = X + j0;
            dim3 g(nb4 * groups);
#define TRSV_SUB(NRV)                                                                             \
    if (upper)                                                                                    \
        trsv_block_kernel<T, NRV, true><<<g, 256, 0, st>>>(Ad, lda, nj, (const T*)h->d_dinvU + boff, (const T*)h->d_wU + boff, \
                                                           nullptr, 0, nullptr, Xd, ldx, nrhs, sy, epoch, nb4); \
    else                                                                                          \
        trsv_block_kernel<T, NRV, false><<<g, 256, 0, st>>>(Ad, lda, nj, (const T*)h->d_dinvL + boff, (const T*)h->d_wL + boff, \
                                                            nullptr, 0, nullptr, Xd, ldx, nrhs, sy, epoch, nb4);
            if (NR == 4) { TRSV_SUB(4) } else { TRSV_SUB(8) }
#undef TRSV_SUB
            LAUNCH_CHECK(h);
            return 0;
        };
        for (int j = 0; j < nd; ++j) {            // forward: L
            if ((rc = diag_solve(j, false))) return rc;
            const int j0 = j * db, j1 = std::min(n, j0 + db);
            if (j1 < n) {
                rc = launch_gemm(h, st, n - j1, nrhs, j1 - j0, A + (int64_t)j0 * lda + j1, lda, X + j0, ldx, X + j1, ldx);
                if (rc) return rc;
            }
        }
        for (int j = nd - 1; j >= 0; --j) {       // backward: U
            if ((rc = diag_solve(j, true))) return rc;
            const int j0 = j * db, j1 = std::min(n, j0 + db);
            if (j0 > 0) {
                rc = launch_gemm(h, st, j0, nrhs, j1 - j0, A + (int64_t)j0 * lda, lda, X + j0, ldx, X, ldx);
                if (rc) return rc;
            }
        }
        return 0;
    }
    dim3 grid(nblk * groups);
    for (int upper = 0; upper < 2; ++upper) {
        if (h->trsv_epoch > (1u << 30)) {
            CU_TRY(h, cudaMemsetAsync(h->d_tflags, 0, h->cap_tflag_bytes, st));
            h->trsv_epoch = 0;
        }
        const unsigned epoch = ++h->trsv_epoch;
        const T* dinv = (const T*)(upper ? h->d_dinvU : h->d_dinvL);
        const T* wmat = (const T*)(upper ? h->d_wU : h->d_wL);
#define TRSV_LAUNCH(NRV)                                                                          \
    if (upper)                                                                                    \
        trsv_block_kernel<T, NRV, true><<<grid, 256, 0, st>>>(A, lda, n, dinv, wmat, nullptr, 0, nullptr, \
                                                              X, ldx, nrhs, sy, epoch, nblk);     \
    else                                                                                          \
        trsv_block_kernel<T, NRV, false><<<grid, 256, 0, st>>>(A, lda, n, dinv, wmat, B, ldb, h->d_perm, \
                                                               X, ldx, nrhs, sy, epoch, nblk);
        if (NR == 1) { TRSV_LAUNCH(1) }
        else if (NR == 4) { TRSV_LAUNCH(4) }
        else { TRSV_LAUNCH(8) }
#undef TRSV_LAUNCH
        LAUNCH_CHECK(h);
    }
    return 0;
}

// X = A^{-1} B on the device with the cached factors, all in the factor type T.
// B and X may alias.
template <typename T>
static int getrs_device(b200lu_handle* h, const T* B, int64_t ldb, T* X, int64_t ldx, int nrhs, bool trans = false) {
    const int n = (int)h->n;
    cudaStream_t st = h->s_main;
    const T* src = B;
    int64_t lds = ldb;
    if ((const void*)B == (const void*)X && !trans) {
        // the row gather through the permutation is not an in-place operation: stage B
        int rc = ensure_rhs(h, nrhs);
        if (rc) return rc;
        CU_TRY(h, cudaMemcpy2DAsync(h->d_B, (size_t)n * sizeof(T), B, (size_t)ldb * sizeof(T),
                                    (size_t)n * sizeof(T), nrhs, cudaMemcpyDeviceToDevice, st));
        src = (const T*)h->d_B;
        lds = n;
    }
    return trsv_sweeps<T>(h, (const T*)h->dA, h->ldd, n, src, lds, X, ldx, nrhs, trans);
}

// MIXED: FP32 factors + FP64 residual refinement, one right-hand side at a time.
// MIXED with a matrix right-hand side: all columns refined together — FP32 getrs of the whole
// block (blocked TRSM on the tensor cores), residual R = B - A X as ONE FP64 GEMM on the DMMA
// kernel (A is read once per 64 columns instead of once per column), per-column norms, stop when
// the worst column has converged or stagnates.
static int refine_solve_block_device(b200lu_handle* h, const double* B, int64_t ldb, double* X,
                                     int64_t ldx, int nrhs) {
    const int n = (int)h->n;
    cudaStream_t st = h->s_main;
    const int maxit = (int)h->opt[B200LU_OPT_REFINE_MAXIT];
    const double eps = 2.220446049250313e-16;
    if (nrhs > h->cap_cscal) {
        CU_TRY(h, cudaStreamSynchronize(st));
        free_dev(h->d_cscal);
        if (h->h_cscal) cudaFreeHost(h->h_cscal);
        CU_TRY(h, cudaMalloc((void**)&h->d_cscal, (size_t)2 * nrhs * sizeof(double)));
        CU_TRY(h, cudaMallocHost((void**)&h->h_cscal, (size_t)2 * nrhs * sizeof(double)));
        h->cap_cscal = nrhs;
    }
    // scratch (the lower halves of d_B / d_X are free here, see b200lu_solve): R (FP64), then
    // two FP32 blocks W32 (rhs / correction input) and C32 (solution / correction output)
    int rc = ensure_rhs(h, nrhs + 1);
    if (rc) return rc;
    const int64_t ldr = ((n + 3) / 4) * 4;
    if ((size_t)ldr * nrhs * 8 > (size_t)h->cap_n * h->cap_rhs * 8) return set_err(h, 2, "refinement scratch too small");
    double* R = (double*)h->d_B;
    float* W32 = (float*)h->d_X;
    float* C32 = W32 + (size_t)ldr * nrhs;
    const dim3 g2(cdiv(n, 256), grid_y(nrhs));
    cast2d_kernel<double, float><<<g2, 256, 0, st>>>(B, ldb, W32, ldr, n, nrhs);
    LAUNCH_CHECK(h);
    rc = getrs_device<float>(h, W32, ldr, C32, ldr, nrhs);
    if (rc) return rc;
    cast2d_kernel<float, double><<<g2, 256, 0, st>>>(C32, ldr, X, ldx, n, nrhs);
    LAUNCH_CHECK(h);
    double prev = 1e300;
    h->last_refine_iters = 0;
    for (int it = 0; it < maxit; ++it) {
        // R = B - A X
        CU_TRY(h, cudaMemcpy2DAsync(R, (size_t)ldr * 8, B, (size_t)ldb * 8, (size_t)n * 8, nrhs, cudaMemcpyDeviceToDevice, st));
        rc = launch_gemm(h, st, n, nrhs, n, h->dA64, h->ldd, X, ldx, R, ldr);
        if (rc) return rc;
        colsumsq_kernel<<<std::min(nrhs, 4096), 256, 0, st>>>(R, ldr, n, h->d_cscal, nrhs);
        LAUNCH_CHECK(h);
        colsumsq_kernel<<<std::min(nrhs, 4096), 256, 0, st>>>(X, ldx, n, h->d_cscal + nrhs, nrhs);
        LAUNCH_CHECK(h);
        CU_TRY(h, cudaMemcpyAsync(h->h_cscal, h->d_cscal, (size_t)2 * nrhs * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
        double worst = 0.0;
        for (int c = 0; c < nrhs; ++c) {
            const double berr = sqrt(h->h_cscal[c]) / (h->normA_F * sqrt(h->h_cscal[nrhs + c]) + 1e-300);
            if (!(berr == berr)) return set_err(h, 3, "refinement produced NaN");
            worst = std::max(worst, berr);
        }
        h->last_refine_iters = std::max(h->last_refine_iters, it);
        if (!(worst > 2.0 * eps) || !(worst < 0.5 * prev)) break;
        prev = worst;
        cast2d_kernel<double, float><<<g2, 256, 0, st>>>(R, ldr, W32, ldr, n, nrhs);
        LAUNCH_CHECK(h);
        rc = getrs_device<float>(h, W32, ldr, C32, ldr, nrhs);
        if (rc) return rc;
        axpy_f32_cols_kernel<<<g2, 256, 0, st>>>(X, ldx, C32, ldr, n, nrhs);
        LAUNCH_CHECK(h);
        h->last_refine_iters = std::max(h->last_refine_iters, it + 1);
    }
    return 0;
}

// trans: op(A) = A^T — the same loop with the transposed FP32 sweeps (U^T, L^T, P^T scatter) and the
// residual r = b - A^T x (one contiguous dot product per column of A), one right-hand side at a time.
static int refine_solve_device(b200lu_handle* h, const double* B, int64_t ldb, double* X,
                               int64_t ldx, int nrhs, bool trans = false) {
    const int n = (int)h->n;
    cudaStream_t st = h->s_main;
    // matrix right-hand side: refine the block as a whole when the GEMM operands are 16-byte aligned
    // and B does not alias X (the residual needs the original B in every sweep)
    if (!trans && nrhs >= 4 && 2 * nrhs <= n && (const void*)B != (const void*)X && (ldx % 2) == 0 && (n % 2) == 0 &&
        (reinterpret_cast<uintptr_t>(X) % 16) == 0)
        return refine_solve_block_device(h, B, ldb, X, ldx, nrhs);
    const int maxit = (int)h->opt[B200LU_OPT_REFINE_MAXIT];
    const double eps = 2.220446049250313e-16;
    int rc = ensure_rhs(h, trans ? 2 : 1);
    if (rc) return rc;
    if ((rc = ensure_rpart(h, n))) return rc;
    float* w32 = (float*)h->d_X;  // n floats of scratch
    // the copy of b: column 0 of d_B, or column 1 when the transposed sweeps use column 0 for x = P^T z
    double* const bcopy = (double*)h->d_B + (trans ? h->cap_n : 0);
    h->last_refine_iters = 0;
    for (int c = 0; c < nrhs; ++c) {
        const double* b = B + (int64_t)c * ldb;
        double* x = X + (int64_t)c * ldx;
        // x0 = fl64( solve32( fl32(b) ) )
        cast2d_kernel<double, float><<<dim3(cdiv(n, 256), 1), 256, 0, st>>>(b, n, h->d_r32, n, n, 1);
        LAUNCH_CHECK(h);
        rc = getrs_device<float>(h, h->d_r32, n, w32, n, 1, trans);
        if (rc) return rc;
        // b may alias x: keep a copy of b in d_r's sibling (d_B, FP64)
        CU_TRY(h, cudaMemcpyAsync(bcopy, b, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        cast2d_kernel<float, double><<<dim3(cdiv(n, 256), 1), 256, 0, st>>>(w32, n, x, n, n, 1);
        LAUNCH_CHECK(h);
        double prev = 1e300;
        for (int it = 0; it < maxit; ++it) {
            // r = b - A x in FP64
            CU_TRY(h, cudaMemcpyAsync(h->d_r, bcopy, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
            if (trans) {
                residual_gemvT_kernel<double><<<cdiv(n, 8), 256, 0, st>>>(h->dA64, h->ldd, n, x, h->d_r);
            } else {
                const int cchunk = residual_chunk(n), nch = cdiv(n, cchunk);
                residual_gemv_kernel<double><<<dim3(cdiv(n, 256), nch), 256, 0, st>>>(h->dA64, h->ldd, n, x, h->d_rpart, cchunk);
                LAUNCH_CHECK(h);
                residual_finish_kernel<<<cdiv(n, 256), 256, 0, st>>>(h->d_rpart, n, nch, h->d_r);
            }
            LAUNCH_CHECK(h);
            sumsq_kernel<<<1, 1024, 0, st>>>(h->d_r, n, h->d_scal + 0);
            LAUNCH_CHECK(h);
            sumsq_kernel<<<1, 1024, 0, st>>>(x, n, h->d_scal + 1);
            LAUNCH_CHECK(h);
            CU_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 2 * sizeof(double),
                                      cudaMemcpyDeviceToHost, st));
            CU_TRY(h, cudaStreamSynchronize(st));
            const double rn = sqrt(h->h_scal[0]), xn = sqrt(h->h_scal[1]);
            const double berr = rn / (h->normA_F * xn + 1e-300);
            h->last_refine_iters = std::max(h->last_refine_iters, it);
            // converged to FP64 working accuracy, or stagnating
            if (!(berr > 2.0 * eps) || !(berr < 0.5 * prev)) {
                if (!(berr == berr)) return set_err(h, 3, "refinement produced NaN");
                break;
            }
            prev = berr;
            cast2d_kernel<double, float><<<dim3(cdiv(n, 256), 1), 256, 0, st>>>(h->d_r, n, h->d_r32, n, n, 1);
            LAUNCH_CHECK(h);
            rc = getrs_device<float>(h, h->d_r32, n, w32, n, 1, trans);
            if (rc) return rc;
            axpy_f32_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, w32, n);
            LAUNCH_CHECK(h);
            h->last_refine_iters = std::max(h->last_refine_iters, it + 1);
        }
    }
    return 0;
}

// ------------------------------------------------------------------- C ABI --
extern "C" {

int b200lu_version(void) { return B200LU_VERSION; }
int64_t b200lu_launch_count(void) { return (int64_t)g_launch_count.load(std::memory_order_relaxed); }

int b200lu_create(b200lu_handle** out, int dtype, int ngpus, const int* devices) {
    if (!out) return -1;
    *out = nullptr;
    if (dtype < 0 || dtype > 2) return -2;
    if (ngpus > 1) return team_create(out, dtype, ngpus, devices);   // one process, ngpus GPUs (dist.inc)
    if (ngpus != 1) return -3;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return 1;  // no CPU fallback
    const int dev = devices ? devices[0] : 0;
    if (dev < 0 || dev >= ndev) return -4;
    if (cudaSetDevice(dev) != cudaSuccess) return 1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 1;
    if (prop.major < 10) return 1;  // sm_100a code only
    b200lu_handle* h = new b200lu_handle();
    h->dtype = dtype;
    h->dev = dev;
    h->sms = prop.multiProcessorCount;
    h->opt[B200LU_OPT_NB] = 256;
    h->opt[B200LU_OPT_LOOKAHEAD] = 1;
    h->opt[B200LU_OPT_REFINE_MAXIT] = 10;
    h->opt[B200LU_OPT_PANEL_CTAS] = PANEL_GMAX;
    h->opt[B200LU_OPT_SOLVE_NRHS_TILE] = 8;
    h->opt[B200LU_OPT_PROFILE] = 0;
    h->opt[B200LU_OPT_PANEL_RPT] = 0;
    h->opt[B200LU_OPT_GEMM_CFG] = 3;
    h->opt[B200LU_OPT_PANEL_MODE] = 0;
    h->opt[B200LU_OPT_SGEMM_MODE] = 0;
    h->opt[B200LU_OPT_TRSV_MODE] = 0;
    h->opt[B200LU_OPT_STREAM_H2D] = 1;
    h->opt[B200LU_OPT_MAPPED_RHS] = 1;
    h->opt[B200LU_OPT_KEEP_A] = 0;
    h->opt[B200LU_OPT_BATCHED_MODE] = 0;
    h->opt[B200LU_OPT_HOST_REGISTER] = 0;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaStreamCreateWithPriority(&h->s_main, cudaStreamNonBlocking, lo) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&h->s_panel, cudaStreamNonBlocking, hi) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev_h2d) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev_a) == cudaSuccess && cudaEventCreate(&h->ev_b) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev_c) == cudaSuccess && cudaEventCreate(&h->ev_d) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_next, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_panelsync, sizeof(PanelMail)) == cudaSuccess;
    ok = ok && cudaMemset(h->d_panelsync, 0, sizeof(PanelMail)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_info, 64) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_deverr, 64) == cudaSuccess;
    ok = ok && cudaMemset(h->d_deverr, 0, 64) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_scal, 64) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&h->h_small, 64) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&h->h_scal, 64) == cudaSuccess;
    if (ok && getenv("B200LU_PANEL_DBG")) {
        ok = cudaMalloc((void**)&h->d_pdbg, 4096 * 24 * sizeof(long long)) == cudaSuccess;
        ok = ok && cudaMemset(h->d_pdbg, 0, 4096 * 24 * sizeof(long long)) == cudaSuccess;
    }
    if (!ok) {
        b200lu_destroy(h);
        return 1;
    }
    *out = h;
    return 0;
}

static void host_unregister(b200lu_handle* h);
void b200lu_destroy(b200lu_handle* h) {
    if (!h) return;
    if (h->team) {   // the sub-handles own every device resource
        team_destroy(h);
        host_unregister(h);
        delete h;
        return;
    }
    cudaSetDevice(h->dev);
    dist_destroy(h);
    if (h->s_main) cudaStreamSynchronize(h->s_main);
    if (h->s_panel) cudaStreamSynchronize(h->s_panel);
    if (h->s_copy) cudaStreamSynchronize(h->s_copy);
    host_unregister(h);
    if (h->d_pdbg) {
        const int cnt = std::min(h->pdbg_n, 4096);
        std::vector<long long> t((size_t)cnt * 24);
        cudaMemcpy(t.data(), h->d_pdbg, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const int step = std::max(1, cnt / 96);   // ~96 lines over the whole run (B200LU_PANEL_DBG=1)
        for (int i = 0; i < cnt; i += step) {
            const long long* s = &t[(size_t)i * 24];
            fprintf(stderr, "[pdbg] launch %d m=%lld G=%lld: load %lld loop %lld store %lld swaps %lld exit %lld cycles\n", i, s[6], s[7],
                    s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4]);
            if (s[16] | s[17]) fprintf(stderr, "[pdbg]    last sub-block prologue: load rows %lld | L11 + TRSM (phase A) %lld | rank-k update (phase B) %lld cycles; whole launch %lld\n", s[16], s[17], s[18], s[19]);
            if (s[8] | s[9]) fprintf(stderr, "[pdbg]    last sub-block, sums over its columns: publish-prep %lld | barrier+cta-reduce+send %lld | wait %lld | cluster-reduce %lld | freeze+positions+scale %lld | rank-1 update %lld | argmax-finish+stage %lld | loop-overhead %lld\n", s[8], s[9], s[10], s[11], s[14], s[15], s[12], s[13]);
        }
        cudaFree(h->d_pdbg);
    }
    if (h->d_t3dbg) {
        long long t[16 * 8];
        cudaMemcpy(t, h->d_t3dbg, sizeof(t), cudaMemcpyDeviceToHost);
        for (int r = 0; r < 16; ++r)
            if (t[r * 8 + 4])
                fprintf(stderr, "[trsv3] chain CTA %d: %lld rows; cycles per row: far sums+operands %lld | dinv + older x %lld | wait x(t-1) %lld | step %lld | issue next %lld ; loop %lld cycles in %lld ns = %.0f MHz\n",
                        r, t[r * 8 + 4], t[r * 8] / t[r * 8 + 4], t[r * 8 + 1] / t[r * 8 + 4], t[r * 8 + 2] / t[r * 8 + 4], t[r * 8 + 3] / t[r * 8 + 4],
                        t[r * 8 + 5] / t[r * 8 + 4], t[r * 8 + 6], t[r * 8 + 7], 1e3 * (double)t[r * 8 + 6] / (double)std::max(1LL, t[r * 8 + 7]));
        cudaFree(h->d_t3dbg);
    }
    free_dev(h->dA); free_dev(h->dA64); free_dev(h->d_ipiv); free_dev(h->d_perm);
    free_dev(h->d_info); free_dev(h->d_deverr); free_dev(h->d_plans); free_dev(h->d_panelsync); free_dev(h->d_split);
    free_dev(h->d_dinvL); free_dev(h->d_dinvU); free_dev(h->d_wL); free_dev(h->d_wU); free_dev(h->d_wLt); free_dev(h->d_wUt); free_dev(h->d_tflags); free_dev(h->d_tticket); free_dev(h->d_t2items); free_dev(h->d_t2x); free_dev(h->d_t2p); free_dev(h->d_t2ticket);
    free_dev(h->d_t3items); free_dev(h->d_t3x); free_dev(h->d_t3p); free_dev(h->d_t3ticket);
    free_dev(h->dA_keep);
    free_dev(h->d_B); free_dev(h->d_X); free_dev(h->d_r); free_dev(h->d_r32); free_dev(h->d_scal); free_dev(h->d_cscal); free_dev(h->d_rpart);
    if (h->h_cscal) cudaFreeHost(h->h_cscal);
    free_dev(h->dB_LU); free_dev(h->dB_ipiv); free_dev(h->dB_info); free_dev(h->dB_perm); free_dev(h->dB_in);
    free_dev(h->dB_rhs); free_dev(h->dB_x);
    if (h->h_ipiv) cudaFreeHost(h->h_ipiv);
    if (h->h_rhs) cudaFreeHost(h->h_rhs);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->hB_stage) cudaFreeHost(h->hB_stage);
    for (cudaEvent_t e : h->ev_panel) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->chain_ev) cudaEventDestroy(e);
    cudaEvent_t evs[] = {h->ev_a, h->ev_b, h->ev_c, h->ev_d, h->ev_fork, h->ev_next, h->ev_h2d};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (h->s_main) cudaStreamDestroy(h->s_main);
    if (h->s_panel) cudaStreamDestroy(h->s_panel);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    delete h;
}

const char* b200lu_last_error(const b200lu_handle* h) { return h ? h->err : "null handle"; }

double b200lu_last_timing(const b200lu_handle* h, int phase) {
    if (!h || phase < 0 || phase >= B200LU_T_COUNT) return -1.0;
    return h->timing[phase];
}

double b200lu_last_counter(const b200lu_handle* h, int which) {
    if (!h || which < 0 || which >= B200LU_C_COUNT) return -1.0;
    if (which == B200LU_C_REFINE_ITERS) return (double)h->last_refine_iters;
    return h->counters[which];
}

int b200lu_debug_gemm_sub(b200lu_handle* h, int64_t M, int64_t N, int64_t K, const void* dA, int64_t lda,
                          const void* dB, int64_t ldb, void* dC, int64_t ldc) {
    if (!h) return -1;
    TEAM_ONLY_HOST(h);
    if (M < 0 || N < 0 || K < 0 || !dA || !dB || !dC) return set_err(h, -2, "bad gemm arguments");
    CU_TRY(h, cudaSetDevice(h->dev));
    int rc;
    if (h->dtype == B200LU_F64)
        rc = launch_gemm(h, h->s_main, (int)M, (int)N, (int)K, (const double*)dA, lda, (const double*)dB, ldb, (double*)dC, ldc);
    else
        rc = launch_gemm(h, h->s_main, (int)M, (int)N, (int)K, (const float*)dA, lda, (const float*)dB, ldb, (float*)dC, ldc);
    if (rc) return rc;
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    int de = 0;
    CU_TRY(h, cudaMemcpy(&de, h->d_deverr, sizeof(int), cudaMemcpyDeviceToHost));
    if (de) return set_err(h, 1, "device-side error %d in gemm", de);
    return 0;
}

int b200lu_probe_peak(b200lu_handle* h, int kind, double* out) {
    if (!h || !out) return -1;
    if (h->team) return b200lu_probe_peak(team_first(h), kind, out);
    CU_TRY(h, cudaSetDevice(h->dev));
    cudaStream_t st = h->s_main;
    int sms = 0;
    CU_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->dev));
    double* sink = nullptr;
    CU_TRY(h, cudaMalloc((void**)&sink, 1 << 20));
    float best = 1e30f;
    if (kind == B200LU_PEAK_FP64_DMMA || kind == B200LU_PEAK_FP64_DFMA) {
        const int iters = 4096, ctas = sms * 4, thr = 256;
        for (int rep = 0; rep < 5; ++rep) {
            CU_TRY(h, cudaEventRecord(h->ev_a, st));
            if (kind == B200LU_PEAK_FP64_DMMA) probe_dmma_kernel<<<ctas, thr, 0, st>>>(sink, iters);
            else probe_dfma_kernel<<<ctas, thr, 0, st>>>(sink, iters);
            LAUNCH_CHECK(h);
            CU_TRY(h, cudaEventRecord(h->ev_b, st));
            CU_TRY(h, cudaStreamSynchronize(st));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
            if (rep > 0) best = std::min(best, ms);
        }
        // per thread per iteration: DMMA probe = 16 m8n8k4 (256 FMA / 32 lanes each); DFMA probe = 16 FMAs
        const double fma_per_thread_iter = (kind == B200LU_PEAK_FP64_DMMA) ? 16.0 * 8.0 : 16.0;
        *out = 2.0 * fma_per_thread_iter * iters * (double)ctas * thr / (best * 1e-3) / 1e12;
    } else if (kind == B200LU_PEAK_HBM_COPY) {
        const size_t bytes = (size_t)1 << 30;
        void *a = nullptr, *b = nullptr;
        CU_TRY(h, cudaMalloc(&a, bytes));
        CU_TRY(h, cudaMalloc(&b, bytes));
        for (int rep = 0; rep < 5; ++rep) {
            CU_TRY(h, cudaEventRecord(h->ev_a, st));
            CU_TRY(h, cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, st));
            CU_TRY(h, cudaEventRecord(h->ev_b, st));
            CU_TRY(h, cudaStreamSynchronize(st));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
            if (rep > 0) best = std::min(best, ms);
        }
        cudaFree(a);
        cudaFree(b);
        *out = 2.0 * (double)bytes / (best * 1e-3) / 1e9;
    } else {
        cudaFree(sink);
        return set_err(h, -2, "unknown peak kind");
    }
    cudaFree(sink);
    return 0;
}

int b200lu_set_option(b200lu_handle* h, int option, int64_t value) {
    if (!h) return -1;
    if (option < 0 || option >= B200LU_OPT_COUNT) return -2;
    if (option == B200LU_OPT_NB) {
        if (value < 16 || value > 256 || value % 16 != 0) return -3;
    }
    if (option == B200LU_OPT_PANEL_CTAS && (value < 1 || value > PANEL_GMAX)) return -3;
    if (option == B200LU_OPT_REFINE_MAXIT && value < 0) return -3;
    if (option == B200LU_OPT_PANEL_RPT && (value < 0 || value > 2)) return -3;
    if (option == B200LU_OPT_GEMM_CFG && (value < 0 || value > 3)) return -3;
    if (option == B200LU_OPT_PANEL_MODE && (value < 0 || value > 2)) return -3;
    if (option == B200LU_OPT_SGEMM_MODE && (value < 0 || value > 2)) return -3;
    if (option == B200LU_OPT_TRSV_MODE && (value < 0 || value > 3)) return -3;
    if (option == B200LU_OPT_STREAM_H2D && (value < 0 || value > 1)) return -3;
    if (option == B200LU_OPT_HOST_REGISTER && (value < 0 || value > 1)) return -3;
    if (option == B200LU_OPT_MAPPED_RHS && (value < 0 || value > 1)) return -3;
    if (option == B200LU_OPT_KEEP_A && (value < 0 || value > 1)) return -3;
    if (option == B200LU_OPT_BATCHED_MODE && (value < 0 || value > 1)) return -3;
    if (option == B200LU_OPT_NB) h->nb_user = true;
    if (h->team) return team_set_option(h, option, value);
    h->opt[option] = value;
    return 0;
}
int64_t b200lu_get_option(const b200lu_handle* h, int option) {
    if (!h || option < 0 || option >= B200LU_OPT_COUNT) return -1;
    return h->opt[option];
}

// B200LU_OPT_KEEP_A: a device copy of A (same layout as dA) taken before the factorization
// overwrites it, so that the residual check of src/factorization.jl:127-156 can run on the device
// (b200lu_residual_norms).  MIXED handles keep A in dA64 anyway.
static int keep_copy_of_A(b200lu_handle* h, int64_t n) {
    h->keep_valid = false;
    if (!h->opt[B200LU_OPT_KEEP_A] || h->dtype == B200LU_MIXED) return 0;
    const int64_t need = h->ldd * n * (int64_t)elem_size(h);
    if (need > h->cap_keep) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dA_keep);
        h->cap_keep = 0;
        CU_TRY(h, cudaMalloc(&h->dA_keep, (size_t)need));
        h->cap_keep = need;
    }
    CU_TRY(h, cudaMemcpyAsync(h->dA_keep, h->dA, (size_t)need, cudaMemcpyDeviceToDevice, h->s_main));
    h->keep_valid = true;
    return 0;
}

// factor whatever already sits in h->dA (and dA64 for MIXED)
static int factor_resident(b200lu_handle* h, int64_t n, int64_t* info, int nchunks = 0) {
    int rc;
    h->factored = false;
    h->solve_ready = false;
    h->solve_ready_t = false;
    h->prof_used = 0;
    h->prof_flops = 0.0;
    CU_TRY(h, cudaEventRecord(h->ev_b, h->s_main));
    if (h->dtype == B200LU_F64) {
        rc = getrf_device<double>(h, (double*)h->dA, h->ldd, (int)n, nchunks);
    } else {
        if (h->dtype == B200LU_MIXED) {
            if ((rc = ensure_rpart(h, n))) return rc;
            CU_TRY(h, cudaMemsetAsync(h->d_scal, 0, 4 * sizeof(double), h->s_main));
            sumsq2d_kernel<<<dim3(cdiv(n, 256), 256), 256, 0, h->s_main>>>(h->dA64, h->ldd, (int)n, h->d_rpart);
            LAUNCH_CHECK(h);
            sum_partials_kernel<<<1, 1024, 0, h->s_main>>>(h->d_rpart, (long long)cdiv(n, 256) * 256, h->d_scal + 2);
            LAUNCH_CHECK(h);
            cast2d_kernel<double, float><<<dim3(cdiv(n, 256), 512), 256, 0, h->s_main>>>(
                h->dA64, h->ldd, (float*)h->dA, h->ldd, (int)n, (int)n);
            LAUNCH_CHECK(h);
        }
        rc = getrf_device<float>(h, (float*)h->dA, h->ldd, (int)n, nchunks);
    }
    if (rc) return rc;
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    CU_TRY(h, cudaMemcpyAsync(h->h_small, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaMemcpyAsync(h->h_small + 1, h->d_deverr, sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
    if (h->dtype == B200LU_MIXED)
        CU_TRY(h, cudaMemcpyAsync(h->h_scal + 2, h->d_scal + 2, sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    if (h->opt[B200LU_OPT_LOOKAHEAD]) CU_TRY(h, cudaStreamSynchronize(h->s_panel));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
    h->timing[B200LU_T_FACTOR] = ms;
    if (h->h_small[1] != 0) {
        cudaMemset(h->d_deverr, 0, sizeof(int));
        return set_err(h, 4, "device watchdog fired during getrf (code %d)", h->h_small[1]);
    }
    if (h->dtype == B200LU_MIXED) h->normA_F = sqrt(h->h_scal[2]);
    {
        double gms = 0.0;
        for (int i = 0; i + 1 < h->prof_used; i += 2) {
            float t = 0.f;
            cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]);
            gms += t;
        }
        h->timing[B200LU_T_GEMM] = gms;
        h->counters[B200LU_C_GEMM_FLOPS] = h->prof_flops;
        h->counters[B200LU_C_GEMM_LAUNCHES] = h->prof_used / 2;
    }
    h->info = h->h_small[0];
    h->factored = true;
    if (info) *info = h->info;
    return 0;
}

// B200LU_OPT_HOST_REGISTER: page-lock the caller's matrix once; a buffer that is already pinned (cudaMallocHost,
// cudaHostRegister by the caller, or by an earlier call) is left alone.  A failed registration is not an error:
// the upload then goes through the driver's staging like any pageable copy.
static void host_unregister(b200lu_handle* h) {
    if (h->reg_ptr) {
        cudaHostUnregister(h->reg_ptr);
        (void)cudaGetLastError();
        h->reg_ptr = nullptr;
        h->reg_bytes = 0;
    }
}
static void host_register(b200lu_handle* h, const void* A_host, size_t bytes) {
    if (!h->opt[B200LU_OPT_HOST_REGISTER] || !A_host || bytes == 0) return;
    if (h->reg_ptr == A_host && h->reg_bytes >= bytes) return;
    cudaSetDevice(h->dev);
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, A_host) != cudaSuccess) {
        (void)cudaGetLastError();
        return;
    }
    if (at.type != cudaMemoryTypeUnregistered && h->reg_ptr != A_host) return;   // pinned by somebody else
    host_unregister(h);
    if (cudaHostRegister(const_cast<void*>(A_host), bytes, cudaHostRegisterPortable) == cudaSuccess) {
        h->reg_ptr = const_cast<void*>(A_host);
        h->reg_bytes = bytes;
    } else {
        (void)cudaGetLastError();
    }
}

int b200lu_factor(b200lu_handle* h, int64_t n, const void* A_host, int64_t lda, int64_t* ipiv_out,
                  int64_t* info) {
    if (!h) return -1;
    if (n > 0 && A_host && lda >= n && n <= 131072)
        host_register(h, A_host, ((size_t)lda * (size_t)(n - 1) + (size_t)n) * iface_size(h));
    if (h->team) return team_factor(h, n, A_host, lda, ipiv_out, info);
    if (n < 0 || n > 131072) return set_err(h, -2, "n out of range");
    if (!A_host && n > 0) return set_err(h, -3, "A is NULL");
    if (lda < std::max<int64_t>(1, n)) return set_err(h, -4, "lda < max(1,n)");
    if (info) *info = 0;
    if (n == 0) { h->n = 0; h->factored = true; h->info = 0; return 0; }
    CU_TRY(h, cudaSetDevice(h->dev));
    int rc = ensure_capacity(h, n);
    if (rc) return rc;
    const size_t is = iface_size(h);
    void* dst = (h->dtype == B200LU_MIXED) ? (void*)h->dA64 : h->dA;
    CU_TRY(h, cudaEventRecord(h->ev_a, h->s_main));
    // streamed upload: column chunks on the copy stream, the factorization starts under them
    const int nb = (int)h->opt[B200LU_OPT_NB];
    int nchunks = 0;
    if (h->opt[B200LU_OPT_STREAM_H2D] && h->opt[B200LU_OPT_LOOKAHEAD] && h->dtype != B200LU_MIXED && n >= 2048 &&
        !h->opt[B200LU_OPT_KEEP_A]) {   // the kept copy is taken from the complete upload
        static const int want = getenv("B200LU_H2D_CHUNKS") ? std::max(2, atoi(getenv("B200LU_H2D_CHUNKS"))) : 8;
        const int cw = cdiv(cdiv((int)n, want), nb) * nb;
        const int first = std::min(cw, 2 * nb);   // a short first chunk: the first panel starts early
        nchunks = 1 + cdiv((int)n - first, cw);
        if (nchunks < 2) nchunks = 0;
        else {
            h->chunk_end.resize(nchunks);
            while ((int)h->ev_chunk.size() < nchunks) {
                cudaEvent_t e;
                CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->ev_chunk.push_back(e);
            }
            CU_TRY(h, cudaStreamWaitEvent(h->s_copy, h->ev_a, 0));
            for (int c = 0; c < nchunks; ++c) {
                const int64_t c0 = c == 0 ? 0 : first + (int64_t)(c - 1) * cw;
                const int64_t c1 = c == 0 ? first : std::min<int64_t>(n, c0 + cw);
                h->chunk_end[c] = (int)c1;
                CU_TRY(h, cudaMemcpy2DAsync((char*)dst + (size_t)c0 * h->ldd * is, (size_t)h->ldd * is,
                                            (const char*)A_host + (size_t)c0 * lda * is, (size_t)lda * is,
                                            (size_t)n * is, (size_t)(c1 - c0), cudaMemcpyHostToDevice, h->s_copy));
                CU_TRY(h, cudaEventRecord(h->ev_chunk[c], h->s_copy));
            }
            CU_TRY(h, cudaEventRecord(h->ev_h2d, h->s_copy));
        }
    }
    if (nchunks == 0)
        CU_TRY(h, cudaMemcpy2DAsync(dst, (size_t)h->ldd * is, A_host, (size_t)lda * is, (size_t)n * is,
                                    (size_t)n, cudaMemcpyHostToDevice, h->s_main));
    rc = keep_copy_of_A(h, n);
    if (rc) return rc;
    rc = factor_resident(h, n, info, nchunks);
    if (rc) {
        if (nchunks) cudaStreamSynchronize(h->s_copy);   // the caller's buffer must not be read after we return
        return rc;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, nchunks ? h->ev_h2d : h->ev_b);
    h->timing[B200LU_T_H2D] = ms;
    if (ipiv_out) {
        rc = b200lu_get_ipiv(h, ipiv_out);
        if (rc) return rc;
    }
    return 0;
}

int b200lu_factor_device(b200lu_handle* h, int64_t n, const void* A_dev, int64_t lda, int64_t* info) {
    if (!h) return -1;
    TEAM_ONLY_HOST(h);
    if (n < 0 || n > 131072) return set_err(h, -2, "n out of range");
    if (!A_dev && n > 0) return set_err(h, -3, "A is NULL");
    if (lda < std::max<int64_t>(1, n)) return set_err(h, -4, "lda < max(1,n)");
    if (info) *info = 0;
    if (n == 0) { h->n = 0; h->factored = true; h->info = 0; return 0; }
    CU_TRY(h, cudaSetDevice(h->dev));
    int rc = ensure_capacity(h, n);
    if (rc) return rc;
    const size_t is = iface_size(h);
    void* dst = (h->dtype == B200LU_MIXED) ? (void*)h->dA64 : h->dA;
    CU_TRY(h, cudaEventRecord(h->ev_a, h->s_main));
    if (A_dev != dst)
        CU_TRY(h, cudaMemcpy2DAsync(dst, (size_t)h->ldd * is, A_dev, (size_t)lda * is, (size_t)n * is,
                                    (size_t)n, cudaMemcpyDeviceToDevice, h->s_main));
    rc = keep_copy_of_A(h, n);
    if (rc) return rc;
    rc = factor_resident(h, n, info);
    if (rc) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
    h->timing[B200LU_T_H2D] = ms;  // device-to-device staging of A into the factor buffer
    return 0;
}

static int check_solve_args(b200lu_handle* h, char trans, int64_t nrhs, const void* B, int64_t ldb,
                            void* X, int64_t ldx) {
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');   // real types: 'C' == 'T'
    if (!tr && trans != 'N' && trans != 'n') return set_err(h, -2, "trans='%c' is not one of N, T, C", trans);
    if (nrhs < 0) return set_err(h, -3, "nrhs < 0");
    if (!h->factored) return set_err(h, 3, "no factorization cached");
    if (h->info != 0) return set_err(h, 3, "cached factorization is singular (info=%lld)", (long long)h->info);
    if (h->n > 0 && nrhs > 0) {
        if (!B) return set_err(h, -4, "B is NULL");
        if (ldb < h->n) return set_err(h, -5, "ldb < n");
        if (!X) return set_err(h, -6, "X is NULL");
        if (ldx < h->n) return set_err(h, -7, "ldx < n");
    }
    return 0;
}

static int finish_solve(b200lu_handle* h) {
    CU_TRY(h, cudaMemcpyAsync(h->h_small + 1, h->d_deverr, sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    if (h->h_small[1] != 0) {
        cudaMemset(h->d_deverr, 0, sizeof(int));
        return set_err(h, 4, "device watchdog fired during getrs (code %d)", h->h_small[1]);
    }
    return 0;
}

int b200lu_solve_device(b200lu_handle* h, char trans, int64_t nrhs, const void* B_dev, int64_t ldb,
                        void* X_dev, int64_t ldx) {
    TEAM_ONLY_HOST(h);
    int rc = check_solve_args(h, trans, nrhs, B_dev, ldb, X_dev, ldx);
    if (rc) return rc;
    if (h->n == 0 || nrhs == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->dev));
    CU_TRY(h, cudaEventRecord(h->ev_b, h->s_main));
    if (h->dtype == B200LU_F64)
        rc = getrs_device<double>(h, (const double*)B_dev, ldb, (double*)X_dev, ldx, (int)nrhs, trans != 'N' && trans != 'n');
    else if (h->dtype == B200LU_F32)
        rc = getrs_device<float>(h, (const float*)B_dev, ldb, (float*)X_dev, ldx, (int)nrhs, trans != 'N' && trans != 'n');
    else
        rc = refine_solve_device(h, (const double*)B_dev, ldb, (double*)X_dev, ldx, (int)nrhs, trans != 'N' && trans != 'n');
    if (rc) return rc;
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    rc = finish_solve(h);
    if (rc) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
    h->timing[B200LU_T_SOLVE] = ms;
    return 0;
}

int b200lu_solve(b200lu_handle* h, char trans, int64_t nrhs, const void* B_host, int64_t ldb,
                 void* X_host, int64_t ldx) {
    if (h && h->team) return team_solve(h, trans, nrhs, B_host, ldb, X_host, ldx);
    int rc = check_solve_args(h, trans, nrhs, B_host, ldb, X_host, ldx);
    if (rc) return rc;
    if (h->n == 0 || nrhs == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->dev));
    const bool tr = !(trans == 'N' || trans == 'n');
    if (nrhs == 1 && !tr && h->dtype != B200LU_MIXED && h->n >= 256 && h->opt[B200LU_OPT_MAPPED_RHS]) {
        // one right-hand side: the getrs kernels read b and write x in mapped host memory themselves
        const int64_t n = h->n;
        const size_t is = iface_size(h);
        if (n > h->cap_hrhs) {
            CU_TRY(h, cudaStreamSynchronize(h->s_main));
            if (h->h_rhs) cudaFreeHost(h->h_rhs);
            h->h_rhs = nullptr;
            h->cap_hrhs = 0;
            CU_TRY(h, cudaHostAlloc((void**)&h->h_rhs, (size_t)2 * h->cap_n * 8, cudaHostAllocMapped));
            CU_TRY(h, cudaHostGetDevicePointer((void**)&h->d_rhs_map, h->h_rhs, 0));
            h->cap_hrhs = h->cap_n;
        }
        memcpy(h->h_rhs, B_host, (size_t)n * is);
        char* dB = h->d_rhs_map;
        char* dX = h->d_rhs_map + (size_t)h->cap_hrhs * 8;
        CU_TRY(h, cudaEventRecord(h->ev_b, h->s_main));
        if (h->dtype == B200LU_F64) rc = getrs_device<double>(h, (const double*)dB, n, (double*)dX, n, 1, false);
        else rc = getrs_device<float>(h, (const float*)dB, n, (float*)dX, n, 1, false);
        if (rc) return rc;
        CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
        rc = finish_solve(h);
        if (rc) return rc;
        memcpy(X_host, h->h_rhs + (size_t)h->cap_hrhs * 8, (size_t)n * is);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
        h->timing[B200LU_T_SOLVE] = ms;
        h->timing[B200LU_T_H2D] = 0.0;   // no copies: the kernels touch the mapped buffer
        h->timing[B200LU_T_D2H] = 0.0;
        return 0;
    }
    // d_B/d_X are also scratch of the in-place / refinement paths: keep the host
    // staging in the upper half of a 2*nrhs allocation
    rc = ensure_rhs(h, 2 * nrhs + 2);
    if (rc) return rc;
    const size_t is = iface_size(h);
    const int64_t n = h->n;
    char* dBs = (char*)h->d_B + (size_t)h->cap_n * (nrhs + 1) * 8;
    char* dXs = (char*)h->d_X + (size_t)h->cap_n * (nrhs + 1) * 8;
    CU_TRY(h, cudaEventRecord(h->ev_a, h->s_main));
    CU_TRY(h, cudaMemcpy2DAsync(dBs, (size_t)n * is, B_host, (size_t)ldb * is, (size_t)n * is,
                                (size_t)nrhs, cudaMemcpyHostToDevice, h->s_main));
    CU_TRY(h, cudaEventRecord(h->ev_b, h->s_main));
    if (h->dtype == B200LU_F64)
        rc = getrs_device<double>(h, (const double*)dBs, n, (double*)dXs, n, (int)nrhs, trans != 'N' && trans != 'n');
    else if (h->dtype == B200LU_F32)
        rc = getrs_device<float>(h, (const float*)dBs, n, (float*)dXs, n, (int)nrhs, trans != 'N' && trans != 'n');
    else
        rc = refine_solve_device(h, (const double*)dBs, n, (double*)dXs, n, (int)nrhs, trans != 'N' && trans != 'n');
    if (rc) return rc;
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    CU_TRY(h, cudaMemcpy2DAsync(X_host, (size_t)ldx * is, dXs, (size_t)n * is, (size_t)n * is,
                                (size_t)nrhs, cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaEventRecord(h->ev_d, h->s_main));
    rc = finish_solve(h);
    if (rc) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_b); h->timing[B200LU_T_H2D] = ms;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c); h->timing[B200LU_T_SOLVE] = ms;
    cudaEventElapsedTime(&ms, h->ev_c, h->ev_d); h->timing[B200LU_T_D2H] = ms;
    return 0;
}

// Residual norms on the device: resid[c] = ||B[:, c] - A X[:, c]||_2 and bnorm[c] = ||B[:, c]||_2 in FP64,
// with the copy of A kept by B200LU_OPT_KEEP_A (F64/F32) or the FP64 copy of the MIXED handle.
int b200lu_residual_norms(b200lu_handle* h, int64_t nrhs, const void* B_host, int64_t ldb,
                          const void* X_host, int64_t ldx, double* resid_out, double* bnorm_out) {
    if (!h) return -1;
    if (h->team) return set_err(h, 3, "the residual check on the device needs a single-GPU handle (no copy of A is kept on a multi-GPU handle)");
    if (nrhs < 0) return set_err(h, -2, "nrhs < 0");
    if (!h->factored) return set_err(h, 3, "no factorization cached");
    const bool have_A = (h->dtype == B200LU_MIXED) ? (h->dA64 != nullptr) : h->keep_valid;
    if (!have_A) return set_err(h, 3, "no copy of A on the device: set B200LU_OPT_KEEP_A = 1 before b200lu_factor");
    const int64_t n = h->n;
    if (n > 0 && nrhs > 0) {
        if (!B_host) return set_err(h, -3, "B is NULL");
        if (ldb < n) return set_err(h, -4, "ldb < n");
        if (!X_host) return set_err(h, -5, "X is NULL");
        if (ldx < n) return set_err(h, -6, "ldx < n");
        if (!resid_out || !bnorm_out) return set_err(h, -7, "output is NULL");
    }
    if (nrhs == 0) return 0;
    if (n == 0) {
        for (int64_t c = 0; c < nrhs; ++c) resid_out[c] = bnorm_out[c] = 0.0;
        return 0;
    }
    CU_TRY(h, cudaSetDevice(h->dev));
    cudaStream_t st = h->s_main;
    int rc = ensure_rhs(h, 2 * nrhs + 2);
    if (rc) return rc;
    if ((rc = ensure_rpart(h, n))) return rc;
    if (nrhs > h->cap_cscal) {
        CU_TRY(h, cudaStreamSynchronize(st));
        free_dev(h->d_cscal);
        if (h->h_cscal) cudaFreeHost(h->h_cscal);
        h->h_cscal = nullptr;
        h->cap_cscal = 0;
        CU_TRY(h, cudaMalloc((void**)&h->d_cscal, (size_t)2 * nrhs * sizeof(double)));
        CU_TRY(h, cudaMallocHost((void**)&h->h_cscal, (size_t)2 * nrhs * sizeof(double)));
        h->cap_cscal = (int)nrhs;
    }
    const size_t is = iface_size(h);
    // interface-typed staging in the upper halves of d_B / d_X, FP64 working copies R (= B, then the
    // residual) and X64 in the lower halves
    char* dBs = (char*)h->d_B + (size_t)h->cap_n * (nrhs + 1) * 8;
    char* dXs = (char*)h->d_X + (size_t)h->cap_n * (nrhs + 1) * 8;
    double* R = (double*)h->d_B;
    double* X64 = (double*)h->d_X;
    CU_TRY(h, cudaEventRecord(h->ev_a, st));
    CU_TRY(h, cudaMemcpy2DAsync(dBs, (size_t)n * is, B_host, (size_t)ldb * is, (size_t)n * is, (size_t)nrhs, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpy2DAsync(dXs, (size_t)n * is, X_host, (size_t)ldx * is, (size_t)n * is, (size_t)nrhs, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaEventRecord(h->ev_b, st));
    const dim3 g2(cdiv(n, 256), grid_y(nrhs));
    if (h->dtype == B200LU_F32) {
        cast2d_kernel<float, double><<<g2, 256, 0, st>>>((const float*)dBs, n, R, n, (int)n, (int)nrhs);
        LAUNCH_CHECK(h);
        cast2d_kernel<float, double><<<g2, 256, 0, st>>>((const float*)dXs, n, X64, n, (int)n, (int)nrhs);
        LAUNCH_CHECK(h);
    } else {
        CU_TRY(h, cudaMemcpyAsync(R, dBs, (size_t)n * nrhs * 8, cudaMemcpyDeviceToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(X64, dXs, (size_t)n * nrhs * 8, cudaMemcpyDeviceToDevice, st));
    }
    const int gn = (int)std::min<int64_t>(nrhs, 4096);
    colsumsq_kernel<<<gn, 256, 0, st>>>(R, n, (int)n, h->d_cscal + nrhs, (int)nrhs);     // ||b||^2 before R becomes the residual
    LAUNCH_CHECK(h);
    const int cchunk = residual_chunk(n), nch = cdiv(n, cchunk);
    for (int64_t c = 0; c < nrhs; ++c) {
        const dim3 gr(cdiv(n, 256), nch);
        if (h->dtype == B200LU_F64)
            residual_gemv_kernel<double><<<gr, 256, 0, st>>>((const double*)h->dA_keep, h->ldd, (int)n, X64 + c * n, h->d_rpart, cchunk);
        else if (h->dtype == B200LU_F32)
            residual_gemv_kernel<float><<<gr, 256, 0, st>>>((const float*)h->dA_keep, h->ldd, (int)n, X64 + c * n, h->d_rpart, cchunk);
        else
            residual_gemv_kernel<double><<<gr, 256, 0, st>>>(h->dA64, h->ldd, (int)n, X64 + c * n, h->d_rpart, cchunk);
        LAUNCH_CHECK(h);
        residual_finish_kernel<<<cdiv(n, 256), 256, 0, st>>>(h->d_rpart, (int)n, nch, R + c * n);
        LAUNCH_CHECK(h);
    }
    colsumsq_kernel<<<gn, 256, 0, st>>>(R, n, (int)n, h->d_cscal, (int)nrhs);
    LAUNCH_CHECK(h);
    CU_TRY(h, cudaEventRecord(h->ev_c, st));
    CU_TRY(h, cudaMemcpyAsync(h->h_cscal, h->d_cscal, (size_t)2 * nrhs * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaEventRecord(h->ev_d, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    for (int64_t c = 0; c < nrhs; ++c) {
        resid_out[c] = sqrt(h->h_cscal[c]);
        bnorm_out[c] = sqrt(h->h_cscal[nrhs + c]);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_b); h->timing[B200LU_T_H2D] = ms;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c); h->timing[B200LU_T_SOLVE] = ms;
    cudaEventElapsedTime(&ms, h->ev_c, h->ev_d); h->timing[B200LU_T_D2H] = ms;
    return 0;
}

int b200lu_get_factors(b200lu_handle* h, void* LU_host, int64_t ldlu) {
    if (!h) return -1;
    if (h->team) return team_get_factors(h, LU_host, ldlu);
    if (!h->factored) return set_err(h, 3, "no factorization cached");
    if (!LU_host) return set_err(h, -2, "LU is NULL");
    if (ldlu < h->n) return set_err(h, -3, "ldlu < n");
    if (h->n == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->dev));
    const size_t es = elem_size(h);
    CU_TRY(h, cudaMemcpy2DAsync(LU_host, (size_t)ldlu * es, h->dA, (size_t)h->ldd * es,
                                (size_t)h->n * es, (size_t)h->n, cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    return 0;
}

int b200lu_get_ipiv(b200lu_handle* h, int64_t* ipiv_out) {
    if (!h) return -1;
    if (h->team) {
        if (!h->factored) return set_err(h, 3, "no factorization cached");
        b200lu_handle* s0 = team_first(h);
        const int rc = b200lu_get_ipiv(s0, ipiv_out);
        return rc ? set_err(h, rc, "%s", s0->err) : 0;
    }
    if (!h->factored && !h->factored_dist) return set_err(h, 3, "no factorization cached");
    if (!ipiv_out) return set_err(h, -2, "ipiv is NULL");
    if (h->n == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->dev));
    // widen on the device into d_X-free scratch: reuse d_plans? no — a dedicated pass via pinned ints
    const int n = (int)h->n;
    int* tmp = (int*)h->h_ipiv;  // pinned, n*8 bytes: first n ints as staging
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    CU_TRY(h, cudaMemcpyAsync(tmp + n, h->d_ipiv, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaEventRecord(h->ev_d, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    for (int i = 0; i < n; ++i) ipiv_out[i] = (int64_t)tmp[n + i] + 1;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_c, h->ev_d);
    h->timing[B200LU_T_D2H] = ms;
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ batched --
// pinned host staging for the int32 pivots / info of a batch (grown on demand, reused by warm calls)
static int ensure_batched_stage(b200lu_handle* h, int64_t count) {
    if (count <= h->hb_cap) return 0;
    if (h->hB_stage) cudaFreeHost(h->hB_stage);
    h->hB_stage = nullptr;
    h->hb_cap = 0;
    CU_TRY(h, cudaMallocHost((void**)&h->hB_stage, (size_t)count * sizeof(int)));
    h->hb_cap = count;
    return 0;
}

static int ensure_batched(b200lu_handle* h, int64_t batch, int64_t n) {
    const size_t es = elem_size(h) == 8 ? 8 : 4;
    const size_t fs = (h->dtype == B200LU_F64) ? 8 : 4;
    (void)es;
    const int64_t need = batch * n * n * (int64_t)fs;
    if (need > h->b_cap_bytes) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dB_LU);
        CU_TRY(h, cudaMalloc(&h->dB_LU, (size_t)need));
        h->b_cap_bytes = need;
    }
    if (batch * n > h->b_cap_batch_n) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dB_ipiv);
        free_dev(h->dB_info);
        free_dev(h->dB_perm);
        CU_TRY(h, cudaMalloc((void**)&h->dB_perm, (size_t)(batch * n) * sizeof(int)));
        CU_TRY(h, cudaMalloc((void**)&h->dB_ipiv, (size_t)(batch * n) * sizeof(int)));
        CU_TRY(h, cudaMalloc((void**)&h->dB_info, (size_t)(batch * n) * sizeof(int)));
        h->b_cap_batch_n = batch * n;
    }
    h->b_batch = batch;
    h->b_n = n;
    return 0;
}

// warp-per-system kernel (batched.cuh): factor, optionally with the first solve fused in
template <typename T, int NMAX, bool SOLVE>
static int launch_batched_warp(b200lu_handle* h, const BwArgs<T>& a, int64_t batch) {
    auto kern = getrf_batched_warp_kernel<T, NMAX, SOLVE>;
    constexpr size_t smem = bw_smem_bytes<T, NMAX>();
    static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
    if (!attr_set) {
        CU_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    kern<<<(unsigned)batch, 32, smem, h->s_main>>>(a);
    LAUNCH_CHECK(h);
    return 0;
}

// B (strideB) / X (strideX) non-null: the first right-hand side is solved in the same kernel (n <= 64 only)
template <typename T>
static int batched_factor_launch(b200lu_handle* h, const T* A, int64_t lda, int64_t strideA, const T* B = nullptr,
                                 int64_t strideB = 0, T* X = nullptr, int64_t strideX = 0) {
    const int n = (int)h->b_n;
    const int64_t batch = h->b_batch;
    T* LU = (T*)h->dB_LU;
    cudaStream_t st = h->s_main;
    if (n <= 64 && h->opt[B200LU_OPT_BATCHED_MODE] == 0) {
        BwArgs<T> a{A, lda, strideA, LU, n, (long long)n * n, h->dB_ipiv, h->dB_perm, h->dB_info, n, B, strideB, X, strideX};
        if (B) return n <= 32 ? launch_batched_warp<T, 32, true>(h, a, batch) : launch_batched_warp<T, 64, true>(h, a, batch);
        return n <= 32 ? launch_batched_warp<T, 32, false>(h, a, batch) : launch_batched_warp<T, 64, false>(h, a, batch);
    }
    if (n > 64) {
        // 65 ... BATCHED_SMEM_NMAX rows: the system lives in shared memory, one CTA each
        const size_t smem = (size_t)n * (n | 1) * sizeof(T);
        static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */
        if (!attr_set) {
            CU_TRY(h, cudaFuncSetAttribute(getrf_batched_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((size_t)BATCHED_SMEM_NMAX * (BATCHED_SMEM_NMAX | 1) * sizeof(T))));
            attr_set = true;
        }
        getrf_batched_smem_kernel<T><<<(unsigned)batch, BATCHED_SMEM_NT, smem, st>>>(A, lda, strideA, LU, n, (int64_t)n * n, h->dB_ipiv, h->dB_perm, h->dB_info, n);
    } else if (n <= 16)
        getrf_batched_kernel<T, 16><<<(unsigned)batch, 32, 0, st>>>(A, lda, strideA, LU, n, (int64_t)n * n, h->dB_ipiv, h->dB_perm, h->dB_info, n);
    else if (n <= 32)
        getrf_batched_kernel<T, 32><<<(unsigned)batch, 32, 0, st>>>(A, lda, strideA, LU, n, (int64_t)n * n, h->dB_ipiv, h->dB_perm, h->dB_info, n);
    else
        getrf_batched_kernel<T, 64><<<(unsigned)batch, 64, 0, st>>>(A, lda, strideA, LU, n, (int64_t)n * n, h->dB_ipiv, h->dB_perm, h->dB_info, n);
    LAUNCH_CHECK(h);
    return 0;
}

template <typename T>
static int batched_solve_launch(b200lu_handle* h, int nrhs, const T* B, int64_t ldb, int64_t strideB,
                                T* X, int64_t ldx, int64_t strideX, bool trans = false) {
    const int n = (int)h->b_n;
    const int64_t batch = h->b_batch;
    const T* LU = (const T*)h->dB_LU;
    cudaStream_t st = h->s_main;
    // WPC systems (warps) per CTA: about 32 KB of staged factors per CTA
#define GETRS_B(NMAXV, WPCV)                                                                      \
    {                                                                                             \
        const unsigned grid = (unsigned)((batch + (WPCV) - 1) / (WPCV));                           \
        if ((NMAXV) > 64) {   /* more than 48 KB of dynamic shared memory: opt in once */          \
            static std::atomic<bool> attr_dev[64]; std::atomic<bool>& attr_set = attr_dev[h->dev & 63];   /* function attributes are per device */                                                         \
            if (!attr_set) {                                                                      \
                CU_TRY(h, cudaFuncSetAttribute(getrs_batched_kernel<T, NMAXV, WPCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)((size_t)(WPCV) * (NMAXV) * (NMAXV) * sizeof(T)))); \
                CU_TRY(h, cudaFuncSetAttribute(getrs_batched_trans_kernel<T, NMAXV, WPCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)((size_t)(WPCV) * (NMAXV) * ((NMAXV) + 1) * sizeof(T)))); \
                attr_set = true;                                                                  \
            }                                                                                     \
        }                                                                                         \
        if (trans) {                                                                              \
            const size_t smem = (size_t)(WPCV) * (NMAXV) * ((NMAXV) + 1) * sizeof(T);              \
            getrs_batched_trans_kernel<T, NMAXV, WPCV><<<grid, 32 * (WPCV), smem, st>>>(           \
                LU, n, (int64_t)n * n, h->dB_perm, B, ldb, strideB, X, ldx, strideX, n, nrhs, batch); \
        } else {                                                                                  \
            const size_t smem = (size_t)(WPCV) * (NMAXV) * (NMAXV) * sizeof(T);                    \
            getrs_batched_kernel<T, NMAXV, WPCV><<<grid, 32 * (WPCV), smem, st>>>(                 \
                LU, n, (int64_t)n * n, h->dB_perm, B, ldb, strideB, X, ldx, strideX, n, nrhs, batch); \
        }                                                                                         \
    }
    if (n > 64) GETRS_B(BATCHED_SMEM_NMAX, 1)
    else if (n <= 16) GETRS_B(16, 8)
    else if (n <= 32) GETRS_B(32, 4)
    else if (sizeof(T) == 4) GETRS_B(64, 2)
    else GETRS_B(64, 1)
#undef GETRS_B
    LAUNCH_CHECK(h);
    return 0;
}

extern "C" {

static int check_batched_args(b200lu_handle* h, int64_t batch, int64_t n, const void* A, int64_t lda,
                              int64_t strideA) {
    if (!h) return -1;
    if (h->dtype == B200LU_MIXED) return set_err(h, -1, "batched mode supports F64 and F32 handles");
    if (batch < 0) return set_err(h, -2, "batch < 0");
    if (n < 0 || n > BATCHED_SMEM_NMAX) return set_err(h, -3, "batched n must be in [0, %d]", BATCHED_SMEM_NMAX);
    if (batch > 0 && n > 0) {
        if (!A) return set_err(h, -4, "A is NULL");
        if (lda < n) return set_err(h, -5, "lda < n");
        if (batch > 1 && strideA < lda * (n - 1) + n) return set_err(h, -6, "strideA overlaps");
    }
    return 0;
}

static int factor_batched_device_impl(b200lu_handle* h, int64_t batch, int64_t n, const void* A_dev, int64_t lda,
                                      int64_t strideA, const void* B_dev, int64_t strideB, void* X_dev, int64_t strideX,
                                      int64_t* any_info) {
    int rc = check_batched_args(h, batch, n, A_dev, lda, strideA);
    if (rc) return rc;
    if (any_info) *any_info = 0;
    h->b_factored = false;
    if (batch == 0 || n == 0) { h->b_batch = batch; h->b_n = n; h->b_factored = true; return 0; }
    CU_TRY(h, cudaSetDevice(h->dev));
    rc = ensure_batched(h, batch, n);
    if (rc) return rc;
    CU_TRY(h, cudaEventRecord(h->ev_b, h->s_main));
    // the first right-hand side rides in the factorization kernel when the warp kernel runs (n <= 64);
    // otherwise it is an ordinary getrs launch behind the factorization
    const bool fuse = B_dev && n <= 64 && h->opt[B200LU_OPT_BATCHED_MODE] == 0;
    if (h->dtype == B200LU_F64)
        rc = batched_factor_launch<double>(h, (const double*)A_dev, lda, strideA, fuse ? (const double*)B_dev : nullptr, strideB,
                                           fuse ? (double*)X_dev : nullptr, strideX);
    else
        rc = batched_factor_launch<float>(h, (const float*)A_dev, lda, strideA, fuse ? (const float*)B_dev : nullptr, strideB,
                                          fuse ? (float*)X_dev : nullptr, strideX);
    if (rc) return rc;
    if (B_dev && !fuse) {
        if (h->dtype == B200LU_F64)
            rc = batched_solve_launch<double>(h, 1, (const double*)B_dev, n, strideB, (double*)X_dev, n, strideX, false);
        else
            rc = batched_solve_launch<float>(h, 1, (const float*)B_dev, n, strideB, (float*)X_dev, n, strideX, false);
        if (rc) return rc;
    }
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    h->b_factored = true;
    if (any_info) {
        // count failures: max over info on the host side is done by the host path; here a cheap flag
        rc = ensure_batched_stage(h, batch);
        if (rc) return rc;
        CU_TRY(h, cudaMemcpyAsync(h->hB_stage, h->dB_info, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        int64_t bad = 0;
        for (int64_t i = 0; i < batch; ++i) bad += h->hB_stage[i] != 0;
        *any_info = bad;
    } else {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
    h->timing[B200LU_T_FACTOR] = ms;
    return 0;
}

int b200lu_factor_batched_device(b200lu_handle* h, int64_t batch, int64_t n, const void* A_dev,
                                 int64_t lda, int64_t strideA, int64_t* any_info) {
    TEAM_ONLY_HOST(h);
    return factor_batched_device_impl(h, batch, n, A_dev, lda, strideA, nullptr, 0, nullptr, 0, any_info);
}

int b200lu_factor_solve_batched_device(b200lu_handle* h, int64_t batch, int64_t n, const void* A_dev, int64_t lda,
                                       int64_t strideA, const void* B_dev, int64_t strideB, void* X_dev, int64_t strideX,
                                       int64_t* any_info) {
    TEAM_ONLY_HOST(h);
    if (h && batch > 0 && n > 0) {
        if (!B_dev || !X_dev) return set_err(h, -7, "B or X is NULL");
        if (batch > 1 && (strideB < n || strideX < n)) return set_err(h, -8, "strideB / strideX < n");
    }
    return factor_batched_device_impl(h, batch, n, A_dev, lda, strideA, B_dev, strideB, X_dev, strideX, any_info);
}

static int parse_trans(b200lu_handle* h, char trans, bool* tr) {
    *tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');   // real element types: 'C' == 'T'
    if (!*tr && trans != 'N' && trans != 'n') return set_err(h, -2, "trans='%c' is not one of N, T, C", trans);
    return 0;
}

static int solve_batched_device_impl(b200lu_handle* h, bool trans, int64_t nrhs, const void* B_dev, int64_t ldb,
                                     int64_t strideB, void* X_dev, int64_t ldx, int64_t strideX) {
    if (!h->b_factored) return set_err(h, 3, "no batched factorization cached");
    if (nrhs < 0) return set_err(h, -2, "nrhs < 0");
    if (h->b_batch == 0 || h->b_n == 0 || nrhs == 0) return 0;
    if (!B_dev || !X_dev) return set_err(h, -3, "B or X is NULL");
    if (ldb < h->b_n || ldx < h->b_n) return set_err(h, -4, "ldb/ldx < n");
    CU_TRY(h, cudaSetDevice(h->dev));
    int rc;
    CU_TRY(h, cudaEventRecord(h->ev_b, h->s_main));
    if (h->dtype == B200LU_F64)
        rc = batched_solve_launch<double>(h, (int)nrhs, (const double*)B_dev, ldb, strideB, (double*)X_dev, ldx, strideX, trans);
    else
        rc = batched_solve_launch<float>(h, (int)nrhs, (const float*)B_dev, ldb, strideB, (float*)X_dev, ldx, strideX, trans);
    if (rc) return rc;
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
    h->timing[B200LU_T_SOLVE] = ms;
    return 0;
}

int b200lu_solve_batched_device(b200lu_handle* h, int64_t nrhs, const void* B_dev, int64_t ldb,
                                int64_t strideB, void* X_dev, int64_t ldx, int64_t strideX) {
    if (!h) return -1;
    TEAM_ONLY_HOST(h);
    return solve_batched_device_impl(h, false, nrhs, B_dev, ldb, strideB, X_dev, ldx, strideX);
}

int b200lu_solve_batched_trans_device(b200lu_handle* h, char trans, int64_t nrhs, const void* B_dev, int64_t ldb,
                                      int64_t strideB, void* X_dev, int64_t ldx, int64_t strideX) {
    if (!h) return -1;
    TEAM_ONLY_HOST(h);
    bool tr = false;
    int rc = parse_trans(h, trans, &tr);
    if (rc) return rc;
    return solve_batched_device_impl(h, tr, nrhs, B_dev, ldb, strideB, X_dev, ldx, strideX);
}

int b200lu_factor_batched(b200lu_handle* h, int64_t batch, int64_t n, const void* A_host, int64_t lda,
                          int64_t strideA, int64_t* ipiv_out, int64_t* info_out) {
    if (h && h->team) return team_factor_batched(h, batch, n, A_host, lda, strideA, ipiv_out, info_out);
    int rc = check_batched_args(h, batch, n, A_host, lda, strideA);
    if (rc) return rc;
    h->b_factored = false;
    if (batch == 0 || n == 0) { h->b_batch = batch; h->b_n = n; h->b_factored = true; return 0; }
    CU_TRY(h, cudaSetDevice(h->dev));
    const size_t is = iface_size(h);
    // stage A compactly (lda -> n) so one 2D copy moves the whole batch when lda == n, strideA == n*n
    const int64_t in_bytes = batch * n * n * (int64_t)is;
    if (in_bytes > h->b_cap_in) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dB_in);
        CU_TRY(h, cudaMalloc(&h->dB_in, (size_t)in_bytes));
        h->b_cap_in = in_bytes;
    }
    CU_TRY(h, cudaEventRecord(h->ev_a, h->s_main));
    if (lda == n && (strideA == n * n || batch == 1)) {
        CU_TRY(h, cudaMemcpyAsync(h->dB_in, A_host, (size_t)in_bytes, cudaMemcpyHostToDevice, h->s_main));
    } else {
        for (int64_t i = 0; i < batch; ++i)
            CU_TRY(h, cudaMemcpy2DAsync((char*)h->dB_in + (size_t)(i * n * n) * is, (size_t)n * is,
                                        (const char*)A_host + (size_t)(i * strideA) * is, (size_t)lda * is,
                                        (size_t)n * is, (size_t)n, cudaMemcpyHostToDevice, h->s_main));
    }
    rc = b200lu_factor_batched_device(h, batch, n, h->dB_in, n, n * n, nullptr);
    if (rc) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
    h->timing[B200LU_T_H2D] = ms;
    if (ipiv_out || info_out) return b200lu_get_factors_batched(h, nullptr, 0, 0, ipiv_out, info_out);
    return 0;
}

// getrf of the batch + getrs of ONE right-hand side per system in the same kernel; the factors stay
// cached like after b200lu_factor_batched (the `solve!` of a fresh BlockDiagonal cache)
int b200lu_factor_solve_batched(b200lu_handle* h, int64_t batch, int64_t n, const void* A_host, int64_t lda,
                                int64_t strideA, const void* B_host, int64_t strideB, void* X_host, int64_t strideX,
                                int64_t* ipiv_out, int64_t* info_out) {
    if (h && h->team) return team_factor_solve_batched(h, batch, n, A_host, lda, strideA, B_host, strideB, X_host, strideX, ipiv_out, info_out);
    int rc = check_batched_args(h, batch, n, A_host, lda, strideA);
    if (rc) return rc;
    h->b_factored = false;
    if (batch == 0 || n == 0) { h->b_batch = batch; h->b_n = n; h->b_factored = true; return 0; }
    if (!B_host || !X_host) return set_err(h, -7, "B or X is NULL");
    if (batch > 1 && (strideB < n || strideX < n)) return set_err(h, -8, "strideB / strideX < n");
    CU_TRY(h, cudaSetDevice(h->dev));
    const size_t is = iface_size(h);
    const int64_t in_bytes = batch * n * n * (int64_t)is;
    if (in_bytes > h->b_cap_in) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dB_in);
        h->b_cap_in = 0;
        CU_TRY(h, cudaMalloc(&h->dB_in, (size_t)in_bytes));
        h->b_cap_in = in_bytes;
    }
    const int64_t rbytes = batch * n * (int64_t)is;
    if (rbytes > h->b_cap_rhs) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dB_rhs);
        free_dev(h->dB_x);
        h->b_cap_rhs = 0;
        CU_TRY(h, cudaMalloc(&h->dB_rhs, (size_t)rbytes));
        CU_TRY(h, cudaMalloc(&h->dB_x, (size_t)rbytes));
        h->b_cap_rhs = rbytes;
    }
    CU_TRY(h, cudaEventRecord(h->ev_a, h->s_main));
    if (lda == n && (strideA == n * n || batch == 1)) {
        CU_TRY(h, cudaMemcpyAsync(h->dB_in, A_host, (size_t)in_bytes, cudaMemcpyHostToDevice, h->s_main));
    } else {
        for (int64_t i = 0; i < batch; ++i)
            CU_TRY(h, cudaMemcpy2DAsync((char*)h->dB_in + (size_t)(i * n * n) * is, (size_t)n * is,
                                        (const char*)A_host + (size_t)(i * strideA) * is, (size_t)lda * is,
                                        (size_t)n * is, (size_t)n, cudaMemcpyHostToDevice, h->s_main));
    }
    CU_TRY(h, cudaMemcpy2DAsync(h->dB_rhs, (size_t)n * is, B_host, (size_t)strideB * is, (size_t)n * is, (size_t)batch,
                                cudaMemcpyHostToDevice, h->s_main));
    rc = factor_batched_device_impl(h, batch, n, h->dB_in, n, n * n, h->dB_rhs, n, h->dB_x, n, nullptr);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpy2DAsync(X_host, (size_t)strideX * is, h->dB_x, (size_t)n * is, (size_t)n * is, (size_t)batch,
                                cudaMemcpyDeviceToHost, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
    h->timing[B200LU_T_H2D] = ms;
    if (ipiv_out || info_out) return b200lu_get_factors_batched(h, nullptr, 0, 0, ipiv_out, info_out);
    return 0;
}

int b200lu_get_factors_batched(b200lu_handle* h, void* LU_host, int64_t lda, int64_t strideA,
                               int64_t* ipiv_out, int64_t* info_out) {
    if (!h) return -1;
    if (h->team) return team_get_factors_batched(h, LU_host, lda, strideA, ipiv_out, info_out);
    if (!h->b_factored) return set_err(h, 3, "no batched factorization cached");
    const int64_t batch = h->b_batch, n = h->b_n;
    if (batch == 0 || n == 0) return 0;
    CU_TRY(h, cudaSetDevice(h->dev));
    const size_t fs = (h->dtype == B200LU_F64) ? 8 : 4;
    if (LU_host) {
        if (lda < n) return set_err(h, -3, "lda < n");
        if (lda == n && strideA == n * n) {
            CU_TRY(h, cudaMemcpy(LU_host, h->dB_LU, (size_t)(batch * n * n) * fs, cudaMemcpyDeviceToHost));
        } else {
            for (int64_t i = 0; i < batch; ++i)
                CU_TRY(h, cudaMemcpy2D((char*)LU_host + (size_t)(i * strideA) * fs, (size_t)lda * fs,
                                       (const char*)h->dB_LU + (size_t)(i * n * n) * fs, (size_t)n * fs,
                                       (size_t)n * fs, (size_t)n, cudaMemcpyDeviceToHost));
        }
    }
    if (ipiv_out || info_out) {
        int rc = ensure_batched_stage(h, batch * n);
        if (rc) return rc;
    }
    if (ipiv_out) {
        const int64_t cnt = batch * n;
        CU_TRY(h, cudaMemcpyAsync(h->hB_stage, h->dB_ipiv, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        for (int64_t i = 0; i < cnt; ++i) ipiv_out[i] = (int64_t)h->hB_stage[i] + 1;
    }
    if (info_out) {
        CU_TRY(h, cudaMemcpyAsync(h->hB_stage, h->dB_info, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        for (int64_t i = 0; i < batch; ++i) info_out[i] = h->hB_stage[i];
    }
    return 0;
}

static int solve_batched_impl(b200lu_handle* h, bool trans, int64_t nrhs, const void* B_host, int64_t ldb,
                              int64_t strideB, void* X_host, int64_t ldx, int64_t strideX) {
    if (!h->b_factored) return set_err(h, 3, "no batched factorization cached");
    if (nrhs < 0) return set_err(h, -2, "nrhs < 0");
    const int64_t batch = h->b_batch, n = h->b_n;
    if (batch == 0 || n == 0 || nrhs == 0) return 0;
    if (!B_host || !X_host) return set_err(h, -3, "B or X is NULL");
    if (ldb < n || ldx < n) return set_err(h, -4, "ldb/ldx < n");
    CU_TRY(h, cudaSetDevice(h->dev));
    const size_t is = iface_size(h);
    const int64_t bytes = batch * n * nrhs * (int64_t)is;
    if (bytes > h->b_cap_rhs) {
        CU_TRY(h, cudaStreamSynchronize(h->s_main));
        free_dev(h->dB_rhs);
        free_dev(h->dB_x);
        CU_TRY(h, cudaMalloc(&h->dB_rhs, (size_t)bytes));
        CU_TRY(h, cudaMalloc(&h->dB_x, (size_t)bytes));
        h->b_cap_rhs = bytes;
    }
    CU_TRY(h, cudaEventRecord(h->ev_a, h->s_main));
    const bool compactB = (ldb == n) && (strideB == n * nrhs || batch == 1);
    if (compactB) {
        CU_TRY(h, cudaMemcpyAsync(h->dB_rhs, B_host, (size_t)bytes, cudaMemcpyHostToDevice, h->s_main));
    } else {
        for (int64_t i = 0; i < batch; ++i)
            CU_TRY(h, cudaMemcpy2DAsync((char*)h->dB_rhs + (size_t)(i * n * nrhs) * is, (size_t)n * is,
                                        (const char*)B_host + (size_t)(i * strideB) * is, (size_t)ldb * is,
                                        (size_t)n * is, (size_t)nrhs, cudaMemcpyHostToDevice, h->s_main));
    }
    int rc = solve_batched_device_impl(h, trans, nrhs, h->dB_rhs, n, n * nrhs, h->dB_x, n, n * nrhs);
    if (rc) return rc;
    const bool compactX = (ldx == n) && (strideX == n * nrhs || batch == 1);
    CU_TRY(h, cudaEventRecord(h->ev_c, h->s_main));
    if (compactX) {
        CU_TRY(h, cudaMemcpyAsync(X_host, h->dB_x, (size_t)bytes, cudaMemcpyDeviceToHost, h->s_main));
    } else {
        for (int64_t i = 0; i < batch; ++i)
            CU_TRY(h, cudaMemcpy2DAsync((char*)X_host + (size_t)(i * strideX) * is, (size_t)ldx * is,
                                        (const char*)h->dB_x + (size_t)(i * n * nrhs) * is, (size_t)n * is,
                                        (size_t)n * is, (size_t)nrhs, cudaMemcpyDeviceToHost, h->s_main));
    }
    CU_TRY(h, cudaEventRecord(h->ev_d, h->s_main));
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_c, h->ev_d);
    h->timing[B200LU_T_D2H] = ms;
    return 0;
}

int b200lu_solve_batched(b200lu_handle* h, int64_t nrhs, const void* B_host, int64_t ldb,
                         int64_t strideB, void* X_host, int64_t ldx, int64_t strideX) {
    if (!h) return -1;
    if (h->team) return team_solve_batched(h, false, 'N', nrhs, B_host, ldb, strideB, X_host, ldx, strideX);
    return solve_batched_impl(h, false, nrhs, B_host, ldb, strideB, X_host, ldx, strideX);
}

int b200lu_solve_batched_trans(b200lu_handle* h, char trans, int64_t nrhs, const void* B_host, int64_t ldb,
                               int64_t strideB, void* X_host, int64_t ldx, int64_t strideX) {
    if (!h) return -1;
    if (h->team) return team_solve_batched(h, true, trans, nrhs, B_host, ldb, strideB, X_host, ldx, strideX);
    bool tr = false;
    int rc = parse_trans(h, trans, &tr);
    if (rc) return rc;
    return solve_batched_impl(h, tr, nrhs, B_host, ldb, strideB, X_host, ldx, strideX);
}

// --------------------------------------------------------------- synthetic --
int b200lu_fill_uniform_device(b200lu_handle* h, void* A_dev, int64_t lda, int64_t n, int64_t ncols,
                               int64_t first_global_col, int64_t col_block, int64_t col_block_stride,
                               uint64_t seed, double diag_shift) {
    if (!h) return -1;
    TEAM_ONLY_HOST(h);
    if (!A_dev || lda < n || n <= 0 || ncols <= 0 || col_block <= 0) return set_err(h, -2, "bad fill arguments");
    CU_TRY(h, cudaSetDevice(h->dev));
    dim3 grid(cdiv(n, 256), (unsigned)std::min<int64_t>(ncols, 4096));
    if (h->dtype == B200LU_F32)
        fill_uniform_kernel<float><<<grid, 256, 0, h->s_main>>>((float*)A_dev, lda, n, ncols, first_global_col,
                                                                 col_block, col_block_stride, seed, diag_shift);
    else
        fill_uniform_kernel<double><<<grid, 256, 0, h->s_main>>>((double*)A_dev, lda, n, ncols, first_global_col,
                                                                  col_block, col_block_stride, seed, diag_shift);
    LAUNCH_CHECK(h);
    CU_TRY(h, cudaStreamSynchronize(h->s_main));
    return 0;
}

}  // extern "C"

#include "dist.inc"

