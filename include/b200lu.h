/*
 * b200lu.h — C ABI of libb200lu.so: a B200-native (sm_100a) dense partially
 * pivoted LU factor-and-solve path for the LinearSolve.jl interface.
 *
 * This is the drop-in boundary: the entry points below are what the
 * reference's Julia side would bind with `ccall`, in the same shape in which
 * it binds LAPACK today.  Citations are relative to the reference tree
 * (SciML/LinearSolve.jl v5.12.0):
 *
 *   b200lu_factor        replaces  dgetrf_/sgetrf_ as called at
 *                                  src/openblas.jl:131-154 (openblas_getrf!)
 *                                  and LAPACK.getrf! at src/factorization.jl:632-637
 *   b200lu_solve         replaces  dgetrs_/sgetrs_ as called at
 *                                  src/openblas.jl:247-278 (openblas_getrs!)
 *                                  and `_smart_lu_ldiv!` src/factorization.jl:601-611
 *   b200lu_factor_batched / b200lu_solve_batched
 *                        replace   the per-block `lu!` / `ldiv!` loop of
 *                                  ext/LinearSolveBlockDiagonalsExt.jl:119-125,183-205
 *   dtype B200LU_F32 / B200LU_MIXED
 *                        replace   the sgetrf/sgetrs pair of
 *                                  src/openblas.jl:487-543 (OpenBLAS32MixedLUFactorization);
 *                                  B200LU_MIXED adds the FP64 refinement loop the
 *                                  north star asks for (a superset of the reference).
 *   *_device variants    are the "GPUArray interface" surface
 *                                  (docs/src/tutorials/gpu.md:65-92): same calls on
 *                                  raw device pointers, no PCIe in the timed path.
 *
 * Conventions (identical to LAPACK as the reference calls it):
 *   - matrices are column-major with a leading dimension in ELEMENTS;
 *   - ipiv is the 1-based LAPACK interchange sequence, int64 (BlasInt == Int64,
 *     src/LinearSolve.jl:52-63); row k was swapped with row ipiv[k];
 *   - info > 0  : index (1-based) of the FIRST exactly-zero pivot; the
 *                 factorization still runs to completion
 *                 (src/generic_lufact.jl:121-123, src/blocked_lufact.jl:93-122);
 *     info == 0 : success;
 *   - every function returns a status: 0 = ok, < 0 = bad argument (LAPACK
 *     style: -i = i-th argument), > 0 = CUDA/NCCL/runtime failure
 *     (text via b200lu_last_error).  Nothing throws or longjmps.
 *   - calls block until the outputs are written (ccall semantics).  A handle is
 *     single-threaded; distinct handles may be driven from distinct host threads
 *     (test/Core/basictests.jl:1331-1358).
 *   - warm calls (same n / same batch geometry) allocate nothing
 *     (test/Core/direct_blas_refactorization.jl:32-37, test/qa/allocations.jl:7-18).
 *
 * There is no CPU fallback behind this ABI: without a CUDA device
 * b200lu_create fails with a non-zero status.
 */
#ifndef B200LU_H
#define B200LU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200lu_handle b200lu_handle;

/* element type of the factorization */
enum {
    B200LU_F64 = 0,   /* FP64 getrf + getrs                                  */
    B200LU_F32 = 1,   /* FP32 getrf + getrs (interface data is float)        */
    B200LU_MIXED = 2  /* FP64 interface; FP32 factor + FP64 iterative refine */
};

/* phases for b200lu_last_timing (milliseconds, CUDA-event measured) */
enum {
    B200LU_T_H2D = 0,      /* host -> device copy of A (factor) or B (solve)   */
    B200LU_T_FACTOR = 1,   /* getrf on the device (with B200LU_OPT_STREAM_H2D the
                              upload runs underneath it: H2D and FACTOR overlap) */
    B200LU_T_SOLVE = 2,    /* getrs (+ refinement) on the device               */
    B200LU_T_D2H = 3,      /* device -> host copy of ipiv/info or X            */
    B200LU_T_GEMM = 4,     /* of FACTOR: sum of trailing-update GEMM launches;
                              only with B200LU_OPT_PROFILE = 1                 */
    /* distributed getrf with B200LU_OPT_PROFILE = 1: the three phases of the critical chain, summed over the
       panels THIS rank factored / pushed (CUDA events on its panel stream) */
    B200LU_T_PANEL = 5,    /* panel factorizations (recursive panel of the blocks this rank owns)  */
    B200LU_T_LOOKAHEAD = 6,/* update k of the block this rank factors next (laswp + TRSM + GEMM)   */
    B200LU_T_PUSH = 7,     /* hand-off of the factored panels (peer stores / ncclBroadcast)        */
    B200LU_T_COUNT = 8
};

/* counters for b200lu_last_counter (filled by the last factor call) */
enum {
    B200LU_C_GEMM_FLOPS = 0,    /* 2*M*N*K summed over the profiled GEMM launches */
    B200LU_C_GEMM_LAUNCHES = 1, /* number of profiled GEMM launches               */
    B200LU_C_REFINE_ITERS = 2,  /* MIXED: refinement sweeps of the last solve     */
    B200LU_C_COUNT = 3
};

/* kinds for b200lu_probe_peak: on-device micro-benchmarks used as roofline
 * denominators next to MEASURED_PEAKS.json */
enum {
    B200LU_PEAK_FP64_DMMA = 0,  /* TFLOP/s of a register-resident DMMA.8x8x4 loop */
    B200LU_PEAK_FP64_DFMA = 1,  /* TFLOP/s of a register-resident DFMA loop       */
    B200LU_PEAK_HBM_COPY = 2    /* GB/s (read+write) of a 1 GiB device copy       */
};

/* tuning / test knobs for b200lu_set_option */
enum {
    B200LU_OPT_NB = 0,          /* outer panel width (multiple of 16, <= 256); default 256, and 128 for one
                                   system over 8 or more GPUs unless the caller sets it              */
    B200LU_OPT_LOOKAHEAD = 1,   /* 0/1: factor panel k+1 under trailing update k */
    B200LU_OPT_REFINE_MAXIT = 2,/* B200LU_MIXED: max refinement sweeps (def. 10) */
    B200LU_OPT_PANEL_CTAS = 3,  /* max CTAs of the cooperative base-panel kernel */
    B200LU_OPT_SOLVE_NRHS_TILE = 4, /* right-hand sides per triangular sweep     */
    B200LU_OPT_PROFILE = 5,     /* 1: bracket every trailing GEMM with CUDA events */
    B200LU_OPT_PANEL_RPT = 6,   /* rows per thread in the base panel: 0 auto, 1, 2 */
    B200LU_OPT_GEMM_CFG = 7,    /* FP64 trailing-update tile configuration 0..2, 3 = auto (default) */
    B200LU_OPT_PANEL_MODE = 8,  /* base panel: 0 auto = the left-looking fused cluster/DSMEM kernel (2-4 sub-blocks of
                                   8 / 16 / 32 columns per launch; FP64 panels of up to 32768 rows), L2 mailbox above;
                                   1 always L2 mailbox; 2 = the round-1 cluster kernels, one block per launch     */
    B200LU_OPT_SGEMM_MODE = 9,  /* FP32 trailing update: 0 auto (tcgen05 3xTF32 kernel for large
                                   updates, FFMA otherwise), 1 always FFMA          */
    B200LU_OPT_TRSV_MODE = 10,  /* single-RHS getrs: 0 auto (n >= 6144: mode 3, else mode 2;
                                   transposed solves: mode 2), 1 = one CTA per block row,
                                   2 = 2-D work items on a persistent grid, 3 = the dependency
                                   chain in one thread-block cluster over DSMEM + worker CTAs
                                   on the far blocks (n >= 1024)                          */
    B200LU_OPT_STREAM_H2D = 11, /* b200lu_factor from a HOST matrix (F64/F32, n >= 2048): 1 (default)
                                   = upload A in column chunks on a copy stream and start
                                   factoring as soon as the first chunk has landed (late chunks
                                   are caught up left-looking when they arrive); 0 = copy all of
                                   A, then factor.  The factors are the same either way.  */
    B200LU_OPT_MAPPED_RHS = 12, /* b200lu_solve with ONE right-hand side (F64/F32, trans = 'N', n >= 256):
                                   1 (default) = b and x travel through a page-locked, device-mapped
                                   staging buffer that the getrs kernels read and write directly over
                                   PCIe (no copy-engine launches); 0 = H2D / D2H copies             */
    B200LU_OPT_KEEP_A = 13,     /* 1 = b200lu_factor / b200lu_factor_device keep a device copy of A (F64/F32 handles;
                                   +n^2 elements of HBM, one device-to-device copy, and the host upload
                                   is no longer streamed under the factorization) so that
                                   b200lu_residual_norms can check a solution on the device; default 0.
                                   MIXED handles always hold A in FP64.                              */
    B200LU_OPT_BATCHED_MODE = 14, /* batched getrf of systems of <= 64 rows: 0 (default) = one WARP per system, the
                                   system in shared memory, 8-column register panels (and b200lu_factor_solve_batched
                                   fuses the first getrs into it); 1 = the round-1 kernel, one row per thread    */
    B200LU_OPT_HOST_REGISTER = 15, /* b200lu_factor from a HOST matrix in pageable memory: 1 = page-lock the caller's
                                   buffer once (cudaHostRegister, portable across the handle's GPUs) and keep it
                                   registered until another buffer is factored or the handle is destroyed, so that
                                   every factorization from that buffer gets the streamed, DMA-direct upload a pinned
                                   buffer gets.  For callers that own the buffer for the lifetime of the cache
                                   (reference: a LinearCache with alias_A = false, src/common.jl:818-842) and cannot
                                   allocate pinned memory themselves.  Default 0: pageable buffers are staged by the
                                   driver.  The registration costs about as much as one pageable upload.  The
                                   buffer must stay allocated while it is registered: destroy the handle (or factor
                                   from another buffer) before freeing it.                                        */
    B200LU_OPT_COUNT = 16
};

/* library/ABI version: major*10000 + minor*100 + patch */
int b200lu_version(void);

/* Number of kernels launched by this library (all handles) since load; the
 * bench reads it before/after the timed region to report gpu_launches. */
int64_t b200lu_launch_count(void);

/*
 * Create a handle bound to `ngpus` devices; `devices` may be NULL (device 0 .. ngpus-1).
 *   ngpus == 1 : the single-GPU path.  For one-PROCESS-per-GPU multi-GPU runs create such a
 *                handle on the local device and attach a communicator with b200lu_comm_init.
 *   ngpus in 2..16 : ONE process drives all the GPUs (what a Julia `solve!` can reach; SURVEY §5
 *                "single process, 8 devices").  The handle owns one sub-handle and one host
 *                thread per GPU; every pair of devices must be peers (NVLink / NVSwitch).
 *                b200lu_factor distributes the host matrix itself (1-D block-cyclic columns of
 *                width B200LU_OPT_NB: each GPU pulls its own column blocks over its own PCIe
 *                link), the owner of a panel stores it into its peers' memory (no NCCL),
 *                b200lu_solve runs the distributed getrs (trans = 'N'), b200lu_factor_batched /
 *                b200lu_solve_batched shard the batch index over the GPUs (no communication).
 *                F64 and F32 handles; the *_device entry points, b200lu_residual_norms and the
 *                comm/dist entry points are single-GPU-handle calls (status -1 / 3 here).
 * Returns 5 when the devices are not all peers of each other.
 */
int b200lu_create(b200lu_handle** h, int dtype, int ngpus, const int* devices);
void b200lu_destroy(b200lu_handle* h);
const char* b200lu_last_error(const b200lu_handle* h);
double b200lu_last_timing(const b200lu_handle* h, int phase);
double b200lu_last_counter(const b200lu_handle* h, int which);
/* test hook: C -= A * B on DEVICE pointers with the handle's trailing-update kernel (column-major;
   element type = the handle's factor type).  Mirrors the reference's `_blocked_lu_schur!`
   (src/blocked_lufact.jl:186-620) so the kernel can be checked against a plain FP32/FP64 matmul. */
int b200lu_debug_gemm_sub(b200lu_handle* h, int64_t M, int64_t N, int64_t K, const void* dA, int64_t lda,
                          const void* dB, int64_t ldb, void* dC, int64_t ldc);
int b200lu_probe_peak(b200lu_handle* h, int kind, double* out);
int b200lu_set_option(b200lu_handle* h, int option, int64_t value);
int64_t b200lu_get_option(const b200lu_handle* h, int option);

/*
 * getrf.  A_host: n x n column-major, leading dimension lda, element type per
 * the handle's dtype (double for F64 and MIXED, float for F32).  A_host is NOT
 * overwritten (like CudaOffloadLUFactorization, ext/LinearSolveCUDAExt.jl:82);
 * the factors stay on the device inside the handle (use b200lu_get_factors to
 * read them back).  ipiv_out (length n) may be NULL.  *info as above.
 */
int b200lu_factor(b200lu_handle* h, int64_t n, const void* A_host, int64_t lda,
                  int64_t* ipiv_out, int64_t* info);

/*
 * getrs with the cached factors: op(A) X = B, trans in {'N','T','C'}.
 * B_host: n x nrhs, X_host: n x nrhs (may alias B_host).  Element type: double
 * for F64/MIXED, float for F32.  Returns 1000+info-style failure (status 3) if
 * the cached factorization is singular (info > 0) or absent.  'T'/'C' (the
 * reference's `solve!(cache; adjoint = true)`, src/common.jl:1012-1027; LAPACK
 * getrs trans argument, src/openblas.jl:247-278) reuse the same factors:
 * U^T y = b, L^T z = y, x = P^T z; every handle type (B200LU_MIXED refines the
 * transposed system: FP32 transposed sweeps + FP64 residual b - A^T x, one
 * right-hand side at a time).
 */
int b200lu_solve(b200lu_handle* h, char trans, int64_t nrhs,
                 const void* B_host, int64_t ldb, void* X_host, int64_t ldx);

/* Same two calls on device-resident data (pointers valid on the handle's
 * device).  factor_device copies A_dev into the handle's own factor buffer
 * unless A_dev IS that buffer (see b200lu_device_matrix). */
int b200lu_factor_device(b200lu_handle* h, int64_t n, const void* A_dev, int64_t lda,
                         int64_t* info);
int b200lu_solve_device(b200lu_handle* h, char trans, int64_t nrhs,
                        const void* B_dev, int64_t ldb, void* X_dev, int64_t ldx);

/*
 * The a-posteriori residual check of the reference (`_check_residual_safety`,
 * src/factorization.jl:127-156: ||A u - b|| <= abstol + reltol ||b||) on the device:
 * resid_out[c] = ||B[:, c] - A X[:, c]||_2 and bnorm_out[c] = ||B[:, c]||_2, accumulated
 * in FP64, for host B and X (n x nrhs, interface element type).  Needs the matrix itself:
 * B200LU_OPT_KEEP_A switched on at factor time for F64/F32 handles (status 3 otherwise).
 */
int b200lu_residual_norms(b200lu_handle* h, int64_t nrhs, const void* B_host, int64_t ldb,
                          const void* X_host, int64_t ldx, double* resid_out, double* bnorm_out);

/* Copy factors (L\U packed like LAPACK) / pivots of the cached factorization
 * to the host: parity tests, adjoint reuse (src/adjoint_factorization.jl). */
int b200lu_get_factors(b200lu_handle* h, void* LU_host, int64_t ldlu);
int b200lu_get_ipiv(b200lu_handle* h, int64_t* ipiv_out);

/*
 * Many independent small systems (BlockDiagonal surface): `batch` matrices of
 * n x n (n <= 160: register kernels up to 64 rows, a shared-memory kernel above),
 * matrix i at A + i*strideA elements, column-major with
 * leading dimension lda.  ipiv: batch*n (1-based), info: batch entries.
 * The factors stay on the device; solve_batched applies them to B
 * (n x nrhs per system, system i at B + i*strideB).
 */
int b200lu_factor_batched(b200lu_handle* h, int64_t batch, int64_t n,
                          const void* A_host, int64_t lda, int64_t strideA,
                          int64_t* ipiv_out, int64_t* info_out);
int b200lu_solve_batched(b200lu_handle* h, int64_t nrhs,
                         const void* B_host, int64_t ldb, int64_t strideB,
                         void* X_host, int64_t ldx, int64_t strideX);
/* The same getrs with op(A_i) = A_i^T ('T'/'C'; 'N' is b200lu_solve_batched): the adjoint
 * solve of a BlockDiagonal problem with the cached per-block factors
 * (`solve!(cache; adjoint = true)`, src/common.jl:1012-1027). */
int b200lu_solve_batched_trans(b200lu_handle* h, char trans, int64_t nrhs,
                               const void* B_host, int64_t ldb, int64_t strideB,
                               void* X_host, int64_t ldx, int64_t strideX);
int b200lu_factor_batched_device(b200lu_handle* h, int64_t batch, int64_t n,
                                 const void* A_dev, int64_t lda, int64_t strideA,
                                 int64_t* any_info);
int b200lu_solve_batched_device(b200lu_handle* h, int64_t nrhs,
                                const void* B_dev, int64_t ldb, int64_t strideB,
                                void* X_dev, int64_t ldx, int64_t strideX);
int b200lu_solve_batched_trans_device(b200lu_handle* h, char trans, int64_t nrhs,
                                      const void* B_dev, int64_t ldb, int64_t strideB,
                                      void* X_dev, int64_t ldx, int64_t strideX);
/* getrf of the batch AND getrs of one right-hand side per system (B: n values per system at
 * B + i*strideB, X likewise) in the SAME kernel for n <= 64 — what `solve!` on a fresh BlockDiagonal
 * cache does (per-block lu! followed by per-block ldiv!, ext/LinearSolveBlockDiagonalsExt.jl:119-125,
 * 183-205): A is read once, the factors are written once and stay cached for later
 * b200lu_solve_batched calls. */
int b200lu_factor_solve_batched(b200lu_handle* h, int64_t batch, int64_t n,
                                const void* A_host, int64_t lda, int64_t strideA,
                                const void* B_host, int64_t strideB, void* X_host, int64_t strideX,
                                int64_t* ipiv_out, int64_t* info_out);
int b200lu_factor_solve_batched_device(b200lu_handle* h, int64_t batch, int64_t n,
                                       const void* A_dev, int64_t lda, int64_t strideA,
                                       const void* B_dev, int64_t strideB, void* X_dev, int64_t strideX,
                                       int64_t* any_info);
int b200lu_get_factors_batched(b200lu_handle* h, void* LU_host, int64_t lda,
                               int64_t strideA, int64_t* ipiv_out, int64_t* info_out);

/*
 * One-process-per-GPU multi-GPU (1-D block-cyclic columns).  Rank 0 calls
 * b200lu_comm_unique_id, the host language broadcasts the 128 bytes (torch.distributed / MPI /
 * files), every rank calls b200lu_comm_init (nranks == 1 needs no id).  Afterwards
 * b200lu_factor_dist / b200lu_solve_dist (collective calls) operate on this rank's local column
 * blocks: global column block j (width nb) lives on rank j % nranks at local block index
 * j / nranks; lda must be a multiple of 16 bytes, A 16-byte aligned; the local matrix is factored
 * in place.  The factored panel travels by peer stores into device windows the ranks map with
 * cudaIpc (the NCCL communicator only carries the handle exchange); when the windows cannot be
 * mapped, or with B200LU_DIST_MODE=nccl in the environment, the panel is ncclBroadcast instead
 * (b200lu_dist_transport: 1 = peer stores, 0 = NCCL).  b200lu_solve_dist: B and X replicated
 * (n x nrhs on every rank), the factors stay distributed (no n x n replica, warm calls allocate
 * nothing); it needs the peer-store transport (status 5 otherwise).
 */
int b200lu_comm_unique_id(void* id128);
int b200lu_comm_init(b200lu_handle* h, const void* id128, int rank, int nranks);
int b200lu_dist_local_cols(const b200lu_handle* h, int64_t n, int64_t* ncols_local);
int b200lu_dist_transport(const b200lu_handle* h);
int b200lu_factor_dist(b200lu_handle* h, int64_t n, const void* Aloc_dev, int64_t lda,
                       int64_t* info);
int b200lu_solve_dist(b200lu_handle* h, int64_t nrhs, const void* B_dev, int64_t ldb,
                      void* X_dev, int64_t ldx);

/* Synthetic-input helper for benches/tests at sizes that do not fit the host:
 * fills an n x ncols column-major device block with U[0,1) (counter-based,
 * reproducible: element (i, j_global) depends only on seed, i, j_global), and
 * adds `diag_shift` where i == j_global. */
int b200lu_fill_uniform_device(b200lu_handle* h, void* A_dev, int64_t lda, int64_t n,
                               int64_t ncols, int64_t first_global_col,
                               int64_t col_block, int64_t col_block_stride,
                               uint64_t seed, double diag_shift);

#ifdef __cplusplus
}
#endif
#endif /* B200LU_H */
