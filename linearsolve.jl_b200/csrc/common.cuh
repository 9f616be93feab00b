// common.cuh — shared device helpers for the b200lu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "b200lu kernels are written for sm_100a (B200) only"
#endif

namespace b200lu {

// Global launch counter (bench reports gpu_launches from it).
extern std::atomic<unsigned long long> g_launch_count;   // handles on different host threads count concurrently
#define B200LU_COUNT_LAUNCH() (::b200lu::g_launch_count.fetch_add(1, std::memory_order_relaxed))

// ---- error word written by device code (spin-wait watchdogs etc.) ----------
enum DevErr : int {
    DEV_OK = 0,
    DEV_ERR_PANEL_TIMEOUT = 1,
    DEV_ERR_TRSV_TIMEOUT = 2,
    DEV_ERR_GEMM_TIMEOUT = 3,
    DEV_ERR_PEER_TIMEOUT = 4,   // a peer GPU's panel / right-hand side never arrived
};

// ~2 s at 1.9 GHz: a spin loop that runs this long means a lost CTA, not work.
__device__ constexpr long long kSpinTimeoutCycles = 4000000000LL;

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// L2-only (L1-bypassing) loads for data another CTA has just produced.
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) {
    return __ldcg(p);
}

// ---- cp.async (LDGSTS) 16-byte copies with zero-fill predicate --------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- FP64 tensor-core MMA (DMMA.8x8x4 is the native sm_100 shape) -----------
// D(8x8) += A(8x4, row) * B(4x8, col).  Lane T holds a = A[T/4][T%4],
// b = B[T%4][T/4], c0,c1 = C[T/4][2*(T%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double neg_bits(double x) {
    // sign flip on the integer pipe (keeps the FP64 pipe for the MMAs)
    long long v = __double_as_longlong(x) ^ 0x8000000000000000LL;
    return __longlong_as_double(v);
}

template <typename T>
__device__ __forceinline__ T tabs(T x);
template <>
__device__ __forceinline__ double tabs<double>(double x) { return fabs(x); }
template <>
__device__ __forceinline__ float tabs<float>(float x) { return fabsf(x); }

template <typename T>
__device__ __forceinline__ T tfma(T a, T b, T c);
template <>
__device__ __forceinline__ double tfma<double>(double a, double b, double c) { return fma(a, b, c); }
template <>
__device__ __forceinline__ float tfma<float>(float a, float b, float c) { return fmaf(a, b, c); }

template <typename T>
__device__ __forceinline__ T shfl(T v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
template <typename T>
__device__ __forceinline__ T shfl_xor(T v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace b200lu
