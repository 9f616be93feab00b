// util_kernels.cuh — small bandwidth kernels around the LU path: precision
// casts for the FP32-factor mode (reference `A_32 .= T32.(A)`,
// src/openblas.jl:497-500), norms/axpy for the FP64 refinement loop, iota, and
// the counter-based synthetic fill used by benches at sizes that do not fit
// the host.
#pragma once
#include "common.cuh"

namespace b200lu {

// gridDim.y carries a column index in several kernels here: the hardware limit of that dimension is
// 65535, so launchers clamp it and the kernels stride over the columns
inline unsigned grid_y(long long cols) { return (unsigned)(cols < 1 ? 1 : (cols > 32768 ? 32768 : cols)); }

__global__ void iota_kernel(int* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

template <typename TS, typename TD>
__global__ void cast2d_kernel(const TS* __restrict__ S, long long lds, TD* __restrict__ D,
                              long long ldd, int rows, int cols) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int c = blockIdx.y; c < cols; c += gridDim.y)
        D[(long long)c * ldd + r] = (TD)S[(long long)c * lds + r];
}

// Sums of squares for the refinement loop and the residual check.  Every reduction here has ONE fixed shape
// (strided per-thread sums, shuffle tree, shared-memory tree, a single writer): the norms — and with them the
// sweep count of the refinement — are the same bits on every run.
__device__ __forceinline__ double block_sum_fixed(double s) {
    __shared__ double sh[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __syncthreads();   // sh may still be read by the previous call
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    return s;   // valid in thread 0
}

// out[0] = sum x^2 (one CTA)
__global__ void sumsq_kernel(const double* __restrict__ x, long long n, double* out) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s = fma(x[i], x[i], s);
    s = block_sum_fixed(s);
    if (threadIdx.x == 0) out[0] = s;
}

// Frobenius norm^2 of an n x n column-major matrix with leading dimension lda, stage 1: CTA (bx, by) sums its
// 256-row strip over the columns by, by + gridDim.y, ... into part[by * gridDim.x + bx]
__global__ void sumsq2d_kernel(const double* __restrict__ A, long long lda, int n, double* __restrict__ part) {
    double s = 0.0;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        for (int c = blockIdx.y; c < n; c += gridDim.y) {
            const double v = A[(long long)c * lda + r];
            s = fma(v, v, s);
        }
    }
    s = block_sum_fixed(s);
    if (threadIdx.x == 0) part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = s;
}
// stage 2 (one CTA): out[0] = sum of the np partial sums, fixed order
__global__ void sum_partials_kernel(const double* __restrict__ part, long long np, double* out) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < np; i += blockDim.x) s += part[i];
    s = block_sum_fixed(s);
    if (threadIdx.x == 0) out[0] = s;
}

// per-column sums of squares of an n x ncols block: out[c] = sum_i A[i, c]^2 (one CTA per column)
__global__ void colsumsq_kernel(const double* __restrict__ A, long long lda, int n, double* __restrict__ out, int ncols) {
    for (int c = blockIdx.x; c < ncols; c += gridDim.x) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double v = A[(long long)c * lda + i];
            s = fma(v, v, s);
        }
        s = block_sum_fixed(s);
        if (threadIdx.x == 0) out[c] = s;
    }
}

// X[:, c] += double(D[:, c]) for an n x ncols block
__global__ void axpy_f32_cols_kernel(double* __restrict__ X, long long ldx, const float* __restrict__ D, long long ldd, int n, int ncols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (long long c = blockIdx.y; c < ncols; c += gridDim.y) X[c * ldx + i] += (double)D[c * ldd + i];
}

__global__ void axpy_f32_kernel(double* __restrict__ x, const float* __restrict__ d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += (double)d[i];
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// A[i, jl] = U[0,1)(seed, i, j_global) (+ diag_shift on the diagonal);
// local column jl maps to global column
//   first_global_col + (jl / col_block) * col_block_stride + jl % col_block.
template <typename T>
__global__ void fill_uniform_kernel(T* __restrict__ A, long long lda, long long n, long long ncols,
                                    long long first_global_col, long long col_block,
                                    long long col_block_stride, unsigned long long seed,
                                    double diag_shift) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (long long jl = blockIdx.y; jl < ncols; jl += gridDim.y) {
        const long long jg = first_global_col + (jl / col_block) * col_block_stride + jl % col_block;
        const unsigned long long h =
            splitmix64(seed ^ splitmix64((unsigned long long)jg * 0x100000001B3ULL + (unsigned long long)i));
        double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
        if (i == jg) u += diag_shift;
        A[jl * lda + i] = (T)u;
    }
}

// ---- roofline probes: register-resident FP64 issue loops (no memory traffic) ----
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* sink, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(double* sink, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[threadIdx.x] = s;
}

}  // namespace b200lu
