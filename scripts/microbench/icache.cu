// icache.cu — cost of straight-line code larger than the instruction caches, at the occupancy of
// the panel kernel (8 warps per SM, all in the same phase).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache.bin icache.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int N>
__global__ void __launch_bounds__(256, 1) k(int iters, double x, double y, double* out, long long* cyc) {
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < iters; ++r) {
#pragma unroll
        for (int n = 0; n < N; ++n) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fma(acc[i], x, y + n);   // distinct immediates keep the copies apart
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int N>
void run(double* out, long long* cyc) {
    const int iters = 4096 / N < 4 ? 4 : 4096 / N;
    k<N><<<16, 256>>>(iters, 1.0000001, 1e-9, out, cyc);
    cudaDeviceSynchronize();
    k<N><<<16, 256>>>(iters, 1.0000001, 1e-9, out, cyc);
    cudaDeviceSynchronize();
    const double ninstr = (double)iters * N * 32;
    printf("body %5d DFMA (~%4d KB): %.2f cycles per DFMA warp-instruction per warp (8 warps/SM) -> %.2f per scheduler slot\n", N * 32, N * 32 * 16 / 1024 + N * 2 * 16 / 1024,
           cyc[0] / ninstr, cyc[0] / ninstr / 2);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 16 * 256 * 8); cudaMallocManaged(&cyc, 64);
    run<1>(out, cyc); run<4>(out, cyc); run<8>(out, cyc); run<16>(out, cyc); run<32>(out, cyc); run<48>(out, cyc); run<64>(out, cyc); run<96>(out, cyc); run<128>(out, cyc); run<256>(out, cyc);
    cudaError_t e = cudaGetLastError(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
