// cluster_sched.cu — how long does a 16-CTA cluster kernel (512 threads, needs a whole SM's
// registers) wait for placement while a GEMM-like low-priority grid holds every SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_sched.bin cluster_sched.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 2) k_filler(long long cycles, unsigned* sink) {
    extern __shared__ unsigned char fsm[];
    long long t0 = clock64();
    unsigned a = threadIdx.x;
    while (clock64() - t0 < cycles) a = a * 1664525u + 1013904223u;
    if (a == 0x12345678u) *sink = a + fsm[0];
}
__global__ void __launch_bounds__(512, 1) k_cluster(long long cycles, unsigned* sink) {
    __shared__ unsigned char big[20 * 1024];
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    long long t0 = clock64();
    unsigned a = threadIdx.x;
    while (clock64() - t0 < cycles) a = a * 1664525u + 1013904223u;
    if (a == 0x12345678u) *sink = a + big[threadIdx.x];
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__global__ void __launch_bounds__(256, 1) k_plain(long long cycles, unsigned* sink) {
    __shared__ unsigned char big[20 * 1024];
    long long t0 = clock64();
    unsigned a = threadIdx.x;
    while (clock64() - t0 < cycles) a = a * 1664525u + 1013904223u;
    if (a == 0x12345678u) *sink = a + big[threadIdx.x];
}

static void launch_cluster(int G, cudaStream_t st, long long cycles, unsigned* sink) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_cluster, cycles, sink));
}

int main() {
    CK(cudaSetDevice(0));
    unsigned* sink; CK(cudaMalloc(&sink, 64));
    int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s_lo, s_hi, s_same;
    CK(cudaStreamCreateWithPriority(&s_lo, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&s_hi, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&s_same, cudaStreamNonBlocking, lo));
    CK(cudaFuncSetAttribute(k_filler, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    CK(cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaEvent_t e0, e1, f0, f1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
    const long long work = 40000;       // ~20 us of "panel" work per launch
    const long long fill_cta = 60000;   // ~30 us per filler CTA
    const int nlaunch = 20;
    // warm up
    k_filler<<<296, 128, 110 * 1024, s_lo>>>(1000, sink); launch_cluster(16, s_hi, 1000, sink); k_plain<<<32, 256, 0, s_hi>>>(1000, sink);
    CK(cudaDeviceSynchronize());
    for (int mode = 0; mode < 3; ++mode) {          // 0: no filler, 1: filler + high priority, 2: filler + same priority
        for (int kind = 0; kind < 4; ++kind) {      // 0: cluster 16, 1: cluster 8, 2: plain 32 CTAs x256, 3: cluster 4
            cudaStream_t st = (mode == 2) ? s_same : s_hi;
            CK(cudaEventRecord(f0, s_lo));
            if (mode) k_filler<<<148 * 2 * 30, 128, 110 * 1024, s_lo>>>(fill_cta, sink);   // 30 waves x 30 us = 0.9 ms
            CK(cudaEventRecord(f1, s_lo));
            CK(cudaEventRecord(e0, st));
            for (int i = 0; i < nlaunch; ++i) {
                if (kind == 0) launch_cluster(16, st, work, sink);
                else if (kind == 1) launch_cluster(8, st, work, sink);
                else if (kind == 3) launch_cluster(4, st, work, sink);
                else k_plain<<<32, 256, 0, st>>>(work, sink);
            }
            CK(cudaEventRecord(e1, st));
            CK(cudaDeviceSynchronize());
            float ms, fms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaEventElapsedTime(&fms, f0, f1));
            const char* kn[] = {"cluster16x512", "cluster8x512 ", "plain 32x256 ", "cluster4x512 "};
            const char* mn[] = {"no filler          ", "filler, high prio  ", "filler, same prio  "};
            printf("%s %s: %7.1f us per launch (work itself ~%.1f us); filler took %.2f ms\n", mn[mode], kn[kind], ms * 1e3 / nlaunch, work / 1965.0, fms);
        }
    }
    return 0;
}
