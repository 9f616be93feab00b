// latency.cu — dependent-chain latencies (cycles) of the instructions on the panel kernel's
// critical path, one warp and eight warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency.bin latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#define N 512
__global__ void k(long long* out, double* sink, int dummy) {
    __shared__ double sm[64];
    __shared__ unsigned long long mbar;
    const int lane = threadIdx.x & 31;
    unsigned u = threadIdx.x * 2654435761u + dummy;
    double d = 1.0 + threadIdx.x * 1e-3;
    long long t0, t1;
    int slot = 0;
    sm[lane] = d; sm[lane + 32] = d;
    __syncthreads();
#define REC() if (threadIdx.x == 0) out[slot] = (t1 - t0); slot++;
    // redux.sync max chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = __reduce_max_sync(0xffffffffu, u ^ lane) + i;
    t1 = clock64(); REC();
    // ballot chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = __ballot_sync(0xffffffffu, (u >> (lane & 7)) & 1) + i;
    t1 = clock64(); REC();
    // shfl chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = __shfl_sync(0xffffffffu, u, (lane + 1) & 31) + i;
    t1 = clock64(); REC();
    // lds chain (pointer chasing through smem)
    unsigned idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = ((unsigned)__double_as_longlong(sm[idx & 63]) + idx) & 63;
    t1 = clock64(); REC();
    u += idx;
    // dfma chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) d = fma(d, 1.0000001, 1e-9);
    t1 = clock64(); REC();
    // double division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) d = 1.0 / (d + 1.5);
    t1 = clock64(); REC();
    // __syncthreads chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) __syncthreads();
    t1 = clock64(); REC();
    // dsetp/sel chain (compare + select on doubles)
    double e = d;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { e = (fabs(e) > d) ? e * 0.5 : e + 1.0; }
    t1 = clock64(); REC();
    // st.async to self + mbarrier wait
    if (threadIdx.x == 0) {
        unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        unsigned dst = (unsigned)__cvta_generic_to_shared(&sm[0]);
        unsigned rdst, rmb;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(rdst) : "r"(dst));
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(rmb) : "r"(mb));
        t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8;" ::"r"(mb) : "memory");
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];" ::"r"(rdst), "l"((unsigned long long)i), "r"(rmb) : "memory");
            }
            unsigned ok = 0;
            while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(mb), "r"(i & 1) : "memory");
        }
        t1 = clock64();
        if (threadIdx.x == 0) out[slot] = (t1 - t0) * (N / 64);
    }
    slot++;
    sink[threadIdx.x] = d + e + u;
}
int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 256); cudaMalloc(&sink, 1024 * 8);
    const char* names[] = {"redux.sync(max)+add", "ballot+add", "shfl+add", "LDS pointer chase", "DFMA", "double division (+add)", "__syncthreads", "dsetp+select+dmul/dadd", "st.async(self)+mbarrier wait"};
    for (int nt : {32, 256}) {
        k<<<1, nt>>>(out, sink, 0); cudaDeviceSynchronize();
        k<<<1, nt>>>(out, sink, 1); cudaDeviceSynchronize();
        printf("threads per CTA = %d\n", nt);
        for (int i = 0; i < 9; ++i) printf("  %-32s %7.1f cycles per op\n", names[i], (double)out[i] / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
