// exchange_latency.cu — microbenchmark of the per-column all-to-all candidate exchange of the
// panel factorization (one message of MSG 64-bit packets from every CTA to every CTA per round).
//   v1: L2 mailbox, single phase  (every CTA reads every message; LL packets {data32, flag32})
//   v2: L2 mailbox, two phases    (headers first, then the winner's row) — the panel v5 protocol
//   v3: cluster DSMEM stores + barrier.cluster per round
//   v4: cluster DSMEM stores of LL packets, receivers poll their OWN shared memory (no barrier)
//   v5: like v4, launched while a low-priority filler grid occupies every SM (scheduling latency)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_latency.bin exchange_latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int MSG = 66;      // packets per message (2 header + 64 row words)
constexpr int GMAX = 64;
constexpr long long TIMEOUT = 2000000000LL;

__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned d, unsigned f) {
    unsigned long long v = ((unsigned long long)f << 32) | d;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct Mail { unsigned long long m[2][GMAX][128]; };

// v1: single phase
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_l2_single(Mail* mail, int G, int rounds, unsigned epoch, long long* out, unsigned* sink) {
    __shared__ unsigned s_acc[NT];
    const int tid = threadIdx.x, cta = blockIdx.x;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int par = r & 1;
        const unsigned want = epoch + r + 1;
        if (tid < MSG) ll_store(&mail->m[par][cta][tid], acc + tid, want);
        const int total = G * MSG;
        for (int i = tid; i < total; i += NT) {
            const int c = i / MSG, k = i - c * MSG;
            unsigned long long v = ll_load(&mail->m[par][c][k]);
            long long ts = clock64();
            while ((unsigned)(v >> 32) != want) {
                v = ll_load(&mail->m[par][c][k]);
                if (clock64() - ts > TIMEOUT) { out[1] = -1; return; }
            }
            acc += (unsigned)v;
        }
        s_acc[tid] = acc;
        __syncthreads();
        acc = s_acc[(tid + 1) % NT];
        __syncthreads();
    }
    long long t1 = clock64();
    if (cta == 0 && tid == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
}

// v2: two phases (header poll by warp 0, then NT/2.. threads fetch two rows)
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_l2_two(Mail* mail, int G, int rounds, unsigned epoch, long long* out, unsigned* sink, int sleep_ns) {
    __shared__ unsigned s_win;
    __shared__ unsigned s_rows[128];
    const int tid = threadIdx.x, cta = blockIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int par = r & 1;
        const unsigned want = epoch + r + 1;
        if (tid < MSG) ll_store(&mail->m[par][cta][tid], acc + tid + (tid == 0 ? cta : 0), want);
        if (warp == 0) {
            unsigned best = 0;
            for (int c = lane; c < G; c += 32) {
                unsigned long long v0 = ll_load(&mail->m[par][c][0]);
                unsigned long long v1 = ll_load(&mail->m[par][c][1]);
                long long ts = clock64();
                while ((unsigned)(v0 >> 32) != want) { if (sleep_ns) __nanosleep(sleep_ns); v0 = ll_load(&mail->m[par][c][0]); if (clock64() - ts > TIMEOUT) { out[1] = -1; return; } }
                while ((unsigned)(v1 >> 32) != want) { if (sleep_ns) __nanosleep(sleep_ns); v1 = ll_load(&mail->m[par][c][1]); if (clock64() - ts > TIMEOUT) { out[1] = -1; return; } }
                best = max(best, ((unsigned)v0 & 0xffff) + (unsigned)v1 % 3);
            }
            best = __reduce_max_sync(0xffffffffu, best);
            if (lane == 0) s_win = (best + r) % G;
        }
        __syncthreads();
        const int win = s_win;
        if (tid < 128) {
            const int c = tid < 64 ? win : 0;
            const unsigned long long* src = &mail->m[par][c][2 + (tid & 63)];
            unsigned long long v = ll_load(src);
            long long ts = clock64();
            while ((unsigned)(v >> 32) != want) { v = ll_load(src); if (clock64() - ts > TIMEOUT) { out[1] = -1; return; } }
            s_rows[tid] = (unsigned)v;
        }
        __syncthreads();
        acc += s_rows[tid & 127];
        __syncthreads();
    }
    long long t1 = clock64();
    if (cta == 0 && tid == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(unsigned addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

// v3: DSMEM + cluster barrier
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_dsmem_bar(int rounds, long long* out, unsigned* sink) {
    __shared__ __align__(16) unsigned long long box[2][16][MSG + 2];
    __shared__ unsigned s_acc[NT];
    const int tid = threadIdx.x;
    const unsigned me = cluster_rank(), G = cluster_size();
    unsigned acc = 0;
    cluster_sync_();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int par = r & 1;
        const int total = G * MSG;
        for (int i = tid; i < total; i += NT) {
            const unsigned dst = i / MSG, k = i - dst * MSG;
            st_cluster_u64(mapa(smem_u32(&box[par][me][k]), dst), ((unsigned long long)r << 32) | (acc + k));
        }
        cluster_sync_();
        const int win = (acc + r) % G;
        if (tid < MSG) acc += (unsigned)box[par][win][tid];
        s_acc[tid] = acc;
        __syncthreads();
        acc = s_acc[0] + s_acc[MSG - 1];
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync_();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
}

// v4: DSMEM LL packets, poll own smem
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_dsmem_ll(int rounds, long long* out, unsigned* sink, int hdr_only_first) {
    __shared__ __align__(16) unsigned long long box[2][16][MSG + 2];
    __shared__ unsigned s_acc[NT];
    __shared__ unsigned s_win;
    const int tid = threadIdx.x;
    const unsigned me = cluster_rank(), G = cluster_size();
    for (int i = tid; i < 2 * 16 * (MSG + 2); i += NT) (&box[0][0][0])[i] = 0;
    unsigned acc = 0;
    cluster_sync_();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int par = r & 1;
        const unsigned want = r + 1;
        const int total = G * MSG;
        for (int i = tid; i < total; i += NT) {
            const unsigned dst = i / MSG, k = i - dst * MSG;
            st_cluster_u64(mapa(smem_u32(&box[par][me][k]), dst), ((unsigned long long)want << 32) | ((acc + k) & 0xffff));
        }
        // phase 1: warp 0 polls the G headers (2 packets each) in its own shared memory
        if (tid < 32) {
            unsigned best = 0;
            if (tid < (int)G) {
                volatile unsigned long long* p0 = &box[par][tid][0];
                volatile unsigned long long* p1 = &box[par][tid][1];
                unsigned long long v0 = *p0, v1 = *p1;
                long long ts = clock64();
                while ((unsigned)(v0 >> 32) != want || (unsigned)(v1 >> 32) != want) {
                    v0 = *p0; v1 = *p1;
                    if (clock64() - ts > TIMEOUT) { out[1] = -1; break; }
                }
                best = (unsigned)v0 + (unsigned)v1;
            }
            best = __reduce_max_sync(0xffffffffu, best);
            if (tid == 0) s_win = (best + r) % G;
        }
        __syncthreads();
        const int win = s_win;
        if (tid < MSG) {
            volatile unsigned long long* p = &box[par][win][tid];
            unsigned long long v = *p;
            long long ts = clock64();
            while ((unsigned)(v >> 32) != want) { v = *p; if (clock64() - ts > TIMEOUT) { out[1] = -1; break; } }
            acc += (unsigned)v;
        }
        s_acc[tid] = acc;
        __syncthreads();
        acc = s_acc[0] + s_acc[MSG - 1];
        // double buffering by parity is safe: a peer can only be one round ahead of the slowest
        // reader, because its round r+2 stores need MY round r+1 message first.
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync_();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(128, 2) k_filler(long long cycles, unsigned* sink) {
    extern __shared__ unsigned char fsm[];
    long long t0 = clock64();
    unsigned a = threadIdx.x;
    while (clock64() - t0 < cycles) a = a * 1664525u + 1013904223u;
    if (a == 0x12345678u) *sink = a + fsm[0];
}

template <typename K, typename... Args>
static float launch_cluster(K kern, int G, int NT, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e != cudaSuccess) { printf("  cluster launch G=%d failed: %s\n", G, cudaGetErrorString(e)); cudaGetLastError(); return -1.f; }
    return 0.f;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    Mail* mail; CK(cudaMalloc(&mail, sizeof(Mail))); CK(cudaMemset(mail, 0, sizeof(Mail)));
    long long* out; CK(cudaMallocManaged(&out, 64));
    unsigned* sink; CK(cudaMalloc(&sink, 64));
    const int rounds = 2000;
    unsigned epoch = 0;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int Gs[] = {2, 8, 16, 32, 64};
    for (int G : Gs) {
        for (int rep = 0; rep < 2; ++rep) {
            out[0] = out[1] = 0;
            CK(cudaEventRecord(e0));
            k_l2_single<256><<<G, 256>>>(mail, G, rounds, epoch, out, sink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            epoch += rounds + 8;
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) printf("v1 L2 single-phase   G=%2d NT=256: %7.0f cycles/round  (%.3f us/round by events) err=%lld\n", G, (double)out[0] / rounds, ms * 1e3 / rounds, out[1]);
        }
    }
    for (int sleep_ns : {0, 20}) for (int G : Gs) {
        for (int rep = 0; rep < 2; ++rep) {
            out[0] = out[1] = 0;
            CK(cudaEventRecord(e0));
            k_l2_two<256><<<G, 256>>>(mail, G, rounds, epoch, out, sink, sleep_ns);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            epoch += rounds + 8;
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) printf("v2 L2 two-phase sl=%2d G=%2d NT=256: %7.0f cycles/round  (%.3f us/round) err=%lld\n", sleep_ns, G, (double)out[0] / rounds, ms * 1e3 / rounds, out[1]);
        }
    }
    CK(cudaFuncSetAttribute(k_dsmem_bar<512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_dsmem_ll<512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_dsmem_bar<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_dsmem_ll<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int Cs[] = {2, 4, 8, 16};
    for (int G : Cs) {
        for (int rep = 0; rep < 2; ++rep) {
            out[0] = out[1] = 0;
            CK(cudaEventRecord(e0));
            float r = launch_cluster(k_dsmem_bar<512>, G, 512, 0, rounds, out, sink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && r == 0) printf("v3 DSMEM + cluster barrier G=%2d NT=512: %7.0f cycles/round (%.3f us/round)\n", G, (double)out[0] / rounds, ms * 1e3 / rounds);
        }
    }
    for (int G : Cs) {
        for (int rep = 0; rep < 2; ++rep) {
            out[0] = out[1] = 0;
            CK(cudaEventRecord(e0));
            float r = launch_cluster(k_dsmem_ll<512>, G, 512, 0, rounds, out, sink, 0);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && r == 0) printf("v4 DSMEM LL (poll own smem) G=%2d NT=512: %7.0f cycles/round (%.3f us/round) err=%lld\n", G, (double)out[0] / rounds, ms * 1e3 / rounds, out[1]);
        }
    }
    for (int G : Cs) {
        for (int rep = 0; rep < 2; ++rep) {
            out[0] = out[1] = 0;
            CK(cudaEventRecord(e0));
            float r = launch_cluster(k_dsmem_ll<256>, G, 256, 0, rounds, out, sink, 0);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && r == 0) printf("v4 DSMEM LL (poll own smem) G=%2d NT=256: %7.0f cycles/round (%.3f us/round) err=%lld\n", G, (double)out[0] / rounds, ms * 1e3 / rounds, out[1]);
        }
    }
    // v5: scheduling latency of a short cluster kernel while a filler grid holds every SM
    {
        int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        cudaStream_t s_lo, s_hi;
        CK(cudaStreamCreateWithPriority(&s_lo, cudaStreamNonBlocking, lo));
        CK(cudaStreamCreateWithPriority(&s_hi, cudaStreamNonBlocking, hi));
        CK(cudaFuncSetAttribute(k_filler, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        for (int fill = 0; fill < 2; ++fill) for (int G : {8, 16}) {
            const int nlaunch = 20, short_rounds = 32;
            if (fill) k_filler<<<148 * 2 * 40, 128, 100 * 1024, s_lo>>>(60000, sink);  // ~30 us CTAs, 2 per SM, ~1.2 ms total
            CK(cudaEventRecord(e0, s_hi));
            for (int i = 0; i < nlaunch; ++i) launch_cluster(k_dsmem_ll<512>, G, 512, s_hi, short_rounds, out, sink, 0);
            CK(cudaEventRecord(e1, s_hi));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("v5 %s cluster G=%2d x %d launches of %d rounds: %.1f us per launch\n", fill ? "WITH filler" : "no filler  ", G, nlaunch, short_rounds, ms * 1e3 / nlaunch);
        }
        // same for the classic (non-cluster) 32-CTA spin-synchronised kernel
        for (int fill = 0; fill < 2; ++fill) {
            const int nlaunch = 20, short_rounds = 32;
            if (fill) k_filler<<<148 * 2 * 40, 128, 100 * 1024, s_lo>>>(60000, sink);
            CK(cudaEventRecord(e0, s_hi));
            for (int i = 0; i < nlaunch; ++i) { k_l2_single<256><<<32, 256, 0, s_hi>>>(mail, 32, short_rounds, epoch, out, sink); epoch += short_rounds + 8; }
            CK(cudaEventRecord(e1, s_hi));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("v5 %s L2 single G=32 x %d launches of %d rounds: %.1f us per launch\n", fill ? "WITH filler" : "no filler  ", nlaunch, short_rounds, ms * 1e3 / nlaunch);
        }
    }
    printf("done\n");
    return 0;
}
