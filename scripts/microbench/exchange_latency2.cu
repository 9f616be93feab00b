// exchange_latency2.cu — second round: find the floor of the cluster exchange.
//   b1: barrier.cluster arrive+wait only
//   p1: DSMEM ping-pong (plain st.shared::cluster, poll own smem)       -> one-way latency
//   p2: L2 ping-pong (LL packet)                                          -> one-way latency
//   p3: st.async + mbarrier complete_tx ping-pong
//   m1: all-to-all of 272-byte messages with st.async.v2.f64 + mbarrier tx-count, one wait per round
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_latency2.bin exchange_latency2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr long long TIMEOUT = 2000000000LL;

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
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned addr, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned addr, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(unsigned addr, unsigned parity) {
    if (mbar_try_wait(addr, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(addr, parity)) if (clock64() - t0 > TIMEOUT) return false;
    return true;
}
__device__ __forceinline__ void st_async_v2f64(unsigned raddr, double a, double b, unsigned rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "d"(a), "d"(b), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void st_async_u64(unsigned raddr, unsigned long long a, unsigned rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];"
                 ::"r"(raddr), "l"(a), "r"(rmbar) : "memory");
}

__global__ void __launch_bounds__(512, 1) k_barrier_only(int rounds, long long* out) {
    cluster_sync_();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) cluster_sync_();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// p1: cluster of 2, thread 0 of each CTA
__global__ void __launch_bounds__(32, 1) k_pp_dsmem(int rounds, long long* out) {
    __shared__ volatile unsigned long long slot;
    const unsigned me = cluster_rank();
    if (threadIdx.x == 0) slot = 0;
    cluster_sync_();
    if (threadIdx.x == 0) {
        const unsigned peer = mapa(smem_u32((const void*)&slot), me ^ 1);
        long long t0 = clock64();
        for (int r = 1; r <= rounds; ++r) {
            if (me == 0) {
                st_cluster_u64(peer, r);
                while (slot != (unsigned long long)r) {}
            } else {
                while (slot != (unsigned long long)r) {}
                st_cluster_u64(peer, r);
            }
        }
        long long t1 = clock64();
        if (me == 0) out[0] = t1 - t0;
    }
    cluster_sync_();
}

__global__ void __launch_bounds__(32, 1) k_pp_l2(unsigned long long* slots, int rounds, unsigned base, long long* out) {
    const int me = blockIdx.x;
    if (threadIdx.x == 0) {
        unsigned long long* mine = slots + me * 32;
        unsigned long long* theirs = slots + (me ^ 1) * 32;
        long long t0 = clock64();
        for (int r = 1; r <= rounds; ++r) {
            const unsigned long long want = base + r;
            unsigned long long v;
            if (me == 0) {
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(theirs), "l"(want) : "memory");
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory"); } while (v != want);
            } else {
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory"); } while (v != want);
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(theirs), "l"(want) : "memory");
            }
        }
        long long t1 = clock64();
        if (me == 0) out[0] = t1 - t0;
    }
}

// p3: st.async ping-pong with mbarrier
__global__ void __launch_bounds__(32, 1) k_pp_async(int rounds, long long* out) {
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ __align__(16) unsigned long long slot[2];
    const unsigned me = cluster_rank();
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_();
    if (threadIdx.x == 0) {
        const unsigned pslot = mapa(smem_u32(&slot[0]), me ^ 1);
        const unsigned pbar = mapa(smem_u32(&mbar), me ^ 1);
        const unsigned mybar = smem_u32(&mbar);
        long long t0 = clock64();
        bool ok = true;
        for (int r = 0; r < rounds && ok; ++r) {
            mbar_expect_tx(mybar, 8);
            if (me == 0) {
                st_async_u64(pslot, r, pbar);
                ok = mbar_wait(mybar, r & 1);
            } else {
                ok = mbar_wait(mybar, r & 1);
                st_async_u64(pslot, r, pbar);
            }
        }
        long long t1 = clock64();
        if (me == 0) { out[0] = t1 - t0; out[1] = ok ? (long long)slot[0] : -1; }
    }
    cluster_sync_();
}

// m1: all-to-all with st.async + tx-count.  Message = 17 x 16 bytes (hdr + 32 doubles).
constexpr int MSGV = 17;
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_a2a_async(int rounds, long long* out, double* sink) {
    __shared__ __align__(16) double box[2][16][2 * MSGV];
    __shared__ __align__(16) double stage[2 * MSGV];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int tid = threadIdx.x;
    const unsigned me = cluster_rank(), G = cluster_size();
    if (tid == 0) {
        mbar_init(smem_u32(&mbar[0]), 1);
        mbar_init(smem_u32(&mbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_();
    double acc = 0.0, chk = 0.0;
    bool ok = true;
    // precompute my destination
    const int dst = tid / MSGV, k = tid - dst * MSGV;
    const bool sender = dst < (int)G;
    unsigned raddr[2], rbar[2];
    for (int p = 0; p < 2; ++p) {
        raddr[p] = mapa(smem_u32(&box[p][me][2 * k]), sender ? dst : 0);
        rbar[p] = mapa(smem_u32(&mbar[p]), sender ? dst : 0);
    }
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int par = r & 1;
        // stage my message: value depends on (me, r) so the receiver can check it
        if (tid < 2 * MSGV) stage[tid] = (double)(me * 100000 + r) + tid * 0.001 + acc * 0.0;
        if (tid == 0) mbar_expect_tx(smem_u32(&mbar[par]), G * MSGV * 16);
        __syncthreads();
        if (sender) st_async_v2f64(raddr[par], stage[2 * k], stage[2 * k + 1], rbar[par]);
        if (!mbar_wait(smem_u32(&mbar[par]), (r >> 1) & 1)) { ok = false; break; }
        // every warp reduces the G headers redundantly, then reads the winner's row
        const int lane = tid & 31;
        double hv = lane < (int)G ? box[par][lane][0] : -1.0;
        unsigned key = __double2uint_rd(hv);
        unsigned mx = __reduce_max_sync(0xffffffffu, key);
        const int win = (mx / 100000) % G;  // == G-1
        acc = box[par][win][2 + (lane & 31)];
        chk += acc;
    }
    long long t1 = clock64();
    cluster_sync_();
    if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = ok ? 0 : -1; }
    if (tid == 33) sink[blockIdx.x] = chk;
}

template <typename K, typename... Args>
static int launch_cluster(K kern, int G, int NT, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e != cudaSuccess) { printf("  cluster launch G=%d failed: %s\n", G, cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    return 0;
}

int main() {
    CK(cudaSetDevice(0));
    long long* out; CK(cudaMallocManaged(&out, 64));
    double* sink; CK(cudaMallocManaged(&sink, 16 * 8));
    unsigned long long* slots; CK(cudaMalloc(&slots, 4096)); CK(cudaMemset(slots, 0, 4096));
    const int rounds = 2000;
    CK(cudaFuncSetAttribute(k_barrier_only, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_a2a_async<512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_a2a_async<1024>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int G : {2, 4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) { out[0] = 0; launch_cluster(k_barrier_only, G, 512, 0, rounds, out); CK(cudaDeviceSynchronize()); }
        printf("b1 barrier.cluster only G=%2d NT=512: %6.0f cycles/round\n", G, (double)out[0] / rounds);
    }
    for (int rep = 0; rep < 2; ++rep) { out[0] = 0; launch_cluster(k_pp_dsmem, 2, 32, 0, rounds, out); CK(cudaDeviceSynchronize()); }
    printf("p1 DSMEM ping-pong: RTT %6.0f cycles (one-way %.0f)\n", (double)out[0] / rounds, (double)out[0] / rounds / 2);
    unsigned base = 0;
    for (int rep = 0; rep < 2; ++rep) { out[0] = 0; k_pp_l2<<<2, 32>>>(slots, rounds, base, out); CK(cudaDeviceSynchronize()); base += rounds + 1; }
    printf("p2 L2 ping-pong:    RTT %6.0f cycles (one-way %.0f)\n", (double)out[0] / rounds, (double)out[0] / rounds / 2);
    for (int rep = 0; rep < 2; ++rep) { out[0] = 0; out[1] = 0; launch_cluster(k_pp_async, 2, 32, 0, rounds, out); CK(cudaDeviceSynchronize()); }
    printf("p3 st.async+mbarrier ping-pong: RTT %6.0f cycles (one-way %.0f) last=%lld\n", (double)out[0] / rounds, (double)out[0] / rounds / 2, out[1]);
    for (int G : {2, 4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) { out[0] = 0; out[1] = 0; launch_cluster(k_a2a_async<512>, G, 512, 0, rounds, out, sink); CK(cudaDeviceSynchronize()); }
        // expected chk on thread 33 (lane 1): sum over r of ((G-1)*1000 + r) + (2+1)*0.001
        double expect = 0; for (int r = 0; r < rounds; ++r) expect += (double)((G - 1) * 100000 + r) + 3 * 0.001;
        printf("m1 st.async all-to-all G=%2d NT=512: %6.0f cycles/round err=%lld check %s (%.3f vs %.3f)\n", G, (double)out[0] / rounds, out[1], fabs(sink[0] - expect) < 1e-6 * expect ? "OK" : "MISMATCH", sink[0], expect);
    }
    for (int G : {8, 16}) {
        for (int rep = 0; rep < 2; ++rep) { out[0] = 0; out[1] = 0; launch_cluster(k_a2a_async<1024>, G, 1024, 0, rounds, out, sink); CK(cudaDeviceSynchronize()); }
        printf("m1 st.async all-to-all G=%2d NT=1024: %6.0f cycles/round err=%lld\n", G, (double)out[0] / rounds, out[1]);
    }
    printf("done\n");
    return 0;
}
