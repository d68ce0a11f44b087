// umma_probe.cu — issue-rate microbenchmark of tcgen05.mma (1-CTA vs CTA-pair) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_probe tools/umma_probe.cu
// Each launch: one cluster (1 or 2 CTAs); the leader issues `n_mma` back-to-back MMAs of shape (128*G) x N x 16
// on whatever is in shared memory, commits, waits, and reports cycles per MMA.
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

template <int G>
__global__ void __launch_bounds__(192, 1) probe(int N, int n_mma, int mode, int b_rows_off, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2;   // pollers spin on this one until the leader thread is done
    __shared__ uint32_t tmem_slot;
    uint32_t rank = 0;
    if (G == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(mode >= 3 ? n_mma / 4 + 1 : 1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (G == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (G == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t*)&tmem_slot;
    if (mode >= 6 && warp >= 2 && rank == 0) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar2)) : "memory");
    }
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t id = idesc(128 * G, N);
        const uint64_t a0 = desc_sw128(base), b0 = desc_sw128(base + 128 * 128 + b_rows_off * 128);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            // mode 0: all MMAs accumulate into the same columns; mode 1: alternate between two column ranges;
            // mode 2: same columns, operands advance through 4 k-slices and 4 stages like a real mainloop
            if (mode == 5 && (i & 3) == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t d = tmem + ((mode == 1 && (i & 1)) ? 256u : 0u);
            uint64_t a = a0, b = b0;
            if (mode == 2) { a += (uint64_t)((i & 3) * 2) + (uint64_t)(((i >> 2) & 3) * (32768 >> 4)); b += (uint64_t)((i & 3) * 2) + (uint64_t)(((i >> 2) & 3) * (32768 >> 4)); }
            if (G == 1)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(i) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(i) : "memory");
            // mode 3: a commit after every 4th MMA, as a mainloop does per k-block (cta_group::2: multicast to both CTAs);
            // mode 4: the same commit without multicast (leader's barrier only)
            if (mode >= 3 && (i & 3) == 3) {
                const int mode = 3;
                if (G == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
                else if (mode == 3) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
                else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            }
        }
        long long t1 = clock64();
        if (G == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "h"((uint16_t)(mode == 3 ? 3 : 1)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (blockIdx.x < G) { out[0] = t1 - t0; out[1] = t2 - t0; }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar2)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (G == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        if (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
    }
}

template <int G>
void run(int N, int n_mma, int mode, int b_off, int nclusters = 1) {
    long long* out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G * nclusters); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe<G>, N, n_mma, mode, b_off, out);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sync failed: %s\n", cudaGetErrorString(e)); return; }
    }
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("clusters=%d cta_group::%d M=%3d N=%3d mode=%d n=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)\n", nclusters, G, 128 * G, N, mode, n_mma,
           (double)h[0] / n_mma, (double)h[1] / n_mma, N / 2.0);
    cudaFree(out);
}

int main() {
    const int n = 512;
    for (int mode : {3, 6})
        for (int N : {128, 256})
            for (int ncl : {1, 74}) {
                run<1>(N, n, mode, 0, ncl * 2);
                run<2>(N, n, mode, 0, ncl);
            }
    return 0;
}
