// Microbenchmark: issue rate of tcgen05.mma kind::f16 (BF16 -> FP32) with both operands in shared memory
// (128-byte swizzle, K-major), no TMA traffic, no epilogue: cycles per instruction by shape and cta_group.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t gtime() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// CG = cta_group (1 or 2); N = MMA N; KCH = K chunks (of 64) per tile; a tile = KCH*4 MMAs into one accumulator
template <int CG, int N, int KCH, int SYNC>
__global__ void __launch_bounds__(128, 1) k_mma(int tiles, long long* out) {
    extern __shared__ unsigned char dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = 128 * 64 * 2, B_ROWS = N / CG, B_BYTES = B_ROWS * 64 * 2, NB = 4;
    unsigned char* sA = base;                       // KCH chunks
    unsigned char* sB = sA + KCH * A_BYTES;         // NB stages
    uint64_t* bar = (uint64_t*)(sB + NB * B_BYTES);
    uint32_t* slot = (uint32_t*)(bar + 4);
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < (KCH * A_BYTES + NB * B_BYTES) / 4; i += blockDim.x) {
        uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        ((uint32_t*)base)[i] = (h & 0x3fff3fffu) | 0x3c003c00u;   // bf16 pairs around 1.0
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + 1)), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + 2)), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + 3)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    if (threadIdx.x == 0 && rank == 0) {
        const long long c0 = clock64();
        const uint64_t g0 = gtime();
        int st = 0;
        for (int t = 0; t < tiles; ++t) {
            const uint32_t d = tmem + (uint32_t)(t & 1) * 256;
            for (int c = 0; c < KCH; ++c) {
                const uint32_t a0 = smem_u32(sA + c * A_BYTES), b0 = smem_u32(sB + st * B_BYTES);
                st = (st + 1) % NB;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t acc = (c | ks) != 0;
                    if (CG == 1)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(desc_sw128(a0 + ks * 32)), "l"(desc_sw128(b0 + ks * 32)), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(desc_sw128(a0 + ks * 32)), "l"(desc_sw128(b0 + ks * 32)), "r"(idesc), "r"(acc) : "memory");
                }
            }
            if (SYNC == 3 || SYNC == 4) {   // fire-and-forget commits (nobody waits): steady-state cost of tcgen05.commit
                for (int r = 0; r < (SYNC == 3 ? 1 : 3); ++r) {
                    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 1 + r)) : "memory");
                    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar + 1 + r)), "h"((uint16_t)3) : "memory");
                }
            } else if (SYNC) {   // commit + wait after every tile: per-tile time - KCH*4*N/2 = commit-to-barrier latency
                if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
                else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
                if (SYNC == 1) while (!try_wait(bar + 1, t & 1)) {}
                else while (!try_wait(bar + 2, t & 1)) {}          // SYNC == 2: another warp relays bar+1 -> bar+2
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        const long long c1 = clock64();
        while (!try_wait(bar, 0)) {}
        const long long c2 = clock64();
        const uint64_t g1 = gtime();
        out[blockIdx.x * 4 + 0] = c1 - c0; out[blockIdx.x * 4 + 1] = c2 - c0; out[blockIdx.x * 4 + 2] = (long long)(g1 - g0);
    }
    if (SYNC == 2 && threadIdx.x == 32 && rank == 0) {
        for (int t = 0; t < tiles; ++t) {
            while (!try_wait(bar + 1, t & 1)) {}
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + 2)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    if (threadIdx.x < 32) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int CG, int N, int KCH, int SYNC = 0>
void run(long long* dbuf) {
    const int tiles = 4000;
    const size_t smem = 1024 + KCH * 16384 + 4 * (N / CG) * 128 + 128;
    cudaFuncSetAttribute(k_mma<CG, N, KCH, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {2, 148}) {
        cudaMemset(dbuf, 0, 148 * 4 * 8);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_mma<CG, N, KCH, SYNC>, tiles, dbuf);
            cudaEventRecord(e1);
            cudaError_t e2 = cudaDeviceSynchronize();
            if (e != cudaSuccess || e2 != cudaSuccess) { printf("cg%d N=%d: %s / %s\n", CG, N, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        static long long h[148 * 4];
        cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
        double issue = 0, total = 0, ns = 0; int n = 0;
        for (int i = 0; i < grid; i += CG) { issue += h[i * 4]; total += h[i * 4 + 1]; ns += h[i * 4 + 2]; ++n; }
        issue /= n; total /= n; ns /= n;
        const double nmma = (double)tiles * KCH * 4;
        const double flop = 2.0 * 128 * CG * N * 16 * nmma * (grid / CG);
        printf("%scta_group::%d M=%d N=%d K/tile=%d grid=%3d: %.1f cyc/MMA (issue %.1f), SM clock %.0f MHz, %.1f TFLOP/s (kernel %.3f ms), ideal %d cyc/MMA\n",
               SYNC == 1 ? "[commit+wait per tile] " : SYNC == 2 ? "[commit+relay+wait per tile] " : SYNC == 3 ? "[1 commit per tile, no wait] " : SYNC == 4 ? "[3 commits per tile, no wait] " : "", CG, 128 * CG, N, KCH * 64, grid, total / nmma, issue / nmma, total / ns * 1e3, flop / (ms * 1e-3) / 1e12, ms, N / 2);
    }
}

int main() {
    long long* dbuf; cudaMalloc(&dbuf, 148 * 4 * 8);
    run<1, 256, 2, 3>(dbuf);
    run<2, 256, 2, 3>(dbuf);
    run<1, 256, 2, 4>(dbuf);
    run<2, 256, 2, 4>(dbuf);
    run<1, 256, 2, 1>(dbuf);
    run<2, 256, 2, 1>(dbuf);
    run<1, 256, 2, 2>(dbuf);
    run<2, 256, 2, 2>(dbuf);
    run<1, 256, 2>(dbuf);
    run<2, 256, 2>(dbuf);
    run<1, 128, 2>(dbuf);
    run<2, 128, 2>(dbuf);
    run<1, 256, 4>(dbuf);
    run<2, 256, 4>(dbuf);
    return 0;
}
