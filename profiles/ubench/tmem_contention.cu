// Microbenchmark: do tcgen05.ld reads compete with a running tcgen05.mma for TMEM?  One thread issues MMAs
// (M128 N256 K16, BF16 -> FP32, accumulating into TMEM columns [0, 256)) back to back while NW warps stream
// tcgen05.ld.32x32b.x32 from columns [256, 512).  Reports MMA cycles/instruction and LDTM bytes/clock, alone and together.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tmem_ld_gen.h"
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// mode bit 0: run the MMA stream; bit 1: run the LDTM stream
__global__ void __launch_bounds__(544, 1) k(int mode, int n_mma, int n_ld, int accumulate, long long* out) {
    extern __shared__ unsigned char dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(base + 96 * 1024);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 16) {                       // MMA issuer
        if (lane == 0 && (mode & 1)) {
            const uint64_t da = desc_sw128(smem_u32(base)), db = desc_sw128(smem_u32(base + 32768));
            const long long c0 = clock64();
            for (int i = 0; i < n_mma; ++i)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
            while (!try_wait(bar, 0)) {}
            out[blockIdx.x * 32 + 16] = clock64() - c0;
        }
    } else if (mode & 2) {                  // LDTM streamers: warps 0-15, lane quarter = warp & 3, columns 256 + 64 * (warp >> 2)
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u + 64u * (uint32_t)(warp >> 2);
        uint32_t a[32], b[32], sink = 0;
        const long long c0 = clock64();
        for (int i = 0; i < n_ld; ++i) {
            ld_32x32b_x32(taddr, a);
            ld_32x32b_x32(taddr + 32, b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sink ^= a[0] ^ b[31];
        }
        const long long c1 = clock64();
        if (sink == 0x12345678u) out[4000] = sink;
        if (lane == 0) out[blockIdx.x * 32 + warp] = c1 - c0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 148 * 32 * 8 + 65536);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int n_mma = 8000, n_ld = 4000;
    for (int accumulate = 1; accumulate >= 0; --accumulate)
        for (int mode = 1; mode <= 3; ++mode) {
            cudaMemset(d, 0, 148 * 32 * 8);
            for (int rep = 0; rep < 2; ++rep) k<<<148, 544, 100 * 1024>>>(mode, n_mma, mode == 3 ? n_ld * 4 : n_ld, accumulate, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            static long long h[148 * 32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double mma = 0, ld = 0;
            for (int b = 0; b < 148; ++b) { mma += h[b * 32 + 16]; long long mx = 0; for (int w = 0; w < 16; ++w) mx = h[b * 32 + w] > mx ? h[b * 32 + w] : mx; ld += mx; }
            mma /= 148; ld /= 148;
            const int nl = mode == 3 ? n_ld * 4 : n_ld;
            printf("accumulate=%d %-18s", accumulate, mode == 1 ? "MMA alone" : mode == 2 ? "LDTM alone" : "MMA + LDTM");
            if (mode & 1) printf("  MMA %.1f cyc/instr (ideal 128)", mma / n_mma);
            if (mode & 2) printf("  LDTM %.0f B/clk/SM over its own run (16 warps x 2 x 4 KB per iteration, %.0f cyc/iteration)", 16.0 * nl * 2 * 4096 / ld, ld / nl);
            printf("\n");
        }
    return 0;
}
