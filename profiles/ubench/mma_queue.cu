// Microbenchmark: (1) depth of the tcgen05.mma issue queue: clock after each of 24 back-to-back MMA issues from an
// idle pipe; (2) cost of mbarrier.try_wait on an already-completed phase (+ tcgen05.fence::after_thread_sync).
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
__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ unsigned char dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(base + 96 * 1024);
    uint32_t* slot = (uint32_t*)(bar + 4);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 3; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + b)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        long long ts[40];
        const uint64_t da = desc_sw128(smem_u32(base)), db = desc_sw128(smem_u32(base + 32768));
        ts[0] = clock64();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (i & 1) * 256), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
            ts[i + 1] = clock64();
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        const long long c1 = clock64();
        while (!try_wait(bar, 0)) {}
        const long long c2 = clock64();
        for (int i = 0; i <= 32; ++i) out[i] = ts[i] - ts[0];
        out[40] = c1 - ts[0]; out[41] = c2 - ts[0];
        // (2) completed-phase try_wait cost: bar phase 0 is complete now
        const long long w0 = clock64();
        int acc = 0;
#pragma unroll 1
        for (int i = 0; i < 100; ++i) { acc += try_wait(bar, 0); }
        const long long w1 = clock64();
#pragma unroll 1
        for (int i = 0; i < 100; ++i) { acc += try_wait(bar, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        const long long w2 = clock64();
        long long s = 0;
#pragma unroll 1
        for (int i = 0; i < 100; ++i) { s += clock64(); }
        const long long w3 = clock64();
        out[42] = w1 - w0; out[43] = w2 - w1; out[44] = w3 - w2; out[45] = acc + (s & 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 128, 100 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("clock after issuing MMA i (M128 N256 K16, 128 cyc each), from an idle pipe:\n");
    for (int i = 1; i <= 32; ++i) printf(" %lld", h[i]);
    printf("\ncommit issued at %lld, barrier observed at %lld (32 MMAs = 4096 cyc of work)\n", h[40], h[41]);
    printf("try_wait on a completed phase: %.1f cyc; + tcgen05.fence::after_thread_sync: %.1f cyc; clock64 read: %.1f cyc\n", h[42] / 100.0, h[43] / 100.0, h[44] / 100.0);
    return 0;
}
