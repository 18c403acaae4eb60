// Microbenchmark: tcgen05.ld throughput per SM (bytes / clock) by shape, width, warps and loads in flight.
// Decides whether the filter epilogue of K3 (a full drain of a 128 x 256 fp32 accumulator per MMA tile) is
// bounded by TMEM read bandwidth.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tmem_ld_gen.h"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CASE, int NREG, int DEPTH>
__global__ void __launch_bounds__((NREG * DEPTH > 96 ? 256 : 512), 1) k_ld(int iters, long long* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 1) * 256;
    uint32_t a[NREG], b[DEPTH > 1 ? NREG : 1];
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#define DO(name, n, str, lanes) if (CASE == __COUNTER__ - C0) { name(base, reinterpret_cast<uint32_t(&)[n]>(a)); if (DEPTH > 1) name(base, reinterpret_cast<uint32_t(&)[n]>(b)); }
        constexpr int C0 = __COUNTER__ + 1;
        TMEM_LD_CASES(DO)
#undef DO
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink ^= a[0] ^ a[NREG - 1] ^ b[0];
    }
    const long long t1 = clock64();
    if (sink == 0x12345678u) out[1000] = sink;
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int CASE, int NREG, int DEPTH>
void run(const char* name, long long* dbuf) {
    const int iters = 2000;
    for (int nw : {4, 8, 16}) {
        if (nw * 32 > (NREG * DEPTH > 96 ? 256 : 512)) continue;
        for (int grid : {1, 148}) {
            cudaMemset(dbuf, 0, 148 * 16 * 8);
            k_ld<CASE, NREG, DEPTH><<<grid, nw * 32>>>(iters, dbuf);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s nw=%d grid=%d depth=%d: %s\n", name, nw, grid, DEPTH, cudaGetErrorString(e)); return; }
            static long long h[148 * 16];
            cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid * 16; ++i) mx = h[i] > mx ? h[i] : mx;
            const double bytes = (double)nw * iters * DEPTH * NREG * 128.0;
            printf("%-14s warps=%2d grid=%3d inflight=%d  %8.1f B/clk/SM  (%.0f clk per ld per warp)\n", name, nw, grid, DEPTH, bytes / mx, (double)mx / iters / DEPTH);
        }
    }
}

int main() {
    long long* dbuf;
    cudaMalloc(&dbuf, 148 * 16 * 8 + 16384);
#define RUN(name, n, str, lanes) { constexpr int C = __COUNTER__ - R0; run<C, n, 1>(str, dbuf); if constexpr (n <= 64) run<C, n, 2>(str, dbuf); }
    constexpr int R0 = __COUNTER__ + 1;
    TMEM_LD_CASES(RUN)
    return 0;
}
