// Microbenchmark: issue cost (cycles per warp instruction per SM sub-partition) of the candidate reduction
// instructions of the filter epilogue: FMNMX, FMNMX3, VIMNMX3 (DPX), IMNMX, FSETP(.OR), HMNMX2, F2FP pack.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, float seed, long long* out, float* sink) {
    float x[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) x[i] = seed * (float)(threadIdx.x + i);
    int acc_i = 0;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = sink[threadIdx.x + 32 * i + (blockIdx.x & 1)];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float& a = x[3 * r]; float& b = x[3 * r + 1]; float& c = x[3 * r + 2];
            if (OP == 0) { a = fmaxf(a, b); b = fmaxf(b, c); c = fmaxf(c, a); }                                   // 3 FMNMX
            if (OP == 1) { a = fmaxf(fmaxf(a, y[r]), y[(r + 1) & 7]); b = fmaxf(fmaxf(b, y[(r + 2) & 7]), y[(r + 3) & 7]); c = fmaxf(fmaxf(c, y[(r + 4) & 7]), y[(r + 5) & 7]); }     // 3 FMNMX3
            if (OP == 2) {                                                                                        // 3 VIMNMX3
                int ia = __float_as_int(a), ib = __float_as_int(b), ic = __float_as_int(c);
                ia = __vimax3_s32(ia, ib, ic); ib = __vimax3_s32(ib, ic, ia); ic = __vimax3_s32(ic, ia, ib);
                a = __int_as_float(ia); b = __int_as_float(ib); c = __int_as_float(ic);
            }
            if (OP == 3) {                                                                                        // 3 IMNMX
                int ia = __float_as_int(a), ib = __float_as_int(b), ic = __float_as_int(c);
                ia = max(ia, ib); ib = max(ib, ic); ic = max(ic, ia);
                a = __int_as_float(ia); b = __int_as_float(ib); c = __int_as_float(ic);
            }
            if (OP == 4) { acc_i += (a >= seed) | (b >= seed) | (c >= seed); a += 1.0f; }                         // 3 FSETP + glue
            if (OP == 5) {                                                                                        // 3 HMNMX2 (bf16x2)
                __nv_bfloat162 ha = *reinterpret_cast<__nv_bfloat162*>(&a), hb = *reinterpret_cast<__nv_bfloat162*>(&b), hc = *reinterpret_cast<__nv_bfloat162*>(&c);
                ha = __hmax2(ha, hb); hb = __hmax2(hb, hc); hc = __hmax2(hc, ha);
                a = *reinterpret_cast<float*>(&ha); b = *reinterpret_cast<float*>(&hb); c = *reinterpret_cast<float*>(&hc);
            }
            if (OP == 6) {                                                                                        // 3 F2FP packs
                __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(b, c), p2 = __floats2bfloat162_rn(c, a);
                a = *reinterpret_cast<float*>(&p0); b = *reinterpret_cast<float*>(&p1); c = *reinterpret_cast<float*>(&p2);
            }
            if (OP == 7) { a = fmaf(a, b, c); b = fmaf(b, c, a); c = fmaf(c, a, b); }                             // 3 FFMA (reference)
        }
    }
    const long long t1 = clock64();
    float s = (float)acc_i;
#pragma unroll
    for (int i = 0; i < 24; ++i) s += x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, long long* d, float* sink) {
    const int iters = 2000;
    k<OP><<<148, 512>>>(iters, 1.0f, d, sink);
    k<OP><<<148, 512>>>(iters, 1.0f, d, sink);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    // 16 warps = 4 per sub-partition; 24 ops per warp per iteration
    printf("%-28s %.2f cycles per warp instruction per SMSP (4 warps/SMSP, independent chains of 8)\n", name, m / (iters * 24.0 * 4.0));
}
int main() {
    long long* d; float* sink; cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    run<7>("FFMA", d, sink); run<0>("FMNMX (2-input)", d, sink); run<1>("FMNMX3", d, sink); run<2>("VIMNMX3 (__vimax3_s32)", d, sink);
    run<3>("IMNMX (2-input int max)", d, sink); run<4>("FSETP x3 + OR glue", d, sink); run<5>("HMNMX2.BF16", d, sink); run<6>("F2FP.BF16 pack", d, sink);
    return 0;
}
