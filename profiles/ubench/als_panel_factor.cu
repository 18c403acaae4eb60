// DRAFT for next round (run once on B200: correct -- 1.0e-5 of the fp64 solve -- and 188 us per matrix per SM, i.e. not yet
// faster than the product kernel's column loop; see ubench_als_panel_factor.txt): the d=256 factorisation planned in DESIGN.md section 7,
// K5 (a), as a standalone microbenchmark.  The algorithm and the shared-memory layout are the ones checked on the CPU by
// profiles/ldl_panel_model.py; this file adds the thread mapping:
//   * the 256 x 256 normal matrix lives in shared memory (packed lower triangle, column k's virtual row 0 at the 16-byte
//     aligned offset v(k)); one thread block of 640 threads per matrix (the product kernel's block at d=256, 96 registers);
//   * a 32-column panel is held by 256 threads as 8 x 4 register tiles (rows c0 + 8*rg .., columns c0 + 4*cg ..) and factored
//     in eight 4-column rounds: the thread with the 4x4 diagonal block factors it (A), the round's column owners eliminate
//     it from their rows and publish raw M / scaled L as one float4 per row plus the raw columns into the matrix (B), the
//     panel's remaining columns take the rank-4 update (C); the right-hand side rides along in registers (thread t = row t);
//   * then every trailing 8 x 8 tile (the product kernel's tile map) is loaded from shared memory, takes the panel's rank-32
//     update (4 LDS.128 + 8 FMUL per 64 FFMA per column) and is stored back;
//   * back substitution walks the stored raw columns.
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o als_panel_factor als_panel_factor.cu
// run:    ./als_panel_factor [n_matrices=1184]     (prints max relative error vs a host fp64 solve and us per matrix per SM)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int DP = 256, PW = 32, NT = 640;
constexpr int S_FLOATS = DP * DP / 2 + 2 * DP;          // 33280, see ldl_panel_model.py::check_layout

__host__ __device__ inline int col_origin(int k) {
    const int r = k & 3;
    return k * DP - k * (k + 1) / 2 + 6 * (k >> 2) + r * (r + 1) / 2;
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct Tile { int I, J, r0, c0; };
__device__ __forceinline__ Tile tile_of(int tid) {      // the product kernel's 8x8 tile map (als_solve.cu)
    Tile t;
    const int blk = tid >> 6, u = tid & 63;
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= blk) ++I;
    t.I = I; t.J = blk - I * (I + 1) / 2;
    t.r0 = I * 64 + (u >> 3) * 4;
    t.c0 = t.J * 64 + (u & 7) * 4;
    return t;
}

__global__ void __launch_bounds__(NT, 1) panel_factor_kernel(const float* __restrict__ A_all, const float* __restrict__ b_all,
                                                                         float* __restrict__ x_all) {
    extern __shared__ __align__(16) float sm[];
    float* S = sm;                                       // matrix
    float4* Lt = reinterpret_cast<float4*>(S + S_FLOATS);  // [DP] scaled L of the round's 4 columns, by row
    float4* Rt = Lt + DP;                                // [DP] raw M of the round's 4 columns, by row
    float* pinv = reinterpret_cast<float*>(Rt + DP);     // [DP]
    float* zs = pinv + DP;                               // [DP]
    float* xs = zs + DP;                                 // [DP]
    float* d44 = xs + DP;                                // [16]
    const int tid = threadIdx.x;
    const float* A = A_all + (size_t)blockIdx.x * DP * DP;
    // ---- load the lower triangle (in the product kernel: stored from the accumulator tiles)
    for (int e = tid; e < DP * DP; e += NT) {
        const int r = e / DP, c = e - r * DP;
        if (r >= c) S[col_origin(c) + r] = A[e];
    }
    float z = tid < DP ? b_all[(size_t)blockIdx.x * DP + tid] : 0.f;
    float pinv_mine = 1.f;
    __syncthreads();
    const Tile tt = tile_of(tid);
    const int rg = tid >> 3, cg = tid & 7;               // panel map (tid < 256): 8 rows x 4 columns
    for (int c0 = 0; c0 < DP; c0 += PW) {
        if (tid < 256) {
            const int r0 = c0 + rg * 8, pc0 = c0 + cg * 4;
            const bool live = r0 < DP;
            float a[8][4];
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float* col = S + col_origin(pc0 + q);
                    const float4 lo = *reinterpret_cast<const float4*>(col + r0), hi = *reinterpret_cast<const float4*>(col + r0 + 4);
                    a[0][q] = lo.x; a[1][q] = lo.y; a[2][q] = lo.z; a[3][q] = lo.w;
                    a[4][q] = hi.x; a[5][q] = hi.y; a[6][q] = hi.z; a[7][q] = hi.w;
                }
            }
#pragma unroll
            for (int cgq = 0; cgq < 8; ++cgq) {
                const int j0 = c0 + 4 * cgq;
                const int ri0 = 4 * (cgq & 1);                       // the diagonal thread's first row register
                // ---- A
                if (cg == cgq && rg == (cgq >> 1)) {
                    float a00 = a[ri0][0];
                    float a10 = a[ri0 + 1][0], a11 = a[ri0 + 1][1];
                    float a20 = a[ri0 + 2][0], a21 = a[ri0 + 2][1], a22 = a[ri0 + 2][2];
                    float a30 = a[ri0 + 3][0], a31 = a[ri0 + 3][1], a32 = a[ri0 + 3][2], a33 = a[ri0 + 3][3];
                    const float i0 = __frcp_rn(a00);
                    const float l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
                    a11 = fmaf(-l10, a10, a11); a21 = fmaf(-l20, a10, a21); a31 = fmaf(-l30, a10, a31);
                    a22 = fmaf(-l20, a20, a22); a32 = fmaf(-l30, a20, a32); a33 = fmaf(-l30, a30, a33);
                    const float i1 = __frcp_rn(a11);
                    const float l21 = a21 * i1, l31 = a31 * i1;
                    a22 = fmaf(-l21, a21, a22); a32 = fmaf(-l31, a21, a32); a33 = fmaf(-l31, a31, a33);
                    const float i2 = __frcp_rn(a22);
                    const float l32 = a32 * i2;
                    a33 = fmaf(-l32, a32, a33);
                    const float i3 = __frcp_rn(a33);
                    float4* o = reinterpret_cast<float4*>(d44);
                    o[0] = make_float4(i0, i1, i2, i3);
                    o[1] = make_float4(l10, l20, l30, l21);
                    o[2] = make_float4(l31, l32, 0.f, 0.f);
                    pinv[j0] = i0; pinv[j0 + 1] = i1; pinv[j0 + 2] = i2; pinv[j0 + 3] = i3;
                    // raw entries of the diagonal block (back substitution reads them): M[r][c] = updated a[r][c]
                    float* c0p = S + col_origin(j0);     S[col_origin(j0) + j0] = a00; c0p[j0 + 1] = a10; c0p[j0 + 2] = a20; c0p[j0 + 3] = a30;
                    float* c1p = S + col_origin(j0 + 1); c1p[j0 + 1] = a11; c1p[j0 + 2] = a21; c1p[j0 + 3] = a31;
                    float* c2p = S + col_origin(j0 + 2); c2p[j0 + 2] = a22; c2p[j0 + 3] = a32;
                    S[col_origin(j0 + 3) + j0 + 3] = a33;
                }
                if (tid >= j0 && tid < j0 + 4) zs[tid] = z;
                named_barrier(2, 256);
                // ---- B
                const float4 di = reinterpret_cast<const float4*>(d44)[0];
                const float4 la = reinterpret_cast<const float4*>(d44)[1];
                const float4 lb = reinterpret_cast<const float4*>(d44)[2];
                const float l10 = la.x, l20 = la.y, l30 = la.z, l21 = la.w, l31 = lb.x, l32 = lb.y;
                if (live && cg == cgq) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = r0 + i;
                        if (r > j0 + 3) {
                            const float m0 = a[i][0];
                            const float m1 = fmaf(-m0, l10, a[i][1]);
                            const float m2 = fmaf(-m1, l21, fmaf(-m0, l20, a[i][2]));
                            const float m3 = fmaf(-m2, l32, fmaf(-m1, l31, fmaf(-m0, l30, a[i][3])));
                            Rt[r] = make_float4(m0, m1, m2, m3);
                            Lt[r] = make_float4(m0 * di.x, m1 * di.y, m2 * di.z, m3 * di.w);
                            S[col_origin(j0) + r] = m0; S[col_origin(j0 + 1) + r] = m1;
                            S[col_origin(j0 + 2) + r] = m2; S[col_origin(j0 + 3) + r] = m3;
                        }
                    }
                }
                float zq0, zq1, zq2, zq3;
                {
                    const float4 zz = *reinterpret_cast<const float4*>(zs + j0);
                    zq0 = zz.x;
                    zq1 = fmaf(-l10, zq0, zz.y);
                    zq2 = fmaf(-l21, zq1, fmaf(-l20, zq0, zz.z));
                    zq3 = fmaf(-l32, zq2, fmaf(-l31, zq1, fmaf(-l30, zq0, zz.w)));
                    if (tid >= j0 && tid < j0 + 4) {
                        const int qq = tid - j0;
                        z = qq == 0 ? zq0 : qq == 1 ? zq1 : qq == 2 ? zq2 : zq3;
                        pinv_mine = qq == 0 ? di.x : qq == 1 ? di.y : qq == 2 ? di.z : di.w;
                    }
                }
                named_barrier(2, 256);
                // ---- C
                if (tid > j0 + 3) {
                    const float4 lz = Lt[tid];
                    z = fmaf(-lz.w, zq3, fmaf(-lz.z, zq2, fmaf(-lz.y, zq1, fmaf(-lz.x, zq0, z))));
                }
                if (live && cg > cgq && r0 + 7 > j0 + 3) {
                    float4 mc[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) mc[q] = Rt[pc0 + q];         // row (pc0+q) of the round's raw columns; pc0 + q > j0 + 3
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (r0 + i > j0 + 3) {
                            const float4 lr = Lt[r0 + i];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float v = a[i][q];
                                v = fmaf(-lr.x, mc[q].x, v); v = fmaf(-lr.y, mc[q].y, v);
                                v = fmaf(-lr.z, mc[q].z, v); v = fmaf(-lr.w, mc[q].w, v);
                                a[i][q] = v;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- trailing update with the finished panel: columns >= c0 + PW
        const int cfirst = c0 + PW;
        const bool half0 = tt.c0 >= cfirst, half1 = tt.c0 + 32 >= cfirst;
        if (half1 && tt.r0 + 35 >= cfirst) {
            float acc[8][8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int c = tt.c0 + (jj & 3) + (jj >> 2) * 32;
                const float* col = S + col_origin(c);
                const float4 lo = *reinterpret_cast<const float4*>(col + tt.r0), hi = *reinterpret_cast<const float4*>(col + tt.r0 + 32);
                acc[0][jj] = lo.x; acc[1][jj] = lo.y; acc[2][jj] = lo.z; acc[3][jj] = lo.w;
                acc[4][jj] = hi.x; acc[5][jj] = hi.y; acc[6][jj] = hi.z; acc[7][jj] = hi.w;
            }
            const float* vcol = S + col_origin(c0);
            int k = c0;
#pragma unroll 2
            for (int kk = 0; kk < PW; ++kk) {
                const float pk = pinv[k];
                const float4 ra = *reinterpret_cast<const float4*>(vcol + tt.r0), rb = *reinterpret_cast<const float4*>(vcol + tt.r0 + 32);
                const float4 ca = *reinterpret_cast<const float4*>(vcol + tt.c0), cb = *reinterpret_cast<const float4*>(vcol + tt.c0 + 32);
                const float cr[8] = {ra.x * pk, ra.y * pk, ra.z * pk, ra.w * pk, rb.x * pk, rb.y * pk, rb.z * pk, rb.w * pk};
                const float cc[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(-cr[i], cc[jj], acc[i][jj]);
                ++k;
                vcol += DP - k + (k & 3);
            }
            // store back the lower-triangular entries of the columns that are still open
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int c = tt.c0 + (jj & 3) + (jj >> 2) * 32;
                if ((jj >> 2) == 0 ? half0 : half1) {
                    float* col = S + col_origin(c);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rb0 = tt.r0 + 32 * h;
                        if (rb0 >= c) {
                            *reinterpret_cast<float4*>(col + rb0) = make_float4(acc[4 * h][jj], acc[4 * h + 1][jj], acc[4 * h + 2][jj], acc[4 * h + 3][jj]);
                        } else if (rb0 + 3 >= c) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (rb0 + i >= c) col[rb0 + i] = acc[4 * h + i][jj];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- back substitution: x_r = (z_r - sum_{q > r} M[q][r] x_q) / D_r
    if (tid < DP) {
        const float* mycol = S + col_origin(tid);
        float zz = z, x = 0.f;
        for (int r = DP - 1; r >= 0; --r) {
            if (tid == r) { x = zz * pinv_mine; xs[r] = x; }
            named_barrier(1, DP);
            if (tid < r) zz = fmaf(-mycol[r], xs[r], zz);
        }
        x_all[(size_t)blockIdx.x * DP + tid] = x;
    }
}

static void host_solve(const std::vector<float>& A, const std::vector<float>& b, std::vector<double>& x) {   // fp64 Cholesky-free LDL^T
    std::vector<double> M(A.begin(), A.end()), z(b.begin(), b.end());
    for (int j = 0; j < DP; ++j) {
        const double p = M[(size_t)j * DP + j];
        for (int r = j + 1; r < DP; ++r) {
            const double l = M[(size_t)r * DP + j] / p;
            for (int c = j + 1; c <= r; ++c) M[(size_t)r * DP + c] -= l * M[(size_t)c * DP + j];
            z[r] -= l * z[j];
        }
    }
    x.assign(DP, 0.0);
    for (int r = DP - 1; r >= 0; --r) {
        double t = z[r];
        for (int q = r + 1; q < DP; ++q) t -= M[(size_t)q * DP + r] * x[q];
        x[r] = t / M[(size_t)r * DP + r];
    }
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1184;
    std::vector<float> A((size_t)DP * DP), b(DP);
    srand(1);
    // one SPD matrix of the shape the ALS step produces: 0.01 * V^T V + 0.01 I + 0.99 * Vi^T Vi with uniform(0,1) factors
    std::vector<float> V((size_t)1200 * DP);
    for (auto& v : V) v = (float)rand() / RAND_MAX;
    for (int r = 0; r < DP; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = 0, si = 0;
            for (int i = 0; i < 1200; ++i) { const double p = (double)V[(size_t)i * DP + r] * V[(size_t)i * DP + c]; s += p; if (i < 208) si += p; }
            A[(size_t)r * DP + c] = A[(size_t)c * DP + r] = (float)(0.01 * s + 0.99 * si + (r == c ? 0.01 : 0.0));
        }
    for (int r = 0; r < DP; ++r) { double s = 0; for (int i = 0; i < 208; ++i) s += V[(size_t)i * DP + r]; b[r] = (float)s; }
    std::vector<double> xref;
    host_solve(A, b, xref);
    float *dA, *db, *dx;
    cudaMalloc(&dA, (size_t)n * DP * DP * 4); cudaMalloc(&db, (size_t)n * DP * 4); cudaMalloc(&dx, (size_t)n * DP * 4);
    for (int i = 0; i < n; ++i) {
        cudaMemcpy(dA + (size_t)i * DP * DP, A.data(), (size_t)DP * DP * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(db + (size_t)i * DP, b.data(), DP * 4, cudaMemcpyHostToDevice);
    }
    const size_t smem = (size_t)(S_FLOATS + 8 * DP + 3 * DP + 16) * 4;
    cudaFuncSetAttribute(panel_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    panel_factor_kernel<<<n, NT, smem>>>(dA, db, dx);
    cudaEventRecord(e0);
    panel_factor_kernel<<<n, NT, smem>>>(dA, db, dx);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> x((size_t)n * DP);
    cudaMemcpy(x.data(), dx, x.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int r = 0; r < DP; ++r) scale = fmax(scale, fabs(xref[r]));
    for (int i = 0; i < n; i += (n > 8 ? n / 8 : 1))
        for (int r = 0; r < DP; ++r) worst = fmax(worst, fabs(x[(size_t)i * DP + r] - xref[r]) / scale);
    printf("%d matrices of %d x %d: %.3f ms, %.1f us per matrix per SM (148 SMs); max |x - x64| / max |x64| = %.2e\n", n, DP, DP, ms,
           ms * 1e3 * 148 / n, worst);
    return worst < 1e-4 ? 0 : 2;
}
