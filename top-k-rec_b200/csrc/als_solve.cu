// Alternating least squares for WMF / CER (SURVEY.md 8(f) NEXT-1): one half-step
//     x_r = solve( base + (a-b) * sum_{p in pos(r)} y_p y_p^T + ridge*I ,  a * sum_p y_p + ridge * prior_r )
// for every row r of the solved side (reference: single/cer.py:36-63, single/wmf.py:67-96 -- a Python loop
// calling np.dot + np.linalg.solve per row).
//
// Shape of the kernel (fp32 FMA pipe; the k x k systems are far too small / too many for a library call):
//   * one thread block per row segment.  The d x d normal matrix lives in REGISTERS: the lower triangle of
//     64x64 blocks, two warps per block, an 8x8 register tile per thread (symmetry halves the FMA work).
//   * the positives' factor rows are gathered HBM/L2 -> shared memory with cp.async, 16 rows per stage,
//     3 stages in flight; each staged row is a rank-1 update read with conflict-free 128-bit shared loads.
//   * the matrix is then factored where it is, A = L D L^T, column by column: the owners of column j publish it
//     to shared memory (packed column-major, which also keeps the factor for the back substitution), one
//     barrier, every thread applies the rank-1 update to its register tile.  The right-hand side rides along as
//     an extra row, so forward substitution costs nothing; back substitution walks the packed columns.
//   * rows with more positives than one segment (popular items: up to every user) are split: the segments'
//     partial matrices go to a caller workspace in the threads' own register order and a second kernel sums
//     them in a fixed order (deterministic) and solves.  The shared Gram b*Yr^T Yr uses the same two kernels.
//   * the row's loss terms (cer.py:46,58-63) are formed from x, the right-hand side and the column sums:
//     x^T B x = x^T rhs - ridge |x|^2 for the solved x, so the matrix is not needed again.
#include "common.cuh"
#include <algorithm>

namespace tkr {
namespace als {

constexpr int KC = 16;      // rows per stage
constexpr int STAGES = 3;

template <int NB> struct Geo {
    static constexpr int DP = 64 * NB;                 // padded width
    static constexpr int NBLK = NB * (NB + 1) / 2;     // lower-triangular 64x64 blocks
    static constexpr int NT = NBLK * 64;               // threads per block
    static constexpr int PACKED = DP * (DP + 1) / 2;   // packed lower triangle (column-major)
    static constexpr int PART = NBLK * 4096 + DP;      // floats per partial slot
    static constexpr size_t SMEM = (size_t)(PACKED + STAGES * KC * DP + 2 * DP + 64) * sizeof(float);
};

struct RowArgs {
    const float* Y; float* X;
    const int32_t* idx;
    const float* base; const float* prior;
    const int32_t* seg_row; const int64_t* seg_off; const int32_t* seg_len; const int32_t* seg_slot;
    const int32_t* multi_row; const int32_t* multi_slot0; const int32_t* multi_nslots; const int64_t* multi_total;
    float* partial; double* loss_rows;
    float a, amb, ridge, lreg;
    int d, solve_empty, item_loss;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// thread -> tile geometry
struct Tile {
    int I, J;        // block coordinates (I >= J)
    int r0, c0;      // first row / column of the tile; register index q maps to x0 + (q & 3) + (q >> 2) * 32
    int tx;
};
__device__ __forceinline__ Tile tile_of(int tid) {
    Tile t;
    const int blk = tid >> 6, u = tid & 63;
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= blk) ++I;
    t.I = I; t.J = blk - I * (I + 1) / 2;
    t.tx = u & 7;
    t.r0 = I * 64 + (u >> 3) * 4;
    t.c0 = t.J * 64 + t.tx * 4;
    return t;
}
__device__ __forceinline__ int reg_pos(int x0, int q) { return x0 + (q & 3) + (q >> 2) * 32; }

// acc += sum over rows list[0..n) of y y^T (this thread's tile); ssum += column tid of the same rows.
template <int NB>
__device__ __forceinline__ void gram_accumulate(float (&acc)[8][8], float& ssum, const Tile& t, const float* __restrict__ Y,
                                                int d, const int32_t* __restrict__ list, int64_t n, float* stage) {
    using G = Geo<NB>;
    const int tid = threadIdx.x;
    const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const int64_t nchunks = (n + KC - 1) / KC;
    auto issue = [&](int64_t c) {
        if (c < nchunks) {
            const int rows = (int)min((int64_t)KC, n - c * KC);
            const uint32_t buf = stage_s + (uint32_t)((c % STAGES) * KC * G::DP) * 4u;
            if (vec) {
                const int q = d >> 2;
                for (int e = tid; e < rows * q; e += G::NT) {
                    const int r = e / q, c4 = e - r * q;
                    cp_async16(buf + (uint32_t)(r * G::DP + c4 * 4) * 4u, Y + (size_t)list[c * KC + r] * d + c4 * 4);
                }
            } else {
                for (int e = tid; e < rows * d; e += G::NT) {
                    const int r = e / d, cc = e - r * d;
                    cp_async4(buf + (uint32_t)(r * G::DP + cc) * 4u, Y + (size_t)list[c * KC + r] * d + cc);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (int64_t c = 0; c < nchunks; ++c) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(c + STAGES - 1);
        const float* buf = stage + (c % STAGES) * KC * G::DP;
        const int rows = (int)min((int64_t)KC, n - c * KC);
#pragma unroll 2
        for (int k = 0; k < rows; ++k) {
            const float* y = buf + k * G::DP;
            const float4 a0 = *reinterpret_cast<const float4*>(y + t.r0), a1 = *reinterpret_cast<const float4*>(y + t.r0 + 32);
            const float4 b0 = *reinterpret_cast<const float4*>(y + t.c0), b1 = *reinterpret_cast<const float4*>(y + t.c0 + 32);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (tid < G::DP) ssum += y[tid];
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

template <int NB>
__device__ __forceinline__ void store_partial(const float (&acc)[8][8], float ssum, float* slot) {
    using G = Geo<NB>;
    const int tid = threadIdx.x, blk = tid >> 6, u = tid & 63;
    float4* p = reinterpret_cast<float4*>(slot) + (size_t)blk * 1024 + u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        p[(i * 2) * 64] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        p[(i * 2 + 1) * 64] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    if (tid < G::DP) slot[G::NBLK * 4096 + tid] = ssum;
}
template <int NB>
__device__ __forceinline__ void add_partial(float (&acc)[8][8], float& ssum, const float* slot, int blk, int u, bool with_sum) {
    using G = Geo<NB>;
    const float4* p = reinterpret_cast<const float4*>(slot) + (size_t)blk * 1024 + u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 lo = p[(i * 2) * 64], hi = p[(i * 2 + 1) * 64];
        acc[i][0] += lo.x; acc[i][1] += lo.y; acc[i][2] += lo.z; acc[i][3] += lo.w;
        acc[i][4] += hi.x; acc[i][5] += hi.y; acc[i][6] += hi.z; acc[i][7] += hi.w;
    }
    if (with_sum && (int)threadIdx.x < G::DP) ssum += slot[G::NBLK * 4096 + threadIdx.x];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Loss terms of one row (cer.py:46,58-63 / wmf.py:77,91-96), threads tid < DP hold x, rhs0, s of their column.
template <int NB>
__device__ __forceinline__ void row_loss(const RowArgs& p, int row, int64_t n, bool solved, float x, float rhs0, float ssum,
                                         float pr, float* red) {
    using G = Geo<NB>;
    const int tid = threadIdx.x;
    if (tid < G::DP) {
        const float dx = p.prior ? x - pr : x;
        float v0 = warp_sum(x * x), v1 = warp_sum(x * rhs0), v2 = warp_sum(ssum * x), v3 = warp_sum(dx * dx);
        if ((tid & 31) == 0) { float* o = red + (tid >> 5) * 4; o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3; }
    }
    __syncthreads();
    if (tid == 0) {
        double xx = 0, xr = 0, sx = 0, dd = 0;
        for (int w = 0; w < G::DP / 32; ++w) { xx += red[w * 4]; xr += red[w * 4 + 1]; sx += red[w * 4 + 2]; dd += red[w * 4 + 3]; }
        double loss;
        if (!p.item_loss) loss = 0.5 * (double)p.lreg * xx;
        else {
            loss = 0.5 * (double)p.lreg * dd;
            if (n > 0 && solved) loss += 0.5 * (xr - (double)p.ridge * xx) + 0.5 * (double)n * (double)p.a - (double)p.a * sx;
        }
        p.loss_rows[row] = loss;
    }
}

// Build A from the accumulated Gram, factor it (L D L^T in registers), solve, write the row and its loss.
template <int NB>
__device__ __forceinline__ void finish_row(const RowArgs& p, int row, int64_t n, float (&acc)[8][8], float ssum, const Tile& t,
                                           float* M, float* zs, float* xs, float* red) {
    using G = Geo<NB>;
    const int tid = threadIdx.x, d = p.d;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = reg_pos(t.r0, i);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = reg_pos(t.c0, j);
            float v;
            if (r < d && c < d) v = fmaf(p.amb, acc[i][j], __ldg(p.base + (size_t)r * d + c)) + (r == c ? p.ridge : 0.f);
            else v = (r == c) ? 1.f : 0.f;
            acc[i][j] = v;
        }
    }
    float pr = 0.f, z = 0.f;
    if (tid < d) {
        if (p.prior) pr = p.prior[(size_t)row * d + tid];
        z = fmaf(p.a, ssum, p.ridge * pr);      // a * sum_p y_p + ridge * prior  (cer.py:43,55)
    }
    const float rhs0 = z;
    // Column j of the factor is published at M + j*DP - j(j-1)/2 (packed, column-major); `col - j` is the column's
    // virtual row 0, so a tile addresses its rows / columns with compile-time offsets.  Reads at rows < j land in
    // earlier columns (valid shared memory) and are discarded by the select.
    const int rmax = t.r0 + 35, cmax = t.c0 + 35;
    float* col = M;
    int j = 0;
    for (int Jb = 0; Jb < NB; ++Jb) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll 1
            for (int tq = 0; tq < 8; ++tq) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cj = g * 4 + q;
                    if (t.J == Jb && t.tx == tq && rmax >= j) {
                        float* pw = col - j + t.r0;
                        if (t.r0 >= j) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) pw[(i & 3) + (i >> 2) * 32] = acc[i][cj];
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (t.r0 + (i & 3) + (i >> 2) * 32 >= j) pw[(i & 3) + (i >> 2) * 32] = acc[i][cj];
                        }
                    }
                    if (tid == j) zs[j] = z;
                    __syncthreads();
                    const float inv = 1.0f / col[0];
                    if (rmax > j && cmax > j) {
                        const float* pr = col - j + t.r0;
                        const float* pc = col - j + t.c0;
                        float cr[8], cc[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            cr[i] = pr[(i & 3) + (i >> 2) * 32] * inv;
                            cc[i] = pc[(i & 3) + (i >> 2) * 32];
                        }
                        if (t.r0 <= j || t.c0 <= j) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (t.r0 + (i & 3) + (i >> 2) * 32 <= j) cr[i] = 0.f;
                                if (t.c0 + (i & 3) + (i >> 2) * 32 <= j) cc[i] = 0.f;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(-cr[i], cc[jj], acc[i][jj]);
                    }
                    if (tid > j && tid < G::DP) z = fmaf(-col[tid - j] * inv, zs[j], z);
                    col += G::DP - j;
                    ++j;
                }
            }
        }
    }
    // back substitution: x_r = (z_r - sum_{q > r} M[q][r] x_q) / p_r, walking r downwards; thread c keeps z_c - (partial sum)
    float x = 0.f;
    if (tid < G::DP) {
        const float* mycol = M + ((size_t)tid * G::DP - (size_t)tid * (tid - 1) / 2);
        const float pinv = 1.0f / mycol[0];
        for (int r = G::DP - 1; r >= 0; --r) {
            if (tid == r) { x = z * pinv; xs[r] = x; }
            named_barrier(1, G::DP);
            if (tid < r) z = fmaf(-mycol[r - tid], xs[r], z);
        }
        if (tid < d) p.X[(size_t)row * d + tid] = x;
    }
    __syncthreads();
    row_loss<NB>(p, row, n, true, x, rhs0, ssum, pr, red);
}

template <int NB>
__device__ __forceinline__ void loss_only_row(const RowArgs& p, int row, float* red) {
    const int tid = threadIdx.x;
    float x = 0.f, pr = 0.f;
    if (tid < p.d) {
        x = p.X[(size_t)row * p.d + tid];
        if (p.prior) pr = p.prior[(size_t)row * p.d + tid];
    }
    row_loss<NB>(p, row, 0, false, x, 0.f, 0.f, pr, red);
}

template <int NB> struct Smem {
    using G = Geo<NB>;
    float* M; float* stage; float* zs; float* xs; float* red;
    __device__ explicit Smem(float* base) {
        M = base; stage = M + G::PACKED; stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(stage) + 15) & ~(uintptr_t)15);
        zs = stage + STAGES * KC * G::DP; xs = zs + G::DP; red = xs + G::DP;
    }
};

// kernel 1: one block per segment
template <int NB>
__global__ void __launch_bounds__(Geo<NB>::NT, 1) als_segment_kernel(const RowArgs p) {
    using G = Geo<NB>;
    extern __shared__ __align__(16) float smem_raw[];
    Smem<NB> sm(smem_raw);
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int row = p.seg_row[seg], len = p.seg_len[seg], slot = p.seg_slot[seg];
    if (slot < 0 && len == 0 && !p.solve_empty) {
        if (p.loss_rows) loss_only_row<NB>(p, row, sm.red);
        return;
    }
    for (int e = tid; e < STAGES * KC * G::DP; e += G::NT) sm.stage[e] = 0.f;   // pad columns stay zero
    __syncthreads();
    const Tile t = tile_of(tid);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float ssum = 0.f;
    gram_accumulate<NB>(acc, ssum, t, p.Y, p.d, p.idx + p.seg_off[seg], len, sm.stage);
    if (slot >= 0) {
        store_partial<NB>(acc, ssum, p.partial + (size_t)slot * G::PART);
        return;
    }
    finish_row<NB>(p, row, len, acc, ssum, t, sm.M, sm.zs, sm.xs, sm.red);
}

// kernel 2: one block per split row: ordered sum of its partial slots, then the same finish
template <int NB>
__global__ void __launch_bounds__(Geo<NB>::NT, 1) als_multi_kernel(const RowArgs p) {
    using G = Geo<NB>;
    extern __shared__ __align__(16) float smem_raw[];
    Smem<NB> sm(smem_raw);
    const int m = blockIdx.x, tid = threadIdx.x;
    const int row = p.multi_row[m], slot0 = p.multi_slot0[m], ns = p.multi_nslots[m];
    const Tile t = tile_of(tid);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float ssum = 0.f;
    for (int s = 0; s < ns; ++s) add_partial<NB>(acc, ssum, p.partial + (size_t)(slot0 + s) * G::PART, tid >> 6, tid & 63, true);
    finish_row<NB>(p, row, p.multi_total[m], acc, ssum, t, sm.M, sm.zs, sm.xs, sm.red);
}

// shared Gram: out[d,d] = scale * sum of n_slots partial matrices + ridge * I  (one 64-thread block per 64x64 block)
template <int NB>
__global__ void __launch_bounds__(64) als_gram_reduce_kernel(const float* partial, int n_slots, int d, float scale, float ridge,
                                                            float* out) {
    using G = Geo<NB>;
    const int blk = blockIdx.x, u = threadIdx.x;
    const Tile t = tile_of(blk * 64 + u);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float dummy = 0.f;
    for (int s = 0; s < n_slots; ++s) add_partial<NB>(acc, dummy, partial + (size_t)s * G::PART, blk, u, false);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = reg_pos(t.r0, i);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = reg_pos(t.c0, j);
            if (r < d && c < d) {
                const float v = scale * acc[i][j] + (r == c ? ridge : 0.f);
                out[(size_t)r * d + c] = v;
                if (t.I != t.J) out[(size_t)c * d + r] = v;
            }
        }
    }
}

template <int NB> static int set_attr() {
    static thread_local int done_dev = -1;
    int dev = 0;
    TKR_CUDA(cudaGetDevice(&dev));
    if (done_dev != dev) {
        TKR_CUDA(cudaFuncSetAttribute(als_segment_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NB>::SMEM));
        TKR_CUDA(cudaFuncSetAttribute(als_multi_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NB>::SMEM));
        done_dev = dev;
    }
    return TKR_OK;
}

template <int NB> static int launch_rows(const RowArgs& p, int64_t n_segs, int64_t n_multi, cudaStream_t st) {
    using G = Geo<NB>;
    int rc = set_attr<NB>();
    if (rc) return rc;
    if (n_segs > 0) {
        als_segment_kernel<NB><<<(unsigned)n_segs, G::NT, G::SMEM, st>>>(p);
        TKR_LAUNCH_CHECK();
    }
    if (n_multi > 0) {
        als_multi_kernel<NB><<<(unsigned)n_multi, G::NT, G::SMEM, st>>>(p);
        TKR_LAUNCH_CHECK();
    }
    return TKR_OK;
}

template <int NB> static int launch_gram(const RowArgs& p, int64_t n_segs, int d, float scale, float ridge, float* out, cudaStream_t st) {
    using G = Geo<NB>;
    int rc = set_attr<NB>();
    if (rc) return rc;
    als_segment_kernel<NB><<<(unsigned)n_segs, G::NT, G::SMEM, st>>>(p);
    TKR_LAUNCH_CHECK();
    als_gram_reduce_kernel<NB><<<G::NBLK, 64, 0, st>>>(p.partial, (int)n_segs, d, scale, ridge, out);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

// segment table of the shared Gram: n_segs equal slices of the row list, every one a partial slot
__global__ void gram_plan_kernel(int64_t n, int n_segs, int64_t per, int32_t* seg_row, int64_t* seg_off, int32_t* seg_len, int32_t* seg_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const int64_t lo = min(n, (int64_t)s * per), hi = min(n, lo + per);
    seg_row[s] = 0; seg_off[s] = lo; seg_len[s] = (int32_t)(hi - lo); seg_slot[s] = s;
}

static inline int nb_of(int d) { return (d + 63) / 64; }
static inline size_t part_floats(int d) { const int nb = nb_of(d); return (size_t)nb * (nb + 1) / 2 * 4096 + 64 * nb; }
constexpr int kGramSegs = 2 * kNumSMs;

}  // namespace als
}  // namespace tkr

using namespace tkr;

extern "C" size_t tkr_als_partial_bytes(int32_t d, int64_t n_slots) {
    if (d < 1 || d > 256 || n_slots < 0) return 0;
    return align_up(als::part_floats(d) * sizeof(float) * (size_t)n_slots, 256);
}

extern "C" size_t tkr_als_gram_workspace_bytes(int32_t d) {
    if (d < 1 || d > 256) return 0;
    return tkr_als_partial_bytes(d, als::kGramSegs) + align_up((size_t)als::kGramSegs * 24, 256);
}

extern "C" int tkr_als_gram(const float* Y, int32_t d, const int32_t* rows, int64_t n_rows, float scale, float ridge,
                            float* out, void* ws, size_t ws_bytes, void* stream) {
    TKR_CHECK_ARG(Y && out && ws && (rows || n_rows == 0), "tkr_als_gram: null pointer");
    TKR_CHECK_ARG(d >= 1 && d <= 256, "tkr_als_gram: d=%d outside [1,256]", d);
    TKR_CHECK_ARG(n_rows >= 0, "tkr_als_gram: n_rows < 0");
    if (ws_bytes < tkr_als_gram_workspace_bytes(d) || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_error("tkr_als_gram: workspace of %zu bytes (256-aligned) needed, got %zu", tkr_als_gram_workspace_bytes(d), ws_bytes);
        return TKR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_segs = (int)std::min<int64_t>(als::kGramSegs, std::max<int64_t>(1, (n_rows + als::KC - 1) / als::KC));
    const int64_t per = ((n_rows + n_segs - 1) / n_segs + als::KC - 1) / als::KC * als::KC;
    char* w = (char*)ws;
    als::RowArgs p = {};
    p.partial = (float*)w; w += tkr_als_partial_bytes(d, als::kGramSegs);
    int64_t* seg_off = (int64_t*)w; w += (size_t)als::kGramSegs * 8;
    int32_t* seg_row = (int32_t*)w; w += (size_t)als::kGramSegs * 4;
    int32_t* seg_len = (int32_t*)w; w += (size_t)als::kGramSegs * 4;
    int32_t* seg_slot = (int32_t*)w;
    als::gram_plan_kernel<<<(n_segs + 127) / 128, 128, 0, st>>>(n_rows, n_segs, std::max<int64_t>(per, als::KC), seg_row, seg_off, seg_len, seg_slot);
    TKR_LAUNCH_CHECK();
    p.Y = Y; p.idx = rows; p.d = d;
    p.seg_row = seg_row; p.seg_off = seg_off; p.seg_len = seg_len; p.seg_slot = seg_slot;
    switch (als::nb_of(d)) {
        case 1: return als::launch_gram<1>(p, n_segs, d, scale, ridge, out, st);
        case 2: return als::launch_gram<2>(p, n_segs, d, scale, ridge, out, st);
        case 3: return als::launch_gram<3>(p, n_segs, d, scale, ridge, out, st);
        default: return als::launch_gram<4>(p, n_segs, d, scale, ridge, out, st);
    }
}

extern "C" int tkr_als_solve_rows(const tkr_als_cfg* cfg, const tkr_als_plan* plan, const float* Y, float* X, const int32_t* idx,
                                  const float* base, const float* prior, double* loss_rows, void* partial, size_t partial_bytes,
                                  void* stream) {
    TKR_CHECK_ARG(cfg && plan && Y && X && base, "tkr_als_solve_rows: null pointer");
    TKR_CHECK_ARG(cfg->d >= 1 && cfg->d <= 256, "tkr_als_solve_rows: d=%d outside [1,256]", cfg->d);
    TKR_CHECK_ARG(plan->n_segs >= 0 && plan->n_multi >= 0 && plan->n_slots >= 0, "tkr_als_solve_rows: negative plan counts");
    TKR_CHECK_ARG(plan->n_segs == 0 || (plan->seg_row && plan->seg_off && plan->seg_len && plan->seg_slot), "tkr_als_solve_rows: plan segment arrays missing");
    TKR_CHECK_ARG(plan->n_multi == 0 || (plan->multi_row && plan->multi_slot0 && plan->multi_nslots && plan->multi_total),
                  "tkr_als_solve_rows: plan split-row arrays missing");
    TKR_CHECK_ARG(idx || plan->n_segs == 0, "tkr_als_solve_rows: idx is null");
    if (plan->n_slots > 0 && (!partial || partial_bytes < tkr_als_partial_bytes(cfg->d, plan->n_slots) || (reinterpret_cast<uintptr_t>(partial) & 15))) {
        set_error("tkr_als_solve_rows: partial workspace of %zu bytes needed, got %zu", tkr_als_partial_bytes(cfg->d, plan->n_slots), partial_bytes);
        return TKR_ERR_WORKSPACE;
    }
    als::RowArgs p = {};
    p.Y = Y; p.X = X; p.idx = idx; p.base = base; p.prior = prior;
    p.seg_row = plan->seg_row; p.seg_off = plan->seg_off; p.seg_len = plan->seg_len; p.seg_slot = plan->seg_slot;
    p.multi_row = plan->multi_row; p.multi_slot0 = plan->multi_slot0; p.multi_nslots = plan->multi_nslots; p.multi_total = plan->multi_total;
    p.partial = (float*)partial; p.loss_rows = loss_rows;
    p.a = cfg->a; p.amb = cfg->a - cfg->b; p.ridge = cfg->ridge; p.lreg = cfg->lreg;
    p.d = cfg->d; p.solve_empty = cfg->solve_empty; p.item_loss = cfg->item_loss;
    cudaStream_t st = (cudaStream_t)stream;
    switch (als::nb_of(cfg->d)) {
        case 1: return als::launch_rows<1>(p, plan->n_segs, plan->n_multi, st);
        case 2: return als::launch_rows<2>(p, plan->n_segs, plan->n_multi, st);
        case 3: return als::launch_rows<3>(p, plan->n_segs, plan->n_multi, st);
        default: return als::launch_rows<4>(p, plan->n_segs, plan->n_multi, st);
    }
}
