// Tensor-core GEMM for the VBPR content path (single/vbpr.py:56-57,61 forward; the dE / dc gradients of :73):
//
//     C[m, n] = sum_k A[m, k] * B[n, k]        A fp32 [M x K] row-major,  B fp32 [NP x K] row-major,  NP = 16 * ceil((h + 1) / 16)
//
// at fp32-level accuracy on tcgen05 kind::tf32 by the 3-term split  a*b ~= ah*bh + al*bh + ah*bl  (ah = a with the 13 low
// mantissa bits cleared -- exactly representable in TF32 --, al = TF32(a - ah); the dropped al*bl term is < 2^-21 |a b|).
// Both content GEMMs of a VBPR step have this shape with a skinny N:
//     projection   [F.E | F.c]        A = F   [n_items x d_feat],  B = [E | c]^T  [NP x d_feat]
//     gradient     [F^T.W | F^T.wq]   A = F^T [d_feat x n_items],  B = [W | wq]^T [NP x n_items]   (split-K, red.add)
// A is constant (the item features; F^T is built once), so it streams from HBM through TMA untouched and is split in
// shared memory by the CUDA-core warps; B is small, rebuilt per step, and arrives pre-split (Bhi, Blo).
//
// One CTA per SM, 320 threads:
//   warp 0      TMA producer: per K block of 32 floats one {32 x 128} box of A and one {32 x NP} box each of Bhi / Blo
//               (128-byte swizzle) into a ring of stages; out-of-range rows / columns are zero-filled by the TMA unit
//   warps 2-9   split the landed A tile: al = TF32(a - ah) goes to a second tile of the stage (element-wise on the swizzled
//               bytes: the layout does not matter; a itself stays, the tensor core reads its top 19 bits = ah),
//               fence.proxy.async, arrive
//   warp 1      one thread issues 3 x 4 tcgen05.mma (M=128, N=NP, K=8) per stage into TMEM; K blocks rotate over NACC
//               independent accumulators (summed in the epilogue with round-to-nearest adds) so that the tensor core's
//               own accumulation chain -- which truncates -- stays short; tcgen05.commit frees the stage
//   warps 2-5   epilogue: thread == output row, tcgen05.ld 16 columns at a time, store / red.add
// Bound: HBM (A is read once: 164 MB per GEMM at C3 = 25 us at 6.5 TB/s; the 3 x TF32 math is ~20 GFLOP = ~20 us of the
// tensor pipe, overlapped).
#include "common.cuh"
#include <cuda.h>

namespace tkr {
namespace g3 {

constexpr int BM = 128, BK = 32;                  // BK fp32 = one 128-byte swizzle span
constexpr int A_TILE_BYTES = BM * BK * 4;         // 16 KB
constexpr int SPLIT_WARPS = 8;                    // warps 2..9 split the landed A tile; warps 2..5 also run the epilogue
constexpr int THREADS = 64 + 32 * SPLIT_WARPS;
constexpr int MAX_STAGES = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when all prior MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows at 128 B pitch, 8-row groups 1024 B apart): what TMA writes for a
// {32 fp32, rows} box with CU_TENSOR_MAP_SWIZZLE_128B.  Same descriptor as the BF16 filter (score_tc.cu): start >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 >> 4, version 1, layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

enum { EPI_PROJECT = 0, EPI_GRAD = 1 };

struct Args {
    int M, K, NP, h;              // rows of A, reduction length, padded N, real content width (column h of C = the vector product)
    int stages, nacc;
    int wide;                     // 1: ah . [bh | bl] is ONE instruction of N = 2 NP (both B tiles are contiguous in the stage)
    int keep_raw;                 // 1 (default): the raw A tile IS ah -- kind::tf32 ignores the 13 low mantissa bits of an fp32 operand (measured:
                                  // bit-identical results with and without overwriting a by its TF32-exact part, profiles/r02f_probe_gemm3.json)
    int kb_per_split;             // K blocks per blockIdx.y
    // EPI_PROJECT: out[m * ld + off + n] = C[m][n] (n < h), vec_out[m] = vec_add[m] + C[m][h]
    // EPI_GRAD   : out[m * ld + n] += C[m][n] (red.add),     vec_out[m] += C[m][h]
    float* out; int ld, off;
    float* vec_out; const float* vec_add;
};

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, Args p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t b_tile = (uint32_t)p.NP * 128u;                     // NP rows x 128 B
    const uint32_t stage_bytes = 2u * A_TILE_BYTES + 2u * b_tile;      // [A (-> ah) | al | Bhi | Blo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                    // [stages] TMA bytes landed
    uint64_t* split = bars + MAX_STAGES;      // [stages] A tile split (4 warp arrivals)
    uint64_t* empty = bars + 2 * MAX_STAGES;  // [stages] MMAs of the stage retired
    uint64_t* done = bars + 3 * MAX_STAGES;   //          all MMAs retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int kb_total = (p.K + BK - 1) / BK;
    const int kb0 = blockIdx.y * p.kb_per_split;
    const int kb1 = min(kb_total, kb0 + p.kb_per_split);
    const int nkb = kb1 - kb0;                                          // >= 1 by construction of the grid
    const int acc_cols = p.wide ? 2 * p.NP : p.NP;                       // columns of one accumulator
    const uint32_t tmem_cols = p.nacc * acc_cols <= 128 ? 128u : (p.nacc * acc_cols <= 256 ? 256u : 512u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(split + s, SPLIT_WARPS); mbar_init(empty + s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(empty + s, ph ^ 1u);                          // (first pass: passes at once)
                unsigned char* st = base + (size_t)s * stage_bytes;
                mbar_expect_tx(full + s, (uint32_t)A_TILE_BYTES + 2u * b_tile);
                const int k = (kb0 + it) * BK;
                tma_load_2d(st, &tmA, full + s, k, m0);
                tma_load_2d(st + 2 * A_TILE_BYTES, &tmBhi, full + s, k, 0);
                tma_load_2d(st + 2 * A_TILE_BYTES + b_tile, &tmBlo, full + s, k, 0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(p.NP), idesc2 = idesc_tf32(2 * p.NP);
            const uint64_t desc_hi = make_sw128_desc(0);
            for (int it = 0; it < nkb; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(split + s, ph);
                tc_fence_after();
                const uint32_t st = smem_u32(base + (size_t)s * stage_bytes);
                const uint32_t a_hi = (st & 0x3FFFF) >> 4, a_lo = ((st + A_TILE_BYTES) & 0x3FFFF) >> 4;
                const uint32_t b_hi = ((st + 2 * A_TILE_BYTES) & 0x3FFFF) >> 4, b_lo = ((st + 2 * A_TILE_BYTES + b_tile) & 0x3FFFF) >> 4;
                const int acc = it % p.nacc;
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
                const uint32_t first = it < p.nacc ? 0u : 1u;          // first K block of an accumulator overwrites it
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {                  // K = 8 tf32 = 32 bytes per instruction
                    const uint64_t o = (uint64_t)(ks * 2);
                    if (p.wide) {
                        // columns [0, NP) <- ah.bh, [NP, 2NP) <- ah.bl in one instruction (A is read once for both), then al.bh
                        // on top of the first half: two instructions and 15.5 KB of operand reads per K step instead of three and 19.5
                        umma_tf32(tmem_d, desc_hi | (a_hi + o), desc_hi | (b_hi + o), idesc2, first | (uint32_t)(ks != 0));
                        umma_tf32(tmem_d, desc_hi | (a_lo + o), desc_hi | (b_hi + o), idesc, 1u);
                    } else {
                        umma_tf32(tmem_d, desc_hi | (a_lo + o), desc_hi | (b_hi + o), idesc, first | (uint32_t)(ks != 0));   // small terms first
                        umma_tf32(tmem_d, desc_hi | (a_hi + o), desc_hi | (b_lo + o), idesc, 1u);
                        umma_tf32(tmem_d, desc_hi | (a_hi + o), desc_hi | (b_hi + o), idesc, 1u);
                    }
                }
                umma_commit(empty + s);
            }
            umma_commit(done);
        }
    } else {
        // ===================== A splitter, then epilogue (warps 2-5) =====================
        const int t = threadIdx.x - 64;                                 // 0..32*SPLIT_WARPS-1
        for (int it = 0; it < nkb; ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(full + s, ph);
            float4* a = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
            float4* al = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + A_TILE_BYTES);
#pragma unroll
            for (int j = 0; j < A_TILE_BYTES / 16 / (32 * SPLIT_WARPS); ++j) {
                const int e = t + j * 32 * SPLIT_WARPS;
                const float4 v = a[e];
                float4 hi, lo;
                hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
                lo.x = tf32_hi(v.x - hi.x); lo.y = tf32_hi(v.y - hi.y); lo.z = tf32_hi(v.z - hi.z); lo.w = tf32_hi(v.w - hi.w);
                if (!p.keep_raw) a[e] = hi;
                al[e] = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(split + s);
        }
        if (warp < 6) {                                                 // epilogue: one warp per TMEM lane quarter
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3;                                         // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.NP; c0 += 16) {
            float acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            for (int a = 0; a < p.nacc && a < nkb; ++a) {
                uint32_t v[16];
                tmem_ld16(taddr0 + (uint32_t)(a * acc_cols + c0), v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(v[e]);
                if (p.wide) {
                    tmem_ld16(taddr0 + (uint32_t)(a * acc_cols + p.NP + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(v[e]);
                }
            }
            if (m < p.M) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int n = c0 + e;
                    if (n < p.h) {
                        if (EPI == EPI_PROJECT) p.out[(int64_t)m * p.ld + p.off + n] = acc[e];
                        else if (acc[e] != 0.f) atomicAdd(p.out + (int64_t)m * p.ld + n, acc[e]);
                    } else if (n == p.h) {
                        if (EPI == EPI_PROJECT) p.vec_out[m] = p.vec_add[m] + acc[e];
                        else if (acc[e] != 0.f) atomicAdd(p.vec_out + m, acc[e]);
                    }
                }
            }
        }
        tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// B operand builder: Bhi/Blo[n][k] (n < NP, k < K, row pitch Kp) from S[k][n] = src[k * ld + off + n] (n < h),
// vec[k] (n == h), 0 otherwise; pre-split so that the GEMM takes it as is.  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) build_b_kernel(const float* __restrict__ src, int ld, int off, const float* __restrict__ vec,
                                                      int K, int Kp, int h, int NP, float* __restrict__ Bhi, float* __restrict__ Blo) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;             // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + tx;
        float v = 0.f;
        if (k < K) {
            if (n < h) v = src[(int64_t)k * ld + off + n];
            else if (n == h) v = vec[k];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + tx;
        if (n < NP && k < Kp) {
            const float v = tile[tx][r];
            const float hi = tf32_hi(v);
            Bhi[(int64_t)n * Kp + k] = hi;
            Blo[(int64_t)n * Kp + k] = tf32_hi(v - hi);
        }
    }
}

// Ft[f][i] = F[i][f] (row pitch of Ft = Mp >= M, zero padded)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ F, int M, int Kd, int Mp, float* __restrict__ Ft) {
    __shared__ float tile[32][33];
    const int i0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, f = f0 + tx;
        tile[r][tx] = (i < M && f < Kd) ? F[(int64_t)i * Kd + f] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r, i = i0 + tx;
        if (f < Kd && i < Mp) Ft[(int64_t)f * Mp + i] = tile[tx][r];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// rows x cols fp32, row pitch `pitch` floats (multiple of 4); box = {32 (one swizzle span), box_rows}
static int make_tmap_f32(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available"); return TKR_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld fp32 matrix, pitch %lld", (int)r, (long long)rows, (long long)cols, (long long)pitch); return TKR_ERR_CUDA; }
    return TKR_OK;
}

}  // namespace g3

int g_gemm3_flags = 0;   // experiments (tkr_debug_set_gemm3_flags): 1 = overwrite the A tile with its TF32-exact part, 2 = three separate products

// ---- interface used by vbpr_step.cu -------------------------------------------------------------------------------
int gemm3_np(int h) { return (h + 1 + 15) / 16 * 16; }
bool gemm3_legal(int M, int K, int h, int64_t a_pitch, const void* A) {
    return gemm3_np(h) <= 256 && a_pitch % 4 == 0 && ((uintptr_t)A % 16) == 0 && M >= 1 && K >= 1;
}

int gemm3_build_b(const float* src, int ld, int off, const float* vec, int K, int Kp, int h, float* Bhi, float* Blo, cudaStream_t st) {
    const int NP = gemm3_np(h);
    dim3 grid((unsigned)((Kp + 31) / 32), (unsigned)((NP + 31) / 32));
    g3::build_b_kernel<<<grid, 256, 0, st>>>(src, ld, off, vec, K, Kp, h, NP, Bhi, Blo);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

int gemm3_transpose(const float* F, int M, int Kd, int Mp, float* Ft, cudaStream_t st) {
    dim3 grid((unsigned)((Mp + 31) / 32), (unsigned)((Kd + 31) / 32));
    g3::transpose_kernel<<<grid, 256, 0, st>>>(F, M, Kd, Mp, Ft);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

// C = A . B^T  (A [M x K], pitch a_pitch; Bhi / Blo [NP x K], pitch b_pitch), epilogue `epi` (0 project, 1 gradient)
int gemm3_run(int epi, const float* A, int M, int K, int64_t a_pitch, const float* Bhi, const float* Blo, int64_t b_pitch, int h,
              float* out, int ld, int off, float* vec_out, const float* vec_add, int splits, cudaStream_t st) {
    using namespace g3;
    Args p = {};
    p.M = M; p.K = K; p.NP = gemm3_np(h); p.h = h;
    const size_t stage_bytes = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)p.NP * 128;
    int stages = (int)((227 * 1024 - 2048) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) { set_error("gemm_tf32x3: N = %d does not leave room for two pipeline stages", p.NP); return TKR_ERR_UNSUPPORTED; }
    p.stages = stages;
    p.wide = (2 * p.NP <= 256 && !(g_gemm3_flags & 2)) ? 1 : 0;
    p.keep_raw = (g_gemm3_flags & 1) ? 0 : 1;
    const int acc_cols = p.wide ? 2 * p.NP : p.NP;
    p.nacc = 512 / acc_cols < 4 ? 512 / acc_cols : 4;
    const int kb_total = (K + BK - 1) / BK;
    if (splits < 1) splits = 1;
    if (splits > kb_total) splits = kb_total;
    p.kb_per_split = (kb_total + splits - 1) / splits;
    splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;       // no empty split
    p.out = out; p.ld = ld; p.off = off; p.vec_out = vec_out; p.vec_add = vec_add;
    CUtensorMap tmA, tmBhi, tmBlo;
    if (int rc = make_tmap_f32(&tmA, A, M, K, a_pitch, BM)) return rc;
    if (int rc = make_tmap_f32(&tmBhi, Bhi, p.NP, K, b_pitch, p.NP)) return rc;
    if (int rc = make_tmap_f32(&tmBlo, Blo, p.NP, K, b_pitch, p.NP)) return rc;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)splits);
    if (epi == EPI_PROJECT) {
        TKR_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<EPI_PROJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tf32x3_kernel<EPI_PROJECT><<<grid, THREADS, smem, st>>>(tmA, tmBhi, tmBlo, p);
    } else {
        TKR_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<EPI_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tf32x3_kernel<EPI_GRAD><<<grid, THREADS, smem, st>>>(tmA, tmBhi, tmBlo, p);
    }
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

}  // namespace tkr

extern "C" void tkr_debug_set_gemm3_flags(int32_t f) { tkr::g_gemm3_flags = f; }
