// Hot path 2, tensor-core variant (replaces evaluate.py:78 np.dot + :81 np.argsort + :96-105 walk):
//
//   convert   fp32 U / V rows -> BF16, K padded to a multiple of 64, bias folded in as three extra
//             BF16 columns (hi+mid+lo = the fp32 bias exactly) against 1.0 in U; row norms for the bound
//   filter    score_filter_kernel: TMA (cp.async.bulk.tensor, 128B swizzle) feeds a 4-5 stage smem ring,
//             one elected thread issues tcgen05.mma (M=128, N=256, K=16, BF16 -> FP32 in TMEM), two
//             256-column accumulators ping-pong so the MMA of tile t+1 overlaps the epilogue of tile t;
//             4 epilogue warps read TMEM with tcgen05.ld (thread == user row), keep a register threshold,
//             and append the (rare) survivors -- rated columns excluded -- to a per-row candidate buffer
//             that a warp-wide bitonic sort compacts to the best 64.  The score matrix never leaves the SM.
//   merge     item splits (few-user launches) are merged with topk_merge_kernel on the approximate keys
//   refine    exact fp32 fma-chain scores of the <= 64 candidates, exact (score desc, column desc) top-k,
//             and a certificate: the k-th exact score must beat (64th approximate score + eps_row), where
//             eps_row bounds |bf16 tensor-core score - exact score| for every column.  Certified rows are
//             bit-identical to the exact engine by construction;
//   fallback  the uncertified rows (ties/near-ties at the cut, adversarial norms) are re-done by the exact
//             engine (score_topk_kernel) through a device-side row list.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

namespace tkr {

int launch_exact_rows(const float* U, int64_t nu, const float* V, int64_t ni, int d, const float* bias,
                      const int64_t* rp, const int32_t* ri, int k, int64_t col_offset, const int32_t* row_map,
                      const int32_t* n_rows_dev, int32_t* oi, float* os, cudaStream_t st);

constexpr int FM = 128, FN = 256, FK = 64;        // MMA tile; FK bf16 = one 128-byte swizzle span
constexpr int A_CHUNK_BYTES = FM * FK * 2;        // 16 KB
constexpr int B_STAGE_BYTES = FN * FK * 2;        // 32 KB
constexpr int KPRIME = 64, CAP = 128;             // kept candidates / buffer capacity per row
constexpr int F_EPI_WARPS = 8;                    // two per TMEM lane quarter: each takes half the columns of a tile
constexpr int F_THREADS = 64 + 32 * F_EPI_WARPS;  // warps 0-7 epilogue, warp 8 TMA, warp 9 MMA + TMEM alloc
constexpr int W_TMA = F_EPI_WARPS, W_MMA = F_EPI_WARPS + 1;   // highest warp ids: the SMSP arbiter favours them over the busy epilogue warps
constexpr int MAX_KB = 4;                         // d_pad <= 256

// ------------------------------------------------------------------ PTX wrappers
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
    uint32_t ok;   // (no suspend-time hint: with one the compiler emits NANOSLEEP between polls and the MMA issue lags)
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
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when all prior MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows at 128 B pitch, 8-row groups 1024 B apart): the layout TMA
// writes for a {64 bf16, rows} box with CU_TENSOR_MAP_SWIZZLE_128B.  Descriptor fields per the tcgen05
// matrix-descriptor format: start address >> 4, LBO (unused for swizzled K-major) = 1, SBO = 1024 >> 4,
// version = 1 (bits 46-47), layout type 2 = SWIZZLE_128B (bits 61-63).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FN >> 3) << 17) | ((uint32_t)(FM >> 4) << 24);

// ------------------------------------------------------------------ convert
// One warp per row: fp32 -> bf16 (round to nearest even), zero padded to d_pad; optional bias split
// (V rows) or ones (U rows) in columns d..d+2; fp32 row norm.
__global__ void __launch_bounds__(256) convert_rows_kernel(const float* __restrict__ X, int64_t n, int d, int dpad,
                                                           const float* __restrict__ bias, int ones, int extra,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ norm_out,
                                                           unsigned int* __restrict__ norm_max, unsigned int* __restrict__ bias_max) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const float* x = X + row * d;
    __nv_bfloat16* o = out + row * dpad;
    float ss = 0.f;
    for (int c = lane; c < dpad; c += 32) {
        float v = 0.f;
        if (c < d) { v = x[c]; ss = fmaf(v, v, ss); }
        else if (extra && c < d + 3) {
            if (ones) v = 1.0f;
            else {
                const float bv = bias[row];
                const float hi = __bfloat162float(__float2bfloat16_rn(bv));
                const float r1 = bv - hi;
                const float mid = __bfloat162float(__float2bfloat16_rn(r1));
                v = c == d ? hi : c == d + 1 ? mid : (r1 - mid);
            }
        }
        o[c] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if (lane == 0) {
        const float nrm = sqrtf(ss) * 1.000001f;
        if (norm_out) norm_out[row] = nrm;
        if (norm_max) atomicMax(norm_max, __float_as_uint(nrm));                 // non-negative floats order as uints
        if (bias_max && bias) atomicMax(bias_max, __float_as_uint(fabsf(bias[row])));
    }
}

// ------------------------------------------------------------------ filter
struct FilterParams {
    int64_t nu, ni, col_offset;
    int kb, stages, tiles_per_split;
    const int64_t* rated_indptr;
    const int32_t* rated_idx;
    uint64_t* cand;          // [2 * n_splits][nu][CAP] scratch keys (list = split * 2 + column half)
    long long* dbg;          // optional [gridDim.x*gridDim.y][10][4] cycle counters (profiling aid), may be NULL
    int32_t* out_idx;        // [2 * n_splits][nu][KPRIME] approximate lists, (score desc, col desc)
    float* out_score;
};

__device__ __forceinline__ bool rated_has(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t c) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(idx + mid) < c) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(idx + lo) == c;
}

// Warp-wide bitonic sort (descending) of 128 keys, element e = r*32 + lane held in key[r].
__device__ __forceinline__ void warp_sort128_desc(uint64_t (&key)[4], int lane) {
#pragma unroll
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < 32) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int e = r * 32 + lane;
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, key[r], j);
                    const bool desc = (e & k) == 0, lower = (lane & j) == 0;
                    const uint64_t mx = key[r] > o ? key[r] : o, mn = key[r] > o ? o : key[r];
                    key[r] = (lower == desc) ? mx : mn;
                }
            } else {
                const int jr = j >> 5;   // partner register
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if ((r & jr) == 0) {
                        const int e = r * 32 + lane;
                        const bool desc = (e & k) == 0;
                        const uint64_t a = key[r], b2 = key[r | jr];
                        const uint64_t mx = a > b2 ? a : b2, mn = a > b2 ? b2 : a;
                        key[r] = desc ? mx : mn;
                        key[r | jr] = desc ? mn : mx;
                    }
                }
            }
        }
    }
}

// Sort row buffer `buf` (n <= CAP valid keys) and keep the best KPRIME in place; returns the KPRIME-th score
// (or -inf when fewer are valid) to every lane.
__device__ __noinline__ float warp_compact(uint64_t* buf, int n, int lane, uint64_t (&key)[4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = r * 32 + lane;
        key[r] = e < n ? __ldcg(buf + e) : 0ull;
    }
    warp_sort128_desc(key, lane);
    buf[lane] = key[0];
    buf[32 + lane] = key[1];
    const uint64_t last = __shfl_sync(0xffffffffu, key[1], 31);
    __syncwarp();
    return last ? ord_to_f32((uint32_t)(last >> 32)) : -INFINITY;
}

// Fast path: a max tree over the thread's 32 scores and one compare against its threshold; the warp votes and
// moves on when no lane has a hit (uniform branch, no divergence).
// Slow path (rare once the thresholds have risen): for each hit lane the warp transposes that lane's 32 scores
// with shuffles so that lane t holds column t, tests them against the row's threshold in parallel, drops rated
// columns (parallel binary searches) and appends the survivors to the row's buffer with ballot-derived slots.
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], int64_t c0, int64_t ni, int64_t col_offset,
                                           const int32_t* __restrict__ rated_idx, bool has_rated, int64_t r_lo, int64_t r_hi,
                                           uint64_t* buf, int& cnt, float tau, int lane) {
    float m[8];
#pragma unroll
    for (int g = 0; g < 8; ++g)
        m[g] = fmaxf(fmaxf(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])), fmaxf(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])));
    const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
    unsigned hits = __ballot_sync(0xffffffffu, mx >= tau);
    while (hits) {
        const int src = __ffs(hits) - 1;
        hits &= hits - 1;
        float x = 0.f;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const float y = __shfl_sync(0xffffffffu, __uint_as_float(v[t]), src);
            if (lane == t) x = y;
        }
        const float tau_s = __shfl_sync(0xffffffffu, tau, src);
        const int cnt_s = __shfl_sync(0xffffffffu, cnt, src);
        uint64_t* buf_s = (uint64_t*)__shfl_sync(0xffffffffu, (unsigned long long)buf, src);
        const int64_t col = c0 + lane;
        bool pass = x >= tau_s && col < ni;
        const int32_t gc = (int32_t)(col + col_offset);
        if (has_rated) {
            const int64_t lo = __shfl_sync(0xffffffffu, r_lo, src), hi = __shfl_sync(0xffffffffu, r_hi, src);
            if (pass) pass = !rated_has(rated_idx, lo, hi, gc);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (pass) buf_s[cnt_s + __popc(bal & ((1u << lane) - 1u))] = make_key(x + 0.0f, gc);
        if (lane == src) cnt += __popc(bal);
    }
}

// Rows whose buffer could overflow on the next chunk are compacted to their best KPRIME (warp-wide sort),
// which also raises their threshold.
__device__ __forceinline__ void compact_if_needed(uint64_t* buf, int& cnt, float& tau, int lane, uint64_t (&skey)[4], volatile float* tau_pub) {
    unsigned fullm = __ballot_sync(0xffffffffu, cnt > CAP - 32);
    while (fullm) {
        const int src = __ffs(fullm) - 1;
        fullm &= fullm - 1;
        const int n_src = __shfl_sync(0xffffffffu, cnt, src);
        uint64_t* b_src = (uint64_t*)__shfl_sync(0xffffffffu, (unsigned long long)buf, src);
        const float nt = warp_compact(b_src, n_src, lane, skey);
        __syncwarp();
        if (lane == src) { cnt = KPRIME; tau = fmaxf(tau, nt); *tau_pub = tau; }
    }
}

__global__ void __launch_bounds__(F_THREADS, 1) score_filter_kernel(const __grid_constant__ CUtensorMap tmU,
                                                                    const __grid_constant__ CUtensorMap tmV, FilterParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024 B alignment
    unsigned char* sA = base;
    unsigned char* sB = sA + (size_t)p.kb * A_CHUNK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * B_STAGE_BYTES);
    uint64_t* full = bars;                 // [stages]  TMA -> MMA
    uint64_t* empty = bars + 8;            // [stages]  MMA -> TMA
    uint64_t* tfull = bars + 16;           // [2]       MMA -> epilogue
    uint64_t* tempty = bars + 18;          // [2]       epilogue -> MMA
    uint64_t* afull = bars + 20;           //           U tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
    float* tau_sh = reinterpret_cast<float*>(bars + 22);   // [2][FM]: thresholds the two column halves publish to each other

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * FM;
    const int split = blockIdx.y;
    const int64_t ntiles = (p.ni + FN - 1) / FN;
    const int64_t t0 = (int64_t)split * p.tiles_per_split;
    const int64_t t1 = (t0 + p.tiles_per_split < ntiles) ? t0 + p.tiles_per_split : ntiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, F_EPI_WARPS); }
        mbar_init(afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < 2 * FM; t += F_THREADS) tau_sh[t] = -INFINITY;
    if (warp == W_MMA) {   // whole TMEM: two 256-column fp32 accumulators (1 CTA per SM, smem-limited)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(afull, (uint32_t)p.kb * A_CHUNK_BYTES);
            for (int c = 0; c < p.kb; ++c) tma_load_2d(sA + (size_t)c * A_CHUNK_BYTES, &tmU, afull, c * FK, (int)row0);
            uint32_t it = 0;
            long long dbg_wait0 = 0;
            const long long dbg_t0 = clock64();
            for (int64_t tile = t0; tile < t1; ++tile) {
                for (int c = 0; c < p.kb; ++c, ++it) {
                    const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
                    const long long w0 = clock64();
                    mbar_wait(empty + s, ph ^ 1);
                    dbg_wait0 += clock64() - w0;
                    mbar_expect_tx(full + s, B_STAGE_BYTES);
                    tma_load_2d(sB + (size_t)s * B_STAGE_BYTES, &tmV, full + s, c * FK, (int)(tile * FN));
                }
            }
            if (p.dbg) { long long* o = p.dbg + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 10 + warp) * 4; o[0] = clock64() - dbg_t0; o[1] = dbg_wait0; }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            mbar_wait(afull, 0);
            tc_fence_after();
            uint32_t it = 0;
            int tl = 0;
            long long dbg_w_e = 0, dbg_w_f = 0;
            const long long dbg_t0 = clock64();
            for (int64_t tile = t0; tile < t1; ++tile, ++tl) {
                const int acc = tl & 1;
                const long long w0 = clock64();
                mbar_wait(tempty + acc, ((tl >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                dbg_w_e += clock64() - w0;
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * FN;
                for (int c = 0; c < p.kb; ++c, ++it) {
                    const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
                    const long long w1 = clock64();
                    mbar_wait(full + s, ph);                          // V chunk landed
                    dbg_w_f += clock64() - w1;
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)c * A_CHUNK_BYTES), b0 = smem_u32(sB + (size_t)s * B_STAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < FK / 16; ++ks)              // K = 16 bf16 = 32 bytes per instruction
                        umma_bf16(tmem_d, make_sw128_desc(a0 + ks * 32), make_sw128_desc(b0 + ks * 32), kIdescBf16, (c | ks) != 0);
                    umma_commit(empty + s);                           // smem stage reusable once these MMAs retire
                }
                umma_commit(tfull + acc);                             // accumulator ready for the epilogue
            }
            if (p.dbg) { long long* o = p.dbg + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 10 + warp) * 4; o[0] = clock64() - dbg_t0; o[1] = dbg_w_e; o[2] = dbg_w_f; }
        }
    } else {
        // ===================== epilogue: thread == (user row, column half) =====================
        const int q = warp & 3;                                       // TMEM lane quarter this warp may read
        const int half = warp >> 2;                                   // columns [half*128, half*128 + 128) of every tile
        const int row = q * 32 + lane;
        const bool row_ok = row0 + row < p.nu;
        const int list = split * 2 + half;                            // each (split, half) produces its own candidate list
        uint64_t* buf = p.cand + ((size_t)list * p.nu + (size_t)(row_ok ? row0 + row : 0)) * CAP;
        const int64_t r_lo = (row_ok && p.rated_indptr) ? __ldg(p.rated_indptr + row0 + row) : 0;
        const int64_t r_hi = (row_ok && p.rated_indptr) ? __ldg(p.rated_indptr + row0 + row + 1) : 0;
        float tau = (row_ok && !(p.dbg && p.dbg[0] == -12345)) ? -INFINITY : INFINITY;   // padded rows never collect (dbg: no row collects)
        const int64_t ni = p.ni, col_offset = p.col_offset;
        const int32_t* rated_idx = p.rated_idx;
        const bool has_rated = p.rated_indptr != nullptr;
        volatile float* tau_mine = tau_sh + half * FM + row;
        volatile float* tau_other = tau_sh + (half ^ 1) * FM + row;
        int cnt = 0;
        uint64_t skey[4];
        constexpr int NCHUNK = FN / 2 / 32;                           // 4 chunks of 32 columns per thread per tile
        int tl = 0;
        long long dbg_w = 0, dbg_c = 0, dbg_s = 0, dbg_k = 0;
        const long long dbg_t0 = clock64();
        for (int64_t tile = t0; tile < t1; ++tile, ++tl) {
            const int acc = tl & 1;
            const long long w0 = clock64();
            mbar_wait(tfull + acc, (tl >> 1) & 1);
            dbg_w += clock64() - w0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * FN + (uint32_t)half * (FN / 2);
            const int64_t cbase = tile * FN + half * (FN / 2);
            // the 64-th best of EITHER half bounds the row's 64-th best from below: adopt the partner's threshold
            tau = fmaxf(tau, *tau_other);
            // two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is scanned
            uint32_t va[32], vb[32];
            tmem_ld32(taddr, va);
#pragma unroll 1
            for (int cc = 0; cc < NCHUNK; cc += 2) {
                long long c0_ = clock64();
                tmem_ld_wait();
                tmem_ld32(taddr + (cc + 1) * 32, vb);
                long long c1_ = clock64();
                scan_chunk(va, cbase + cc * 32, ni, col_offset, rated_idx, has_rated, r_lo, r_hi, buf, cnt, tau, lane);
                long long c2_ = clock64();
                compact_if_needed(buf, cnt, tau, lane, skey, tau_mine);
                long long c3_ = clock64();
                dbg_c += c1_ - c0_; dbg_s += c2_ - c1_; dbg_k += c3_ - c2_;
                c0_ = clock64();
                tmem_ld_wait();
                if (cc + 2 < NCHUNK) tmem_ld32(taddr + (cc + 2) * 32, va);
                c1_ = clock64();
                scan_chunk(vb, cbase + (cc + 1) * 32, ni, col_offset, rated_idx, has_rated, r_lo, r_hi, buf, cnt, tau, lane);
                c2_ = clock64();
                compact_if_needed(buf, cnt, tau, lane, skey, tau_mine);
                c3_ = clock64();
                dbg_c += c1_ - c0_; dbg_s += c2_ - c1_; dbg_k += c3_ - c2_;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
        }
        if (p.dbg && lane == 0) { long long* o = p.dbg + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 10 + warp) * 4; o[0] = dbg_s; o[1] = dbg_w; o[2] = dbg_c; o[3] = dbg_k; }
        // final: sorted best-KPRIME list of every (row, half) of this warp
        __syncwarp();
        for (int src = 0; src < 32; ++src) {
            const int n_src = __shfl_sync(0xffffffffu, cnt, src);
            uint64_t* b_src = (uint64_t*)__shfl_sync(0xffffffffu, (unsigned long long)buf, src);
            const int64_t grow = row0 + q * 32 + src;
            if (grow >= p.nu) continue;                               // warp-uniform
            warp_compact(b_src, n_src, lane, skey);
            const int64_t o = ((int64_t)list * p.nu + grow) * KPRIME;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint64_t key = skey[r];
                p.out_idx[o + r * 32 + lane] = key ? (int32_t)(uint32_t)key : -1;
                p.out_score[o + r * 32 + lane] = key ? ord_to_f32((uint32_t)(key >> 32)) : -INFINITY;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------ refine
// One warp per user row: exact fp32 fma-chain scores of the <= KPRIME candidates, exact top-k, certificate.
__global__ void __launch_bounds__(256) score_refine_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nu,
                                                           int d, const float* __restrict__ bias, int64_t col_offset,
                                                           const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_score,
                                                           const float* __restrict__ unorm, const unsigned int* __restrict__ vnorm_max,
                                                           const unsigned int* __restrict__ bias_max, float coef, int k,
                                                           int32_t* __restrict__ out_idx, float* __restrict__ out_score,
                                                           int32_t* __restrict__ fail_rows, int32_t* __restrict__ n_fail) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* us = reinterpret_cast<float*>(smem_raw) + (size_t)warp * d;
    uint64_t* keys = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)8 * d) + (size_t)warp * KPRIME;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= nu) return;
    for (int c = lane; c < d; c += 32) us[c] = U[row * d + c];
    __syncwarp();
    uint64_t mykey[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int32_t gc = cand_idx[row * KPRIME + h * 32 + lane];
        uint64_t key = 0;
        if (gc >= 0) {
            const int64_t lc = gc - col_offset;
            const float* v = V + lc * d;
            float acc = 0.f;
            for (int c = 0; c < d; ++c) acc = fmaf(us[c], __ldg(v + c), acc);   // ascending-index chain = the oracle's definition
            if (bias != nullptr) acc = acc + __ldg(bias + lc);
            key = make_key(acc + 0.0f, gc);
        }
        mykey[h] = key;
        keys[h * 32 + lane] = key;
    }
    __syncwarp();
    int rank[2] = {0, 0};
    for (int e = 0; e < KPRIME; ++e) {
        const uint64_t o = keys[e];
        rank[0] += o > mykey[0];
        rank[1] += o > mykey[1];
    }
    for (int p2 = lane; p2 < k; p2 += 32) { out_idx[row * k + p2] = -1; out_score[row * k + p2] = -INFINITY; }
    __syncwarp();
    float kth = -INFINITY;     // exact score at rank k-1 (if that many candidates exist)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (mykey[h] != 0 && rank[h] < k) {
            out_idx[row * k + rank[h]] = (int32_t)(uint32_t)mykey[h];
            out_score[row * k + rank[h]] = ord_to_f32((uint32_t)(mykey[h] >> 32));
        }
        if (mykey[h] != 0 && rank[h] == k - 1) kth = ord_to_f32((uint32_t)(mykey[h] >> 32));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, s));
    if (lane == 0) {
        // Everything the filter dropped has approximate score <= D = the KPRIME-th approximate score, hence exact
        // score <= D + eps.  Certified iff nothing was dropped, or the k-th exact score is strictly above that.
        const bool dropped = cand_idx[row * KPRIME + KPRIME - 1] >= 0;
        bool ok = !dropped;
        if (dropped) {
            const float D = cand_score[row * KPRIME + KPRIME - 1];
            const float eps = coef * unorm[row] * __uint_as_float(*vnorm_max) + 1e-6f * __uint_as_float(*bias_max) + 1e-30f;
            ok = kth > D + eps + fabsf(D) * 1e-6f;
        }
        if (!ok) fail_rows[atomicAdd(n_fail, 1)] = (int32_t)row;
    }
}

// ------------------------------------------------------------------ host side
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

// rows x dpad bf16, row-major; box = {64 (one swizzle span), box_rows}
static int make_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int dpad, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return TKR_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)FK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TKR_ERR_CUDA; }
    return TKR_OK;
}

long long* g_filter_dbg = nullptr;   // set through tkr_debug_set_filter_counters (profiling aid)

struct TcPlan {
    int dpad, kb, stages, ns, tps;
    size_t smem;
    size_t o_ubf, o_vbf, o_unorm, o_scal, o_cand, o_sidx, o_sscore, o_midx, o_mscore, o_fail, total;
};

static bool tc_plan(int64_t nu, int64_t ni, int d, int k, bool has_bias, TcPlan* P) {
    const int dext = d + (has_bias ? 3 : 0);
    P->dpad = (dext + FK - 1) / FK * FK;
    P->kb = P->dpad / FK;
    if (P->kb > MAX_KB || k > 48 || nu <= 0 || ni <= 0) return false;
    P->stages = P->kb >= 4 ? 4 : 5;
    P->smem = 1024 + (size_t)P->kb * A_CHUNK_BYTES + (size_t)P->stages * B_STAGE_BYTES + 256 + 2 * FM * 4;
    const int64_t row_tiles = (nu + FM - 1) / FM, ntiles = (ni + FN - 1) / FN;
    // one CTA per SM; split the items only when the user tiles alone cannot fill the chip
    int64_t ns = row_tiles >= kNumSMs ? 1 : (kNumSMs + row_tiles - 1) / row_tiles;
    const int64_t maxs = ntiles / 16 > 0 ? ntiles / 16 : 1;
    if (ns > maxs) ns = maxs;
    if (ns > 16) ns = 16;
    P->tps = (int)((ntiles + ns - 1) / ns);
    P->ns = (int)((ntiles + P->tps - 1) / P->tps);
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 1024); return r; };
    P->o_ubf = take((size_t)nu * P->dpad * 2);
    P->o_vbf = take((size_t)ni * P->dpad * 2);
    P->o_unorm = take((size_t)nu * 4);
    P->o_scal = take(64);                                   // [0] vnorm_max  [1] bias_max  [2] n_fail
    P->o_cand = take((size_t)2 * P->ns * nu * CAP * 8);            // one list per (item split, column half)
    P->o_sidx = take((size_t)2 * P->ns * nu * KPRIME * 4);
    P->o_sscore = take((size_t)2 * P->ns * nu * KPRIME * 4);
    P->o_midx = take((size_t)nu * KPRIME * 4);
    P->o_mscore = take((size_t)nu * KPRIME * 4);
    P->o_fail = take((size_t)nu * 4);
    P->total = o;
    return true;
}

}  // namespace tkr

using namespace tkr;

extern "C" void tkr_debug_set_filter_counters(long long* dev_buf) { g_filter_dbg = dev_buf; }

extern "C" size_t tkr_score_topk_tc_workspace_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k, int32_t has_bias) {
    TcPlan P;
    if (!tc_plan(nu, ni, d, k, has_bias != 0, &P)) return tkr_score_topk_workspace_bytes(nu, ni, d, k);
    return P.total + 1024;
}

extern "C" int tkr_score_topk_tc(const float* U, int64_t nu, const float* V, int64_t ni, int32_t d, const float* bias,
                                 const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset,
                                 int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int32_t* n_fallback_rows,
                                 void* stream) {
    TKR_CHECK_ARG(U && V && out_idx && out_score, "U, V and the outputs must not be NULL");
    TKR_CHECK_ARG(nu >= 0 && ni >= 1 && d >= 1 && k >= 1, "bad nu/ni/d/k");
    TKR_CHECK_ARG(rated_indptr == nullptr || rated_idx != nullptr, "rated_indptr without rated_idx");
    TKR_CHECK_ARG(ni + col_offset < ((int64_t)1 << 31), "global column index exceeds int32");
    if (nu == 0) return TKR_OK;
    TcPlan P;
    if (!tc_plan(nu, ni, d, k, bias != nullptr, &P))   // shape outside the tensor-core filter: the exact engine does it all
        return tkr_score_topk(U, nu, V, ni, d, bias, rated_indptr, rated_idx, k, col_offset, out_idx, out_score, ws, ws_bytes, stream);
    if (ws == nullptr || ws_bytes < P.total + 1024) { set_error("score_topk_tc workspace too small: have %zu, need %zu", ws_bytes, P.total + 1024); return TKR_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* Ubf = (__nv_bfloat16*)(w + P.o_ubf);
    __nv_bfloat16* Vbf = (__nv_bfloat16*)(w + P.o_vbf);
    float* unorm = (float*)(w + P.o_unorm);
    unsigned int* scal = (unsigned int*)(w + P.o_scal);
    uint64_t* cand = (uint64_t*)(w + P.o_cand);
    int32_t* sidx = (int32_t*)(w + P.o_sidx); float* sscore = (float*)(w + P.o_sscore);
    int32_t* midx = (int32_t*)(w + P.o_midx); float* mscore = (float*)(w + P.o_mscore);
    int32_t* fail = (int32_t*)(w + P.o_fail);

    TKR_CUDA(cudaMemsetAsync(scal, 0, 64, st));
    const int has_bias = bias != nullptr;
    convert_rows_kernel<<<(unsigned)((nu + 7) / 8), 256, 0, st>>>(U, nu, d, P.dpad, nullptr, 1, has_bias, Ubf, unorm, nullptr, nullptr);
    TKR_LAUNCH_CHECK();
    convert_rows_kernel<<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(V, ni, d, P.dpad, bias, 0, has_bias, Vbf, nullptr, scal + 0, scal + 1);
    TKR_LAUNCH_CHECK();

    CUtensorMap tmU, tmV;
    if (int rc = make_tmap(&tmU, Ubf, nu, P.dpad, FM)) return rc;
    if (int rc = make_tmap(&tmV, Vbf, ni, P.dpad, FN)) return rc;
    FilterParams fp;
    fp.nu = nu; fp.ni = ni; fp.col_offset = col_offset; fp.kb = P.kb; fp.stages = P.stages; fp.tiles_per_split = P.tps;
    fp.rated_indptr = rated_indptr; fp.rated_idx = rated_idx; fp.cand = cand;
    fp.out_idx = sidx; fp.out_score = sscore;
    fp.dbg = g_filter_dbg;
    TKR_CUDA(cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    dim3 grid((unsigned)((nu + FM - 1) / FM), (unsigned)P.ns);
    score_filter_kernel<<<grid, F_THREADS, P.smem, st>>>(tmU, tmV, fp);
    TKR_LAUNCH_CHECK();
    if (int rc = tkr_topk_merge(sidx, sscore, 2 * P.ns, nu, KPRIME, midx, mscore, stream)) return rc;

    // |bf16 tensor-core score - exact fma-chain score| <= coef * |u| * |v|: two roundings to 8-bit significands
    // (2^-8 + 2^-18 on every product, Cauchy-Schwarz over the row) + fp32 accumulation slack on both sides.
    const float coef = 0.00390625f * 1.01f + (float)(d + 8) * 9.5367431640625e-7f;
    const size_t rsmem = (size_t)8 * d * 4 + (size_t)8 * KPRIME * 8;
    score_refine_kernel<<<(unsigned)((nu + 7) / 8), 256, rsmem, st>>>(U, V, nu, d, bias, col_offset, midx, mscore, unorm, scal + 0, scal + 1,
                                                                     coef, k, out_idx, out_score, fail, (int32_t*)(scal + 2));
    TKR_LAUNCH_CHECK();
    // uncertified rows -> exact engine, driven by the device-side row list (no host round trip)
    if (int rc = launch_exact_rows(U, nu, V, ni, d, bias, rated_indptr, rated_idx, k, col_offset, fail, (const int32_t*)(scal + 2), out_idx, out_score, st)) return rc;
    if (n_fallback_rows != nullptr) TKR_CUDA(cudaMemcpyAsync(n_fallback_rows, scal + 2, 4, cudaMemcpyDeviceToDevice, st));
    return TKR_OK;
}
