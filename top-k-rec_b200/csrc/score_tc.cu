// Hot path 2, tensor-core variant (replaces evaluate.py:78 np.dot + :81 np.argsort + :96-105 walk):
//
//   convert   fp32 U / V rows -> BF16, K padded to a multiple of 64, bias folded in as three extra
//             BF16 columns (hi+mid+lo = the fp32 bias exactly) against 1.0 in U; row norms for the bound
//   filter    score_filter_kernel, one CTA pair (cluster of 2) per 256 user rows, 1 CTA per SM:
//             * TMA (cp.async.bulk.tensor, 128B swizzle) feeds a 4-stage smem ring; each CTA stages its own 128 user
//               rows and half of every 256-item tile, and credits the leader CTA's mbarrier;
//             * one thread of the leader issues tcgen05.mma.cta_group::2 (M=256 over the two SMs, N=256, K=16,
//               BF16 -> FP32 in TMEM) into two 256-column accumulators that ping-pong; completion is multicast to both
//               CTAs with tcgen05.commit.  The issuing thread does ONE mbarrier wait per stage: the epilogues'
//               accumulator hand-back arrives on the same barrier as the TMA bytes (a try_wait costs ~160 cycles even
//               when the phase is complete, a tcgen05.mma issue ~55: profiles/ubench/);
//             * 16 epilogue warps (thread == user row x 64-column quarter) drain first and scan later: tcgen05.ld the
//               64 scores to registers, hand the accumulator back (~300 cycles after the tile completed), then run a
//               3-input max tree against the row threshold; a thread with a hit dumps its 32-score chunk to a
//               shared-memory block ring;
//             * 4 selection warps (one per TMEM lane quarter) consume the blocks four at a time, drop rated columns,
//               append 64-bit keys to the row's candidate buffer and compact it to the best 64 with a warp-wide
//               bitonic sort, raising the threshold;
//             * the sweep starts with seed tiles spread evenly over the items (no hand-offs; each thread tracks its 4
//               largest chunk maxima) so the threshold starts near the final one.
//             The score matrix never leaves the SM.
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
#include <stddef.h>

namespace tkr {

int launch_exact_rows(const float* U, int64_t nu, const float* V, int64_t ni, int d, const float* bias,
                      const int64_t* rp, const int32_t* ri, int k, int64_t col_offset, const int32_t* row_map,
                      const int32_t* n_rows_dev, int32_t* oi, float* os, void* ws, size_t ws_bytes, cudaStream_t st);
size_t exact_rows_workspace_bytes(int64_t ni, int k);

constexpr int FM = 128;                           // user rows per CTA = its 128 TMEM lanes; a CTA pair covers 256
constexpr int FN = 256, FK = 64;                  // item columns per MMA tile; FK bf16 = one 128-byte swizzle span
constexpr int A_CHUNK_BYTES = FM * FK * 2;        // 16 KB: this CTA's rows of one K chunk of U
constexpr int B_HALF = FN / 2;                    // V rows (tile columns) each CTA of the pair stages; the MMA reads both halves
constexpr int B_CHUNK_BYTES = B_HALF * FK * 2;    // 16 KB per CTA per K chunk
// smem ring depth is 4 stages (8 when a tile is a single K chunk); a stage holds cps K chunks
constexpr int MAX_CPS = 2;                        // K chunks per stage: 2 (d_pad >= 128) or 1
constexpr int KPRIME = 64, CAP = 128;             // kept candidates / buffer capacity per row
constexpr int F_EPI_WARPS = 16;                   // four per TMEM lane quarter: each takes 64 of a tile's 256 columns
constexpr int F_CQ = F_EPI_WARPS / 4;             // column quarters
constexpr int F_SEL_WARPS = 4;                    // selection warps: one per lane quarter, own the rows' candidate lists
constexpr int F_WARPS = F_SEL_WARPS + F_EPI_WARPS + 2;
// Warp roles by id: the SMSP arbiter favours high warp ids, so the latency-tolerant selection warps get the
// lowest ids, the TMEM-draining epilogue the middle ones, and the two single-thread issuers the highest.
constexpr int W_SEL = 0, W_EPI = F_SEL_WARPS, W_TMA = F_SEL_WARPS + F_EPI_WARPS, W_MMA = W_TMA + 1;
constexpr int F_THREADS = 32 * F_WARPS;            // warps 0-3 selection, 4-19 epilogue, 20 TMA, 21 MMA + TMEM alloc
constexpr int NBLK = 32;                          // hand-off blocks per selection warp
constexpr int COL_BITS = 26;                      // a sweep (item split) spans < 2^26 tile-space columns

// Hand-off from the epilogue (which must keep pace with the MMA) to the selection warps (which do the rare,
// latency-bound work).  An epilogue thread whose 32-score chunk contains a score that reaches its row's threshold
// dumps the chunk into a block (8 vector stores) and publishes a header word: 1 << 31 | row-in-quarter << 26 |
// sweep-relative column of the first score.  No fence on the producer side (same-thread shared-memory stores are
// performed in order); header 0 = empty block, the consumer clears it before it advances `tail`.  The selection
// warp then tests the block one score per lane -- the cost of finding the hits is off the epilogue's path.
struct SelShared {
    float data[NBLK][32];
    uint32_t hdr[NBLK];
    int head;                     // next reservation (atomic add by producers)
    volatile int tail;            // reservations consumed
    volatile int done;            // producer warps finished (F_CQ per selection warp)
    int pad;
};
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
// Waiters that are not on the MMA issue path let the hardware suspend them (up to `ns` per try) instead of
// spinning: their polls would otherwise take a fifth of the SM's issue slots.
__device__ __forceinline__ bool mbar_try_wait_suspend(uint32_t bar_addr, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar_addr, uint32_t parity) {   // shared-space address of the barrier
    uint32_t spins = 0;
    while (!mbar_try_wait_suspend(bar_addr, parity, 20000u)) {
        if (++spins > (1u << 20)) __trap();
    }
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// CTA-pair plumbing (cta_group::2): both CTAs of a cluster stage operands in their own shared memory, the
// leader (cluster rank 0) issues the MMAs for both, and completion events are multicast to both.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {   // same smem offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// (default .release.cta semantics: a cluster-scope release here costs a fence per arrival; the data the MMA
// depends on is ordered by tcgen05.fence + tcgen05.wait::ld, not by this arrive)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA into this CTA's shared memory, transaction bytes credited to the barrier at cluster address `bar_cluster`
// (the leader's): the leader's MMA thread waits once for both halves.
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on `bar` (same offset) in both CTAs of the pair when all prior MMAs of this thread are done.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
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
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.  cta_group::2: M = 256 (128 rows from each CTA), N = 256.
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FN >> 3) << 17) | ((uint32_t)((2 * FM) >> 4) << 24);

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
    int kb, cps, stages, tiles_per_split;   // K chunks of 64, chunks per smem stage, ring depth
    int seed_tiles;          // tiles of each sweep (evenly spread) scanned first in seed mode (0 = off)
    int seed_ext;            // 1: the seed tiles come from ANOTHER table (tensor map tmS: the whole item table, of which this sweep's own
    int seed_ext_stride;     //    table is the first shard), every seed_ext_stride-th tile of it
    int seed_rank;           // 4 or 3: the row's seed threshold is the smallest of the column quarters' seed_rank-th largest chunk maxima
    int cap_trigger;         // a row's candidate buffer is compacted to its best KPRIME once it would exceed this many keys (<= CAP)
    float* out_tau0;         // [n_splits][nu] seed threshold of each row (-inf when seeding is off)
    // A sweep may be one SEGMENT of a longer one that continues on another GPU (item-sharded ring, tkr_score_topk_tc_segment):
    // resume = the rows' thresholds / candidate counts come from st_tau / st_cnt (candidates already in `cand`), no seeding;
    // suspend = they are written back there at the end instead of sorting the lists out.
    int resume, suspend;
    float* st_tau;
    int32_t* st_cnt;
    const int64_t* rated_indptr;
    const int32_t* rated_idx;
    uint64_t* cand;          // [n_splits][nu][CAP] scratch keys
    long long* dbg;          // optional [gridDim.x*gridDim.y][14][4] cycle counters (profiling aid), may be NULL
    int32_t* out_idx;        // [n_splits][nu][KPRIME] approximate lists, (score desc, col desc)
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

template <bool DBG> __device__ __forceinline__ long long tick() { return DBG ? clock64() : 0ll; }
// MODE 0 = product; 1 = product + role cycle counters; 2 = counters, epilogue never reads TMEM (TMA + MMA ceiling);
// 3 = counters, epilogue drains TMEM but does not scan (TMEM read ceiling); 4 = 2 without V loads (barrier handshake
// only); 5 = full scan but thresholds pinned at +inf after seeding (no hand-offs).  2-5 produce no lists (profiling only).

// Epilogue fast path: a max tree over the thread's 32 scores (groups of 4) and one compare against its row's
// threshold.  On a hit, each group whose max passes is handed to the row's selection warp through its ring
// (all hit lanes of the warp do this in one short divergent pass).
__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
    // 3-input max tree (FMNMX3): 32 -> 11 -> 4 -> 2 -> 1
    float a[11];
#pragma unroll
    for (int t = 0; t < 10; ++t) a[t] = fmaxf(fmaxf(__uint_as_float(v[3 * t]), __uint_as_float(v[3 * t + 1])), __uint_as_float(v[3 * t + 2]));
    a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float b0 = fmaxf(fmaxf(a[0], a[1]), a[2]), b1 = fmaxf(fmaxf(a[3], a[4]), a[5]), b2 = fmaxf(fmaxf(a[6], a[7]), a[8]), b3 = fmaxf(a[9], a[10]);
    return fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
}

// Shared-space accessors for the hand-off rings: through a generic pointer the compiler emits generic
// ATOM.E.ADD.STRONG.GPU / LD / ST (hundreds of cycles per atomic); these are ATOMS / LDS / STS.
__device__ __forceinline__ int sh_atomic_inc(uint32_t addr) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}
__device__ __forceinline__ int sh_ld_volatile(uint32_t addr) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sh_st_volatile_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sh_st_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Publish one 32-score chunk to the row quarter's selection warp (see SelShared); ring = its shared-space address.
__device__ __forceinline__ void dump_chunk(const uint32_t (&v)[32], uint32_t hdr, uint32_t ring) {
    const int slot = sh_atomic_inc(ring + (uint32_t)offsetof(SelShared, head));
    while (slot - sh_ld_volatile(ring + (uint32_t)offsetof(SelShared, tail)) >= NBLK) __nanosleep(64);   // all blocks in use
    const uint32_t blk = (uint32_t)slot & (NBLK - 1);
    const uint32_t dst = ring + blk * 128u;
#pragma unroll
    for (int i = 0; i < 8; ++i) sh_st_v4(dst + 16u * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    sh_st_volatile_u32(ring + (uint32_t)offsetof(SelShared, hdr) + 4u * blk, hdr);
}

// Epilogue scan of one tile slice (64 scores of one row): two max trees and a compare each against the row threshold.
__device__ __forceinline__ void scan_tile(const uint32_t (&va)[32], const uint32_t (&vb)[32], uint32_t hdr, float tau, uint32_t ring) {
    const float ma = max32(va), mb = max32(vb);
    if (fmaxf(ma, mb) >= tau) {
        if (ma >= tau) dump_chunk(va, hdr, ring);
        if (mb >= tau) dump_chunk(vb, hdr + 32, ring);
    }
}

// Seed mode (first tiles of a sweep): nothing is handed off; the thread only tracks its 4 largest chunk maxima.
// s4 is then an actual score with at least 4 scores >= it among the columns the thread has seen: a cheap rigorous
// seed for the row threshold that skips the hit-heavy warm-up of a running top-k.
__device__ __forceinline__ void seed_chunk(const uint32_t (&v)[32], float& s1, float& s2, float& s3, float& s4) {
    float a = max32(v);
    float t;
    t = fmaxf(s1, a); a = fminf(s1, a); s1 = t;
    t = fmaxf(s2, a); a = fminf(s2, a); s2 = t;
    t = fmaxf(s3, a); a = fminf(s3, a); s3 = t;
    s4 = fmaxf(s4, a);
}

// Rated filter of one row's candidates (evaluate.py:98 `if liid not in rated[uid]`): a candidate whose global column
// is in the row's rated CSR slice [lo, hi) is erased (key 0 sorts last).  Done here, on whole buffers -- 4 keys per lane,
// their binary searches in lockstep so the dependent loads overlap -- instead of per candidate on the selection
// warps' critical path; the buffer may therefore hold rated columns between two compactions.
__device__ __forceinline__ void erase_rated(uint64_t (&key)[4], const int32_t* __restrict__ rated_idx, long long lo0, long long hi0) {
    const int n = (int)(hi0 - lo0);
    if (n <= 0) return;
    const int32_t* rl = rated_idx + lo0;
    int lo[4], hi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { lo[r] = 0; hi[r] = key[r] ? n : 0; }
    while ((lo[0] < hi[0]) | (lo[1] < hi[1]) | (lo[2] < hi[2]) | (lo[3] < hi[3])) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (lo[r] < hi[r]) {
                const int mid = (lo[r] + hi[r]) >> 1;
                if (__ldg(rl + mid) < (int32_t)(uint32_t)key[r]) lo[r] = mid + 1; else hi[r] = mid;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (key[r] && lo[r] < n && __ldg(rl + lo[r]) == (int32_t)(uint32_t)key[r]) key[r] = 0ull;
}

// Selection warp: sort the row buffer (n <= CAP keys, global memory) in registers, rated columns erased first, and keep
// the best KPRIME in place.  `key` holds the sorted keys of elements r*32+lane afterwards.
__device__ __forceinline__ void sel_sort(uint64_t* buf, int n, int lane, uint64_t (&key)[4], const int32_t* rated_idx, long long rlo, long long rhi) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = r * 32 + lane;
        key[r] = e < n ? __ldcg(buf + e) : 0ull;
    }
    if (rated_idx != nullptr) erase_rated(key, rated_idx, rlo, rhi);
    warp_sort128_desc(key, lane);
    buf[lane] = key[0];
    buf[32 + lane] = key[1];
    __syncwarp();
}
// Out of line (one copy of the 1100-instruction sort): returns the KPRIME-th score (or -inf) to every lane and the
// number of candidates kept (<= KPRIME) through *kept.
__device__ __noinline__ float sel_compact(uint64_t* buf, int n, int lane, const int32_t* rated_idx, long long rlo, long long rhi, int* kept) {
    uint64_t key[4];
    sel_sort(buf, n, lane, key, rated_idx, rlo, rhi);
    const unsigned live = __ballot_sync(0xffffffffu, key[0] != 0ull), live1 = __ballot_sync(0xffffffffu, key[1] != 0ull);
    *kept = __popc(live) + __popc(live1);
    const uint64_t last = __shfl_sync(0xffffffffu, key[1], 31);
    return last ? ord_to_f32((uint32_t)(last >> 32)) : -INFINITY;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmS, FilterParams p) {
    constexpr bool DBG = MODE != 0;
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B needs 1024 B alignment; both CTAs of the pair compute the same offsets (the MMA addresses the
    // peer's operands by the leader's offsets)
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;
    unsigned char* sB = sA + (size_t)p.kb * A_CHUNK_BYTES;
    const int cps = p.cps, spt = p.kb / p.cps;                          // K chunks per stage, stages per tile
    const uint32_t smask = (uint32_t)p.stages - 1u;                     // stages is a power of two and a multiple of spt
    const uint32_t sshift = 31u - (uint32_t)__clz(p.stages);
    const uint32_t stage_bytes = (uint32_t)cps * B_CHUNK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * stage_bytes);
    // full[s]: stage s landed in both CTAs (the leader's copy is the live one).  The first stage of a tile also
    // collects the 2 * F_EPI_WARPS arrivals with which the epilogues of both CTAs hand back the accumulator that
    // tile will overwrite, so the MMA thread does ONE barrier wait per stage and none per accumulator
    // (a try_wait costs ~160 cycles even on a completed phase; the issuing thread has ~1000 cycles per tile).
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + 8;            // [STAGES]  MMA -> TMA, multicast to both CTAs
    uint64_t* tfull = bars + 16;           // [2]       MMA -> epilogue, multicast to both CTAs
    uint64_t* afull = bars + 20;           //           U tiles of both CTAs landed (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
    volatile float* tau_sh = reinterpret_cast<float*>(bars + 22);        // [FM] row thresholds (selection warps write, epilogue reads)
    int* cnt_sh = reinterpret_cast<int*>(bars + 22) + FM;                // [FM] candidates buffered per row
    long long* rlo_sh = reinterpret_cast<long long*>(cnt_sh + FM);       // [FM] rated CSR range of each row
    long long* rhi_sh = rlo_sh + FM;
    float* seed_sh = reinterpret_cast<float*>(rhi_sh + FM);              // [F_CQ][FM] per-column-quarter seeds
    SelShared* sel = reinterpret_cast<SelShared*>(seed_sh + F_CQ * FM);  // [F_SEL_WARPS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                            // 0 = leader (issues the pair's MMAs)
    const int64_t row0 = (int64_t)blockIdx.x * FM;
    const int split = blockIdx.y;
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    const int64_t ntiles = (p.ni + FN - 1) / FN;
    const int64_t t0 = (int64_t)split * p.tiles_per_split;
    const int64_t t1 = (t0 + p.tiles_per_split < ntiles) ? t0 + p.tiles_per_split : ntiles;
    // Tile sequence of a sweep: T0 seed tiles spread evenly over [t0, t1) (so the seed is an unbiased sample whatever
    // the item order), scanned without hand-offs, then every tile of [t0, t1) in order.
    const int ntl = (int)(t1 - t0);
    const int T0 = (p.seed_tiles > 0 && (p.seed_ext || ntl >= 4 * p.seed_tiles)) ? p.seed_tiles : 0;
    const int nseq = ntl + T0;
    const int seed_stride = T0 > 0 ? ntl / T0 : 1;
    auto tile_rel = [&](int tl) -> int { return tl < T0 ? tl * seed_stride : tl - T0; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full + s, (s % spt) == 0 ? 1 + 2 * F_EPI_WARPS : 1);   // stages % spt == 0: stage s opens a tile iff s % spt == 0
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) mbar_init(tfull + a, 1);
        mbar_init(afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < FM; t += F_THREADS) {
        const bool ok = row0 + t < p.nu;   // padded rows never collect
        tau_sh[t] = ok ? (p.resume ? p.st_tau[row0 + t] : -INFINITY) : INFINITY;
        if (row0 + t < p.nu && T0 == 0 && !p.resume) p.out_tau0[(int64_t)split * p.nu + row0 + t] = -INFINITY;
        cnt_sh[t] = 0;
        rlo_sh[t] = (ok && p.rated_indptr) ? __ldg(p.rated_indptr + row0 + t) : 0;
        rhi_sh[t] = (ok && p.rated_indptr) ? __ldg(p.rated_indptr + row0 + t + 1) : 0;
    }
    for (int t = threadIdx.x; t < F_SEL_WARPS * (int)(sizeof(SelShared) / 4); t += F_THREADS) reinterpret_cast<int*>(sel)[t] = 0;
    if (warp == W_MMA) {   // whole TMEM of both SMs: two 256-column fp32 accumulators each (1 CTA per SM, smem-limited)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                    // barriers of both CTAs initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ===================== TMA producer (both CTAs: own U rows, own half of every V tile) =====================
        if (lane == 0) {
            const uint32_t afull_leader = mapa_shared(smem_u32(afull), 0);
            if (rank == 0) mbar_expect_tx(afull, 2u * (uint32_t)p.kb * A_CHUNK_BYTES);
            for (int c = 0; c < p.kb; ++c) tma_load_2d_pair(sA + (size_t)c * A_CHUNK_BYTES, &tmU, afull_leader, c * FK, (int)row0);
            uint32_t it = 0;
            const uint32_t empty_a = smem_u32(empty);
            long long dbg_wait0 = 0;
            const long long dbg_t0 = tick<DBG>();
            for (int tl = 0; tl < nseq; ++tl) {
                // (seed tiles of the first segment of a sharded sweep: a sample of the WHOLE table, so that the thresholds start where
                //  a whole sweep's would -- nothing is collected from seed tiles, their columns do not matter)
                const bool ext = p.seed_ext && tl < T0;
                const CUtensorMap* const tmB = ext ? &tmS : &tmV;
                const int vrow = (ext ? tl * p.seed_ext_stride * FN : (int)((t0 + tile_rel(tl)) * FN)) + (int)rank * B_HALF;
                for (int j = 0; j < spt; ++j, ++it) {
                    const uint32_t s = it & smask, ph = (it >> sshift) & 1;
                    const long long w0 = tick<DBG>();
                    mbar_wait_relaxed(empty_a + 8u * s, ph ^ 1);
                    dbg_wait0 += tick<DBG>() - w0;
                    if (MODE == 4) {                                   // probe: no V traffic, barrier handshake only
                        if (rank == 0) mbar_arrive(full + s);
                        continue;
                    }
                    if (rank == 0) mbar_expect_tx(full + s, 2u * stage_bytes);
                    const uint32_t fbar = mapa_shared(smem_u32(full + s), 0);
                    for (int c = 0; c < cps; ++c)
                        tma_load_2d_pair(sB + (size_t)s * stage_bytes + (size_t)c * B_CHUNK_BYTES, tmB, fbar, (j * cps + c) * FK, vrow);
                }
            }
            if (p.dbg) { long long* o = p.dbg + ((size_t)cta_lin * F_WARPS + warp) * 4; o[0] = tick<DBG>() - dbg_t0; o[1] = dbg_wait0; }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer (one thread of the leader CTA) =====================
        if (lane == 0 && rank == 0) {
            mbar_wait(afull, 0);
            tc_fence_after();
            uint32_t it = 0;
            long long dbg_w_f = 0;
            const long long dbg_t0 = tick<DBG>();
            unsigned long long dbg_g0 = 0, dbg_g1 = 0;
            if (DBG) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
            // descriptors differ only in the start-address field (bits 0-13, units of 16 B): hoist the constant part
            const uint64_t desc_hi = make_sw128_desc(0);
            const uint32_t a_lo = (smem_u32(sA) & 0x3FFFF) >> 4, b_lo = (smem_u32(sB) & 0x3FFFF) >> 4;
            for (int tl = 0; tl < nseq; ++tl) {
                const uint32_t tmem_d = tmem_base + (uint32_t)(tl & 1) * FN;
                for (int j = 0; j < spt; ++j, ++it) {
                    const uint32_t s = it & smask, ph = (it >> sshift) & 1;
                    const long long w1 = tick<DBG>();
                    mbar_wait(full + s, ph);          // V stage landed in both CTAs (+ accumulator handed back, if j == 0)
                    dbg_w_f += tick<DBG>() - w1;
                    tc_fence_after();
                    const uint32_t a0 = a_lo + (uint32_t)(j * cps) * (A_CHUNK_BYTES >> 4), b0 = b_lo + s * (stage_bytes >> 4);
#pragma unroll
                    for (int c = 0; c < MAX_CPS; ++c) {
                        if (c < cps) {
#pragma unroll
                            for (int ks = 0; ks < FK / 16; ++ks)      // K = 16 bf16 = 32 bytes per instruction
                                umma_bf16_pair(tmem_d, desc_hi | (uint64_t)(a0 + c * (A_CHUNK_BYTES >> 4) + ks * 2),
                                               desc_hi | (uint64_t)(b0 + c * (B_CHUNK_BYTES >> 4) + ks * 2), kIdescBf16, (j | c | ks) != 0);
                        }
                    }
                    umma_commit_pair(empty + s);                      // smem stage reusable (both CTAs) once these MMAs retire
                }
                umma_commit_pair(tfull + (tl & 1));                   // accumulator ready for both epilogues
            }
            if (DBG) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g1));
            if (p.dbg) { long long* o = p.dbg + ((size_t)cta_lin * F_WARPS + warp) * 4; o[0] = tick<DBG>() - dbg_t0; o[1] = 0; o[2] = dbg_w_f; o[3] = (long long)(dbg_g1 - dbg_g0); }
        }
    } else {
        if (warp >= W_EPI) {
        // ===================== epilogue: thread == (user row, column quarter) =====================
        // Drain first, scan later: the thread copies its 64 scores of the tile from TMEM to registers, hands the
        // accumulator back to the MMA thread at once (~300 cycles after the tile completed, so the two accumulators
        // keep the tensor pipe busy), and only then scans the registers -- throughput-bound work with a whole
        // tile time to finish, hand-offs included.
        const int q = warp & 3;                                       // TMEM lane quarter this warp may read
        const int cq = (warp - W_EPI) >> 2;                           // columns [cq*64, cq*64 + 64) of every tile
        const int row = q * 32 + lane;
        SelShared* hs = sel + q;
        const uint32_t ring = smem_u32(hs);
        const uint32_t tau_addr = smem_u32(const_cast<float*>(tau_sh) + row);
        // accumulator hand-back: after tile tl, arrive on the leader's full barrier of the first stage of tile tl + 2
        const uint32_t full_leader = mapa_shared(smem_u32(full), 0);
        if (lane == 0) {                                              // tiles 0 and 1 find their accumulators free
            mbar_arrive_cluster(full_leader);
            mbar_arrive_cluster(full_leader + 8u * ((uint32_t)spt & smask));
        }
        long long dbg_w = 0, dbg_c = 0, dbg_s = 0, dbg_l = 0;
        float s1 = -INFINITY, s2 = -INFINITY, s3 = -INFINITY, s4 = -INFINITY;
        // loop-carried addresses instead of per-tile recomputation (the kernel runs at its register cap)
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cq * 64;
        const uint32_t hdr0 = 0x80000000u | ((uint32_t)lane << COL_BITS) | (uint32_t)(cq * 64);
        uint32_t hb_stage = (uint32_t)(2 * spt) & smask;              // first stage of tile tl + 2
        uint32_t colrel = 0;                                          // sweep-relative first column of tile tl
        for (int tl = 0; tl < nseq; ++tl) {
            const uint32_t acc = (uint32_t)tl & 1u;
            if (T0 > 0 && tl == T0) {
                // end of seed mode: each thread holds its 4th largest chunk maximum; the row threshold is the smallest
                // of the four column quarters' (>= 16 scores of the seed tiles reach it)
                seed_sh[cq * FM + row] = p.seed_rank == 3 ? s3 : s4;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * F_EPI_WARPS) : "memory");
                if (cq == 0) {
                    const float tau0 = fminf(fminf(seed_sh[row], seed_sh[FM + row]), fminf(seed_sh[2 * FM + row], seed_sh[3 * FM + row]));
                    if (row0 + row < p.nu) {
                        tau_sh[row] = fmaxf(tau_sh[row], tau0);
                        p.out_tau0[(int64_t)split * p.nu + row0 + row] = tau0;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * F_EPI_WARPS) : "memory");
            }
            const long long w0 = tick<DBG>();
            // one warp polls the mbarrier, the other fifteen sleep on a named barrier (16 polling warps cost a third
            // of the SM's issue slots); ids 2/3 alternate with the accumulator so a fast warp cannot lap a slow one
            if (warp == W_EPI) mbar_wait(tfull + acc, ((uint32_t)tl >> 1) & 1u);
            asm volatile("bar.sync %0, %1;" ::"r"(2u + acc), "n"(32 * F_EPI_WARPS) : "memory");
            const long long w1 = tick<DBG>();
            dbg_w += w1 - w0;
            tc_fence_after();
            uint32_t va[32], vb[32];
            if (MODE != 2 && MODE != 4) {
                const uint32_t taddr = taddr0 + acc * FN;
                tmem_ld32(taddr, va);
                tmem_ld32(taddr + 32, vb);
                tmem_ld_wait();
            }
            dbg_l += tick<DBG>() - w1;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(full_leader + 8u * hb_stage);
            hb_stage = (hb_stage + (uint32_t)spt) & smask;
            const long long w2 = tick<DBG>();
            dbg_c += w2 - w1;
            if (MODE == 2 || MODE == 4) continue;
            if (MODE == 3) { s1 = fmaxf(s1, __uint_as_float(va[0] ^ va[31] ^ vb[0] ^ vb[31])); continue; }
            if (tl < T0) {
                seed_chunk(va, s1, s2, s3, s4);
                seed_chunk(vb, s1, s2, s3, s4);
            } else {
                const float tau = MODE == 5 ? INFINITY : __int_as_float(sh_ld_volatile(tau_addr));   // a stale value only costs extra hand-offs
                // block header: this thread's row and first column of the tile (sweep-relative column < 2^COL_BITS)
                scan_tile(va, vb, hdr0 + colrel, tau, ring);
                colrel += FN;
            }
            dbg_s += tick<DBG>() - w2;
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sh_atomic_inc(ring + (uint32_t)offsetof(SelShared, done)); }
        if (p.dbg && lane == 0) { long long* o = p.dbg + ((size_t)cta_lin * F_WARPS + warp) * 4; o[0] = dbg_s; o[1] = dbg_w; o[2] = dbg_c; o[3] = dbg_l; }
        } else {
        // ===================== selection: one warp per lane quarter owns the candidate lists of its 32 rows =====================
        // (explicit shared-space accesses: through generic volatile pointers every poll is an LD.E.STRONG.SYS)
        const int q = warp - W_SEL;
        const uint32_t ring = smem_u32(sel + q);
        const uint32_t hdr_a = ring + (uint32_t)offsetof(SelShared, hdr), head_a = ring + (uint32_t)offsetof(SelShared, head);
        const uint32_t tail_a = ring + (uint32_t)offsetof(SelShared, tail), done_a = ring + (uint32_t)offsetof(SelShared, done);
        const uint32_t tau_a = smem_u32(const_cast<float*>(tau_sh)) + 128u * (uint32_t)q;
        const int64_t sweep_col0 = t0 * FN;
        const uint32_t ni_rel = (uint32_t)(p.ni - sweep_col0 < ((int64_t)1 << COL_BITS) ? p.ni - sweep_col0 : ((int64_t)1 << COL_BITS));   // valid sweep-relative columns
        const int32_t gc_base = (int32_t)(sweep_col0 + p.col_offset);   // global column of sweep-relative column 0
        const int32_t* rated_idx = p.rated_idx;
        const bool has_rated = p.rated_indptr != nullptr;
        uint64_t* bufq = p.cand + ((size_t)split * p.nu + (size_t)row0 + (size_t)q * 32) * CAP;   // row r of the quarter at bufq + r*CAP
        uint64_t skey[4];
        int tail = 0;
        const bool my_row_ok = row0 + q * 32 + lane < p.nu;
        int cnt_reg = (p.resume && my_row_ok) ? p.st_cnt[row0 + q * 32 + lane] : 0;   // lane r: candidates buffered for row r of the quarter
        float tau_reg = my_row_ok ? (p.resume ? p.st_tau[row0 + q * 32 + lane] : -INFINITY) : INFINITY;   // lane r: its threshold (mirrored in tau_sh for the epilogue)
        bool tau_seeded = T0 == 0;
        long long dbg_idle = 0, dbg_busy = 0, dbg_n = 0, dbg_cmp = 0;
        for (;;) {
            // lane l looks at block tail + l; the published prefix is consumed block by block, one score per lane
            const long long i0 = tick<DBG>();
            uint32_t h;
            int n;
            bool finished = false;
            for (;;) {
                h = (uint32_t)sh_ld_volatile(hdr_a + 4u * (uint32_t)((tail + lane) & (NBLK - 1)));
                const unsigned bal = __ballot_sync(0xffffffffu, h != 0u);
                n = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;
                if (n > 0) break;
                int fin = 0;
                if (lane == 0) fin = (sh_ld_volatile(done_a) == F_CQ && sh_ld_volatile(head_a) == tail) ? 1 : 0;
                fin = __shfl_sync(0xffffffffu, fin, 0);
                if (fin) { finished = true; break; }
                __nanosleep(200);                                              // idle: do not steal issue slots from the epilogue
            }
            if (finished) break;
            const long long i1 = tick<DBG>();
            dbg_idle += i1 - i0;
            if (!tau_seeded) {      // blocks only appear after seed mode: pick up the seed thresholds the epilogue stored
                tau_reg = fmaxf(tau_reg, __int_as_float(sh_ld_volatile(tau_a + 4u * (uint32_t)lane)));
                tau_seeded = true;
            }
            // Blocks are taken four at a time with their loads and votes issued together: the per-block chain
            // (load, compare, vote, count, store) is latency-bound, and one selection warp -- which issues an
            // instruction every ~6 cycles at best -- has to keep up with four epilogue warps.  Row counts and
            // thresholds live in lane registers (lane r <-> row r of the quarter).
            for (int b0 = 0; b0 < n; b0 += 4) {
                uint32_t hb[4], bal[4];
                float x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hb[j] = __shfl_sync(0xffffffffu, h, (b0 + j) & 31);
                    x[j] = __int_as_float(sh_ld_volatile(ring + 128u * (uint32_t)((tail + b0 + j) & (NBLK - 1)) + 4u * (uint32_t)lane));
                }
                bool pass[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float tr = __shfl_sync(0xffffffffu, tau_reg, (int)(hb[j] >> COL_BITS));   // shfl takes the row index mod 32
                    const uint32_t crel = (hb[j] & ((1u << COL_BITS) - 1u)) + (uint32_t)lane;       // sweep-relative column
                    pass[j] = b0 + j < n && crel < ni_rel && x[j] >= tr;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) bal[j] = __ballot_sync(0xffffffffu, pass[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (bal[j] == 0u) continue;                                // (also skips the blocks past n)
                    const int rq = (int)((hb[j] >> COL_BITS) & 31u);           // row within the quarter
                    int cnt = __shfl_sync(0xffffffffu, cnt_reg, rq);
                    uint64_t* buf = bufq + rq * CAP;
                    if (cnt + __popc(bal[j]) > p.cap_trigger) {               // would overflow: keep the best KPRIME first
                        const long long k0 = tick<DBG>();
                        __syncwarp();
                        const int row = q * 32 + rq;
                        const float nt = sel_compact(buf, cnt, lane, has_rated ? rated_idx : nullptr, rlo_sh[row], rhi_sh[row], &cnt);
                        const float tr = fmaxf(__shfl_sync(0xffffffffu, tau_reg, rq), nt);
                        if (lane == rq) { tau_reg = tr; sh_st_volatile_u32(tau_a + 4u * (uint32_t)rq, __float_as_uint(tr)); }
                        bal[j] = __ballot_sync(0xffffffffu, ((bal[j] >> lane) & 1u) != 0u && x[j] >= tr);
                        dbg_cmp += tick<DBG>() - k0;
                    }
                    if ((bal[j] >> lane) & 1u) {
                        const uint32_t crel = (hb[j] & ((1u << COL_BITS) - 1u)) + (uint32_t)lane;
                        buf[cnt + __popc(bal[j] & ((1u << lane) - 1u))] = make_key(x[j] + 0.0f, gc_base + (int32_t)crel);
                    }
                    if (lane == rq) cnt_reg = cnt + __popc(bal[j]);
                }
            }
            if (lane < n) sh_st_volatile_u32(hdr_a + 4u * (uint32_t)((tail + lane) & (NBLK - 1)), 0u);   // blocks free again
            __syncwarp();                                                      // clears ordered before the tail store below
            tail += n;
            if (lane == 0) sh_st_volatile_u32(tail_a, (uint32_t)tail);
            dbg_busy += tick<DBG>() - i1; dbg_n += n;
        }
        if (p.suspend) {   // the sweep continues elsewhere: hand the rows' state back (their candidates are in `cand` already)
            if (my_row_ok) {
                if (!tau_seeded) tau_reg = fmaxf(tau_reg, __int_as_float(sh_ld_volatile(tau_a + 4u * (uint32_t)lane)));   // (no block ever arrived: seeds not picked up yet)
                p.st_cnt[row0 + q * 32 + lane] = cnt_reg;
                p.st_tau[row0 + q * 32 + lane] = tau_reg;
            }
        } else
        // final: sorted best-KPRIME list of every row of this quarter
        for (int r = 0; r < 32; ++r) {
            const int row = q * 32 + r;
            const int64_t grow = row0 + row;
            if (grow >= p.nu) break;
            const int cn = __shfl_sync(0xffffffffu, cnt_reg, r);
            sel_sort(bufq + (size_t)r * CAP, cn < CAP ? cn : CAP, lane, skey, has_rated ? rated_idx : nullptr, rlo_sh[row], rhi_sh[row]);
            const int64_t o = ((int64_t)split * p.nu + grow) * KPRIME;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const uint64_t key = skey[h2];
                p.out_idx[o + h2 * 32 + lane] = key ? (int32_t)(uint32_t)key : -1;
                p.out_score[o + h2 * 32 + lane] = key ? ord_to_f32((uint32_t)(key >> 32)) : -INFINITY;
            }
        }
        if (p.dbg && lane == 0) { long long* o = p.dbg + ((size_t)cta_lin * F_WARPS + warp) * 4; o[0] = dbg_busy; o[1] = dbg_idle; o[2] = dbg_n; o[3] = dbg_cmp; }
        }
    }
    // Neither CTA may exit (or free TMEM) while its peer can still signal its barriers or read its operands.
    tc_fence_before();
    cluster_sync_all();
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------ refine
// One warp per user row: exact fp32 fma-chain scores of the <= KPRIME candidates, exact top-k, certificate.
// The candidate item rows are gathered with coalesced 128-bit loads (one row per warp instruction, RR of them in
// flight) into a padded shared-memory tile; each lane then runs the ascending-index fma chain of ITS candidate out of
// shared memory (the chain order is the oracle's definition of a score, so the reduction cannot be split over lanes).
template <int RR>   // candidates staged per round (16, or 8 for wide rows): 8 warps x RR rows x 3 CTAs per SM must fit
__global__ void __launch_bounds__(256) score_refine_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nu,
                                                           int d, const float* __restrict__ bias, int64_t col_offset,
                                                           const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_score,
                                                           const float* __restrict__ unorm, const unsigned int* __restrict__ vnorm_max,
                                                           const unsigned int* __restrict__ bias_max, const float* __restrict__ tau0,
                                                           int n_splits, float coef, int k, int vec,
                                                           int32_t* __restrict__ out_idx, float* __restrict__ out_score,
                                                           int32_t* __restrict__ fail_rows, int32_t* __restrict__ n_fail) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = d + 4;                                              // floats; keeps 16-byte alignment, conflict-free LDS.128
    float* us = reinterpret_cast<float*>(smem_raw) + (size_t)warp * pitch;
    uint64_t* keys = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)8 * pitch) + (size_t)warp * KPRIME;
    float* vs = reinterpret_cast<float*>(reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)8 * pitch) + (size_t)8 * KPRIME)
                + (size_t)warp * RR * pitch;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= nu) return;
    for (int c = lane; c < d; c += 32) us[c] = U[row * d + c];
    // candidate e = h * 32 + lane is scored by this lane
    uint64_t mykey[2];
    int32_t gc[2];
    float acc[2] = {0.f, 0.f};
    gc[0] = cand_idx[row * KPRIME + lane]; gc[1] = cand_idx[row * KPRIME + 32 + lane];
    if (vec) {
        constexpr int ROUNDS = KPRIME / RR;
        const int nv = d >> 2;                                            // float4 per row
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            __syncwarp();
            // stage candidates [r*RR, r*RR + RR): lanes stride over the float4 of one row at a time
            const int32_t gsel = ((r * RR) >> 5) ? gc[1] : gc[0];          // (a round never straddles the two halves)
            for (int c0 = 0; c0 < nv; c0 += 32) {                          // 32 float4 (512 B) of every staged row per pass
                const int c = c0 + lane;
                float4 buf[RR];                                            // all RR row loads in flight before the first store
#pragma unroll
                for (int e = 0; e < RR; ++e) {
                    const int32_t g = __shfl_sync(0xffffffffu, gsel, (r * RR + e) & 31);
                    const float4* src = reinterpret_cast<const float4*>(V + (int64_t)(g >= 0 ? g - col_offset : 0) * d);
                    if (c < nv) buf[e] = __ldg(src + c);
                }
#pragma unroll
                for (int e = 0; e < RR; ++e)
                    if (c < nv) reinterpret_cast<float4*>(vs + (size_t)e * pitch)[c] = buf[e];
            }
            __syncwarp();
            // the RR lanes that own this round's candidates run their chains
            const int h = (r * RR) >> 5, l0 = (r * RR) & 31;
            if (lane >= l0 && lane < l0 + RR) {
                const float4* mine = reinterpret_cast<const float4*>(vs + (size_t)(lane - l0) * pitch);
                float a = 0.f;
                for (int c = 0; c < nv; ++c) {
                    const float4 v4 = mine[c];
                    const float4 u4 = *reinterpret_cast<const float4*>(us + 4 * c);
                    a = fmaf(u4.x, v4.x, a); a = fmaf(u4.y, v4.y, a); a = fmaf(u4.z, v4.z, a); a = fmaf(u4.w, v4.w, a);   // ascending index
                }
                acc[h] = a;
            }
        }
    } else {
        __syncwarp();
        const float* v0 = V + (int64_t)(gc[0] >= 0 ? gc[0] - col_offset : 0) * d;
        const float* v1 = V + (int64_t)(gc[1] >= 0 ? gc[1] - col_offset : 0) * d;
        for (int c = 0; c < d; ++c) { acc[0] = fmaf(us[c], __ldg(v0 + c), acc[0]); acc[1] = fmaf(us[c], __ldg(v1 + c), acc[1]); }
    }
    if (bias != nullptr) {
        if (gc[0] >= 0) acc[0] = acc[0] + __ldg(bias + (gc[0] - col_offset));
        if (gc[1] >= 0) acc[1] = acc[1] + __ldg(bias + (gc[1] - col_offset));
    }
    mykey[0] = gc[0] >= 0 ? make_key(acc[0] + 0.0f, gc[0]) : 0ull;
    mykey[1] = gc[1] >= 0 ? make_key(acc[1] + 0.0f, gc[1]) : 0ull;
    keys[lane] = mykey[0];
    keys[32 + lane] = mykey[1];
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
        // (a seeded sweep also dropped whatever scored below its seed threshold tau0)
        float D = cand_idx[row * KPRIME + KPRIME - 1] >= 0 ? cand_score[row * KPRIME + KPRIME - 1] : -INFINITY;
        for (int sp = 0; sp < n_splits; ++sp) D = fmaxf(D, tau0[(int64_t)sp * nu + row]);
        const bool dropped = D > -INFINITY;
        bool ok = !dropped;
        if (dropped) {
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
int g_seed_div = 12;                 // a sweep seeds on its first 1/g_seed_div tiles (tkr_debug_set_seed_div)
int g_filter_mode = 1;               // kernel MODE used while the counters are set (tkr_debug_set_filter_mode)
int g_seed_rank = 0, g_cap_trigger = 0;   // 0 = automatic (tkr_debug_set_filter_tuning)

struct TcPlan {
    int dpad, kb, cps, stages, ns, tps;
    size_t smem;
    int seed_tiles;
    size_t o_ubf, o_vbf, o_unorm, o_scal, o_cand, o_sidx, o_sscore, o_midx, o_mscore, o_fail, o_tau0, o_fb, fb_bytes, total;
};

static bool tc_plan(int64_t nu, int64_t ni, int d, int k, bool has_bias, TcPlan* P) {
    const int dext = d + (has_bias ? 3 : 0);
    P->dpad = (dext + FK - 1) / FK * FK;
    P->kb = P->dpad / FK;
    if (P->kb > MAX_KB || k > 48 || nu <= 0 || ni <= 0) return false;
    if (P->kb == 3) { P->kb = 4; P->dpad = 4 * FK; }       // stages per tile must divide the ring depth
    P->cps = P->kb >= 2 ? 2 : 1;                           // kb in {1, 2, 4}: stages per tile = kb / cps in {1, 1, 2}
    P->stages = P->kb == 1 ? 8 : 4;
    P->smem = 1024 + (size_t)P->kb * A_CHUNK_BYTES + (size_t)P->stages * P->cps * B_CHUNK_BYTES + 256 + FM * (4 + 4 + 8 + 8 + 8) + F_SEL_WARPS * sizeof(SelShared);
    const int64_t row_ctas = 2 * ((nu + 2 * FM - 1) / (2 * FM)), ntiles = (ni + FN - 1) / FN;   // CTAs come in pairs (256 user rows)
    // one CTA per SM; split the items only when the user tiles alone cannot fill the chip
    int64_t ns = row_ctas >= kNumSMs ? 1 : (kNumSMs + row_ctas - 1) / row_ctas;
    const int64_t maxs = ntiles / 16 > 0 ? ntiles / 16 : 1;
    if (ns > maxs) ns = maxs;
    if (ns > 32) ns = 32;
    const int64_t max_tps = ((int64_t)1 << COL_BITS) / FN - 1;     // hand-off messages carry sweep-relative columns
    if ((ntiles + ns - 1) / ns > max_tps) ns = (ntiles + max_tps - 1) / max_tps;
    P->tps = (int)((ntiles + ns - 1) / ns);
    P->ns = (int)((ntiles + P->tps - 1) / P->tps);
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += align_up(b, 1024); return r; };
    P->o_ubf = take((size_t)nu * P->dpad * 2);
    P->o_vbf = take((size_t)ni * P->dpad * 2);
    P->o_unorm = take((size_t)nu * 4);
    P->o_scal = take(64);                                   // [0] vnorm_max  [1] bias_max  [2] n_fail
    P->o_cand = take((size_t)P->ns * nu * CAP * 8);                // one candidate list per (item split, row)
    P->o_sidx = take((size_t)P->ns * nu * KPRIME * 4);
    P->o_sscore = take((size_t)P->ns * nu * KPRIME * 4);
    P->o_midx = take((size_t)nu * KPRIME * 4);
    P->o_mscore = take((size_t)nu * KPRIME * 4);
    P->o_fail = take((size_t)nu * 4);
    P->o_tau0 = take((size_t)P->ns * nu * 4);
    P->fb_bytes = exact_rows_workspace_bytes(ni, k);
    P->o_fb = take(P->fb_bytes + 256);
    // seed on 1/g_seed_div of a sweep (>= 16 seed scores reach the threshold: expected rank ~16 * g_seed_div of the
    // sweep); sweeps under 256 tiles run unseeded
    // (short sweeps -- item shards of a multi-GPU run -- seed on 1/8: the candidate traffic of a sweep is nearly
    // independent of its length, so it weighs more there; measured on 131 072 / 262 144 items: 1.64 -> 1.41 / 1.86 -> 1.62 ms)
    const int div = (g_seed_div == 12 && P->tps < 2048) ? 8 : g_seed_div;
    P->seed_tiles = P->tps >= 256 ? (P->tps / div < 2048 ? P->tps / div : 2048) : 0;
    P->total = o;
    return true;
}

}  // namespace tkr

using namespace tkr;

extern "C" void tkr_debug_set_filter_counters(long long* dev_buf) { g_filter_dbg = dev_buf; }
extern "C" void tkr_debug_set_seed_div(int32_t div) { g_seed_div = div >= 4 && div <= 1024 ? div : 12; }
extern "C" void tkr_debug_set_filter_tuning(int32_t seed_rank, int32_t cap_trigger) {
    g_seed_rank = (seed_rank == 3 || seed_rank == 4) ? seed_rank : 0;
    g_cap_trigger = (cap_trigger >= KPRIME + 8 && cap_trigger <= CAP) ? cap_trigger : 0;
}
extern "C" void tkr_debug_set_filter_mode(int32_t mode) { g_filter_mode = mode >= 1 && mode <= 5 ? mode : 1; }
// CTA pairs of the filter kernel that can be resident at once on the current device (74 on a full B200), or < 0.
extern "C" int32_t tkr_debug_filter_max_pairs(int32_t d) {
    TcPlan P;
    if (!tc_plan(256, 4096, d, 30, false, &P)) return -1;
    if (cudaFuncSetAttribute(score_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148, 1, 1); cfg.blockDim = dim3(F_THREADS, 1, 1); cfg.dynamicSmemBytes = P.smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, score_filter_kernel<0>, &cfg) != cudaSuccess) return -3;
    return n;
}

extern "C" size_t tkr_score_topk_tc_workspace_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k, int32_t has_bias) {
    TcPlan P;
    if (!tc_plan(nu, ni, d, k, has_bias != 0, &P)) return tkr_score_topk_workspace_bytes(nu, ni, d, k);
    return P.total + 1024;
}

namespace tkr {
// state of a sweep that travels from GPU to GPU (tkr_score_topk_tc_segment): [cand nu*CAP u64 | tau nu f32 | cnt nu i32 | tau0 nu f32 | scal 16 u32]
struct SegState { uint64_t* cand; float* tau; int32_t* cnt; float* tau0; unsigned int* scal; size_t total; };
static SegState seg_state(void* state, int64_t nu) {
    SegState S;
    char* p = (char*)state;
    size_t o = 0;
    S.cand = (uint64_t*)(p + o); o += align_up((size_t)nu * CAP * 8, 256);
    S.tau = (float*)(p + o); o += align_up((size_t)nu * 4, 256);
    S.cnt = (int32_t*)(p + o); o += align_up((size_t)nu * 4, 256);
    S.tau0 = (float*)(p + o); o += align_up((size_t)nu * 4, 256);
    S.scal = (unsigned int*)(p + o); o += 256;
    S.total = o;
    return S;
}
// running maxima of the item norms / |bias| over the segments seen so far (non-negative floats: unsigned compare)
__global__ void seg_scal_kernel(unsigned int* __restrict__ st, const unsigned int* __restrict__ mine, int first) {
    if (threadIdx.x < 2) st[threadIdx.x] = first ? mine[threadIdx.x] : max(st[threadIdx.x], mine[threadIdx.x]);
}

// seg == nullptr: the whole pipeline on one shard (tkr_score_topk_tc).  Otherwise one segment of a sweep: first / last say
// whether the sweep starts / ends here; only the last segment refines (against V_full, whose rows the keys' global columns
// index) and re-does the uncertified rows.
static int tc_run(const float* U, int64_t nu, const float* V, int64_t ni, int32_t d, const float* bias,
                  const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset,
                  int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int32_t* n_fallback_rows,
                  int32_t items_prepared, void* stream, const SegState* seg, int first, int last,
                  const float* V_full, int64_t ni_full, const float* bias_full) {
    TcPlan P;
    if (!tc_plan(nu, ni, d, k, bias != nullptr, &P)) { set_error("score_topk_tc: shape outside the tensor-core filter"); return TKR_ERR_UNSUPPORTED; }
    if (seg != nullptr && P.ns != 1) { set_error("score_topk_tc_segment needs user batches that fill the chip without item splits (>= %d rows)", kNumSMs * FM / 2); return TKR_ERR_UNSUPPORTED; }
    if (ws == nullptr || ws_bytes < P.total + 1024) { set_error("score_topk_tc workspace too small: have %zu, need %zu", ws_bytes, P.total + 1024); return TKR_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* Ubf = (__nv_bfloat16*)(w + P.o_ubf);
    __nv_bfloat16* Vbf = (__nv_bfloat16*)(w + P.o_vbf);
    float* unorm = (float*)(w + P.o_unorm);
    unsigned int* scal = (unsigned int*)(w + P.o_scal);
    uint64_t* cand = seg ? seg->cand : (uint64_t*)(w + P.o_cand);
    int32_t* sidx = (int32_t*)(w + P.o_sidx); float* sscore = (float*)(w + P.o_sscore);
    int32_t* midx = (int32_t*)(w + P.o_midx); float* mscore = (float*)(w + P.o_mscore);
    int32_t* fail = (int32_t*)(w + P.o_fail);
    const bool do_tail = seg == nullptr || last;

    const int has_bias = bias != nullptr;
    // scal: [0] max item norm, [1] max |bias| (both belong to the prepared item table), [2] uncertified rows
    if (items_prepared) TKR_CUDA(cudaMemsetAsync(scal + 2, 0, 4, st));
    else TKR_CUDA(cudaMemsetAsync(scal, 0, 64, st));
    convert_rows_kernel<<<(unsigned)((nu + 7) / 8), 256, 0, st>>>(U, nu, d, P.dpad, nullptr, 1, has_bias, Ubf, unorm, nullptr, nullptr);
    TKR_LAUNCH_CHECK();
    if (!items_prepared) {   // the BF16 item table + its norms stay valid in the workspace for later user batches
        convert_rows_kernel<<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(V, ni, d, P.dpad, bias, 0, has_bias, Vbf, nullptr, scal + 0, scal + 1);
        TKR_LAUNCH_CHECK();
    }
    if (seg != nullptr) { seg_scal_kernel<<<1, 32, 0, st>>>(seg->scal, scal, first); TKR_LAUNCH_CHECK(); }
    // A sharded sweep that STARTS here seeds its thresholds on a sample of the whole table (BF16 copy at the end of the segment
    // workspace, built with the shard's): a shard-only seed starts ~8x further down the ranking at 8 shards, and the first
    // segment then paid for it in hand-offs (1.28 ms against 0.47 for a middle segment, profiles/r02_probe_segments.json).
    __nv_bfloat16* Sbf = nullptr;
    size_t seed_bytes = 0;
    const int64_t ntiles_full = (ni_full + FN - 1) / FN;
    if (seg != nullptr && V_full != nullptr && ni_full > ni && (bias == nullptr) == (bias_full == nullptr) && ntiles_full >= 256) {
        seed_bytes = align_up((size_t)ni_full * P.dpad * 2, 1024) + 1024;
        char* end = (char*)ws + ws_bytes;
        if ((size_t)(end - (w + align_up(P.total, 1024))) >= seed_bytes + exact_rows_workspace_bytes(ni_full, k)) {
            Sbf = (__nv_bfloat16*)(((uintptr_t)(end - seed_bytes) + 1023) & ~(uintptr_t)1023);
            if (!items_prepared) {
                convert_rows_kernel<<<(unsigned)((ni_full + 7) / 8), 256, 0, st>>>(V_full, ni_full, d, P.dpad, bias_full, 0, has_bias, Sbf, nullptr, nullptr, nullptr);
                TKR_LAUNCH_CHECK();
            }
        } else seed_bytes = 0;                                 // (a workspace sized by an older caller: shard-only seeding)
    }

    CUtensorMap tmU, tmV, tmS;
    if (int rc = make_tmap(&tmU, Ubf, nu, P.dpad, FM)) return rc;
    if (int rc = make_tmap(&tmV, Vbf, ni, P.dpad, B_HALF)) return rc;
    tmS = tmV;
    if (Sbf != nullptr && first) { if (int rc = make_tmap(&tmS, Sbf, ni_full, P.dpad, B_HALF)) return rc; }
    FilterParams fp = {};
    fp.nu = nu; fp.ni = ni; fp.col_offset = col_offset; fp.kb = P.kb; fp.cps = P.cps; fp.stages = P.stages; fp.tiles_per_split = P.tps;
    fp.rated_indptr = rated_indptr; fp.rated_idx = rated_idx; fp.cand = cand;
    fp.out_idx = P.ns > 1 ? sidx : midx; fp.out_score = P.ns > 1 ? sscore : mscore;
    fp.dbg = g_filter_dbg;
    fp.seed_tiles = (seg != nullptr && !first) ? 0 : P.seed_tiles;
    if (Sbf != nullptr && first) {                            // the whole table's own seeding: 1 / g_seed_div of its tiles
        fp.seed_ext = 1;
        fp.seed_tiles = (int)(ntiles_full / g_seed_div < 2048 ? ntiles_full / g_seed_div : 2048);
        fp.seed_ext_stride = (int)(ntiles_full / fp.seed_tiles);
    }
    fp.out_tau0 = seg ? seg->tau0 : (float*)(w + P.o_tau0);
    fp.resume = (seg != nullptr && !first) ? 1 : 0; fp.suspend = (seg != nullptr && !last) ? 1 : 0;
    fp.st_tau = seg ? seg->tau : nullptr; fp.st_cnt = seg ? seg->cnt : nullptr;
    // (tuning aids; measured on item shards of 2^17 .. 2^20 items, profiles/r02_filter_roles.txt: seeding on the 3rd largest
    // chunk maximum sends a handful of rows per batch to the exact fallback, which costs more than the hand-offs it saves,
    // and compacting at 96 keys is a wash -- the defaults stay 4 / CAP)
    fp.seed_rank = g_seed_rank ? g_seed_rank : 4;
    fp.cap_trigger = g_cap_trigger ? g_cap_trigger : CAP;
    dim3 grid((unsigned)(2 * ((nu + 2 * FM - 1) / (2 * FM))), (unsigned)P.ns);   // clusters of 2 along x
#define TKR_FILTER_LAUNCH(MODE)                                                                                                 \
    do {                                                                                                                        \
        TKR_CUDA(cudaFuncSetAttribute(score_filter_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem)); \
        score_filter_kernel<MODE><<<grid, F_THREADS, P.smem, st>>>(tmU, tmV, tmS, fp);                                       \
    } while (0)
    if (fp.dbg == nullptr) TKR_FILTER_LAUNCH(0);
    else if (g_filter_mode == 2) TKR_FILTER_LAUNCH(2);
    else if (g_filter_mode == 3) TKR_FILTER_LAUNCH(3);
    else if (g_filter_mode == 4) TKR_FILTER_LAUNCH(4);
    else if (g_filter_mode == 5) TKR_FILTER_LAUNCH(5);
    else TKR_FILTER_LAUNCH(1);
    TKR_LAUNCH_CHECK();
    if (fp.dbg != nullptr && g_filter_mode >= 2) return TKR_OK;   // ceiling probes: no lists were produced
    if (!do_tail) return TKR_OK;                                 // the sweep continues on the next shard
    if (P.ns > 1)
        if (int rc = tkr_topk_merge(sidx, sscore, P.ns, nu, KPRIME, midx, mscore, stream)) return rc;

    // exact re-scoring reads the table the keys' columns index: this shard's (col_offset), or -- at the end of a sweep
    // over several shards -- the whole table
    const float* Vr = seg ? V_full : V;
    const float* br = seg ? bias_full : bias;
    const int64_t cor = seg ? 0 : col_offset, nir = seg ? ni_full : ni;
    const unsigned int* vmax = seg ? seg->scal + 0 : scal + 0;
    const unsigned int* bmax = seg ? seg->scal + 1 : scal + 1;
    // |bf16 tensor-core score - exact fma-chain score| <= coef * |u| * |v|: two roundings to 8-bit significands
    // (2^-8 + 2^-18 on every product, Cauchy-Schwarz over the row) + fp32 accumulation slack on both sides.
    const float coef = 0.00390625f * 1.01f + (float)(d + 8) * 9.5367431640625e-7f;
    {
        const int vec = (d % 4 == 0 && ((uintptr_t)Vr % 16) == 0 && ((uintptr_t)U % 16) == 0) ? 1 : 0;
        const int rr = d <= 128 ? 16 : 8;                                // staged candidates per round
        const size_t rsmem = (size_t)8 * (d + 4) * 4 + (size_t)8 * KPRIME * 8 + (vec ? (size_t)8 * rr * (d + 4) * 4 : 0);
#define TKR_REFINE(RR)                                                                                                            \
        do {                                                                                                                       \
            TKR_CUDA(cudaFuncSetAttribute(score_refine_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));       \
            score_refine_kernel<RR><<<(unsigned)((nu + 7) / 8), 256, rsmem, st>>>(U, Vr, nu, d, br, cor, midx, mscore, unorm, vmax, bmax, \
                fp.out_tau0, P.ns, coef, k, vec, out_idx, out_score, fail, (int32_t*)(scal + 2));                                  \
        } while (0)
        if (rr == 16) TKR_REFINE(16); else TKR_REFINE(8);
#undef TKR_REFINE
    }
    TKR_LAUNCH_CHECK();
    // uncertified rows -> exact engine, driven by the device-side row list (no host round trip)
    const size_t fb_need = exact_rows_workspace_bytes(nir, k);
    char* fbp = w + P.o_fb;
    size_t fbb = P.fb_bytes;
    if (seg != nullptr) {   // the fallback scratch of a segment workspace follows the shard-sized plan (it must hold the whole table's splits)
        fbp = w + align_up(P.total, 1024);
        fbb = (size_t)((char*)ws + ws_bytes - fbp) - seed_bytes;   // (the seed table sits at the end)
    }
    if (fb_need > fbb) { set_error("score_topk_tc: workspace was sized for a smaller item table (fallback needs %zu bytes, has %zu)", fb_need, fbb); return TKR_ERR_WORKSPACE; }
    if (int rc = launch_exact_rows(U, nu, Vr, nir, d, br, rated_indptr, rated_idx, k, cor, fail, (const int32_t*)(scal + 2), out_idx, out_score, fbp, fbb, st)) return rc;
    if (n_fallback_rows != nullptr) TKR_CUDA(cudaMemcpyAsync(n_fallback_rows, scal + 2, 4, cudaMemcpyDeviceToDevice, st));
    return TKR_OK;
}
}  // namespace tkr

extern "C" int tkr_score_topk_tc(const float* U, int64_t nu, const float* V, int64_t ni, int32_t d, const float* bias,
                                 const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset,
                                 int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int32_t* n_fallback_rows,
                                 int32_t items_prepared, void* stream) {
    TKR_CHECK_ARG(U && V && out_idx && out_score, "U, V and the outputs must not be NULL");
    TKR_CHECK_ARG(nu >= 0 && ni >= 1 && d >= 1 && k >= 1, "bad nu/ni/d/k");
    TKR_CHECK_ARG(rated_indptr == nullptr || rated_idx != nullptr, "rated_indptr without rated_idx");
    TKR_CHECK_ARG(ni + col_offset < ((int64_t)1 << 31), "global column index exceeds int32");
    if (nu == 0) return TKR_OK;
    TcPlan P;
    if (!tc_plan(nu, ni, d, k, bias != nullptr, &P))   // shape outside the tensor-core filter: the exact engine does it all
        return tkr_score_topk(U, nu, V, ni, d, bias, rated_indptr, rated_idx, k, col_offset, out_idx, out_score, ws, ws_bytes, stream);
    return tc_run(U, nu, V, ni, d, bias, rated_indptr, rated_idx, k, col_offset, out_idx, out_score, ws, ws_bytes, n_fallback_rows, items_prepared, stream,
                  nullptr, 1, 1, nullptr, 0, nullptr);
}

// One SEGMENT of a tensor-core sweep over an item table sharded across GPUs (see include/topkrec.h).
extern "C" size_t tkr_score_topk_tc_state_bytes(int64_t nu) { return nu > 0 ? seg_state(nullptr, nu).total : 0; }
extern "C" size_t tkr_score_topk_tc_segment_workspace_bytes(int64_t nu, int64_t ni_shard, int64_t ni_full, int32_t d, int32_t k, int32_t has_bias) {
    const size_t a = tkr_score_topk_tc_workspace_bytes(nu, ni_shard, d, k, has_bias);
    TcPlan P;
    size_t seed = 0;                                          // BF16 copy of the whole table: the first segment's seed sample
    if (tc_plan(nu, ni_shard, d, k, has_bias != 0, &P) && ni_full > ni_shard) seed = align_up((size_t)ni_full * P.dpad * 2, 1024) + 2048;
    return a + exact_rows_workspace_bytes(ni_full, k) + 4096 + seed;
}
extern "C" int tkr_score_topk_tc_segment(const float* U, int64_t nu, const float* V_shard, int64_t ni_shard, int32_t d, const float* bias_shard,
                                         const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset, void* state,
                                         int32_t first, int32_t last, const float* V_full, int64_t ni_full, const float* bias_full,
                                         int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int32_t* n_fallback_rows,
                                         int32_t items_prepared, void* stream) {
    TKR_CHECK_ARG(U && V_shard && state, "U, V_shard and state must not be NULL");
    TKR_CHECK_ARG(nu >= 1 && ni_shard >= 1 && d >= 1 && k >= 1, "bad nu/ni/d/k");
    TKR_CHECK_ARG(rated_indptr == nullptr || rated_idx != nullptr, "rated_indptr without rated_idx");
    TKR_CHECK_ARG(!last || (V_full && out_idx && out_score && ni_full >= ni_shard), "the last segment needs V_full and the outputs");
    TKR_CHECK_ARG((bias_shard == nullptr) == (!last || bias_full == nullptr) || !last, "bias_shard and bias_full must be given together");
    TKR_CHECK_ARG(ni_shard + col_offset < ((int64_t)1 << 31), "global column index exceeds int32");
    TKR_CHECK_ARG(((uintptr_t)state % 256) == 0, "state must be 256-byte aligned");
    const SegState S = seg_state(state, nu);
    return tc_run(U, nu, V_shard, ni_shard, d, bias_shard, rated_indptr, rated_idx, k, col_offset, out_idx, out_score, ws, ws_bytes, n_fallback_rows,
                  items_prepared, stream, &S, first, last, V_full, ni_full, bias_full);
}
