// Hot path 2, exact variant: U.V^T for a user tile fused with the rated-filtered running
// top-k, so the score matrix never leaves the SM (replaces evaluate.py:78 np.dot, :81
// np.argsort and the :96-105 walk of the reference).
//
// Scores are bit-defined: fp32 fma chain over ascending feature index,
// + bias, + 0.0f.  tcgen05 has no fp32 x fp32 mode, so this kernel runs the contraction on
// the CUDA cores (register-tiled, smem-staged) and is the *exact* engine: it is what parity
// is asserted on, and the refine/fallback stage of the tensor-core filter.
//
// CTA = BM user rows x all columns of one item split.  U tile resident in smem (k-major),
// V streamed through a double-buffered smem tile; each thread owns an (BM/16) x 4 register
// tile.  Selection: per-row threshold (the current k-th key) in smem; a score that beats it is
// checked against the user's rated CSR, then pushed into a per-row candidate buffer; after
// each column tile the rows that received candidates are re-ranked by a warp (rank counting,
// keys are unique) and the threshold rises.  Expected pushes per row ~ k ln(Ni/k).
#include "common.cuh"
#include <math.h>

namespace tkr {

constexpr int BN = 64;      // columns per tile
constexpr int BK = 32;      // features per smem stage
constexpr int BS = BN + 4;  // Bs row pitch (floats): keeps float4 alignment, spreads banks

template <int BM, int KCAP>
struct ScoreSmem {
    // dynamic smem layout (bytes): As[dpad][BM] | Bs[2][BK][BS] | T[BM][KCAP] | C[BM][BN] | cnt | nvalid | tau
    static size_t bytes(int dpad) {
        return (size_t)dpad * BM * 4 + 2 * BK * BS * 4 + (size_t)BM * KCAP * 8 + (size_t)BM * BN * 8 + BM * 4 * 4;
    }
};

__device__ __forceinline__ bool rated_contains(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t c) {
    const int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(idx + mid) < c) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(idx + lo) == c;
}

// Re-rank one row: T (nvalid sorted keys) U C (n new keys) -> top-k in T.  One warp.
template <int KCAP>
__device__ __forceinline__ void rerank_row(uint64_t* __restrict__ T, const uint64_t* __restrict__ C, int nvalid, int n,
                                           int k, int lane, int* nvalid_out, float* tau_out) {
    const int tot = nvalid + n;
    constexpr int MAXE = (KCAP + BN + 31) / 32;
    uint64_t key[MAXE]; int rank[MAXE];
#pragma unroll
    for (int q = 0; q < MAXE; ++q) {
        const int e = lane + q * 32;
        key[q] = 0; rank[q] = KCAP;
        if (e < tot) {
            const uint64_t ke = e < nvalid ? T[e] : C[e - nvalid];
            int r = 0;
            for (int j = 0; j < nvalid; ++j) r += (T[j] > ke);
            for (int j = 0; j < n; ++j) r += (C[j] > ke);
            key[q] = ke; rank[q] = r;
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < MAXE; ++q)
        if (rank[q] < k) T[rank[q]] = key[q];
    __syncwarp();
    const int nv = tot < k ? tot : k;
    if (lane == 0) {
        *nvalid_out = nv;
        *tau_out = nv == k ? ord_to_f32((uint32_t)(T[k - 1] >> 32)) : -INFINITY;
    }
}

template <int BM, int KCAP, bool VEC4>
__global__ void __launch_bounds__(256) score_topk_kernel(
    const float* __restrict__ U, int64_t nu, const float* __restrict__ V, int64_t ni, int d, int dpad,
    const float* __restrict__ bias, const int64_t* __restrict__ rated_indptr, const int32_t* __restrict__ rated_idx,
    int k, int64_t col_offset, int tiles_per_split, const int32_t* __restrict__ row_map, const int32_t* __restrict__ n_rows_dev,
    int64_t row_begin, int64_t row_limit, int64_t out_rows, int32_t* __restrict__ out_idx, float* __restrict__ out_score) {
    constexpr int RM = BM / 16;  // rows per thread
    // optional indirection: logical row r of this launch is row row_map[r] of U / rated / out, and only
    // the first *n_rows_dev logical rows exist (the tensor-core path hands its uncertified rows over this way)
    // Logical rows [row_begin, min(n_rows, row_limit)) are processed.  out_rows > 0: lists are written at the LOGICAL
    // row (stride out_rows per split) for a later merge; otherwise at the actual row (stride nu).
    int64_t n_rows = n_rows_dev ? (int64_t)*n_rows_dev : nu;
    if (row_limit > 0 && n_rows > row_limit) n_rows = row_limit;
    if (row_begin + (int64_t)blockIdx.x * BM >= n_rows) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);
    float* Bs = As + (size_t)dpad * BM;
    uint64_t* T = reinterpret_cast<uint64_t*>(Bs + 2 * BK * BS);
    uint64_t* C = T + (size_t)BM * KCAP;
    int* cnt = reinterpret_cast<int*>(C + (size_t)BM * BN);
    int* nvalid = cnt + BM;
    float* tau = reinterpret_cast<float*>(nvalid + BM);
    int* arow = reinterpret_cast<int*>(tau + BM);   // actual row per logical row, -1 = none

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = row_begin + (int64_t)blockIdx.x * BM;
    const int split = blockIdx.y;
    const int64_t ntiles = (ni + BN - 1) / BN;
    const int64_t tile_beg = (int64_t)split * tiles_per_split;
    int64_t tile_end = tile_beg + tiles_per_split;
    if (tile_end > ntiles) tile_end = ntiles;

    for (int idx = tid; idx < BM; idx += 256) {
        const int64_t lr = row0 + idx;
        arow[idx] = lr < n_rows ? (row_map ? row_map[lr] : (int)lr) : -1;
    }
    __syncthreads();
    // ---- stage the U tile, k-major: As[kk][r]; lanes walk rows so the stores are conflict-free
    for (int idx = tid; idx < BM * (dpad / 4); idx += 256) {
        const int r = idx % BM, kq = (idx / BM) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (arow[r] >= 0) {
            const float* p = U + (int64_t)arow[r] * d + kq;
            if (VEC4) { if (kq < d) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } }
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (kq + e < d) v[e] = p[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) As[(size_t)(kq + e) * BM + r] = v[e];
    }
    for (int idx = tid; idx < BM * KCAP; idx += 256) T[idx] = 0;
    for (int idx = tid; idx < BM; idx += 256) { cnt[idx] = 0; nvalid[idx] = 0; tau[idx] = -INFINITY; }
    __syncthreads();

    const int nkc = dpad / BK;
    // V-tile loader mapping: col = tid % 64, two float4 along k at kq = (tid/64)*4 and +16
    const int lcol = tid & 63, lkq = (tid >> 6) * 4;

    for (int64_t tile = tile_beg; tile < tile_end; ++tile) {
        const int64_t n0 = tile * BN;
        float acc[RM][4];
#pragma unroll
        for (int r = 0; r < RM; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

        float pre[2][4];
        auto gload = [&](int kc) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kq = kc * BK + lkq + h * 16;
#pragma unroll
                for (int e = 0; e < 4; ++e) pre[h][e] = 0.f;
                if (n0 + lcol < ni) {
                    const float* p = V + (n0 + lcol) * d + kq;
                    if (VEC4) { if (kq < d) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); pre[h][0] = t.x; pre[h][1] = t.y; pre[h][2] = t.z; pre[h][3] = t.w; } }
                    else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (kq + e < d) pre[h][e] = __ldg(p + e);
                    }
                }
            }
        };
        auto sstore = [&](int buf) {
            float* B = Bs + buf * BK * BS;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 4; ++e) B[(lkq + h * 16 + e) * BS + lcol] = pre[h][e];
        };
        gload(0); sstore(0);
        __syncthreads();
        for (int kc = 0; kc < nkc; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < nkc) gload(kc + 1);
            const float* A = As + (size_t)kc * BK * BM + ty * RM;
            const float* B = Bs + buf * BK * BS + tx * 4;
#pragma unroll 8
            for (int kk = 0; kk < BK; ++kk) {
                float a[RM];
                if constexpr (RM == 8) {
                    const float4 a0 = *reinterpret_cast<const float4*>(A + kk * BM), a1 = *reinterpret_cast<const float4*>(A + kk * BM + 4);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                } else if constexpr (RM == 4) {
                    const float4 a0 = *reinterpret_cast<const float4*>(A + kk * BM);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                } else {
                    const float2 a0 = *reinterpret_cast<const float2*>(A + kk * BM);
                    a[0] = a0.x; a[1] = a0.y;
                }
                const float4 bv = *reinterpret_cast<const float4*>(B + kk * BS);
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    acc[r][0] = fmaf(a[r], bv.x, acc[r][0]);
                    acc[r][1] = fmaf(a[r], bv.y, acc[r][1]);
                    acc[r][2] = fmaf(a[r], bv.z, acc[r][2]);
                    acc[r][3] = fmaf(a[r], bv.w, acc[r][3]);
                }
            }
            if (kc + 1 < nkc) sstore(buf ^ 1);
            __syncthreads();
        }

        // ---- selection epilogue
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias != nullptr) {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (n0 + tx * 4 + c < ni) bv[c] = __ldg(bias + n0 + tx * 4 + c);
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            const int row = ty * RM + r;
            const float t = tau[row];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float s = acc[r][c];
                if (bias != nullptr) s = s + bv[c];
                s = s + 0.0f;
                if (s >= t) {
                    const int64_t lc = n0 + tx * 4 + c;
                    if (lc < ni && arow[row] >= 0) {
                        const int32_t gc = (int32_t)(lc + col_offset);
                        const uint64_t key = make_key(s, gc);
                        const bool full = nvalid[row] == k;
                        if (!full || key > T[(size_t)row * KCAP + k - 1]) {
                            bool rated = false;
                            if (rated_indptr != nullptr)
                                rated = rated_contains(rated_idx, __ldg(rated_indptr + arow[row]), __ldg(rated_indptr + arow[row] + 1), gc);
                            if (!rated) C[(size_t)row * BN + atomicAdd(cnt + row, 1)] = key;
                        }
                    }
                }
            }
        }
        __syncthreads();
        for (int row = warp; row < BM; row += 8) {
            const int n = cnt[row];
            if (n > 0) {
                rerank_row<KCAP>(T + (size_t)row * KCAP, C + (size_t)row * BN, nvalid[row], n, k, lane, nvalid + row, tau + row);
                if (lane == 0) cnt[row] = 0;
            }
        }
        __syncthreads();
    }

    // ---- write the lists (split s of row r at [(s*nu + r)*k])
    for (int idx = tid; idx < BM * k; idx += 256) {
        const int row = idx / k, p = idx % k;
        if (arow[row] >= 0) {
            const uint64_t key = T[(size_t)row * KCAP + p];
            const int64_t o = (out_rows > 0 ? (int64_t)split * out_rows + (row0 - row_begin + row) : (int64_t)split * nu + arow[row]) * k + p;
            out_idx[o] = key ? (int32_t)(uint32_t)key : -1;
            out_score[o] = key ? ord_to_f32((uint32_t)(key >> 32)) : -INFINITY;
        }
    }
}

// Merge n_lists sorted lists per row (warp per row): rank of an element = its position in its
// own list + the number of greater keys in every other list (binary search; keys unique).
__global__ void __launch_bounds__(128) topk_merge_kernel(const int32_t* __restrict__ idx, const float* __restrict__ score,
                                                         int n_lists, int64_t nu, int k, int32_t* __restrict__ out_idx,
                                                         float* __restrict__ out_score, const int32_t* __restrict__ row_map,
                                                         const int32_t* __restrict__ n_rows_dev) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * n_lists * k;
    const int tot = n_lists * k;
    // optional indirection: input row r (of the first *n_rows_dev) is written to output row row_map[r]
    const int64_t n_in = n_rows_dev ? ((int64_t)*n_rows_dev < nu ? (int64_t)*n_rows_dev : nu) : nu;
    for (int64_t row = (int64_t)blockIdx.x * 4 + warp; row < n_in; row += (int64_t)gridDim.x * 4) {
        const int64_t orow = row_map ? (int64_t)row_map[row] : row;
        for (int e = lane; e < tot; e += 32) {
            const int g = e / k, p = e % k;
            const int64_t o = ((int64_t)g * nu + row) * k + p;
            const int32_t c = idx[o];
            keys[e] = c >= 0 ? make_key(score[o], c) : 0;
        }
        for (int p = lane; p < k; p += 32) { out_idx[orow * k + p] = -1; out_score[orow * k + p] = -INFINITY; }
        __syncwarp();
        for (int e = lane; e < tot; e += 32) {
            const uint64_t ke = keys[e];
            if (ke == 0) continue;
            const int g = e / k;
            int rank = e % k;
            for (int g2 = 0; g2 < n_lists && rank < k; ++g2) {
                if (g2 == g) continue;
                const uint64_t* L = keys + g2 * k;  // descending, zero padded
                int lo = 0, hi = k;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (L[mid] > ke) lo = mid + 1; else hi = mid; }
                rank += lo;
            }
            if (rank < k) {
                out_idx[orow * k + rank] = (int32_t)(uint32_t)ke;
                out_score[orow * k + rank] = ord_to_f32((uint32_t)(ke >> 32));
            }
        }
        __syncwarp();
    }
}

static int pick_splits(int64_t nu, int64_t ni, int BM, int* tiles_per_split) {
    const int64_t row_tiles = (nu + BM - 1) / BM, ntiles = (ni + BN - 1) / BN;
    // aim for >= 2 waves of CTAs over the 148 SMs, but keep >= 8 column tiles per split
    int64_t want = (2 * kNumSMs + row_tiles - 1) / row_tiles;
    int64_t maxs = ntiles / 8; if (maxs < 1) maxs = 1;
    if (want > maxs) want = maxs;
    if (want > 64) want = 64;
    if (want < 1) want = 1;
    *tiles_per_split = (int)((ntiles + want - 1) / want);
    return (int)((ntiles + *tiles_per_split - 1) / *tiles_per_split);
}

static int pick_bm(int d) { return d <= 128 ? 128 : d <= 256 ? 64 : 32; }

}  // namespace tkr

using namespace tkr;

extern "C" size_t tkr_score_topk_workspace_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k) {
    if (nu <= 0 || ni <= 0 || d <= 0 || k <= 0) return 0;
    int tps;
    const int ns = pick_splits(nu, ni, pick_bm(d), &tps);
    if (ns == 1) return 256;
    return 2 * align_up((size_t)ns * nu * k * 4, 256);
}

static int merge_lists(const int32_t* idx, const float* score, int32_t n_lists, int64_t nu, int32_t k, int32_t* out_idx,
                       float* out_score, const int32_t* row_map, const int32_t* n_rows_dev, void* stream) {
    TKR_CHECK_ARG(idx && score && out_idx && out_score, "NULL list pointer");
    TKR_CHECK_ARG(n_lists >= 1 && nu >= 0 && k >= 1, "bad n_lists/nu/k");
    TKR_CHECK_ARG((int64_t)n_lists * k <= 4096, "n_lists*k = %lld exceeds 4096", (long long)n_lists * k);
    if (nu == 0) return TKR_OK;
    const size_t smem = (size_t)4 * n_lists * k * 8;
    if (smem > 48 * 1024) TKR_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (nu + 3) / 4;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    topk_merge_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(idx, score, n_lists, nu, k, out_idx, out_score, row_map, n_rows_dev);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

namespace tkr {
int merge_lists_public(const int32_t* idx, const float* score, int n_lists, int64_t nu, int k, int32_t* out_idx, float* out_score, void* stream) {
    return merge_lists(idx, score, n_lists, nu, k, out_idx, out_score, nullptr, nullptr, stream);
}
}  // namespace tkr

extern "C" int tkr_topk_merge(const int32_t* idx, const float* score, int32_t n_lists, int64_t nu, int32_t k,
                              int32_t* out_idx, float* out_score, void* stream) {
    return merge_lists(idx, score, n_lists, nu, k, out_idx, out_score, nullptr, nullptr, stream);
}

template <int BM, int KCAP, bool VEC4>
static int launch_score(const float* U, int64_t nu, const float* V, int64_t ni, int d, const float* bias,
                        const int64_t* rp, const int32_t* ri, int k, int64_t col_offset, int ns, int tps,
                        const int32_t* row_map, const int32_t* n_rows_dev, int64_t row_begin, int64_t row_limit, int64_t out_rows,
                        int64_t grid_rows, int32_t* oi, float* os, cudaStream_t st) {
    const int dpad = (d + BK - 1) / BK * BK;
    const size_t smem = ScoreSmem<BM, KCAP>::bytes(dpad);
    if (smem > 227 * 1024) { set_error("score_topk: d=%d k=%d needs %zu B of shared memory (> 227 KB)", d, k, smem); return TKR_ERR_UNSUPPORTED; }
    auto kern = score_topk_kernel<BM, KCAP, VEC4>;
    TKR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((grid_rows + BM - 1) / BM), (unsigned)ns);
    kern<<<grid, 256, smem, st>>>(U, nu, V, ni, d, dpad, bias, rp, ri, k, col_offset, tps, row_map, n_rows_dev, row_begin, row_limit, out_rows, oi, os);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

static int dispatch_score(const float* U, int64_t nu, const float* V, int64_t ni, int d, const float* bias,
                          const int64_t* rp, const int32_t* ri, int k, int64_t col_offset, int ns, int tps,
                          const int32_t* row_map, const int32_t* n_rows_dev, int64_t row_begin, int64_t row_limit,
                          int64_t out_rows, int64_t grid_rows, int32_t* oi, float* os, cudaStream_t st) {
    if (d > 512) { set_error("score_topk: d=%d > 512 is not supported by the exact kernel", d); return TKR_ERR_UNSUPPORTED; }
    const int BM = pick_bm(d);
    const bool vec4 = (d % 4 == 0) && ((uintptr_t)U % 16 == 0) && ((uintptr_t)V % 16 == 0);
    const int kcap = k <= 32 ? 32 : 64;
    int rc;
#define TKR_SCORE(BM_, KC_, V4_) rc = launch_score<BM_, KC_, V4_>(U, nu, V, ni, d, bias, rp, ri, k, col_offset, ns, tps, row_map, n_rows_dev, row_begin, row_limit, out_rows, grid_rows, oi, os, st)
#define TKR_SCORE_BM(BM_)                                                   \
    do {                                                                    \
        if (kcap == 32) { if (vec4) TKR_SCORE(BM_, 32, true); else TKR_SCORE(BM_, 32, false); } \
        else { if (vec4) TKR_SCORE(BM_, 64, true); else TKR_SCORE(BM_, 64, false); }            \
    } while (0)
    if (BM == 128) TKR_SCORE_BM(128); else if (BM == 64) TKR_SCORE_BM(64); else TKR_SCORE_BM(32);
#undef TKR_SCORE_BM
#undef TKR_SCORE
    return rc;
}

namespace tkr {
constexpr int64_t kFallbackRows = 1024;   // uncertified rows that get the item-split (fast) treatment

static int64_t fallback_splits(int64_t ni, int k) {
    const int64_t ntiles = (ni + BN - 1) / BN;
    int64_t ns = ntiles / 8;
    if (ns > 4096 / k) ns = 4096 / k;      // the merge kernel handles n_lists * k <= 4096 keys per row
    if (ns > 128) ns = 128;
    return ns < 1 ? 1 : ns;
}

size_t exact_rows_workspace_bytes(int64_t ni, int k) {
    const int64_t ns = fallback_splits(ni, k);
    return ns > 1 ? 2 * align_up((size_t)ns * kFallbackRows * k * 4, 256) : 0;
}

// Exact engine over a device-side list of rows (row_map[0 .. *n_rows_dev)), results written in place into
// out[row_map[r]].  The first kFallbackRows listed rows are swept with the items split over many CTAs and
// merged (a handful of rows must not serialise a whole item sweep on one SM); any further rows take the
// unsplit kernel.  CTAs beyond the list exit immediately, so both launches are always issued.
int launch_exact_rows(const float* U, int64_t nu, const float* V, int64_t ni, int d, const float* bias,
                      const int64_t* rp, const int32_t* ri, int k, int64_t col_offset, const int32_t* row_map,
                      const int32_t* n_rows_dev, int32_t* oi, float* os, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int64_t ntiles = (ni + BN - 1) / BN;
    const int64_t ns = fallback_splits(ni, k);
    const size_t half = align_up((size_t)ns * kFallbackRows * k * 4, 256);
    const int64_t head = kFallbackRows < nu ? kFallbackRows : nu;
    if (ns > 1 && ws != nullptr && ws_bytes >= 2 * half) {
        const int tps = (int)((ntiles + ns - 1) / ns);
        const int ns_eff = (int)((ntiles + tps - 1) / tps);
        int32_t* pi = (int32_t*)ws; float* ps = (float*)((char*)ws + half);
        if (int rc = dispatch_score(U, nu, V, ni, d, bias, rp, ri, k, col_offset, ns_eff, tps, row_map, n_rows_dev, 0, head, kFallbackRows,
                                    head, pi, ps, st)) return rc;
        if (int rc = merge_lists(pi, ps, ns_eff, kFallbackRows, k, oi, os, row_map, n_rows_dev, st)) return rc;
        if (nu > head)
            if (int rc = dispatch_score(U, nu, V, ni, d, bias, rp, ri, k, col_offset, 1, (int)ntiles, row_map, n_rows_dev, head, 0, 0,
                                        nu - head, oi, os, st)) return rc;
        return TKR_OK;
    }
    return dispatch_score(U, nu, V, ni, d, bias, rp, ri, k, col_offset, 1, (int)ntiles, row_map, n_rows_dev, 0, 0, 0, nu, oi, os, st);
}
}  // namespace tkr

extern "C" int tkr_score_topk(const float* U, int64_t nu, const float* V, int64_t ni, int32_t d, const float* bias,
                              const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset,
                              int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, void* stream) {
    TKR_CHECK_ARG(U && V && out_idx && out_score, "U, V and the outputs must not be NULL");
    TKR_CHECK_ARG(nu >= 0 && ni >= 1 && d >= 1, "bad nu/ni/d");
    TKR_CHECK_ARG(k >= 1 && k <= 64, "k must be in [1, 64] (got %d)", k);
    TKR_CHECK_ARG(rated_indptr == nullptr || rated_idx != nullptr, "rated_indptr without rated_idx");
    TKR_CHECK_ARG(ni + col_offset < ((int64_t)1 << 31), "global column index exceeds int32");
    if (nu == 0) return TKR_OK;
    if (d > 512) { set_error("score_topk: d=%d > 512 is not supported by the exact kernel", d); return TKR_ERR_UNSUPPORTED; }
    const int BM = pick_bm(d);
    int tps;
    const int ns = pick_splits(nu, ni, BM, &tps);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* oi = out_idx; float* os = out_score;
    if (ns > 1) {
        const size_t half = align_up((size_t)ns * nu * k * 4, 256);
        if (ws == nullptr || ws_bytes < 2 * half) { set_error("score_topk workspace too small: have %zu, need %zu", ws_bytes, 2 * half); return TKR_ERR_WORKSPACE; }
        oi = (int32_t*)ws; os = (float*)((char*)ws + half);
    }
    if (int rc = dispatch_score(U, nu, V, ni, d, bias, rated_indptr, rated_idx, k, col_offset, ns, tps, nullptr, nullptr, 0, 0, 0, nu, oi, os, st)) return rc;
    if (ns > 1) return tkr_topk_merge(oi, os, ns, nu, k, out_idx, out_score, stream);
    return TKR_OK;
}

extern "C" size_t tkr_score_topk_host_device_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k, int64_t n_rated) {
    if (nu <= 0 || ni <= 0 || d <= 0 || k <= 0) return 0;
    size_t n = 0;
    n += align_up((size_t)nu * d * 4, 256) + align_up((size_t)ni * d * 4, 256) + align_up((size_t)ni * 4, 256);
    n += align_up((size_t)(nu + 1) * 8, 256) + align_up((size_t)(n_rated > 0 ? n_rated : 1) * 4, 256);
    n += 2 * align_up((size_t)nu * k * 4, 256);
    const size_t a = tkr_score_topk_workspace_bytes(nu, ni, d, k), b = tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, 1);
    n += (a > b ? a : b) + 1024;
    return n;
}

extern "C" int tkr_score_topk_host(const float* U_host, int64_t nu, const float* V_host, int64_t ni, int32_t d,
                                   const float* bias_host, const int64_t* rated_indptr_host,
                                   const int32_t* rated_idx_host, int32_t k, int32_t* out_idx_host,
                                   float* out_score_host, void* dev, size_t dev_bytes, void* stream) {
    TKR_CHECK_ARG(U_host && V_host && out_idx_host && out_score_host, "NULL host pointer");
    TKR_CHECK_ARG(nu >= 1 && ni >= 1 && d >= 1 && k >= 1, "bad nu/ni/d/k");
    const int64_t n_rated = rated_indptr_host ? rated_indptr_host[nu] - rated_indptr_host[0] : 0;
    const size_t need = tkr_score_topk_host_device_bytes(nu, ni, d, k, n_rated);
    if (dev == nullptr || dev_bytes < need) { set_error("device scratch too small: have %zu, need %zu", dev_bytes, need); return TKR_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    char* p = (char*)dev;
    auto take = [&](size_t b) { char* r = p; p += align_up(b, 256); return r; };
    float* dU = (float*)take((size_t)nu * d * 4);
    float* dV = (float*)take((size_t)ni * d * 4);
    float* dB = (float*)take((size_t)ni * 4);
    int64_t* dP = (int64_t*)take((size_t)(nu + 1) * 8);
    int32_t* dI = (int32_t*)take((size_t)(n_rated > 0 ? n_rated : 1) * 4);
    int32_t* dOi = (int32_t*)take((size_t)nu * k * 4);
    float* dOs = (float*)take((size_t)nu * k * 4);
    void* ws = p;
    const size_t ws_bytes = dev_bytes - (size_t)(p - (char*)dev);
    TKR_CUDA(cudaMemcpyAsync(dU, U_host, (size_t)nu * d * 4, cudaMemcpyHostToDevice, st));
    TKR_CUDA(cudaMemcpyAsync(dV, V_host, (size_t)ni * d * 4, cudaMemcpyHostToDevice, st));
    if (bias_host) TKR_CUDA(cudaMemcpyAsync(dB, bias_host, (size_t)ni * 4, cudaMemcpyHostToDevice, st));
    if (rated_indptr_host) {
        TKR_CHECK_ARG(rated_idx_host != nullptr || n_rated == 0, "rated_idx_host is NULL");
        TKR_CHECK_ARG(rated_indptr_host[0] == 0, "rated_indptr_host must start at 0");
        TKR_CUDA(cudaMemcpyAsync(dP, rated_indptr_host, (size_t)(nu + 1) * 8, cudaMemcpyHostToDevice, st));
        if (n_rated > 0) TKR_CUDA(cudaMemcpyAsync(dI, rated_idx_host, (size_t)n_rated * 4, cudaMemcpyHostToDevice, st));
    }
    // tensor-core filter + exact refine (bit-identical to the exact engine; shapes it does not cover fall through to it)
    if (int rc = tkr_score_topk_tc(dU, nu, dV, ni, d, bias_host ? dB : nullptr, rated_indptr_host ? dP : nullptr,
                                   rated_indptr_host ? dI : nullptr, k, 0, dOi, dOs, ws, ws_bytes, nullptr, 0, stream)) return rc;
    TKR_CUDA(cudaMemcpyAsync(out_idx_host, dOi, (size_t)nu * k * 4, cudaMemcpyDeviceToHost, st));
    TKR_CUDA(cudaMemcpyAsync(out_score_host, dOs, (size_t)nu * k * 4, cudaMemcpyDeviceToHost, st));
    TKR_CUDA(cudaStreamSynchronize(st));
    return TKR_OK;
}
