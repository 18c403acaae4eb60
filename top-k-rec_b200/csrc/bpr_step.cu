// Hot path 1: the synchronous BPR mini-batch step (single/bpr.py:71-101,141 of the
// reference; TF-1.15 RMSProp semantics per SURVEY.md App. A; SGD per old/methods/bpr.py:57-61).
//
// Two kernels per step, both HBM/L2-bandwidth bound (AI ~ 0.33 FLOP/B):
//   bpr_grad_kernel   one warp per 32 triples (lane t owns the scalars of triple t: ids or the fused
//                     sampler draw, biases, flags, loss term); per triple: 128-bit coalesced gather of
//                     U[u], V[i], V[j] into registers (next triple prefetched) -> warp-shuffle dot -> s = sigma(-x) ->
//                     per-occurrence regularised gradients, summed into the per-row
//                     accumulators with vector red.global.add (duplicates summed = the
//                     unique+segment_sum of TF) -> the row is flagged (large batches) or appended
//                     to the step's touched list by its first toucher (small batches).
//   bpr_apply_kernel  one warp per touched row: RMSProp/SGD update from the summed gradient,
//                     re-zeroes the accumulator (the workspace is left clean for the next step).
// All B gradients are therefore taken at the pre-step snapshot and every touched row gets
// exactly ONE optimiser update per step.
// Tables far larger than L2 (MODE_COUNT): the accumulator round trip (read-modify-write in bpr_grad, read + re-zero
// in bpr_apply) would more than double the DRAM traffic of a row, so bpr_count_kernel first counts the occurrences
// of every row in the batch and bpr_grad_kernel updates the rows that occur exactly once in place (no other triple
// reads them, so the pre-step snapshot semantics hold); only duplicated rows take the accumulator path.
#include "bpr_internal.cuh"
#include "bpr_device.cuh"

namespace tkr {

__device__ __forceinline__ bool is_positive(const int32_t* __restrict__ pos, int n, int item) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(pos + mid) < item) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(pos + lo) == item;
}

// The reference redraws the negative forever (bpr.py:163-164); we stop after 2 + 4*(kMaxSampleRounds-1)
// candidates (only reachable when a user likes essentially every item) and keep the last one.
constexpr uint32_t kMaxSampleRounds = 1024;

// One draw of single/bpr.py:159-164.  Philox counter = (draw_lo, draw_hi, round, 0).
__device__ __forceinline__ void sample_triple(const SamplerDev& s, uint64_t draw, int& u, int& i, int& j) {
    uint32_t c[4] = {(uint32_t)draw, (uint32_t)(draw >> 32), 0u, 0u};
    Philox::run(c, s.seed_lo, s.seed_hi);
    u = __ldg(s.tr_users + bounded(c[0], s.n_tr_users));
    const int64_t beg = __ldg(s.pos_indptr + u);
    const int n = (int)(__ldg(s.pos_indptr + u + 1) - beg);
    const int32_t* pos = s.pos_idx + beg;
    i = __ldg(pos + bounded(c[1], (uint32_t)n));
    j = (int)bounded(c[2], s.n_items);
    if (!is_positive(pos, n, j)) return;
    j = (int)bounded(c[3], s.n_items);
    for (uint32_t round = 1; is_positive(pos, n, j) && round < kMaxSampleRounds; ++round) {
        uint32_t r[4] = {(uint32_t)draw, (uint32_t)(draw >> 32), round, 0u};
        Philox::run(r, s.seed_lo, s.seed_hi);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            j = (int)bounded(r[t], s.n_items);
            if (!is_positive(pos, n, j)) return;
        }
    }
}

__global__ void sample_kernel(SamplerDev s, uint64_t first_draw, int64_t n, int32_t* __restrict__ u_out,
                              int32_t* __restrict__ i_out, int32_t* __restrict__ j_out) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        int u, i, j;
        sample_triple(s, first_draw + (uint64_t)t, u, i, j);
        u_out[t] = u; i_out[t] = i; j_out[t] = j;
    }
}

// Rows of one triple held in registers: NCH chunks of 32*VW floats cover a row (d <= 32*VW*NCH).
// INPLACE: also the RMSProp slot rows of the rows this triple will update in place (MODE_COUNT).
template <int VW, int NCH, bool INPLACE>
struct TripleRows {
    Vec<VW> u[NCH], i[NCH], j[NCH];
    Vec<VW> mu[INPLACE ? NCH : 1], mi[INPLACE ? NCH : 1], mj[INPLACE ? NCH : 1];
    __device__ __forceinline__ void load(const float* __restrict__ U, const float* __restrict__ V, int ru, int ri, int rj, int d, int lane) {
        const float* pu = U + (int64_t)ru * d;
        const float* pi = V + (int64_t)ri * d;
        const float* pj = V + (int64_t)rj * d;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int off = (c * 32 + lane) * VW;
            if (off < d) { u[c].load(pu + off); i[c].load(pi + off); j[c].load(pj + off); }
            else {
#pragma unroll
                for (int t = 0; t < VW; ++t) { u[c].v[t] = 0.f; i[c].v[t] = 0.f; j[c].v[t] = 0.f; }
            }
        }
    }
    __device__ __forceinline__ void load_slots(const float* msU, const float* msV, int ru, int ri, int rj, bool su, bool si, bool sj, int d, int lane) {
#pragma unroll
        for (int c = 0; c < (INPLACE ? NCH : 0); ++c) {
            const int off = (c * 32 + lane) * VW;
            if (off < d) {
                if (su) mu[c].load(msU + (int64_t)ru * d + off);
                if (si) mi[c].load(msV + (int64_t)ri * d + off);
                if (sj) mj[c].load(msV + (int64_t)rj * d + off);
            }
        }
    }
};

// MODE_COUNT pre-pass: occurrences of every user / item row in the batch.
__global__ void __launch_bounds__(256) bpr_count_kernel(const int32_t* __restrict__ ub, const int32_t* __restrict__ ib,
                                                        const int32_t* __restrict__ jb, int64_t B, int32_t* __restrict__ cntU,
                                                        int32_t* __restrict__ cntV) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < B; n += (int64_t)gridDim.x * blockDim.x) {
        atomicAdd(cntU + __ldg(ub + n), 1);
        atomicAdd(cntV + __ldg(ib + n), 1);
        atomicAdd(cntV + __ldg(jb + n), 1);
    }
}

// One warp per block of up to 32 consecutive triples.  Lane t owns the scalar side of triple t (ids --
// loaded coalesced or drawn by the fused sampler --, biases, touched flags, bias gradient, loss term);
// the 32 row gathers run through the whole warp one triple after the other, software-pipelined so the
// 128-bit gathers of triple t+1 are in flight while triple t is reduced and scattered.
template <int VW, int NCH, bool L1, bool SAMPLE, bool INPLACE>
__global__ void __launch_bounds__(256, NCH == 1 ? (INPLACE ? 2 : 4) : 1) bpr_grad_kernel(
    tkr_bpr_cfg cfg, const float* __restrict__ U, const float* __restrict__ V, const float* __restrict__ b,
    const int32_t* __restrict__ ub, const int32_t* __restrict__ ib, const int32_t* __restrict__ jb, int64_t B,
    SamplerDev smp, uint64_t first_draw, StepWs ws, int mode, int tpw, StepExtra ex, float* __restrict__ loss_out,
    float* msU, float* msV) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d = cfg.d;
    const bool want_loss = loss_out != nullptr;
    // item-row regularisation applies to the parameter columns only (all of them for plain BPR)
    float lam_i[NCH], lam_j[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const bool par = (c * 32 + lane) * VW < ex.item_cols;
        lam_i[c] = par ? cfg.lambda_i : 0.f;
        lam_j[c] = par ? cfg.lambda_j : 0.f;
    }
    float loss_acc = 0.f;          // per-lane partial; summed over the block at the end
    constexpr unsigned FULL = 0xffffffffu;
    // Popular item rows named by the caller (tkr_bpr_workspace_set_hot_items) are summed here, per block, and added to
    // the global accumulators once at the end: one L2 atomic per block instead of one per occurrence.
    extern __shared__ float hot_smem[];                        // [hmax][d] rows, [hmax] biases, [hmax] dirty flags, [hmax] wq sums
    const int hmax = ex.hot_rows;
    float* hot_b = hot_smem + (size_t)hmax * d;
    float* hot_dirty = hot_b + hmax;
    float* hot_wq = hot_dirty + hmax;
    if (hmax > 0) {
        for (int t = threadIdx.x; t < hmax * (d + 3); t += blockDim.x) hot_smem[t] = 0.f;
        __syncthreads();
    }

    // tpw (<= 32) triples per warp per round: 32 for large batches, fewer when the batch is too small
    // to give every resident warp a full block
    for (int64_t base = warp0 * tpw; base < B; base += nwarps * tpw) {
        const int64_t n = base + lane;
        const bool valid = lane < tpw && n < B;
        const int cnt = (B - base) < tpw ? (int)(B - base) : tpw;
        int u = 0, i = 0, j = 0;
        if (valid) {
            if (SAMPLE) sample_triple(smp, first_draw + (uint64_t)n, u, i, j);
            else { u = __ldg(ub + n); i = __ldg(ib + n); j = __ldg(jb + n); }
        }
        const float bdiff = valid ? __ldg(b + i) - __ldg(b + j) : 0.f;
        const float bi = valid ? __ldg(ex.b_reg + i) : 0.f, bj = valid ? __ldg(ex.b_reg + j) : 0.f;   // regularised bias
        float x_mine = 0.f, s_mine = 0.f;
        // (VBPR as written: weights precomputed per triple; Hogwild: -lr * gradient straight onto the rows.  Neither exists on the
        // in-place streaming route, whose instantiation must not pay for the tests: INPLACE makes both compile-time false.)
        const bool pairwise = !INPLACE && ex.s_emb != nullptr;
        const bool hogwild = !INPLACE && mode == MODE_HOGWILD;
        const float se_mine = (pairwise && valid) ? __ldg(ex.s_emb + n) : 0.f;
        // MODE_COUNT: bit 0/1/2 = the u / i / j row of this triple occurs once in the batch -> updated in place
        int hs_i = 0, hs_j = 0;        // privatised slot + 1 of the item rows (0 = cold)
        if (hmax > 0 && valid) {
            hs_i = ws.hot_slot[i]; hs_j = ws.hot_slot[j];
            if (hs_i > hmax) hs_i = 0;
            if (hs_j > hmax) hs_j = 0;
        }
        int single = 0;
        if (INPLACE && valid) single = (int)(ws.cntU[u] == 1) | ((int)(ws.cntV[i] == 1) << 1) | ((int)(ws.cntV[j] == 1) << 2);
        const bool rms = cfg.optimizer == TKR_OPT_RMSPROP;

        TripleRows<VW, NCH, INPLACE> cur, nxt;
        cur.load(U, V, __shfl_sync(FULL, u, 0), __shfl_sync(FULL, i, 0), __shfl_sync(FULL, j, 0), d, lane);
        if (INPLACE && rms) {
            const int s0 = __shfl_sync(FULL, single, 0);
            cur.load_slots(msU, msV, __shfl_sync(FULL, u, 0), __shfl_sync(FULL, i, 0), __shfl_sync(FULL, j, 0), s0 & 1, s0 & 2, s0 & 4, d, lane);
        }
        for (int t = 0; t < cnt; ++t) {
            const int ru = __shfl_sync(FULL, u, t), ri = __shfl_sync(FULL, i, t), rj = __shfl_sync(FULL, j, t);
            const int sg = INPLACE ? __shfl_sync(FULL, single, t) : 0;
            const int hi = hmax > 0 ? __shfl_sync(FULL, hs_i, t) : 0, hj = hmax > 0 ? __shfl_sync(FULL, hs_j, t) : 0;
            if (t + 1 < cnt) {   // next triple's gathers go out before this one's reduction
                const int nu_ = __shfl_sync(FULL, u, t + 1), ni_ = __shfl_sync(FULL, i, t + 1), nj_ = __shfl_sync(FULL, j, t + 1);
                nxt.load(U, V, nu_, ni_, nj_, d, lane);
                if (INPLACE && rms) {
                    const int s1 = __shfl_sync(FULL, single, t + 1);
                    nxt.load_slots(msU, msV, nu_, ni_, nj_, s1 & 1, s1 & 2, s1 & 4, d, lane);
                }
            }
            float x = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int q = 0; q < VW; ++q) x = fmaf(cur.u[c].v[q], cur.i[c].v[q] - cur.j[c].v[q], x);   // <U_u, V_i - V_j>
            x = warp_sum(x) + __shfl_sync(FULL, bdiff, t);           // x_uij, bpr.py:87-89
            const float s = pairwise ? __shfl_sync(FULL, se_mine, t) : __fdividef(1.0f, 1.0f + __expf(x));      // sigma(-x) = -d/dx log(1+e^-x)
            if (lane == t) { x_mine = x; s_mine = s; }
            if (want_loss) {
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int q = 0; q < VW; ++q)
                        loss_acc += reg_val<L1>(cur.u[c].v[q], cfg.lambda_u) + reg_val<L1>(cur.i[c].v[q], lam_i[c]) + reg_val<L1>(cur.j[c].v[q], lam_j[c]);
            }
            float* gu = ws.GU + (int64_t)ru * d;
            float* gi = ws.GV + (int64_t)ri * d;
            float* gj = ws.GV + (int64_t)rj * d;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int off = (c * 32 + lane) * VW;
                if (off < d) {
                    Vec<VW> a, p, q;
#pragma unroll
                    for (int e = 0; e < VW; ++e) {
                        a.v[e] = fmaf(-s, cur.i[c].v[e] - cur.j[c].v[e], reg_grad<L1>(cur.u[c].v[e], cfg.lambda_u));  // gU   (App. A.2)
                        p.v[e] = fmaf(-s, cur.u[c].v[e], reg_grad<L1>(cur.i[c].v[e], lam_i[c]));                      // gV_i
                        q.v[e] = fmaf(s, cur.u[c].v[e], reg_grad<L1>(cur.j[c].v[e], lam_j[c]));                       // gV_j
                    }
                    if (hogwild) {
#pragma unroll
                        for (int e = 0; e < VW; ++e) { a.v[e] *= -cfg.lr; p.v[e] *= -cfg.lr; q.v[e] *= -cfg.lr; }
                    }
                    if (INPLACE && (sg & 1)) {   // the only occurrence of this row: optimiser update in place
#pragma unroll
                        for (int e = 0; e < VW; ++e) opt_update(cfg, a.v[e], cur.u[c].v[e], cur.mu[c].v[e]);
                        cur.u[c].store(const_cast<float*>(U) + (int64_t)ru * d + off);
                        if (rms) cur.mu[c].store(msU + (int64_t)ru * d + off);
                    } else a.red_add(gu + off);
                    if (INPLACE && (sg & 2)) {
#pragma unroll
                        for (int e = 0; e < VW; ++e) opt_update(cfg, p.v[e], cur.i[c].v[e], cur.mi[c].v[e]);
                        cur.i[c].store(const_cast<float*>(V) + (int64_t)ri * d + off);
                        if (rms) cur.mi[c].store(msV + (int64_t)ri * d + off);
                    } else if (hi) {
#pragma unroll
                        for (int e = 0; e < VW; ++e) sh_red_add(hot_smem + (size_t)(hi - 1) * d + off + e, p.v[e]);
                    } else p.red_add(gi + off);
                    if (INPLACE && (sg & 4)) {
#pragma unroll
                        for (int e = 0; e < VW; ++e) opt_update(cfg, q.v[e], cur.j[c].v[e], cur.mj[c].v[e]);
                        cur.j[c].store(const_cast<float*>(V) + (int64_t)rj * d + off);
                        if (rms) cur.mj[c].store(msV + (int64_t)rj * d + off);
                    } else if (hj) {
#pragma unroll
                        for (int e = 0; e < VW; ++e) sh_red_add(hot_smem + (size_t)(hj - 1) * d + off + e, q.v[e]);
                    } else q.red_add(gj + off);
                }
            }
            cur = nxt;
        }
        // lane-parallel scalar tail: lane t finishes triple t
        if (valid) {
            if (mode == MODE_COUNT || hogwild) {
                // the counts are the touched flags / nothing to flag
            } else if (mode == MODE_DENSE) {
                ws.cntU[u] = 1; ws.tchV[i] = 1.0f; ws.tchV[j] = 1.0f;
            } else {   // first toucher of a row appends it to the step's touched list
                if (atomicAdd(ws.cntU + u, 1) == 0) ws.listU[atomicAdd(ws.n_touched + 0, 1)] = u;
                if (atomicAdd(ws.cntV + i, 1) == 0) ws.listV[atomicAdd(ws.n_touched + 1, 1)] = i;
                if (atomicAdd(ws.cntV + j, 1) == 0) ws.listV[atomicAdd(ws.n_touched + 1, 1)] = j;
            }
            if (pairwise) s_mine = __ldg(ex.s_bias + n);           // the bias / wq path of the [B, B] graph sums over the other index
            const float gsc = hogwild ? -cfg.lr : 1.0f;
            const float gbi = gsc * (-s_mine + reg_grad<L1>(bi, cfg.lambda_b)), gbj = gsc * (s_mine + reg_grad<L1>(bj, cfg.lambda_b));
            if (hs_i) { sh_red_add(hot_b + hs_i - 1, gbi); hot_dirty[hs_i - 1] = 1.f; } else atomicAdd(ws.Gb + i, gbi);
            if (hs_j) { sh_red_add(hot_b + hs_j - 1, gbj); hot_dirty[hs_j - 1] = 1.f; } else atomicAdd(ws.Gb + j, gbj);
            if (ex.wq != nullptr) {
                if (hs_i) sh_red_add(hot_wq + hs_i - 1, -s_mine); else atomicAdd(ex.wq + i, -s_mine);
                if (hs_j) sh_red_add(hot_wq + hs_j - 1, s_mine); else atomicAdd(ex.wq + j, s_mine);
            }
            if (want_loss)   // log(1+e^-x) = max(-x,0) + log(1 + e^-|x|)   (pairwise: the B*B data terms were summed by vbpr_pair_kernel)
                loss_acc += (pairwise ? 0.f : fmaxf(-x_mine, 0.f) + __logf(1.0f + __expf(-fabsf(x_mine)))) + reg_val<L1>(bi, cfg.lambda_b) + reg_val<L1>(bj, cfg.lambda_b);
        }
    }
    if (hmax > 0) {    // flush the block's privatised sums: one vector red per chunk per dirty row
        __syncthreads();
        for (int sl = threadIdx.x >> 5; sl < hmax; sl += blockDim.x >> 5) {
            if (hot_dirty[sl] == 0.f) continue;
            const int id = ws.hot_ids[sl];
            for (int off = lane * VW; off < d; off += 32 * VW) {
                Vec<VW> a;
#pragma unroll
                for (int e = 0; e < VW; ++e) a.v[e] = hot_smem[(size_t)sl * d + off + e];
                a.red_add(ws.GV + (int64_t)id * d + off);
            }
            if (lane == 0) { atomicAdd(ws.Gb + id, hot_b[sl]); if (ex.wq != nullptr) atomicAdd(ex.wq + id, hot_wq[sl]); }
        }
    }
    if (!want_loss) return;
    // block reduction of the per-lane partials -> one atomic per block
    __shared__ float sl[8];
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) sl[threadIdx.x >> 5] = loss_acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? sl[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_out, v);
    }
}

// MODE_LIST : warps [0, nU) update the listed user rows, [nU, nU+nV) the listed item rows (+ their bias);
//             the last block to finish re-arms the list counters (no extra launch).
// MODE_DENSE: warp w looks at row w of [users | items] and updates it if its flag is set
//             (items_dense_only: users still come from the list -- unused today).
template <int VW>
__global__ void __launch_bounds__(256) bpr_apply_kernel(tkr_bpr_cfg cfg, float* __restrict__ U, float* __restrict__ V,
                                                        float* __restrict__ b, float* __restrict__ msU,
                                                        float* __restrict__ msV, float* __restrict__ msb, StepWs ws,
                                                        int mode, StepExtra ex) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d = cfg.d;
    if (mode == MODE_COUNT) {
        // 32 rows per warp step (coalesced count reads); rows that occur once were already updated by the gradient
        // kernel -- only their bias (items) and the count remain; duplicated rows take the accumulator update.
        const int64_t nu32 = ((int64_t)cfg.n_users + 31) / 32, ni32 = ((int64_t)cfg.n_items + 31) / 32;
        for (int64_t w = warp0; w < nu32 + ni32; w += nwarps) {
            const bool users = w < nu32;
            const int64_t r = (users ? w : w - nu32) * 32 + lane;
            const bool in = r < (users ? cfg.n_users : cfg.n_items);
            int32_t* cnt = users ? ws.cntU : ws.cntV;
            const int c = in ? cnt[r] : 0;
            unsigned todo = __ballot_sync(0xffffffffu, c >= 2);
            if (c != 0) {
                cnt[r] = 0;
                if (!users) { apply_row<1>(cfg, b + r, msb + r, ws.Gb + r, 1, 0); if (ex.wq) ex.wq[r] = 0.0f; }
            }
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t rr = (users ? w : w - nu32) * 32 + l;
                if (users) apply_row<VW>(cfg, U + rr * d, msU + rr * d, ws.GU + rr * d, d, lane);
                else apply_row<VW>(cfg, V + rr * d, msV + rr * d, ws.GV + rr * d, ex.item_cols, lane, d);
            }
        }
        return;
    }
    if (mode == MODE_DENSE) {
        const int64_t total = (int64_t)cfg.n_users + cfg.n_items;
        for (int64_t w = warp0; w < total; w += nwarps) {
            if (w < cfg.n_users) {
                const int r = (int)w;
                if (ws.cntU[r] == 0) continue;
                apply_row<VW>(cfg, U + (int64_t)r * d, msU + (int64_t)r * d, ws.GU + (int64_t)r * d, d, lane);
                __syncwarp();
                if (lane == 0) ws.cntU[r] = 0;
            } else {
                const int r = (int)(w - cfg.n_users);
                if (ws.tchV[r] == 0.0f) continue;
                apply_row<VW>(cfg, V + (int64_t)r * d, msV + (int64_t)r * d, ws.GV + (int64_t)r * d, ex.item_cols, lane, d);
                __syncwarp();
                if (lane == 0) { apply_row<1>(cfg, b + r, msb + r, ws.Gb + r, 1, 0); ws.tchV[r] = 0.0f; if (ex.wq) ex.wq[r] = 0.0f; }
            }
        }
        return;
    }
    const int nU = ws.n_touched[0], nV = ws.n_touched[1];
    for (int64_t w = warp0; w < (int64_t)nU + nV; w += nwarps) {
        if (w < nU) {
            const int r = ws.listU[w];
            apply_row<VW>(cfg, U + (int64_t)r * d, msU + (int64_t)r * d, ws.GU + (int64_t)r * d, d, lane);
            if (lane == 0) ws.cntU[r] = 0;
        } else {
            const int r = ws.listV[w - nU];
            apply_row<VW>(cfg, V + (int64_t)r * d, msV + (int64_t)r * d, ws.GV + (int64_t)r * d, ex.item_cols, lane, d);
            if (lane == 0) { apply_row<1>(cfg, b + r, msb + r, ws.Gb + r, 1, 0); ws.cntV[r] = 0; if (ex.wq) ex.wq[r] = 0.0f; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ws.n_touched + 2, 1) == (int)gridDim.x - 1) {
            ws.n_touched[0] = 0; ws.n_touched[1] = 0; ws.n_touched[2] = 0;
        }
    }
}

struct WsLayout { size_t GU, cntU, listU, n_touched, GV, Gb, tchV, cntV, listV, hotV, stage, flow, total; };

WsLayout ws_layout(const tkr_bpr_cfg* cfg, int64_t B) {
    WsLayout L;
    const size_t d = cfg->d, nu = cfg->n_users, ni = cfg->n_items;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    L.GU = take(nu * d * 4);
    L.cntU = take(nu * 4);
    L.listU = take((size_t)(B < (int64_t)nu ? B : (int64_t)nu) * 4);
    L.n_touched = take(16);
    // [GV | Gb | tchV] is one contiguous fp32 region: a single all-reduce covers it in data-parallel mode
    L.GV = o; o += ni * d * 4;
    L.Gb = o; o += ni * 4;
    L.tchV = o; o += ni * 4;
    o = align_up(o, 256);
    L.cntV = take(ni * 4);
    L.listV = take((size_t)(2 * B < (int64_t)ni ? 2 * B : (int64_t)ni) * 4);
    L.hotV = take(ni * 4 + TKR_MAX_HOT * 4);
    // small batches (persistent multi-step kernel): room for kStageTriples sampled triples, drawn a chunk of steps ahead
    L.stage = take(B <= kPersistMaxBatch ? 256 + (size_t)3 * kStageTriples * 4 : 0);   // (+ the kernel's barrier words)
    L.flow = take(B <= kPersistMaxBatch ? bpr_flow_bytes(cfg) : 0);
    L.total = o;
    return L;
}

int bpr_carve(const tkr_bpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes, StepWs* out) {
    const WsLayout L = ws_layout(cfg, B);
    if (ws == nullptr || ws_bytes < L.total) { set_error("bpr workspace too small: have %zu, need %zu", ws_bytes, L.total); return TKR_ERR_WORKSPACE; }
    if ((uintptr_t)ws % 256 != 0) { set_error("bpr workspace must be 256-byte aligned"); return TKR_ERR_WORKSPACE; }
    char* p = (char*)ws;
    out->GU = (float*)(p + L.GU); out->GV = (float*)(p + L.GV); out->Gb = (float*)(p + L.Gb); out->tchV = (float*)(p + L.tchV);
    out->cntU = (int32_t*)(p + L.cntU); out->cntV = (int32_t*)(p + L.cntV);
    out->n_touched = (int32_t*)(p + L.n_touched);
    out->listU = (int32_t*)(p + L.listU); out->listV = (int32_t*)(p + L.listV);
    out->hot_slot = (int32_t*)(p + L.hotV); out->hot_ids = out->hot_slot + cfg->n_items;
    out->sync = B <= kPersistMaxBatch ? (uint32_t*)(p + L.stage) : nullptr;
    out->stage = B <= kPersistMaxBatch ? (int32_t*)(p + L.stage + 256) : nullptr;
    out->flow = B <= kPersistMaxBatch ? p + L.flow : nullptr;
    return TKR_OK;
}

int bpr_check_cfg(const tkr_bpr_cfg* cfg, int64_t B) {
    TKR_CHECK_ARG(cfg != nullptr, "cfg is NULL");
    TKR_CHECK_ARG(cfg->n_users > 0 && cfg->n_items > 0 && cfg->d > 0, "n_users, n_items, d must be positive");
    TKR_CHECK_ARG(B > 0 && B < (int64_t)1 << 31, "batch must be in [1, 2^31)");
    TKR_CHECK_ARG(cfg->optimizer == TKR_OPT_RMSPROP || cfg->optimizer == TKR_OPT_SGD, "unknown optimizer %d", cfg->optimizer);
    return TKR_OK;
}

static inline int64_t grid_cap() { return (int64_t)kNumSMs * 8; }   // 8 CTAs x 8 warps = 64 resident warps per SM

// Dense flags pay a scan of every row in the apply kernel; worth it once a batch touches a good share of
// the rows (and mandatory in data-parallel mode, where the item flags travel with the all-reduce).
int bpr_pick_mode(const tkr_bpr_cfg* cfg, int64_t B, int data_parallel) {
    return (data_parallel || 3 * B >= ((int64_t)cfg->n_users + cfg->n_items) / 8) ? MODE_DENSE : MODE_LIST;
}

// MODE_COUNT pays a counting pass and a scan of every row's count; it wins when the state is far beyond L2 (every
// accumulator access is a DRAM round trip) and most rows of a batch occur once.  Single-GPU tkr_bpr_step only.
int g_count_mode = -1;   // -1 auto, 0 never, 1 whenever legal (tkr_debug_set_count_mode; tests exercise both paths)
static int pick_step_mode(const tkr_bpr_cfg* cfg, int64_t B, bool explicit_triples) {
    const int base = bpr_pick_mode(cfg, B, 0);
    if (!explicit_triples || cfg->l1 || g_count_mode == 0) return base;
    if (g_count_mode == 1) return MODE_COUNT;
    const int64_t rows = (int64_t)cfg->n_users + cfg->n_items;
    const bool beyond_l2 = rows * cfg->d * 4 * 3 > ((int64_t)384 << 20);
    return (beyond_l2 && 3 * B <= rows) ? MODE_COUNT : base;
}

template <int VW, int NCH>
static void launch_grad(const tkr_bpr_cfg* cfg, const float* U, const float* V, const float* b, const int32_t* u,
                        const int32_t* i, const int32_t* j, int64_t B, const SamplerDev& smp, uint64_t first_draw,
                        const StepWs& ws, int mode, const StepExtra& ex, float* loss, cudaStream_t st, float* msU, float* msV) {
    // triples per warp: spread small batches over all resident warps, cap at one per lane
    const int64_t max_warps = grid_cap() * 8;
    int64_t tpw64 = (B + max_warps - 1) / max_warps;
    const int tpw = tpw64 > 32 ? 32 : (int)tpw64;
    int64_t blocks = ((B + tpw - 1) / tpw + 7) / 8;
    if (blocks > grid_cap()) blocks = grid_cap();
    const size_t hot_bytes = (size_t)ex.hot_rows * (cfg->d + 3) * sizeof(float);
#define TKR_K(L1_, S_, IP_) bpr_grad_kernel<VW, NCH, L1_, S_, IP_><<<(unsigned)blocks, 256, hot_bytes, st>>>(*cfg, U, V, b, u, i, j, B, smp, first_draw, ws, mode, tpw, ex, loss, msU, msV)
    if (mode == MODE_COUNT) TKR_K(false, false, true);   // (bpr_pick_mode only returns it for l2 + explicit triples)
    else if (cfg->l1) { if (u == nullptr) TKR_K(true, true, false); else TKR_K(true, false, false); }
    else { if (u == nullptr) TKR_K(false, true, false); else TKR_K(false, false, false); }
#undef TKR_K
}

void bpr_launch_apply(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                      int64_t B, const StepWs& ws, int mode, const StepExtra& ex, cudaStream_t st) {
    const int d = cfg->d;
    int64_t rows = mode == MODE_COUNT ? ((int64_t)cfg->n_users + cfg->n_items) / 32 + 2
                 : mode == MODE_DENSE ? (int64_t)cfg->n_users + cfg->n_items
                                      : (B < cfg->n_users ? B : cfg->n_users) + (2 * B < cfg->n_items ? 2 * B : cfg->n_items);
    int64_t blocks = (rows + 7) / 8;
    if (blocks > grid_cap()) blocks = grid_cap();
    const int a = ex.item_cols;   // vector width must divide both the row pitch and the parameter-column count
    if (d % 4 == 0 && a % 4 == 0) bpr_apply_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*cfg, U, V, b, msU, msV, msb, ws, mode, ex);
    else if (d % 2 == 0 && a % 2 == 0) bpr_apply_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(*cfg, U, V, b, msU, msV, msb, ws, mode, ex);
    else bpr_apply_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*cfg, U, V, b, msU, msV, msb, ws, mode, ex);
}

int bpr_dispatch_grad(const tkr_bpr_cfg* cfg, const float* U, const float* V, const float* b, const int32_t* u,
                      const int32_t* i, const int32_t* j, int64_t B, const SamplerDev& smp, uint64_t first_draw,
                      const StepWs& ws, int mode, const StepExtra& ex, float* loss, cudaStream_t st, float* msU, float* msV) {
    const int d = cfg->d;
    if (mode == MODE_COUNT) {
        if (u == nullptr || cfg->l1 || ex.item_cols != d || ex.wq != nullptr || (cfg->optimizer == TKR_OPT_RMSPROP && !(msU && msV))) {
            set_error("internal: MODE_COUNT needs explicit triples, l2 regularisation, plain BPR rows and the RMSProp slots");
            return TKR_ERR_INVALID;
        }
        int64_t cb = (B + 255) / 256;
        if (cb > grid_cap()) cb = grid_cap();
        bpr_count_kernel<<<(unsigned)cb, 256, 0, st>>>(u, i, j, B, ws.cntU, ws.cntV);
        TKR_LAUNCH_CHECK();
    }
    const int a = ex.item_cols;
    const int vw = (d % 4 == 0 && a % 4 == 0) ? 4 : (d % 2 == 0 && a % 2 == 0) ? 2 : 1;   // widest vector the row pitch allows
    const int nch = (d + 32 * vw - 1) / (32 * vw);
    if (nch > 8) { set_error("d=%d is too wide for the register-resident gather (max %d)", d, 32 * vw * 8); return TKR_ERR_UNSUPPORTED; }
    const int nchp = nch <= 1 ? 1 : nch <= 2 ? 2 : nch <= 4 ? 4 : 8;
#define TKR_GRAD(VW, NCH) launch_grad<VW, NCH>(cfg, U, V, b, u, i, j, B, smp, first_draw, ws, mode, ex, loss, st, msU, msV)
    if (vw == 4) { if (nchp == 1) TKR_GRAD(4, 1); else if (nchp == 2) TKR_GRAD(4, 2); else if (nchp == 4) TKR_GRAD(4, 4); else TKR_GRAD(4, 8); }
    else if (vw == 2) { if (nchp == 1) TKR_GRAD(2, 1); else if (nchp == 2) TKR_GRAD(2, 2); else if (nchp == 4) TKR_GRAD(2, 4); else TKR_GRAD(2, 8); }
    else { if (nchp == 1) TKR_GRAD(1, 1); else if (nchp == 2) TKR_GRAD(1, 2); else if (nchp == 4) TKR_GRAD(1, 4); else TKR_GRAD(1, 8); }
#undef TKR_GRAD
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

size_t bpr_ws_total(const tkr_bpr_cfg* cfg, int64_t B) { return ws_layout(cfg, B).total; }

}  // namespace tkr

using namespace tkr;

extern "C" void tkr_debug_set_count_mode(int32_t m) { g_count_mode = m < -1 || m > 1 ? -1 : m; }

extern "C" size_t tkr_bpr_workspace_bytes(const tkr_bpr_cfg* cfg, int64_t B) {
    if (cfg == nullptr || B <= 0 || cfg->n_users <= 0 || cfg->n_items <= 0 || cfg->d <= 0) return 0;
    return ws_layout(cfg, B).total;
}

extern "C" int tkr_bpr_workspace_layout(const tkr_bpr_cfg* cfg, int64_t B, int64_t* offsets) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    TKR_CHECK_ARG(offsets != nullptr, "offsets is NULL");
    const WsLayout L = ws_layout(cfg, B);
    const size_t v[TKR_WS_NFIELDS] = {L.GU, L.cntU, L.listU, L.n_touched, L.GV, L.Gb, L.tchV, L.cntV, L.listV, L.hotV, L.stage, L.total};
    for (int t = 0; t < TKR_WS_NFIELDS; ++t) offsets[t] = (int64_t)v[t];
    return TKR_OK;
}

extern "C" int tkr_bpr_workspace_init(const tkr_bpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    StepWs v;
    if (int rc = bpr_carve(cfg, B, ws, ws_bytes, &v)) return rc;
    TKR_CUDA(cudaMemsetAsync(ws, 0, ws_layout(cfg, B).total, (cudaStream_t)stream));
    TKR_CUDA(cudaMemsetAsync(v.hot_ids, 0xff, TKR_MAX_HOT * 4, (cudaStream_t)stream));    // no privatised rows
    return TKR_OK;
}

namespace tkr {
__global__ void set_hot_kernel(int32_t* __restrict__ hot_slot, int32_t* __restrict__ hot_ids, int n_items, const int32_t* __restrict__ ids, int n) {
    // one block: clear the previous set, then install the new one
    for (int s = threadIdx.x; s < TKR_MAX_HOT; s += blockDim.x) {
        const int old = hot_ids[s];
        if (old >= 0 && old < n_items) hot_slot[old] = 0;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < TKR_MAX_HOT; s += blockDim.x) {
        const int id = s < n ? ids[s] : -1;
        hot_ids[s] = id;
        if (id >= 0) hot_slot[id] = s + 1;
    }
}
}  // namespace tkr

extern "C" int tkr_bpr_workspace_set_hot_items(const tkr_bpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes,
                                               const int32_t* item_ids_host, int32_t n, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    StepWs v;
    if (int rc = bpr_carve(cfg, B, ws, ws_bytes, &v)) return rc;
    TKR_CHECK_ARG(n >= 0 && n <= TKR_MAX_HOT && (n == 0 || item_ids_host != nullptr), "n must be in [0, %d]", TKR_MAX_HOT);
    for (int a = 0; a < n; ++a) {
        TKR_CHECK_ARG(item_ids_host[a] >= 0 && item_ids_host[a] < cfg->n_items, "hot item id %d out of range", item_ids_host[a]);
        for (int c = 0; c < a; ++c) TKR_CHECK_ARG(item_ids_host[c] != item_ids_host[a], "hot item ids must be distinct");
    }
    cudaStream_t st = (cudaStream_t)stream;
    // the ids ride in the (unused while idle) head of listV; the kernel copies them into place
    if (n > 0) TKR_CUDA(cudaMemcpyAsync(v.listV, item_ids_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    TKR_CUDA(cudaStreamSynchronize(st));            // the host array may be a temporary
    set_hot_kernel<<<1, 64, 0, st>>>(v.hot_slot, v.hot_ids, cfg->n_items, v.listV, n);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

namespace tkr {
int bpr_make_sampler(const tkr_sampler* smp, SamplerDev* out) {
    TKR_CHECK_ARG(smp != nullptr, "sampler is NULL");
    TKR_CHECK_ARG(smp->tr_users && smp->pos_indptr && smp->pos_idx, "sampler tables are NULL");
    TKR_CHECK_ARG(smp->n_tr_users > 0 && smp->n_items > 0, "sampler needs n_tr_users > 0 and n_items > 0");
    out->tr_users = smp->tr_users; out->pos_indptr = smp->pos_indptr; out->pos_idx = smp->pos_idx;
    out->n_tr_users = (uint32_t)smp->n_tr_users; out->n_items = (uint32_t)smp->n_items;
    out->seed_lo = (uint32_t)smp->seed; out->seed_hi = (uint32_t)(smp->seed >> 32);
    return TKR_OK;
}
}  // namespace tkr

extern "C" int tkr_bpr_sample(const tkr_sampler* smp, uint64_t first_draw, int64_t n, int32_t* u_out, int32_t* i_out,
                              int32_t* j_out, void* stream) {
    SamplerDev s;
    if (int rc = bpr_make_sampler(smp, &s)) return rc;
    TKR_CHECK_ARG(n >= 0 && u_out && i_out && j_out, "bad output arguments");
    if (n == 0) return TKR_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s, first_draw, n, u_out, i_out, j_out);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

static int check_state(const tkr_bpr_cfg* cfg, const float* U, const float* V, const float* b) {
    TKR_CHECK_ARG(U && V && b, "U, V, b must not be NULL");
    (void)cfg;
    return TKR_OK;
}

extern "C" int tkr_bpr_grad(const tkr_bpr_cfg* cfg, const float* U, const float* V, const float* b, const int32_t* u,
                            const int32_t* i, const int32_t* j, int64_t B, const tkr_sampler* smp, uint64_t first_draw,
                            float* loss_out, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    if (int rc = check_state(cfg, U, V, b)) return rc;
    SamplerDev sd = {};
    if (u == nullptr) {
        if (int rc = bpr_make_sampler(smp, &sd)) return rc;
        TKR_CHECK_ARG(smp->n_items == cfg->n_items, "sampler n_items != cfg n_items");
    } else {
        TKR_CHECK_ARG(i && j, "i, j must not be NULL when u is given");
    }
    StepWs w;
    if (int rc = bpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    return bpr_dispatch_grad(cfg, U, V, b, u, i, j, B, sd, first_draw, w, bpr_pick_mode(cfg, B, data_parallel), plain_extra(cfg, b), loss_out, (cudaStream_t)stream);
}

extern "C" int tkr_bpr_apply(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                             int64_t B, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    if (int rc = check_state(cfg, U, V, b)) return rc;
    TKR_CHECK_ARG(cfg->optimizer == TKR_OPT_SGD || (msU && msV && msb), "RMSProp needs the msU/msV/msb slots");
    StepWs w;
    if (int rc = bpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    bpr_launch_apply(cfg, U, V, b, msU, msV, msb, B, w, bpr_pick_mode(cfg, B, data_parallel), plain_extra(cfg, b), (cudaStream_t)stream);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

extern "C" int tkr_bpr_step(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                            const int32_t* u, const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps,
                            const tkr_sampler* smp, uint64_t first_draw, float* loss_out, void* ws, size_t ws_bytes,
                            void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    if (int rc = check_state(cfg, U, V, b)) return rc;
    TKR_CHECK_ARG(cfg->optimizer == TKR_OPT_SGD || (msU && msV && msb), "RMSProp needs the msU/msV/msb slots");
    TKR_CHECK_ARG(n_steps >= 0, "n_steps < 0");
    SamplerDev sd = {};
    if (u == nullptr) {
        if (int rc = bpr_make_sampler(smp, &sd)) return rc;
        TKR_CHECK_ARG(smp->n_items == cfg->n_items, "sampler n_items != cfg n_items");
    } else {
        TKR_CHECK_ARG(i && j, "i, j must not be NULL when u is given");
    }
    StepWs w;
    if (int rc = bpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = pick_step_mode(cfg, B, u != nullptr);
    const StepExtra ex = plain_extra(cfg, b);
    if (loss_out != nullptr && n_steps > 0) TKR_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float) * (size_t)n_steps, st));
    // Small batches (the reference's batch_size = 256, bpr.py:103): the persistent cluster kernel runs many steps per
    // launch.  With the fused sampler the triples of a chunk of steps are drawn into the workspace first (same draws).
    // (automatic choice: up to one triple per warp of the cluster; between 257 and 1024 triples the two-launch route is still faster)
    if (g_persist_mode != 0 && mode != MODE_COUNT && bpr_persist_legal(cfg, B) && B <= kPersistMaxBatch) {
        // two multi-step kernels: the dataflow kernel (bpr_flow_steps: no grid-wide barriers, every triple waits only for the rows
        // it reads) and the cluster kernel with two barriers per step (bpr_persist_steps).  g_persist_mode 2 / 1 force one of them,
        // -1 picks by batch size (bpr_internal.cuh)
        const bool flow = g_persist_mode == 2 || (g_persist_mode == -1 && (B <= kFlowSmallBatch || B > kFlowLargeBatch));
        const int64_t chunk = flow ? (u ? kFlowTriples : kStageTriples) / B : (u ? n_steps : kStageTriples / B);
        for (int64_t t = 0; t < n_steps; t += chunk) {
            const int64_t ns = n_steps - t < chunk ? n_steps - t : chunk;
            const int32_t *ut = u ? u + t * B : w.stage, *it = u ? i + t * B : w.stage + kStageTriples, *jt = u ? j + t * B : w.stage + 2 * kStageTriples;
            if (u == nullptr) {
                int64_t blocks = (ns * B + 255) / 256;
                if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
                sample_kernel<<<(unsigned)blocks, 256, 0, st>>>(sd, first_draw + (uint64_t)t * (uint64_t)B, ns * B, w.stage, w.stage + kStageTriples, w.stage + 2 * kStageTriples);
                TKR_LAUNCH_CHECK();
            }
            if (flow) { if (int rc = bpr_flow_steps(cfg, U, V, b, msU, msV, msb, ut, it, jt, B, ns, w, loss_out ? loss_out + t : nullptr, st)) return rc; }
            else if (int rc = bpr_persist_steps(cfg, U, V, b, msU, msV, msb, ut, it, jt, B, ns, w, loss_out ? loss_out + t : nullptr, st)) return rc;
        }
        return TKR_OK;
    }
    for (int64_t t = 0; t < n_steps; ++t) {
        const int32_t* ut = u ? u + t * B : nullptr;
        const int32_t* it = u ? i + t * B : nullptr;
        const int32_t* jt = u ? j + t * B : nullptr;
        float* lt = loss_out ? loss_out + t : nullptr;
        if (int rc = bpr_dispatch_grad(cfg, U, V, b, ut, it, jt, B, sd, first_draw + (uint64_t)t * (uint64_t)B, w, mode, ex, lt, st, msU, msV)) return rc;
        bpr_launch_apply(cfg, U, V, b, msU, msV, msb, B, w, mode, ex, st);
        TKR_LAUNCH_CHECK();
    }
    return TKR_OK;
}

extern "C" int tkr_bpr_step_host(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV,
                                 float* msb, const int32_t* u_host, const int32_t* i_host, const int32_t* j_host,
                                 int64_t B, int64_t n_steps, float* loss_host, void* staging, size_t staging_bytes,
                                 void* ws, size_t ws_bytes, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    TKR_CHECK_ARG(u_host && i_host && j_host && n_steps > 0, "host triples are NULL or n_steps <= 0");
    const size_t n = (size_t)B * (size_t)n_steps;
    const size_t need = align_up(n * 4, 256) * 3 + align_up((size_t)n_steps * 4, 256);
    if (staging == nullptr || staging_bytes < need) { set_error("staging too small: have %zu, need %zu", staging_bytes, need); return TKR_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    char* p = (char*)staging;
    int32_t* du = (int32_t*)p; p += align_up(n * 4, 256);
    int32_t* di = (int32_t*)p; p += align_up(n * 4, 256);
    int32_t* dj = (int32_t*)p; p += align_up(n * 4, 256);
    float* dl = (float*)p;
    if (n_steps == 1 || (g_persist_mode != 0 && bpr_persist_legal(cfg, B) && B <= kPersistMaxBatch)) {
        TKR_CUDA(cudaMemcpyAsync(du, u_host, n * 4, cudaMemcpyHostToDevice, st));
        TKR_CUDA(cudaMemcpyAsync(di, i_host, n * 4, cudaMemcpyHostToDevice, st));
        TKR_CUDA(cudaMemcpyAsync(dj, j_host, n * 4, cudaMemcpyHostToDevice, st));
        if (int rc = tkr_bpr_step(cfg, U, V, b, msU, msV, msb, du, di, dj, B, n_steps, nullptr, 0, dl, ws, ws_bytes, stream)) return rc;
    } else {
        // Several steps in one call: the triples of step t+1 travel on a side stream while step t computes
        // (pinned host memory assumed; pageable memory still works, without the overlap).
        int dev = 0;
        TKR_CUDA(cudaGetDevice(&dev));
        static cudaStream_t copy_stream[64] = {};
        static cudaEvent_t ready[64] = {}, idle[64] = {};
        TKR_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal out of range");
        if (copy_stream[dev] == nullptr) {
            TKR_CUDA(cudaStreamCreateWithFlags(&copy_stream[dev], cudaStreamNonBlocking));
            TKR_CUDA(cudaEventCreateWithFlags(&ready[dev], cudaEventDisableTiming));
            TKR_CUDA(cudaEventCreateWithFlags(&idle[dev], cudaEventDisableTiming));
        }
        cudaStream_t cs = copy_stream[dev];
        TKR_CUDA(cudaEventRecord(idle[dev], st));                 // earlier work on `st` may still read the staging buffer
        TKR_CUDA(cudaStreamWaitEvent(cs, idle[dev], 0));
        for (int64_t t = 0; t < n_steps; ++t) {
            const size_t o = (size_t)t * (size_t)B;
            TKR_CUDA(cudaMemcpyAsync(du + o, u_host + o, (size_t)B * 4, cudaMemcpyHostToDevice, cs));
            TKR_CUDA(cudaMemcpyAsync(di + o, i_host + o, (size_t)B * 4, cudaMemcpyHostToDevice, cs));
            TKR_CUDA(cudaMemcpyAsync(dj + o, j_host + o, (size_t)B * 4, cudaMemcpyHostToDevice, cs));
            TKR_CUDA(cudaEventRecord(ready[dev], cs));
            TKR_CUDA(cudaStreamWaitEvent(st, ready[dev], 0));     // (captures this record: one event serves every step)
            if (int rc = tkr_bpr_step(cfg, U, V, b, msU, msV, msb, du + o, di + o, dj + o, B, 1, nullptr, 0, dl + t, ws, ws_bytes, stream)) return rc;
        }
    }
    if (loss_host != nullptr) TKR_CUDA(cudaMemcpyAsync(loss_host, dl, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, st));
    TKR_CUDA(cudaStreamSynchronize(st));
    return TKR_OK;
}

// Barrier-free ("Hogwild") plain SGD: one kernel per step, every occurrence's -lr * gradient is added straight onto the
// parameter rows with red.global.add while other warps read them (SURVEY 8(f) NEXT-4; the update rule of
// old/methods/bpr.py:57-61 without its batch synchrony).  Not bit-reproducible; equal to the synchronous SGD step when no
// row occurs twice in a batch.  No workspace.
extern "C" int tkr_bpr_hogwild(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, const int32_t* u, const int32_t* i, const int32_t* j,
                               int64_t B, int64_t n_steps, const tkr_sampler* smp, uint64_t first_draw, float* loss_out, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    if (int rc = check_state(cfg, U, V, b)) return rc;
    TKR_CHECK_ARG(n_steps >= 0, "n_steps < 0");
    SamplerDev sd = {};
    if (u == nullptr) {
        if (int rc = bpr_make_sampler(smp, &sd)) return rc;
        TKR_CHECK_ARG(smp->n_items == cfg->n_items, "sampler n_items != cfg n_items");
    } else {
        TKR_CHECK_ARG(i && j, "i, j must not be NULL when u is given");
    }
    cudaStream_t st = (cudaStream_t)stream;
    StepWs w = {};
    w.GU = U; w.GV = V; w.Gb = b;                               // the "accumulators" are the parameters themselves
    StepExtra ex{cfg->d, b, nullptr, 0};
    if (loss_out != nullptr && n_steps > 0) TKR_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float) * (size_t)n_steps, st));
    for (int64_t t = 0; t < n_steps; ++t)
        if (int rc = bpr_dispatch_grad(cfg, U, V, b, u ? u + t * B : nullptr, u ? i + t * B : nullptr, u ? j + t * B : nullptr, B, sd,
                                       first_draw + (uint64_t)t * (uint64_t)B, w, MODE_HOGWILD, ex, loss_out ? loss_out + t : nullptr, st)) return rc;
    return TKR_OK;
}
