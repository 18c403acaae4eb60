// Hot path 1 at the reference's own operating point (single/bpr.py:103: batch_size = 256, 3 906 steps per epoch):
// MANY synchronous mini-batch steps per launch.  A step of 256 triples moves ~0.4 MB; as two launches it costs ~14 us of
// launch latency and tail, ~50x what the data movement needs.  Here one thread-block CLUSTER (256 warps on 16 or 8 SMs of
// one GPC, one triple per warp) runs the whole step loop and the two grid-wide synchronisations a synchronous step needs
// (all gradients of a batch are taken at the pre-step snapshot; every touched row gets exactly one update) are
// hardware cluster barriers (~380 cycles each) instead of kernel boundaries:
//
//   per step, warp w <- triple w:   ids -> 128-bit L2 gathers of U[u], V[i], V[j] -> warp-shuffle dot -> s = sigma(-x)
//       -> per-occurrence regularised gradients summed into the per-row accumulators (red.global.add, L2)
//       -> the warp that touches a row FIRST (atomic claim on the row's counter) becomes its updater and already
//          prefetches the row's RMSProp slot                                   -> fence, cluster barrier 1
//       -> updater warps read the summed gradient, apply the optimiser to the row they still hold in registers (no
//          other warp has written it: the snapshot IS the current value), store row + slot, re-zero the accumulator
//          and the counter                                                          -> fence, cluster barrier 2
//
// Same arithmetic per element as bpr_grad_kernel / bpr_apply_kernel (shared helpers), so the two routes agree to the
// order of the atomic sums.  Mutable state is read with ld.global.cg (L2): an SM's L1 may hold a row from an earlier
// step of the same launch.  Used by tkr_bpr_step for batches up to 1024 triples and d <= 256 (wider rows / larger
// batches keep the two-kernel route, which is bandwidth- rather than latency-bound there).
#include "bpr_device.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace tkr {

// 256 warps per cluster either way: 16 CTAs x 16 warps (non-portable cluster size; 128 registers per thread, no spills,
// half the L2->SM traffic per SM) where a GPC can host it, else the portable 8 CTAs x 32 warps (64 registers per thread)
constexpr int kPersistWarps = 256;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Grid-wide barrier of the (co-scheduled) cluster through ONE global counter: the CTA synchronises, its thread 0 arrives
// with a gpu-scope RELEASE (the CTA's gradients / rows are performed at L2 first) and polls with relaxed loads, the CTA
// synchronises again.  No acquire on the waiting side on purpose: every acquire at cluster scope or wider -- the
// hardware cluster barrier's wait included -- invalidates the SM's L1 (CCTL.IVALL), and the first loads behind it then
// is paired with an L1 invalidation (CCTL.IVALL) this kernel has no use for: all mutable data is read with ld.global.cg,
// which never looks at L1.  (Measured: same step time as the hardware barrier, profiles/r02_persist_phases.txt -- the cost of a
// barrier here is the release itself, ~1 000-1 500 cycles of waiting for the CTA's outstanding writes.)
struct GridBarrier {
    uint32_t* counter;
    uint32_t target;            // value of the counter once every CTA has arrived at the current barrier
    uint32_t nctas;
    __device__ __forceinline__ void arrive() {
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    }
    __device__ __forceinline__ void wait() {
        if (threadIdx.x == 0) {
            uint32_t v, spins = 0;
            do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
                if (++spins > (1u << 28)) __trap();              // a protocol bug traps instead of hanging the GPU
            } while ((int32_t)(v - target) < 0);
        }
        __syncthreads();
        target += nctas;
    }
    __device__ __forceinline__ void sync() { arrive(); wait(); }
};

template <int VW> struct CgVec;
template <> struct CgVec<4> { static __device__ __forceinline__ void load(Vec<4>& o, const float* p) { float4 t = __ldcg(reinterpret_cast<const float4*>(p)); o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w; } };
template <> struct CgVec<2> { static __device__ __forceinline__ void load(Vec<2>& o, const float* p) { float2 t = __ldcg(reinterpret_cast<const float2*>(p)); o.v[0] = t.x; o.v[1] = t.y; } };
template <> struct CgVec<1> { static __device__ __forceinline__ void load(Vec<1>& o, const float* p) { o.v[0] = __ldcg(p); } };

template <int VW, int NCH> struct Row {
    Vec<VW> c[NCH];
    __device__ __forceinline__ void load_cg(const float* p, int d, int lane) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int off = (k * 32 + lane) * VW;
            if (off < d) CgVec<VW>::load(c[k], p + off);
            else {
#pragma unroll
                for (int t = 0; t < VW; ++t) c[k].v[t] = 0.f;
            }
        }
    }
};

// updater of one row: summed gradient (L2) -> optimiser -> row + slot written, accumulator re-zeroed
template <int VW, int NCH>
__device__ __forceinline__ void persist_apply(const tkr_bpr_cfg& cfg, Row<VW, NCH>& var, Row<VW, NCH>& ms, float* __restrict__ pvar,
                                              float* __restrict__ pms, float* __restrict__ G, int d, int lane, bool rms) {
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int off = (k * 32 + lane) * VW;
        if (off < d) {
            Vec<VW> g, z;
            CgVec<VW>::load(g, G + off);
#pragma unroll
            for (int t = 0; t < VW; ++t) { opt_update(cfg, g.v[t], var.c[k].v[t], ms.c[k].v[t]); z.v[t] = 0.f; }
            var.c[k].store(pvar + off);
            if (rms) ms.c[k].store(pms + off);
            z.store(G + off);
        }
    }
}

// same with the summed gradient already in registers
template <int VW, int NCH>
__device__ __forceinline__ void persist_apply_loaded(const tkr_bpr_cfg& cfg, Row<VW, NCH>& var, Row<VW, NCH>& ms, const Row<VW, NCH>& g,
                                                     float* __restrict__ pvar, float* __restrict__ pms, float* __restrict__ G, int d, int lane, bool rms) {
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int off = (k * 32 + lane) * VW;
        if (off < d) {
            Vec<VW> z;
#pragma unroll
            for (int t = 0; t < VW; ++t) { opt_update(cfg, g.c[k].v[t], var.c[k].v[t], ms.c[k].v[t]); z.v[t] = 0.f; }
            var.c[k].store(pvar + off);
            if (rms) ms.c[k].store(pms + off);
            z.store(G + off);
        }
    }
}

__device__ __forceinline__ void persist_apply_bias(const tkr_bpr_cfg& cfg, float* __restrict__ b, float* __restrict__ msb, const StepWs& ws,
                                                   int r, float bv, bool rms) {
    float g = __ldcg(ws.Gb + r), m = rms ? __ldcg(msb + r) : 0.f;
    opt_update(cfg, g, bv, m);
    b[r] = bv;
    if (rms) msb[r] = m;
    ws.Gb[r] = 0.f;
    ws.cntV[r] = 0;
}

// one triple, whole warp: x, s = sigma(-x), per-occurrence gradients into the accumulators, loss term (App. A.1-A.2)
template <int VW, int NCH, bool L1>
__device__ __forceinline__ float persist_grad(const tkr_bpr_cfg& cfg, const Row<VW, NCH>& ru, const Row<VW, NCH>& ri, const Row<VW, NCH>& rj,
                                              float bi, float bj, int u, int i, int j, const StepWs& ws, int d, int lane, bool want_reg, float& x_out, float& reg_out) {
    constexpr unsigned FULL = 0xffffffffu;
    float x = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int q = 0; q < VW; ++q) x = fmaf(ru.c[k].v[q], ri.c[k].v[q] - rj.c[k].v[q], x);
    x = warp_sum(x) + __shfl_sync(FULL, bi - bj, 0);
    const float s = __fdividef(1.0f, 1.0f + __expf(x));
    float* gu = ws.GU + (int64_t)u * d;
    float* gi = ws.GV + (int64_t)i * d;
    float* gj = ws.GV + (int64_t)j * d;
    float reg = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int off = (k * 32 + lane) * VW;
        if (off < d) {
            Vec<VW> a, p, q;
#pragma unroll
            for (int e = 0; e < VW; ++e) {
                a.v[e] = fmaf(-s, ri.c[k].v[e] - rj.c[k].v[e], reg_grad<L1>(ru.c[k].v[e], cfg.lambda_u));
                p.v[e] = fmaf(-s, ru.c[k].v[e], reg_grad<L1>(ri.c[k].v[e], cfg.lambda_i));
                q.v[e] = fmaf(s, ru.c[k].v[e], reg_grad<L1>(rj.c[k].v[e], cfg.lambda_j));
                if (want_reg) reg += reg_val<L1>(ru.c[k].v[e], cfg.lambda_u) + reg_val<L1>(ri.c[k].v[e], cfg.lambda_i) + reg_val<L1>(rj.c[k].v[e], cfg.lambda_j);
            }
            a.red_add(gu + off); p.red_add(gi + off); q.red_add(gj + off);
        }
    }
    if (lane == 0) {
        atomicAdd(ws.Gb + i, -s + reg_grad<L1>(bi, cfg.lambda_b));
        atomicAdd(ws.Gb + j, s + reg_grad<L1>(bj, cfg.lambda_b));
    }
    x_out = x;
    reg_out = reg;             // per-lane partial of the regulariser value (summed later, off the critical path)
    return s;
}

// the triple's term of the batch objective (bpr.py:92-99), added to the CTA's shared word; called between the arrive and
// the wait of barrier 1 so that its two SFU chains and its shuffle tree overlap the barrier instead of preceding it
template <bool L1>
__device__ __forceinline__ void persist_loss(const tkr_bpr_cfg& cfg, float x, float reg, float bi, float bj, int lane, float* cta_loss) {
    reg = warp_sum(reg);
    if (lane == 0)
        atomicAdd(cta_loss, reg + fmaxf(-x, 0.f) + __logf(1.0f + __expf(-fabsf(x))) + reg_val<L1>(bi, cfg.lambda_b) + reg_val<L1>(bj, cfg.lambda_b));
}

template <int VW, int NCH, bool L1, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
bpr_persist_kernel(tkr_bpr_cfg cfg, float* __restrict__ U, float* __restrict__ V, float* __restrict__ b, float* __restrict__ msU,
                   float* __restrict__ msV, float* __restrict__ msb, const int32_t* __restrict__ ub, const int32_t* __restrict__ ib,
                   const int32_t* __restrict__ jb, int B, int n_steps, StepWs ws, float* __restrict__ loss_out, long long* __restrict__ dbg) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
    const int d = cfg.d;
    const bool rms = cfg.optimizer == TKR_OPT_RMSPROP;
    const bool want_loss = loss_out != nullptr;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ float cta_loss;                               // this CTA's share of the step's objective
    if (threadIdx.x == 0) cta_loss = 0.f;
    // barrier counter: ws.sync[0]; its value when this launch started was left in ws.sync[1] by the previous launch
    GridBarrier gb;
    gb.counter = ws.sync;
    gb.nctas = gridDim.x;
    gb.target = __ldcg(ws.sync + 1) + gridDim.x;
    __syncthreads();
    // Ordering across the cluster comes from the barrier itself (arrive.release / wait.acquire at cluster scope; the
    // whole grid is one cluster), not from gpu-scope fences: a MEMBAR.GPU in front of each barrier cost more than the
    // rest of the step.
    // ids of the next step are fetched one step ahead (they are immutable; a first read comes from DRAM)
    int un = 0, in_ = 0, jn = 0;
    if (B <= nwarps && warp < B && n_steps > 0) { un = __ldg(ub + warp); in_ = __ldg(ib + warp); jn = __ldg(jb + warp); }

    for (int step = 0; step < n_steps; ++step) {
        const int32_t* us = ub + (int64_t)step * B;
        const int32_t* is = ib + (int64_t)step * B;
        const int32_t* js = jb + (int64_t)step * B;
        if (B <= nwarps) {
            // ---- the common case (one triple per warp): the triple's rows stay in registers across the barrier
            const bool active = warp < B;                    // warp-uniform; idle warps only take part in the barriers
            const bool prof = dbg != nullptr && warp == 0 && lane == 0;
            long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
            if (prof) c0 = clock64();
            int u = 0, i = 0, j = 0, claim = 0;              // claim bit 0/1/2: this warp updates the u / i / j row after the barrier
            float bi = 0.f, bj = 0.f, x_t = 0.f, reg_t = 0.f;
            Row<VW, NCH> ru, ri, rj, mu, mi, mj;
            if (active) {
                u = un; i = in_; j = jn;
                if (step + 1 < n_steps) { un = __ldg(us + B + warp); in_ = __ldg(is + B + warp); jn = __ldg(js + B + warp); }
                ru.load_cg(U + (int64_t)u * d, d, lane);
                ri.load_cg(V + (int64_t)i * d, d, lane);
                rj.load_cg(V + (int64_t)j * d, d, lane);
                if (lane == 0) {
                    bi = __ldcg(b + i); bj = __ldcg(b + j);
                    claim = (int)(atomicAdd(ws.cntU + u, 1) == 0);
                    claim |= (int)(atomicAdd(ws.cntV + i, 1) == 0) << 1;
                    claim |= (int)(atomicAdd(ws.cntV + j, 1) == 0) << 2;
                }
                persist_grad<VW, NCH, L1>(cfg, ru, ri, rj, bi, bj, u, i, j, ws, d, lane, want_loss, x_t, reg_t);
                if (prof) c1 = clock64();
                claim = __shfl_sync(FULL, claim, 0);
            }
            gb.arrive();                                     // barrier 1 (arrive): this CTA's gradients are performed at L2
            if (active && want_loss) persist_loss<L1>(cfg, x_t, reg_t, bi, bj, lane, &cta_loss);
            if (active && rms) {   // slots of the rows this warp will update: fetched while the cluster synchronises
                if (claim & 1) mu.load_cg(msU + (int64_t)u * d, d, lane);
                if (claim & 2) mi.load_cg(msV + (int64_t)i * d, d, lane);
                if (claim & 4) mj.load_cg(msV + (int64_t)j * d, d, lane);
            }
            if (prof) c2 = clock64();
            gb.wait();                                       // barrier 1 (wait): every gradient of the batch is in the accumulators
            if (prof) c3 = clock64();
            if (want_loss && threadIdx.x == 0) { atomicAdd(loss_out + step, cta_loss); cta_loss = 0.f; }   // (next write: after barrier 2)
            if (active) {
                // every load of the update phase goes out before the first result is used: up to three summed-gradient
                // rows and the two bias words are one L2 round trip instead of three to five in a row
                Row<VW, NCH> gu_, gi_, gj_;
                float gbi = 0.f, gbj = 0.f, mbi = 0.f, mbj = 0.f;
                if (claim & 1) gu_.load_cg(ws.GU + (int64_t)u * d, d, lane);
                if (claim & 2) gi_.load_cg(ws.GV + (int64_t)i * d, d, lane);
                if (claim & 4) gj_.load_cg(ws.GV + (int64_t)j * d, d, lane);
                if (lane == 0) {
                    if (claim & 2) { gbi = __ldcg(ws.Gb + i); if (rms) mbi = __ldcg(msb + i); }
                    if (claim & 4) { gbj = __ldcg(ws.Gb + j); if (rms) mbj = __ldcg(msb + j); }
                }
                if (claim & 1) {
                    persist_apply_loaded<VW, NCH>(cfg, ru, mu, gu_, U + (int64_t)u * d, msU + (int64_t)u * d, ws.GU + (int64_t)u * d, d, lane, rms);
                    if (lane == 0) ws.cntU[u] = 0;
                }
                if (claim & 2) {
                    persist_apply_loaded<VW, NCH>(cfg, ri, mi, gi_, V + (int64_t)i * d, msV + (int64_t)i * d, ws.GV + (int64_t)i * d, d, lane, rms);
                    if (lane == 0) { opt_update(cfg, gbi, bi, mbi); b[i] = bi; if (rms) msb[i] = mbi; ws.Gb[i] = 0.f; ws.cntV[i] = 0; }
                }
                if (claim & 4) {                             // (i == j: the second claim on the row failed, bit 2 is clear)
                    persist_apply_loaded<VW, NCH>(cfg, rj, mj, gj_, V + (int64_t)j * d, msV + (int64_t)j * d, ws.GV + (int64_t)j * d, d, lane, rms);
                    if (lane == 0) { opt_update(cfg, gbj, bj, mbj); b[j] = bj; if (rms) msb[j] = mbj; ws.Gb[j] = 0.f; ws.cntV[j] = 0; }
                }
            }
            if (prof) c4 = clock64();
            gb.sync();                                       // barrier 2: every row of the batch is updated
            if (prof) { const long long c5 = clock64(); dbg[0] += c1 - c0; dbg[1] += c2 - c1; dbg[2] += c3 - c2; dbg[3] += c4 - c3; dbg[4] += c5 - c4; dbg[5] += 1; }
        } else {
            // ---- 256 < B <= 1024: a warp takes several triples; first touchers append their rows to the step's lists
            for (int n = warp; n < B; n += nwarps) {
                const int u = __ldg(us + n), i = __ldg(is + n), j = __ldg(js + n);
                Row<VW, NCH> ru, ri, rj;
                ru.load_cg(U + (int64_t)u * d, d, lane);
                ri.load_cg(V + (int64_t)i * d, d, lane);
                rj.load_cg(V + (int64_t)j * d, d, lane);
                float bi = 0.f, bj = 0.f;
                if (lane == 0) {
                    bi = __ldcg(b + i); bj = __ldcg(b + j);
                    if (atomicAdd(ws.cntU + u, 1) == 0) ws.listU[atomicAdd(ws.n_touched + 0, 1)] = u;
                    if (atomicAdd(ws.cntV + i, 1) == 0) ws.listV[atomicAdd(ws.n_touched + 1, 1)] = i;
                    if (atomicAdd(ws.cntV + j, 1) == 0) ws.listV[atomicAdd(ws.n_touched + 1, 1)] = j;
                }
                float x_t, reg_t;
                persist_grad<VW, NCH, L1>(cfg, ru, ri, rj, bi, bj, u, i, j, ws, d, lane, want_loss, x_t, reg_t);
                if (want_loss) persist_loss<L1>(cfg, x_t, reg_t, bi, bj, lane, &cta_loss);
            }
            gb.sync();
            if (want_loss && threadIdx.x == 0) { atomicAdd(loss_out + step, cta_loss); cta_loss = 0.f; }
            const int nU = __ldcg(ws.n_touched + 0), nV = __ldcg(ws.n_touched + 1);
            for (int w = warp; w < nU + nV; w += nwarps) {
                const bool user = w < nU;
                const int r = user ? __ldcg(ws.listU + w) : __ldcg(ws.listV + (w - nU));
                float* pvar = (user ? U : V) + (int64_t)r * d;
                float* pms = (user ? msU : msV) + (int64_t)r * d;
                Row<VW, NCH> var, ms;
                var.load_cg(pvar, d, lane);
                if (rms) ms.load_cg(pms, d, lane);
                persist_apply<VW, NCH>(cfg, var, ms, pvar, pms, (user ? ws.GU : ws.GV) + (int64_t)r * d, d, lane, rms);
                if (lane == 0) {
                    if (user) ws.cntU[r] = 0;
                    else persist_apply_bias(cfg, b, msb, ws, r, __ldcg(b + r), rms);
                }
            }
            gb.sync();                                       // every row updated, every warp has read the list lengths
            if (warp == 0 && lane == 0) { ws.n_touched[0] = 0; ws.n_touched[1] = 0; }
            gb.sync();                                       // ... and they are re-armed before anyone appends again
        }
    }
    // every CTA has passed the last barrier of this launch; the next launch starts counting from here
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.sync[1] = gb.target - gb.nctas;
}

int g_persist_mode = -1;
long long* g_persist_dbg = nullptr;   // device int64[8]: cycles per phase of warp 0 (tkr_debug_set_persist_counters)   // -1 auto (default), 0 never, 1 whenever legal (tkr_debug_set_persist_mode)

bool bpr_persist_legal(const tkr_bpr_cfg* cfg, int64_t B) {
    const int d = cfg->d;
    const int vw = d % 4 == 0 ? 4 : d % 2 == 0 ? 2 : 1;
    return B <= 1024 && d <= 64 * vw;       // NCH <= 2: rows of a triple + their slots stay in registers at 64 registers per thread
}

struct ClusterLaunch {
    cudaLaunchConfig_t lc;
    cudaLaunchAttribute at[1];
    ClusterLaunch(int ctas, int threads, cudaStream_t st) : lc{} {
        lc.gridDim = dim3((unsigned)ctas, 1, 1); lc.blockDim = dim3((unsigned)threads, 1, 1); lc.dynamicSmemBytes = 0; lc.stream = st;
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
    }
};

template <int VW, int NCH, bool L1>
static int launch_persist_l1(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                             const int32_t* i, const int32_t* j, int B, int n_steps, const StepWs& ws, float* loss, cudaStream_t st) {
    auto wide = bpr_persist_kernel<VW, NCH, L1, 512>;     // 16 CTAs x 512 threads
    auto tall = bpr_persist_kernel<VW, NCH, L1, 1024>;    //  8 CTAs x 1024 threads
    static int use_wide = -1;                              // per instantiation; decided once (benign race: same answer)
    if (use_wide < 0) {
        int n = 0;
        ClusterLaunch probe(kPersistWarps / 16, 512, st);
        bool ok = cudaFuncSetAttribute(wide, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                  cudaOccupancyMaxActiveClusters(&n, wide, &probe.lc) == cudaSuccess && n >= 1;
        (void)cudaGetLastError();
        use_wide = ok ? 1 : 0;
    }
    ClusterLaunch cl(use_wide ? kPersistWarps / 16 : kPersistWarps / 32, use_wide ? 512 : 1024, st);
    cudaError_t e = cudaLaunchKernelEx(&cl.lc, use_wide ? wide : tall, *cfg, U, V, b, msU, msV, msb, u, i, j, B, n_steps, ws, loss, g_persist_dbg);
    if (e != cudaSuccess) { set_error("persistent step kernel launch failed: %s", cudaGetErrorString(e)); return TKR_ERR_CUDA; }
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

template <int VW, int NCH>
static int launch_persist(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                          const int32_t* i, const int32_t* j, int B, int n_steps, const StepWs& ws, float* loss, cudaStream_t st) {
    if (cfg->l1) return launch_persist_l1<VW, NCH, true>(cfg, U, V, b, msU, msV, msb, u, i, j, B, n_steps, ws, loss, st);
    return launch_persist_l1<VW, NCH, false>(cfg, U, V, b, msU, msV, msb, u, i, j, B, n_steps, ws, loss, st);
}

// n_steps consecutive steps of B explicit triples each (u/i/j hold n_steps * B ids), one launch
int bpr_persist_steps(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                      const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const StepWs& ws, float* loss, cudaStream_t st) {
    const int d = cfg->d;
    const int vw = d % 4 == 0 ? 4 : d % 2 == 0 ? 2 : 1;
    const int nch = (d + 32 * vw - 1) / (32 * vw);
#define TKR_P(VW, NCH) return launch_persist<VW, NCH>(cfg, U, V, b, msU, msV, msb, u, i, j, (int)B, (int)n_steps, ws, loss, st)
    if (vw == 4) { if (nch == 1) TKR_P(4, 1); else TKR_P(4, 2); }
    if (vw == 2) { if (nch == 1) TKR_P(2, 1); else TKR_P(2, 2); }
    if (nch == 1) TKR_P(1, 1); else TKR_P(1, 2);
#undef TKR_P
}


// ---------------------------------------------------------------------------------------------------------------------
// Dataflow variant: the same synchronous-step semantics WITHOUT grid-wide barriers.  A synchronous step only orders work
// through the rows it touches: an occurrence of row r in step t must read r as the last EARLIER step touching r left it,
// and r's single update of step t must see every occurrence of r in step t.  Both orders are local to the row:
//   * a pre-pass sorts the chunk's 3 B S row references by (row, step) and gives every reference the previous step that
//     touches its row (prev) and the number of occurrences of the row in its own step (n_occ);
//   * warps take triples in (step, index) order from a ticket counter.  A triple waits until the version words of its three
//     rows have reached prev (usually long since), gathers, adds its gradients to the rows' accumulators, fences, and
//     arrives on the rows' counters; the LAST arriver of a row (count == n_occ) -- which still holds the pre-step row in
//     registers -- applies the optimiser, re-zeroes accumulator and counter, fences, and publishes version = step.
// Steps overlap as far as their rows allow; what serialises is the chain of a row touched in consecutive steps (the most
// popular items).  A triple only ever waits for smaller tickets, which running warps hold: no residency assumption.
// (kFlowTriples, bpr_internal.cuh: triples per chunk; the staged sampler draws kStageTriples ahead, explicit triples come 4x as many)
constexpr size_t kFlowTmpBytes = (size_t)4 << 20;             // cub radix-sort scratch (checked at run time)

struct FlowWs {
    uint32_t* ctl;        // [0] ticket counter, [1] global step number of the chunk's first step, [2] steps of the running chunk
    int32_t* ver;         // [n_users + n_items] global step number + 1 of the row's last update by this kernel
    uint64_t* keys_a; uint64_t* keys_b;
    int2* meta;           // [3 * triples] (prev step within the chunk or -1, occurrences of the row in the triple's step)
    void* tmp;
};
size_t bpr_flow_bytes(const tkr_bpr_cfg* cfg) {
    return 256 + align_up(((size_t)cfg->n_users + cfg->n_items) * 4, 256) + 3 * align_up((size_t)3 * kFlowTriples * 8, 256) + kFlowTmpBytes;
}
static FlowWs flow_carve(const tkr_bpr_cfg* cfg, char* p) {
    FlowWs f;
    f.ctl = (uint32_t*)p; p += 256;
    f.ver = (int32_t*)p; p += align_up(((size_t)cfg->n_users + cfg->n_items) * 4, 256);
    f.keys_a = (uint64_t*)p; p += align_up((size_t)3 * kFlowTriples * 8, 256);
    f.keys_b = (uint64_t*)p; p += align_up((size_t)3 * kFlowTriples * 8, 256);
    f.meta = (int2*)p; p += align_up((size_t)3 * kFlowTriples * 8, 256);
    f.tmp = p;
    return f;
}

__global__ void flow_begin_kernel(uint32_t* ctl, uint32_t n_steps) { ctl[1] += ctl[2]; ctl[2] = n_steps; ctl[0] = 0; }

// key = row (users first, then items) : step : reference (3 * index + slot)
__global__ void __launch_bounds__(256) flow_keys_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ i, const int32_t* __restrict__ j,
                                                        int B, int64_t n_refs, uint32_t nu, uint64_t* __restrict__ keys) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= n_refs) return;
    const int64_t n = e / 3;
    const int s = (int)(e - n * 3);
    const uint32_t row = s == 0 ? (uint32_t)u[n] : nu + (uint32_t)(s == 1 ? i[n] : j[n]);
    const uint64_t t = (uint64_t)(n / B), w = (uint64_t)(n % B);
    keys[e] = ((uint64_t)row << 28) | (t << 12) | (w * 3 + (uint64_t)s);
}
__global__ void __launch_bounds__(256) flow_meta_kernel(const uint64_t* __restrict__ keys, int64_t n_refs, int B, int2* __restrict__ meta) {
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= n_refs) return;
    const uint64_t k = keys[p], grp = k >> 12;                // (row, step)
    int64_t lo = p, hi = p;
    while (lo > 0 && (keys[lo - 1] >> 12) == grp) --lo;
    while (hi + 1 < n_refs && (keys[hi + 1] >> 12) == grp) ++hi;
    int prev = -1;
    if (lo > 0 && (keys[lo - 1] >> 28) == (k >> 28)) prev = (int)((keys[lo - 1] >> 12) & 0xffffu);
    const int64_t t = (int64_t)((k >> 12) & 0xffffu), ws = (int64_t)(k & 0xfffu);
    meta[t * 3 * B + ws] = make_int2(prev, (int)(hi - lo + 1));
}

// lanes 0..2 poll the version words of the triple's three rows at once (relaxed loads: the rows are read with ld.cg, issued
// after the poll has seen the version, and the writer fenced between its stores and the version -- an acquire here would only
// add an L1 invalidation per poll); the warp goes on when all three have arrived
__device__ __forceinline__ void flow_wait3(const int32_t* p, uint32_t need, bool waits) {
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t spins = 0;
    bool ok = !waits;
    while (true) {
        if (!ok) {
            int32_t v;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            ok = (int32_t)((uint32_t)v - need) >= 0;          // (step numbers wrap after 2^32 steps: compare differences)
        }
        if (__all_sync(FULL, ok)) break;
        if (++spins > (1u << 22)) __trap();                   // a protocol bug traps instead of hanging the GPU
    }
}
__device__ __forceinline__ void flow_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// One triple, whole warp, for the dataflow kernel: as persist_grad, but a row that occurs ONCE in its step (`direct` bit 0 / 1 / 2
// for the u / i / j row; the pre-pass knows) skips the accumulator round trip altogether -- its gradient is complete in this
// warp's registers, so the optimiser is applied and the row (and its slot) stored right here: no red, no counter, no re-read.
template <int VW, int NCH, bool L1>
__device__ __forceinline__ void flow_grad(const tkr_bpr_cfg& cfg, Row<VW, NCH>& ru, Row<VW, NCH>& ri, Row<VW, NCH>& rj, Row<VW, NCH>& mu, Row<VW, NCH>& mi,
                                          Row<VW, NCH>& mj, float bi, float bj, float mbi, float mbj, int u, int i, int j, const StepWs& ws, float* __restrict__ U,
                                          float* __restrict__ V, float* __restrict__ b, float* __restrict__ msU, float* __restrict__ msV, float* __restrict__ msb,
                                          int d, int lane, int direct, bool rms, bool want_reg, float& x_out, float& reg_out) {
    constexpr unsigned FULL = 0xffffffffu;
    float x = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int q = 0; q < VW; ++q) x = fmaf(ru.c[k].v[q], ri.c[k].v[q] - rj.c[k].v[q], x);
    x = warp_sum(x) + __shfl_sync(FULL, bi - bj, 0);
    const float s = __fdividef(1.0f, 1.0f + __expf(x));
    float reg = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int off = (k * 32 + lane) * VW;
        if (off < d) {
            Vec<VW> a, p, q;
#pragma unroll
            for (int e = 0; e < VW; ++e) {
                a.v[e] = fmaf(-s, ri.c[k].v[e] - rj.c[k].v[e], reg_grad<L1>(ru.c[k].v[e], cfg.lambda_u));
                p.v[e] = fmaf(-s, ru.c[k].v[e], reg_grad<L1>(ri.c[k].v[e], cfg.lambda_i));
                q.v[e] = fmaf(s, ru.c[k].v[e], reg_grad<L1>(rj.c[k].v[e], cfg.lambda_j));
                if (want_reg) reg += reg_val<L1>(ru.c[k].v[e], cfg.lambda_u) + reg_val<L1>(ri.c[k].v[e], cfg.lambda_i) + reg_val<L1>(rj.c[k].v[e], cfg.lambda_j);
            }
            if (direct & 1) {
#pragma unroll
                for (int e = 0; e < VW; ++e) opt_update(cfg, a.v[e], ru.c[k].v[e], mu.c[k].v[e]);
                ru.c[k].store(U + (int64_t)u * d + off);
                if (rms) mu.c[k].store(msU + (int64_t)u * d + off);
            } else a.red_add(ws.GU + (int64_t)u * d + off);
            if (direct & 2) {
#pragma unroll
                for (int e = 0; e < VW; ++e) opt_update(cfg, p.v[e], ri.c[k].v[e], mi.c[k].v[e]);
                ri.c[k].store(V + (int64_t)i * d + off);
                if (rms) mi.c[k].store(msV + (int64_t)i * d + off);
            } else p.red_add(ws.GV + (int64_t)i * d + off);
            if (direct & 4) {
#pragma unroll
                for (int e = 0; e < VW; ++e) opt_update(cfg, q.v[e], rj.c[k].v[e], mj.c[k].v[e]);
                rj.c[k].store(V + (int64_t)j * d + off);
                if (rms) mj.c[k].store(msV + (int64_t)j * d + off);
            } else q.red_add(ws.GV + (int64_t)j * d + off);
        }
    }
    if (lane == 0) {
        const float gbi = -s + reg_grad<L1>(bi, cfg.lambda_b), gbj = s + reg_grad<L1>(bj, cfg.lambda_b);
        if (direct & 2) { opt_update(cfg, gbi, bi, mbi); b[i] = bi; if (rms) msb[i] = mbi; } else atomicAdd(ws.Gb + i, gbi);
        if (direct & 4) { opt_update(cfg, gbj, bj, mbj); b[j] = bj; if (rms) msb[j] = mbj; } else atomicAdd(ws.Gb + j, gbj);
    }
    x_out = x;
    reg_out = reg;
}

template <int VW, int NCH, bool L1>
__global__ void __launch_bounds__(256, 2)
bpr_flow_kernel(tkr_bpr_cfg cfg, float* __restrict__ U, float* __restrict__ V, float* __restrict__ b, float* __restrict__ msU,
                float* __restrict__ msV, float* __restrict__ msb, const int32_t* __restrict__ ub, const int32_t* __restrict__ ib,
                const int32_t* __restrict__ jb, int B, int n_steps, StepWs ws, FlowWs fw, float* __restrict__ loss_out) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int d = cfg.d;
    const bool rms = cfg.optimizer == TKR_OPT_RMSPROP;
    const bool want_loss = loss_out != nullptr;
    const int total = n_steps * B;
    const uint32_t base = __ldcg(fw.ctl + 1);
    int32_t* const verU = fw.ver;
    int32_t* const verV = fw.ver + cfg.n_users;
    int n_next = 0;
    if (lane == 0) n_next = (int)atomicAdd(fw.ctl, 1u);
    n_next = __shfl_sync(FULL, n_next, 0);
    while (n_next < total) {
        const int n = n_next;
        const int u = __ldg(ub + n), i = __ldg(ib + n), j = __ldg(jb + n);
        const int2 du = __ldg(fw.meta + (int64_t)n * 3), di = __ldg(fw.meta + (int64_t)n * 3 + 1), dj = __ldg(fw.meta + (int64_t)n * 3 + 2);
        if (lane == 0) n_next = (int)atomicAdd(fw.ctl, 1u);  // the next ticket's round trip hides behind this triple
        const int t = n / B;
        // the rows as the last earlier step touching them left them: lane 0 / 1 / 2 <-> the u / i / j row
        const int2 dm = lane == 0 ? du : lane == 1 ? di : dj;
        int32_t* const my_ver = lane == 0 ? verU + u : verV + (lane == 1 ? i : j);
        int32_t* const my_cnt = lane == 0 ? ws.cntU + u : ws.cntV + (lane == 1 ? i : j);
        flow_wait3(my_ver, base + (uint32_t)(dm.x + 1), lane < 3 && dm.x >= 0);
        // rows that occur once in their step are updated in place by this warp (bit 0 / 1 / 2 <-> u / i / j)
        const int direct = (int)(du.y == 1) | (int)(di.y == 1) << 1 | (int)(dj.y == 1) << 2;
        Row<VW, NCH> ru, ri, rj, mu, mi, mj;
        ru.load_cg(U + (int64_t)u * d, d, lane);
        ri.load_cg(V + (int64_t)i * d, d, lane);
        rj.load_cg(V + (int64_t)j * d, d, lane);
        if (rms) {
            if (direct & 1) mu.load_cg(msU + (int64_t)u * d, d, lane);
            if (direct & 2) mi.load_cg(msV + (int64_t)i * d, d, lane);
            if (direct & 4) mj.load_cg(msV + (int64_t)j * d, d, lane);
        }
        float bi = 0.f, bj = 0.f, mbi = 0.f, mbj = 0.f, x_t = 0.f, reg_t = 0.f;
        if (lane == 0) {
            bi = __ldcg(b + i); bj = __ldcg(b + j);
            if (rms && (direct & 2)) mbi = __ldcg(msb + i);
            if (rms && (direct & 4)) mbj = __ldcg(msb + j);
        }
        flow_grad<VW, NCH, L1>(cfg, ru, ri, rj, mu, mi, mj, bi, bj, mbi, mbj, u, i, j, ws, U, V, b, msU, msV, msb, d, lane, direct, rms, want_loss, x_t, reg_t);
        if (want_loss) persist_loss<L1>(cfg, x_t, reg_t, bi, bj, lane, loss_out + t);
        flow_fence();                                         // this warp's gradient contributions / in-place updates are performed ...
        __syncwarp();
        const bool is_direct = lane < 3 && ((direct >> lane) & 1);
        if (is_direct) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_ver), "r"(base + (uint32_t)t + 1u) : "memory");   // ... before their versions
        bool last = false;                                    // ... and before it arrives on the other rows' counters (one round trip;
        if (lane < 3 && !is_direct) last = atomicAdd(my_cnt, 1) == dm.y - 1;   //  i == j: two arrivals on one row, either can be the last)
        const int claim = (int)(__ballot_sync(FULL, last) & 7u);
        if (claim) {
            flow_fence();                                     // every occurrence's contribution is visible to the last arriver
            Row<VW, NCH> gu_, gi_, gj_;
            float gbi = 0.f, gbj = 0.f;
            if (rms) {
                if (claim & 1) mu.load_cg(msU + (int64_t)u * d, d, lane);
                if (claim & 2) mi.load_cg(msV + (int64_t)i * d, d, lane);
                if (claim & 4) mj.load_cg(msV + (int64_t)j * d, d, lane);
            }
            if (claim & 1) gu_.load_cg(ws.GU + (int64_t)u * d, d, lane);
            if (claim & 2) gi_.load_cg(ws.GV + (int64_t)i * d, d, lane);
            if (claim & 4) gj_.load_cg(ws.GV + (int64_t)j * d, d, lane);
            if (lane == 0) {
                if (claim & 2) { gbi = __ldcg(ws.Gb + i); if (rms) mbi = __ldcg(msb + i); }
                if (claim & 4) { gbj = __ldcg(ws.Gb + j); if (rms) mbj = __ldcg(msb + j); }
            }
            if (claim & 1) {
                persist_apply_loaded<VW, NCH>(cfg, ru, mu, gu_, U + (int64_t)u * d, msU + (int64_t)u * d, ws.GU + (int64_t)u * d, d, lane, rms);
                if (lane == 0) ws.cntU[u] = 0;
            }
            if (claim & 2) {
                persist_apply_loaded<VW, NCH>(cfg, ri, mi, gi_, V + (int64_t)i * d, msV + (int64_t)i * d, ws.GV + (int64_t)i * d, d, lane, rms);
                if (lane == 0) { opt_update(cfg, gbi, bi, mbi); b[i] = bi; if (rms) msb[i] = mbi; ws.Gb[i] = 0.f; ws.cntV[i] = 0; }
            }
            if (claim & 4) {
                persist_apply_loaded<VW, NCH>(cfg, rj, mj, gj_, V + (int64_t)j * d, msV + (int64_t)j * d, ws.GV + (int64_t)j * d, d, lane, rms);
                if (lane == 0) { opt_update(cfg, gbj, bj, mbj); b[j] = bj; if (rms) msb[j] = mbj; ws.Gb[j] = 0.f; ws.cntV[j] = 0; }
            }
            flow_fence();                                     // rows, slots, re-zeroed accumulators and counters first, then the version
            __syncwarp();
            if (lane < 3 && ((claim >> lane) & 1)) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_ver), "r"(base + (uint32_t)t + 1u) : "memory");
        }
        n_next = __shfl_sync(FULL, n_next, 0);
    }
}

template <int VW, int NCH, bool L1>
static int launch_flow_l1(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                          const int32_t* i, const int32_t* j, int B, int n_steps, const StepWs& ws, const FlowWs& fw, float* loss, cudaStream_t st) {
    auto kern = bpr_flow_kernel<VW, NCH, L1>;
    static int per_sm = -1;                                    // per instantiation; decided once (benign race: same answer)
    if (per_sm < 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 256, 0) != cudaSuccess || n < 1) n = 1;
        (void)cudaGetLastError();
        per_sm = n;
    }
    int64_t blocks = (int64_t)kNumSMs * per_sm;
    const int64_t need = ((int64_t)n_steps * B + 7) / 8;
    if (blocks > need) blocks = need;
    kern<<<(unsigned)blocks, 256, 0, st>>>(*cfg, U, V, b, msU, msV, msb, u, i, j, B, n_steps, ws, fw, loss);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}
template <int VW, int NCH>
static int launch_flow(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                       const int32_t* i, const int32_t* j, int B, int n_steps, const StepWs& ws, const FlowWs& fw, float* loss, cudaStream_t st) {
    if (cfg->l1) return launch_flow_l1<VW, NCH, true>(cfg, U, V, b, msU, msV, msb, u, i, j, B, n_steps, ws, fw, loss, st);
    return launch_flow_l1<VW, NCH, false>(cfg, U, V, b, msU, msV, msb, u, i, j, B, n_steps, ws, fw, loss, st);
}

// n_steps consecutive steps of B explicit triples each (n_steps * B <= kFlowTriples): pre-pass + one launch
int bpr_flow_steps(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                   const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const StepWs& ws, float* loss, cudaStream_t st) {
    if (n_steps <= 0) return TKR_OK;
    if (ws.flow == nullptr || n_steps * B > kFlowTriples || n_steps > 65536 || 3 * B > 4096) {
        set_error("internal: dataflow step kernel called with %lld steps of %lld triples", (long long)n_steps, (long long)B);
        return TKR_ERR_INVALID;
    }
    const FlowWs fw = flow_carve(cfg, ws.flow);
    const int64_t n_refs = 3 * n_steps * B;
    flow_begin_kernel<<<1, 1, 0, st>>>(fw.ctl, (uint32_t)n_steps);
    flow_keys_kernel<<<(unsigned)((n_refs + 255) / 256), 256, 0, st>>>(u, i, j, (int)B, n_refs, (uint32_t)cfg->n_users, fw.keys_a);
    TKR_LAUNCH_CHECK();
    int row_bits = 1;
    while (((int64_t)1 << row_bits) < (int64_t)cfg->n_users + cfg->n_items) ++row_bits;
    cub::DoubleBuffer<uint64_t> keys(fw.keys_a, fw.keys_b);
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, (int)n_refs, 12, 28 + row_bits, st);
    if (e != cudaSuccess || tmp_bytes > kFlowTmpBytes) { set_error("dataflow step kernel: radix sort scratch of %zu bytes (cuda: %s)", tmp_bytes, cudaGetErrorString(e)); return TKR_ERR_CUDA; }
    e = cub::DeviceRadixSort::SortKeys(fw.tmp, tmp_bytes, keys, (int)n_refs, 12, 28 + row_bits, st);
    if (e != cudaSuccess) { set_error("dataflow step kernel: radix sort failed: %s", cudaGetErrorString(e)); return TKR_ERR_CUDA; }
    flow_meta_kernel<<<(unsigned)((n_refs + 255) / 256), 256, 0, st>>>(keys.Current(), n_refs, (int)B, fw.meta);
    TKR_LAUNCH_CHECK();
    const int d = cfg->d;
    const int vw = d % 4 == 0 ? 4 : d % 2 == 0 ? 2 : 1;
    const int nch = (d + 32 * vw - 1) / (32 * vw);
#define TKR_F(VW, NCH) return launch_flow<VW, NCH>(cfg, U, V, b, msU, msV, msb, u, i, j, (int)B, (int)n_steps, ws, fw, loss, st)
    if (vw == 4) { if (nch == 1) TKR_F(4, 1); else TKR_F(4, 2); }
    if (vw == 2) { if (nch == 1) TKR_F(2, 1); else TKR_F(2, 2); }
    if (nch == 1) TKR_F(1, 1); else TKR_F(1, 2);
#undef TKR_F
}

}  // namespace tkr

extern "C" void tkr_debug_set_persist_counters(long long* dev_buf) { tkr::g_persist_dbg = dev_buf; }
extern "C" void tkr_debug_set_persist_mode(int32_t m) { tkr::g_persist_mode = m < -1 || m > 2 ? -1 : m; }
