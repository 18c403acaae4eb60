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

}  // namespace tkr

extern "C" void tkr_debug_set_persist_counters(long long* dev_buf) { tkr::g_persist_dbg = dev_buf; }
extern "C" void tkr_debug_set_persist_mode(int32_t m) { tkr::g_persist_mode = m < -1 || m > 1 ? -1 : m; }
