// Hot path 1 on several GPUs: data-parallel BPR step whose gradient exchange is fused into the
// optimiser update and runs over peer memory (NVLink loads / stores issued by the kernel itself) instead of an NCCL
// all-reduce between two launches.  See include/topkrec.h (tkr_bpr_dp_step) for the contract.
//
// Why: at C2 the item-gradient region is 5.2 MB per step; through torch + NCCL the all-reduce cost 0.41 ms of a 0.86 ms
// step at 8 GPUs (round 1, SCALE: efficiency 0.53) although the wire time is ~10 us.  Here every rank owns the item rows
// r % world == rank: it pulls those rows of every peer's accumulator (reduce-scatter by loads), updates them once, and
// pushes the new parameter row to every replica (all-gather by stores).  Per rank and step: (world-1)/world of
// [GV|Gb|tch] in, the same amount of V out, two flag round trips.
#include "bpr_device.cuh"
#include "peer.cuh"

namespace tkr {

struct DpLayout { size_t V, b, G[2], flags, total, g_floats; };

static DpLayout dp_layout(const tkr_bpr_cfg* cfg) {
    DpLayout L;
    const size_t d = cfg->d, ni = cfg->n_items;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    L.V = take(ni * d * 4);
    L.b = take(ni * 4);
    L.g_floats = ni * d + 2 * ni;
    L.G[0] = take(L.g_floats * 4);
    L.G[1] = take(L.g_floats * 4);
    L.flags = take(kPeerFlagBytes + 64);       // + local words: [0] "barrier A passed" epoch, [1] finished exchange blocks
    L.total = o;
    return L;
}

struct DpArgs {
    tkr_bpr_cfg cfg;
    PeerView pv;
    size_t off_V, off_b, off_G, off_Gnext, off_flags;
    float* U; float* msU; float* msV; float* msb;
    float* GU; int32_t* cntU;
    uint32_t epoch;
    int xblocks;          // blocks [0, xblocks) exchange item rows, the rest apply user rows and re-zero the other G
    size_t g_floats;
};

template <int VW, int W>
__device__ __forceinline__ void dp_item_row(const DpArgs& a, int r, int lane) {
    const tkr_bpr_cfg& cfg = a.cfg;
    const int d = cfg.d;
    const size_t ni = cfg.n_items;
    // touched by anyone?  (the flags are exact small integers in fp32: their sum is order-independent)
    float tch = 0.f;
#pragma unroll
    for (int p = 0; p < W; ++p)
        if (p < a.pv.world) tch += __ldcv(reinterpret_cast<const float*>(a.pv.base[p] + a.off_G) + ni * d + ni + r);
    if (tch == 0.f) return;
    float* Vloc = reinterpret_cast<float*>(a.pv.base[a.pv.rank] + a.off_V) + (size_t)r * d;
    const bool rms = cfg.optimizer == TKR_OPT_RMSPROP;
    for (int off = lane * VW; off < d; off += 32 * VW) {
        Vec<VW> g[W];
#pragma unroll
        for (int p = 0; p < W; ++p)          // all peer loads in flight before the first add (NVLink latency ~1 us)
            if (p < a.pv.world) g[p].load_cv(reinterpret_cast<const float*>(a.pv.base[p] + a.off_G) + (size_t)r * d + off);
        Vec<VW> v, m;
        v.load(Vloc + off);
        if (rms) m.load(a.msV + (size_t)r * d + off);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
            float s = g[0].v[e];
#pragma unroll
            for (int p = 1; p < W; ++p)
                if (p < a.pv.world) s += g[p].v[e];     // fixed rank order: the sum does not depend on who computes it
            opt_update(cfg, s, v.v[e], m.v[e]);
        }
        if (rms) m.store(a.msV + (size_t)r * d + off);
#pragma unroll
        for (int p = 0; p < W; ++p)          // the new row goes to every replica, own copy included
            if (p < a.pv.world) v.store(reinterpret_cast<float*>(a.pv.base[p] + a.off_V) + (size_t)r * d + off);
    }
    if (lane == 0) {
        float gb = 0.f;
#pragma unroll
        for (int p = 0; p < W; ++p)
            if (p < a.pv.world) gb += __ldcv(reinterpret_cast<const float*>(a.pv.base[p] + a.off_G) + ni * d + r);
        float* bloc = reinterpret_cast<float*>(a.pv.base[a.pv.rank] + a.off_b) + r;
        float bv = *bloc, bm = rms ? a.msb[r] : 0.f;
        opt_update(cfg, gb, bv, bm);
        if (rms) a.msb[r] = bm;
#pragma unroll
        for (int p = 0; p < W; ++p)
            if (p < a.pv.world) reinterpret_cast<float*>(a.pv.base[p] + a.off_b)[r] = bv;
    }
}

template <int VW, int W>
__global__ void __launch_bounds__(256) bpr_dp_exchange_kernel(DpArgs a) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int d = a.cfg.d;
    uint32_t* local = reinterpret_cast<uint32_t*>(a.pv.base[a.pv.rank] + a.off_flags + kPeerFlagBytes);
    if ((int)blockIdx.x >= a.xblocks) {
        // ---- local work, overlapped with the exchange: this rank's user rows + re-zeroing of the other accumulator
        const int ub = blockIdx.x - a.xblocks, nub = gridDim.x - a.xblocks;
        const int64_t w0 = (int64_t)ub * 8 + wib, nw = (int64_t)nub * 8;
        for (int64_t r = w0; r < a.cfg.n_users; r += nw) {
            if (a.cntU[r] == 0) continue;
            apply_row<VW>(a.cfg, a.U + r * d, a.msU + r * d, a.GU + r * d, d, lane);
            __syncwarp();
            if (lane == 0) a.cntU[r] = 0;
        }
        // G[(epoch+1) & 1] was last read by the peers in the step before this one; its barrier has passed
        float4* z = reinterpret_cast<float4*>(a.pv.base[a.pv.rank] + a.off_Gnext);
        const size_t n4 = a.g_floats / 4;
        for (size_t t = (size_t)ub * blockDim.x + threadIdx.x; t < n4; t += (size_t)nub * blockDim.x) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ub == 0 && threadIdx.x < (int)(a.g_floats % 4)) reinterpret_cast<float*>(z)[n4 * 4 + threadIdx.x] = 0.f;
        return;
    }
    // ---- barrier A: every rank's gradient kernel has finished (ours has: this kernel follows it in the stream)
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            peer_signal(a.pv, a.off_flags, 0, a.epoch);
            peer_wait(a.pv, a.off_flags, 0, a.epoch);
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(local), "r"(a.epoch) : "memory");
        } else {
            uint32_t v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory"); } while ((int32_t)(v - a.epoch) < 0);
        }
    }
    __syncthreads();
    // ---- owned item rows: r = rank + world * t
    const int world = a.pv.world;
    const int64_t mine = ((int64_t)a.cfg.n_items - a.pv.rank + world - 1) / world;
    for (int64_t t = (int64_t)blockIdx.x * 8 + wib; t < mine; t += (int64_t)a.xblocks * 8)
        dp_item_row<VW, W>(a, (int)(a.pv.rank + world * t), lane);
    // ---- barrier B: every rank has stored its rows into every replica (closes the step: the next gradient kernel reads V)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(local + 1, 1u) == (uint32_t)a.xblocks - 1) {
            local[1] = 0;
            peer_signal(a.pv, a.off_flags, 1, a.epoch);
            peer_wait(a.pv, a.off_flags, 1, a.epoch);
        }
    }
}

}  // namespace tkr

using namespace tkr;

extern "C" int tkr_bpr_dp_layout(const tkr_bpr_cfg* cfg, int64_t* offsets) {
    if (int rc = bpr_check_cfg(cfg, 1)) return rc;
    TKR_CHECK_ARG(offsets != nullptr, "offsets is NULL");
    const DpLayout L = dp_layout(cfg);
    offsets[TKR_DP_V] = (int64_t)L.V; offsets[TKR_DP_B] = (int64_t)L.b; offsets[TKR_DP_G0] = (int64_t)L.G[0];
    offsets[TKR_DP_G1] = (int64_t)L.G[1]; offsets[TKR_DP_FLAGS] = (int64_t)L.flags; offsets[TKR_DP_TOTAL] = (int64_t)L.total;
    return TKR_OK;
}

extern "C" int tkr_bpr_dp_step(const tkr_bpr_cfg* cfg, float* U, float* msU, float* msV, float* msb, const int32_t* u,
                               const int32_t* i, const int32_t* j, int64_t B, const tkr_sampler* smp, uint64_t first_draw,
                               float* loss_out, void* ws, size_t ws_bytes, const tkr_peers* peers, uint64_t epoch, void* stream) {
    if (int rc = bpr_check_cfg(cfg, B)) return rc;
    TKR_CHECK_ARG(U != nullptr, "U must not be NULL");
    TKR_CHECK_ARG(cfg->optimizer == TKR_OPT_SGD || (msU && msV && msb), "RMSProp needs the msU/msV/msb slots");
    TKR_CHECK_ARG(epoch >= 1, "epoch starts at 1");
    DpArgs a = {};
    if (int rc = peer_view_from(peers, &a.pv)) return rc;
    SamplerDev sd = {};
    if (u == nullptr) {
        if (int rc = bpr_make_sampler(smp, &sd)) return rc;
        TKR_CHECK_ARG(smp->n_items == cfg->n_items, "sampler n_items != cfg n_items");
    } else {
        TKR_CHECK_ARG(i && j, "i, j must not be NULL when u is given");
    }
    StepWs w;
    if (int rc = bpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    const DpLayout L = dp_layout(cfg);
    char* mine = a.pv.base[a.pv.rank];
    const size_t ni = cfg->n_items, d = cfg->d;
    float* V = reinterpret_cast<float*>(mine + L.V);
    float* b = reinterpret_cast<float*>(mine + L.b);
    float* G = reinterpret_cast<float*>(mine + L.G[epoch & 1]);
    w.GV = G; w.Gb = G + ni * d; w.tchV = G + ni * d + ni;       // item gradients go to the exchange buffer
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = bpr_dispatch_grad(cfg, U, V, b, u, i, j, B, sd, first_draw, w, MODE_DENSE, plain_extra(cfg, b), loss_out, st)) return rc;

    a.cfg = *cfg;
    a.off_V = L.V; a.off_b = L.b; a.off_G = L.G[epoch & 1]; a.off_Gnext = L.G[(epoch + 1) & 1]; a.off_flags = L.flags;
    a.U = U; a.msU = msU; a.msV = msV; a.msb = msb; a.GU = w.GU; a.cntU = w.cntU;
    a.epoch = (uint32_t)epoch; a.g_floats = L.g_floats;
    const int64_t mine_rows = ((int64_t)ni - a.pv.rank + a.pv.world - 1) / a.pv.world;
    int64_t xb = (mine_rows + 7) / 8;
    if (xb > kNumSMs) xb = kNumSMs;
    if (xb < 1) xb = 1;
    int64_t ub = ((int64_t)cfg->n_users + 7) / 8;
    if (ub > 3 * kNumSMs) ub = 3 * kNumSMs;                         // 4 blocks of 256 threads per SM in total
    a.xblocks = (int)xb;
    const unsigned grid = (unsigned)(xb + ub);
    const int vw = d % 4 == 0 ? 4 : d % 2 == 0 ? 2 : 1;
#define TKR_DP(VW, W) bpr_dp_exchange_kernel<VW, W><<<grid, 256, 0, st>>>(a)
#define TKR_DP_W(VW) do { if (a.pv.world <= 2) TKR_DP(VW, 2); else if (a.pv.world <= 4) TKR_DP(VW, 4); else TKR_DP(VW, 8); } while (0)
    if (vw == 4) TKR_DP_W(4); else if (vw == 2) TKR_DP_W(2); else TKR_DP_W(1);
#undef TKR_DP_W
#undef TKR_DP
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

extern "C" int tkr_bpr_dp_status(const tkr_bpr_cfg* cfg, const tkr_peers* peers, void* stream) {
    if (int rc = bpr_check_cfg(cfg, 1)) return rc;
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    const DpLayout L = dp_layout(cfg);
    uint32_t err = 0;
    TKR_CUDA(cudaMemcpyAsync(&err, pv.base[pv.rank] + L.flags + (size_t)kPeerSlots * TKR_MAX_PEERS * 4, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TKR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (err != 0) { set_error("data-parallel step: a rank did not reach barrier %u within 20 s", err - 1); return TKR_ERR_CUDA; }
    return TKR_OK;
}
