// Hot path 2 on several GPUs: item columns are sharded over the ranks, every rank produces the
// filtered top-k of ALL users of a batch against its shard, and the candidate lists are exchanged all-to-all by user
// slice -- rank r ends up with the world lists of the users of slice r only, merges them, and owns the final lists of
// those users.  The exchange is done by the ranks' own kernels over peer memory (NVLink stores into the owner's merge
// buffer + flag words), double-buffered by batch parity so that the push / merge of batch t overlap the tensor-core
// filter of batch t+1 on another stream.  Replaces round 1's stack + all_gather (every rank received and merged every
// row: 109 MB in per rank and step at 8 GPUs) + two re-layout copies.
#include "peer.cuh"

namespace tkr {

// exchange buffer: 2 parity slots of [idx int32 | score f32], each world * slice_cap * k elements, then the flag block
struct XchLayout { size_t slot_elems, slot_bytes, flags, total; };
static XchLayout xch_layout(int64_t nu_cap, int k, int world) {
    XchLayout L;
    const int64_t slice = (nu_cap + world - 1) / world;
    L.slot_elems = (size_t)world * slice * k;
    L.slot_bytes = align_up(L.slot_elems * 4, 256) * 2;
    L.flags = 2 * L.slot_bytes;
    L.total = L.flags + align_up(kPeerFlagBytes, 256);
    return L;
}

constexpr int SLOT_PUSHED = 2, SLOT_MERGED = 3;     // flag slots (0/1 belong to the data-parallel step)

__global__ void peer_signal_kernel(PeerView pv, size_t flag_off, int slot, uint32_t epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) peer_signal(pv, flag_off, slot, epoch);
}
__global__ void peer_wait_kernel(PeerView pv, size_t flag_off, int slot, uint32_t epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) peer_wait(pv, flag_off, slot, epoch);
}

// lists [nu, k] of this rank -> list `rank` of every owner's slot (owner of row u = u / slice)
__global__ void __launch_bounds__(256) topk_push_kernel(PeerView pv, const int32_t* __restrict__ idx, const float* __restrict__ score,
                                                        int64_t nu, int k, int64_t slice, size_t slot_off, size_t score_off) {
    const int64_t n = nu * k;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = e / k;
        const int c = (int)(e - u * k);
        const int owner = (int)(u / slice);
        const int64_t rows_owner = (nu - owner * slice) < slice ? (nu - owner * slice) : slice;
        const int64_t o = ((int64_t)pv.rank * rows_owner + (u - owner * slice)) * k + c;
        char* dst = pv.base[owner] + slot_off;
        reinterpret_cast<int32_t*>(dst)[o] = idx[e];
        reinterpret_cast<float*>(dst + score_off)[o] = score[e];
    }
}

__global__ void peer_signal_one_kernel(PeerView pv, size_t flag_off, int slot, int target, uint32_t epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { __threadfence_system(); st_release_sys(peer_flag(pv, target, flag_off, slot, pv.rank), epoch); }
}
__global__ void peer_wait_one_kernel(PeerView pv, size_t flag_off, int slot, int source, uint32_t epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const uint64_t t0 = global_timer_ns();
        const uint32_t* f = peer_flag(pv, pv.rank, flag_off, slot, source);
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if (global_timer_ns() - t0 > kPeerTimeoutNs) { atomicExch(peer_error_word(pv, flag_off), 1u + (uint32_t)slot); break; }
            __nanosleep(64);
        }
    }
}

int merge_lists_public(const int32_t* idx, const float* score, int n_lists, int64_t nu, int k, int32_t* out_idx, float* out_score, void* stream);

}  // namespace tkr

using namespace tkr;

extern "C" size_t tkr_topk_exchange_bytes(int64_t nu_cap, int32_t k, int32_t world) {
    if (nu_cap <= 0 || k <= 0 || world <= 0 || world > TKR_MAX_PEERS) return 0;
    return xch_layout(nu_cap, k, world).total;
}

extern "C" int tkr_topk_exchange_push(const int32_t* idx, const float* score, int64_t nu, int64_t nu_cap, int32_t k,
                                      const tkr_peers* peers, uint64_t epoch, void* stream) {
    TKR_CHECK_ARG(idx && score && nu >= 1 && nu <= nu_cap && k >= 1 && epoch >= 1, "bad arguments");
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    const XchLayout L = xch_layout(nu_cap, k, pv.world);
    cudaStream_t st = (cudaStream_t)stream;
    // the slot of this parity was merged two batches ago by every owner?  (first two batches: flags are still 0 >= e-2)
    if (epoch > 2 && pv.world > 1) { peer_wait_kernel<<<1, 32, 0, st>>>(pv, L.flags, SLOT_MERGED, (uint32_t)(epoch - 2)); TKR_LAUNCH_CHECK(); }
    const int64_t slice = (nu + pv.world - 1) / pv.world;
    int64_t blocks = (nu * k + 255) / 256;
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    topk_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(pv, idx, score, nu, k, slice, (epoch & 1) * L.slot_bytes, align_up(L.slot_elems * 4, 256));
    TKR_LAUNCH_CHECK();
    if (pv.world > 1) { peer_signal_kernel<<<1, 32, 0, st>>>(pv, L.flags, SLOT_PUSHED, (uint32_t)epoch); TKR_LAUNCH_CHECK(); }
    return TKR_OK;
}

extern "C" int tkr_topk_exchange_merge(int64_t nu, int64_t nu_cap, int32_t k, const tkr_peers* peers, uint64_t epoch,
                                       int32_t* out_idx, float* out_score, void* stream) {
    TKR_CHECK_ARG(out_idx && out_score && nu >= 1 && nu <= nu_cap && k >= 1 && epoch >= 1, "bad arguments");
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    const XchLayout L = xch_layout(nu_cap, k, pv.world);
    cudaStream_t st = (cudaStream_t)stream;
    if (pv.world > 1) { peer_wait_kernel<<<1, 32, 0, st>>>(pv, L.flags, SLOT_PUSHED, (uint32_t)epoch); TKR_LAUNCH_CHECK(); }
    const int64_t slice = (nu + pv.world - 1) / pv.world;
    int64_t rows = nu - (int64_t)pv.rank * slice;
    if (rows > slice) rows = slice;
    if (rows > 0) {
        const char* slot = pv.base[pv.rank] + (epoch & 1) * L.slot_bytes;
        if (int rc = merge_lists_public(reinterpret_cast<const int32_t*>(slot), reinterpret_cast<const float*>(slot + align_up(L.slot_elems * 4, 256)),
                                        pv.world, rows, k, out_idx, out_score, stream)) return rc;
    }
    if (pv.world > 1) { peer_signal_kernel<<<1, 32, 0, st>>>(pv, L.flags, SLOT_MERGED, (uint32_t)epoch); TKR_LAUNCH_CHECK(); }
    return TKR_OK;
}

extern "C" int tkr_topk_exchange_status(int64_t nu_cap, int32_t k, const tkr_peers* peers, void* stream) {
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    const XchLayout L = xch_layout(nu_cap, k, pv.world);
    uint32_t err = 0;
    TKR_CUDA(cudaMemcpyAsync(&err, pv.base[pv.rank] + L.flags + (size_t)kPeerSlots * TKR_MAX_PEERS * 4, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TKR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (err != 0) { set_error("top-k exchange: a rank did not reach barrier %u within 20 s", err - 1); return TKR_ERR_CUDA; }
    return TKR_OK;
}

// ---- point-to-point flags on a caller-laid-out exchange buffer (the ring of tkr_score_topk_tc_segment uses them) ----
extern "C" size_t tkr_peer_flag_bytes(void) { return align_up(kPeerFlagBytes, 256); }
extern "C" int tkr_peer_signal_to(const tkr_peers* peers, size_t flag_off, int32_t slot, int32_t target, uint64_t epoch, void* stream) {
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    TKR_CHECK_ARG(slot >= 0 && slot < kPeerSlots && target >= 0 && target < pv.world, "bad slot / target");
    peer_signal_one_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pv, flag_off, slot, target, (uint32_t)epoch);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}
extern "C" int tkr_peer_wait_from(const tkr_peers* peers, size_t flag_off, int32_t slot, int32_t source, uint64_t epoch, void* stream) {
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    TKR_CHECK_ARG(slot >= 0 && slot < kPeerSlots && source >= 0 && source < pv.world, "bad slot / source");
    peer_wait_one_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pv, flag_off, slot, source, (uint32_t)epoch);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}
extern "C" int tkr_peer_status(const tkr_peers* peers, size_t flag_off, void* stream) {
    PeerView pv;
    if (int rc = peer_view_from(peers, &pv)) return rc;
    uint32_t err = 0;
    TKR_CUDA(cudaMemcpyAsync(&err, pv.base[pv.rank] + flag_off + (size_t)kPeerSlots * TKR_MAX_PEERS * 4, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TKR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (err != 0) { set_error("peer flag %u was not raised within 20 s", err - 1); return TKR_ERR_CUDA; }
    return TKR_OK;
}
