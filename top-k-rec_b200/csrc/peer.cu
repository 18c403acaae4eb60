// Exchange buffers for the multi-GPU paths: device memory that every rank of one box maps into its own
// address space through CUDA IPC, so the fused step kernels load from and store to peer HBM over NVLink / NVSwitch.
// One process per GPU; the 64-byte handles travel over torch.distributed (any backend) -- see topkrec/peer.py.
#include "peer.cuh"
#include <string.h>

using namespace tkr;

namespace tkr {
int peer_view_from(const tkr_peers* peers, PeerView* out) {
    TKR_CHECK_ARG(peers != nullptr, "peers is NULL");
    TKR_CHECK_ARG(peers->world >= 1 && peers->world <= TKR_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
                  "bad peer table: rank %d of %d (max %d ranks)", peers->rank, peers->world, TKR_MAX_PEERS);
    out->rank = peers->rank; out->world = peers->world;
    for (int p = 0; p < TKR_MAX_PEERS; ++p) {
        out->base[p] = p < peers->world ? (char*)peers->base[p] : nullptr;
        TKR_CHECK_ARG(p >= peers->world || peers->base[p] != nullptr, "peer buffer %d is not mapped", p);
    }
    return TKR_OK;
}
}  // namespace tkr

extern "C" int tkr_peer_alloc(size_t bytes, void** out) {
    TKR_CHECK_ARG(out != nullptr && bytes > 0, "bad arguments");
    void* p = nullptr;
    TKR_CUDA(cudaMalloc(&p, bytes));      // a plain cudaMalloc allocation: what cudaIpcGetMemHandle can export
    TKR_CUDA(cudaMemset(p, 0, bytes));
    *out = p;
    return TKR_OK;
}

extern "C" int tkr_peer_free(void* p) {
    if (p != nullptr) TKR_CUDA(cudaFree(p));
    return TKR_OK;
}

extern "C" int tkr_peer_export(void* p, void* handle_out) {
    TKR_CHECK_ARG(p != nullptr && handle_out != nullptr, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == TKR_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    TKR_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out, &h, sizeof(h));
    return TKR_OK;
}

extern "C" int tkr_peer_import(const void* handle, void** out) {
    TKR_CHECK_ARG(handle != nullptr && out != nullptr, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    TKR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out = p;
    return TKR_OK;
}

extern "C" int tkr_peer_release(void* p) {
    if (p != nullptr) TKR_CUDA(cudaIpcCloseMemHandle(p));
    return TKR_OK;
}
