// Peer-memory plumbing shared by the multi-GPU kernels (one process per GPU; every rank maps every other rank's
// exchange buffer through CUDA IPC, so kernels read and write peer HBM directly over NVLink / NVSwitch).
#pragma once
#include "common.cuh"

namespace tkr {

// Kernel-side view of the ranks' exchange buffers: base[p] is rank p's buffer as mapped into THIS process (base[rank]
// is the local allocation).  All ranks use the same layout, so an offset means the same thing in every base[p].
struct PeerView {
    int rank, world;
    char* base[TKR_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Flag block at `flag_off` of every buffer: uint32 [TKR_PEER_SLOTS][TKR_MAX_PEERS] arrival words + one error word.
// Slot s, word q of rank p's block = "rank q has reached barrier s of epoch <value>".
constexpr int kPeerSlots = 8;
constexpr size_t kPeerFlagBytes = (size_t)(kPeerSlots * TKR_MAX_PEERS + 8) * 4;
constexpr uint64_t kPeerTimeoutNs = 20000000000ull;     // a rank that never arrives must not hang the GPU: 20 s, then error

__device__ __forceinline__ uint32_t* peer_flag(const PeerView& pv, int owner, size_t flag_off, int slot, int word) {
    return reinterpret_cast<uint32_t*>(pv.base[owner] + flag_off) + slot * TKR_MAX_PEERS + word;
}
__device__ __forceinline__ uint32_t* peer_error_word(const PeerView& pv, size_t flag_off) {
    return reinterpret_cast<uint32_t*>(pv.base[pv.rank] + flag_off) + kPeerSlots * TKR_MAX_PEERS;
}

// Called by ONE thread after the work to publish is complete (and fenced at gpu scope by the caller's own
// synchronisation): tell every peer that this rank reached (slot, epoch).
__device__ __forceinline__ void peer_signal(const PeerView& pv, size_t flag_off, int slot, uint32_t epoch) {
    __threadfence_system();
    for (int p = 0; p < pv.world; ++p)
        if (p != pv.rank) st_release_sys(peer_flag(pv, p, flag_off, slot, pv.rank), epoch);
}
// ... and wait until every peer has reached it (epochs only grow; compared modulo 2^32).  Returns false on timeout
// after setting the local error word.
__device__ __forceinline__ bool peer_wait(const PeerView& pv, size_t flag_off, int slot, uint32_t epoch) {
    const uint64_t t0 = global_timer_ns();
    for (int p = 0; p < pv.world; ++p) {
        if (p == pv.rank) continue;
        const uint32_t* f = peer_flag(pv, pv.rank, flag_off, slot, p);
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if (global_timer_ns() - t0 > kPeerTimeoutNs) { atomicExch(peer_error_word(pv, flag_off), 1u + (uint32_t)slot); return false; }
            __nanosleep(64);
        }
    }
    return true;
}

int peer_view_from(const tkr_peers* peers, PeerView* out);

}  // namespace tkr
