// Device helpers shared by the K1 kernels (bpr_step.cu, bpr_dp.cu, bpr_persist.cu).
#pragma once
#include "bpr_internal.cuh"

namespace tkr {

// ---- vector helpers: VW floats per lane per chunk (4 = 128-bit, 2 = 64-bit, 1 = scalar) ----
template <int VW> struct Vec;
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load_cv(const float* p) { float4 t = __ldcv(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    __device__ __forceinline__ void red_add(float* p) const {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};
template <> struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float* p) { float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void load_cv(const float* p) { float2 t = __ldcv(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
    __device__ __forceinline__ void red_add(float* p) const {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    }
};
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = *p; }
    __device__ __forceinline__ void load_cv(const float* p) { v[0] = __ldcv(p); }
    __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
    __device__ __forceinline__ void red_add(float* p) const { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory"); }
};

__device__ __forceinline__ void sh_red_add(float* shared_ptr, float v) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(shared_ptr)), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

template <bool L1> __device__ __forceinline__ float reg_grad(float x, float lam) {
    if (L1) return lam * (float)((x > 0.f) - (x < 0.f));
    return lam * x;
}
template <bool L1> __device__ __forceinline__ float reg_val(float x, float lam) {
    if (L1) return lam * fabsf(x);
    return 0.5f * lam * x * x;
}

// One element of the optimiser update (App. A.4); shared by the apply kernel and the in-place path of the gradient
// kernel so both produce the same bits.
__device__ __forceinline__ void opt_update(const tkr_bpr_cfg& cfg, float g, float& v, float& m) {
    if (cfg.optimizer == TKR_OPT_RMSPROP) {
        m = cfg.rms_decay * m + (1.0f - cfg.rms_decay) * g * g;
        // rsqrt * lr * g: the form of TF's SparseApplyRMSProp kernel (training_ops.cc).  One MUFU.RSQ (<= 2 ulp) instead of
        // an IEEE sqrt and an IEEE divide: in the persistent small-batch kernel those two were 1 800 of a row update's
        // cycles on the critical path; the update term is lr-scaled, 2 ulp of it is ~1e-11 of the parameter.
        v = v - cfg.lr * g * rsqrtf(m + cfg.rms_eps);
    } else {
        v = v - cfg.lr * g;
    }
}

// `d` leading columns are parameters; columns [d, dz) of the accumulator row are only re-zeroed
template <int VW>
__device__ __forceinline__ void apply_row(const tkr_bpr_cfg& cfg, float* __restrict__ var, float* __restrict__ ms,
                                          float* __restrict__ G, int d, int lane, int dz = 0) {
    for (int off = d + lane * VW; off < dz; off += 32 * VW) {
        Vec<VW> z;
#pragma unroll
        for (int t = 0; t < VW; ++t) z.v[t] = 0.f;
        z.store(G + off);
    }
    for (int off = lane * VW; off < d; off += 32 * VW) {
        Vec<VW> g, v, m, z;
        g.load(G + off); v.load(var + off);
        const bool rms = cfg.optimizer == TKR_OPT_RMSPROP;
        if (rms) m.load(ms + off);
#pragma unroll
        for (int t = 0; t < VW; ++t) opt_update(cfg, g.v[t], v.v[t], m.v[t]);
        if (rms) m.store(ms + off);
        v.store(var + off);
#pragma unroll
        for (int t = 0; t < VW; ++t) z.v[t] = 0.f;
        z.store(G + off);
    }
}

}  // namespace tkr
