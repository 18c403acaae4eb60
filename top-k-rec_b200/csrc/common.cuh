// Shared helpers for libtopkrec.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "topkrec.h"

namespace tkr {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define TKR_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            tkr::set_error(__VA_ARGS__);         \
            return TKR_ERR_INVALID;              \
        }                                        \
    } while (0)

#define TKR_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            tkr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return TKR_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define TKR_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            tkr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return TKR_ERR_CUDA;                                                               \
        }                                                                                      \
        tkr::count_launch();                                                                   \
    } while (0)

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Order-preserving map fp32 -> uint32 (larger float <=> larger uint); NaN-free inputs.
__host__ __device__ inline uint32_t f32_to_ord(float s) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(s);
#else
    uint32_t b; memcpy(&b, &s, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float ord_to_f32(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float s; memcpy(&s, &b, 4); return s;
#endif
}
// Ranking key: (score desc, column desc)  <=>  key desc.  key 0 = "empty".
__host__ __device__ inline uint64_t make_key(float s, int32_t col) {
    return ((uint64_t)f32_to_ord(s) << 32) | (uint32_t)col;
}

// Philox4x32-10 (Salmon et al. 2011), the counter-based generator behind tkr_bpr_sample.
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    }
    __host__ __device__ static inline void run(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(M0, c[0], h0, l0);
            mulhilo(M1, c[2], h1, l1);
            uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
            c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
            k0 += W0; k1 += W1;
        }
    }
};

// Uniform integer in [0, n) from 32 random bits (multiply-shift; bias < n / 2^32).
__host__ __device__ inline uint32_t bounded(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

}  // namespace tkr
