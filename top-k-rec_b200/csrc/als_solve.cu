// Alternating least squares for WMF / CER (SURVEY.md 8(f) NEXT-1): one half-step
//     x_r = solve( base + (a-b) * sum_{p in pos(r)} y_p y_p^T + ridge*I ,  a * sum_p y_p + ridge * prior_r )
// for every row r of the solved side (reference: single/cer.py:36-63, single/wmf.py:67-96 -- a Python loop
// calling np.dot + np.linalg.solve per row).
//
// Shape of the kernel (fp32 FMA pipe; the k x k systems are far too small / too many for a library call):
//   * one thread block per row segment.  The d x d normal matrix lives in REGISTERS: the lower triangle of
//     64x64 blocks, two warps per block, an 8x8 register tile per thread (symmetry halves the FMA work).
//   * the positives' factor rows are gathered HBM/L2 -> shared memory with cp.async, 16 rows per stage,
//     3 stages in flight; each staged row is a rank-1 update read with conflict-free 128-bit shared loads.
//   * the matrix is then factored where it is, A = L D L^T (no square roots).  d <= 192: in rounds of four columns -- the
//     thread holding the 4x4 diagonal block factors it, the column owners eliminate it from their rows and publish raw M
//     and scaled L = M/D (a float4 per row), every tile applies the rank-4 update: two barriers per four columns.
//     d <= 256 (96 registers per thread at 640 threads, no room for the rank-4 operands): one column per barrier.
//     Either way the published columns ARE the stored factor (packed, in shared memory) that back substitution walks,
//     and the right-hand side rides along as an extra row, so forward substitution costs nothing.
//   * rows with more positives than one segment (popular items: up to every user) are split: the segments'
//     partial matrices go to a caller workspace in the threads' own register order and a second kernel sums
//     them in a fixed order (deterministic) and solves.  The shared Gram b*Yr^T Yr uses the same two kernels.
//   * the row's loss terms (cer.py:46,58-63) are formed from x, the right-hand side and the column sums:
//     x^T B x = x^T rhs - ridge |x|^2 for the solved x, so the matrix is not needed again.
#include "common.cuh"
#include <algorithm>
#include <type_traits>

namespace tkr {
namespace als {

constexpr int KC_ALIGN = 32; // slices of the shared Gram's row list are multiples of this (>= every Geo<NB>::KC)

template <int NB> struct Geo {
    // gather pipeline: rows per stage x stages.  The per-stage bookkeeping (indices, cp.async issue, barrier) is ~290
    // instructions per warp, a fifth of a 16-row stage's FMA work: at NB = 4 two stages of 32 rows measured 3 % (user rows) to
    // 7 % (item rows) faster than three of 16; at NB = 3 they measured 5 % slower, at NB <= 2 the same -- so 16 x 3 there.
    static constexpr int KC = NB == 4 ? 32 : 16;
    static constexpr int STAGES = NB == 4 ? 2 : 3;
    static constexpr int DP = 64 * NB;                 // padded width
    static constexpr int NBLK = NB * (NB + 1) / 2;     // lower-triangular 64x64 blocks
    // rows of a thread's register tile (columns: always 8).  16 x 8 tiles on 320 threads would halve the shared-memory
    // reads per FMA at NB = 4, but 10 warps put 3 on one scheduler partition (16 K registers): 168 registers per thread
    // at most, not enough next to 128 accumulators (measured: 1.2 KB of spills).  So 8 x 8 everywhere.
    static constexpr int TR = 8;
    static constexpr int RS = 32;                      // row register i sits at r0 + (i & 3) + (i >> 2) * RS
    static constexpr int TPB = 512 / TR;               // threads per 64x64 block
    static constexpr int NT = NBLK * TPB;              // threads per thread block
    static constexpr int MAXREG = NB == 4 ? 96 : 168;  // register budget: 6 / 2 / 1 / 1 resident blocks per SM (the 4-column rounds
                                                       // spill below ~130 registers: 112 -> 584 bytes at NB = 2).  Same budgets as
                                                       // __launch_bounds__(NT, 6 / 2 / 1 / 1), should a compiler reject the pair
    static constexpr int NP = DP / 4;                  // 4-column panels of the factor
    static constexpr int PACKED = DP * DP / 2 + 2 * DP;// floats of the factor: panel p keeps a float4 per row 4p..DP-1
    static constexpr int PART = NBLK * 4096 + DP;      // floats per partial slot
    // blocked factorisation (factor_solve_blocked): panels of PW columns, panel p keeps rows PW*p..DP-1 of its columns k-major
    // (column k of the panel = pitch(p) consecutive floats; the +4 keeps 16-byte alignment and spreads the banks of the
    // back substitution's column walks)
    static constexpr int PW = 16;
    static constexpr int NPAN = DP / PW;
    __host__ __device__ static constexpr int pitch(int p) { return DP - PW * p + 4; }
    __host__ __device__ static constexpr int panel_off(int p) { return PW * (p * (DP + 4) - PW * (p * (p - 1) / 2)); }
    static constexpr int BLOCKED = panel_off(NPAN);    // floats of the blocked factor
    static constexpr int FACTOR = PACKED > BLOCKED ? PACKED : BLOCKED;
    static constexpr int MT_PITCH = DP + 4;            // raw M of the current panel, k-major, in the (free) gather stages
    static constexpr int NTB = NT;                     // threads of a blocked-variant block (warp 0 doubles as the pivot warp)
    static constexpr int MAXREG_B = MAXREG;
    static constexpr size_t SMEM = (size_t)(FACTOR + STAGES * KC * DP + 4 * DP + 2 * DP + 16 + 64 + 3 * DP + PW) * sizeof(float);
};

struct RowArgs {
    const float* Y; float* X;
    const int32_t* idx;
    const float* base; const float* prior;
    const int32_t* seg_row; const int64_t* seg_off; const int32_t* seg_len; const int32_t* seg_slot;
    const int32_t* multi_row; const int32_t* multi_slot0; const int32_t* multi_nslots; const int64_t* multi_total;
    float* partial; double* loss_rows;
    float a, amb, ridge, lreg;
    int d, solve_empty, item_loss;
    int full_diag;              // 1: the upper-triangle quarter of the diagonal blocks is wanted too (the shared Gram's reduce kernel)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// compile-time loop: f(std::integral_constant<int, B>{}) ... f(<E - 1>) -- register arrays indexed by the loop variable stay
// in registers whatever the unroller decides
template <int B, int E, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// thread -> tile geometry
struct Tile {
    int I, J;        // block coordinates (I >= J)
    int r0, c0;      // first row / column of the tile (see row_pos / reg_pos)
    int tx;
};
template <int NB> __device__ __forceinline__ Tile tile_of(int tid) {
    Tile t;
    const int blk = tid / Geo<NB>::TPB, u = tid % Geo<NB>::TPB;
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= blk) ++I;
    t.I = I; t.J = blk - I * (I + 1) / 2;
    t.tx = u & 7;
    t.r0 = I * 64 + (u >> 3) * 4;
    t.c0 = t.J * 64 + t.tx * 4;
    return t;
}
__device__ __forceinline__ int reg_pos(int x0, int q) { return x0 + (q & 3) + (q >> 2) * 32; }           // column register q
template <int NB> __device__ __forceinline__ int row_pos(int r0, int i) { return r0 + (i & 3) + (i >> 2) * Geo<NB>::RS; }   // row register i

// acc += sum over rows list[0..n) of y y^T (this thread's tile); ssum += column tid of the same rows.
template <int NB>
__device__ __forceinline__ void gram_accumulate(float (&acc)[Geo<NB>::TR][8], float& ssum, const Tile& t, const float* __restrict__ Y,
                                                int d, const int32_t* __restrict__ list, int64_t n, float* stage, bool lower_only) {
    using G = Geo<NB>;
    constexpr int EPT = (G::KC * (G::DP / 4) + G::NT - 1) / G::NT;   // 16-byte pieces of a stage per thread (upper bound)
    const int tid = threadIdx.x;
    const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const int64_t nchunks = (n + G::KC - 1) / G::KC;
    // vector path: piece e = tid + i*NT of every stage is (row e / q, floats 4*(e % q)..+3); the row's index is fetched
    // one stage ahead of its cp.async so the gather never waits on the index load
    int prow[EPT], pcol[EPT], pidx[EPT];
    const int q = d >> 2;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int e = tid + i * G::NT;
        prow[i] = (vec && e < G::KC * q) ? e / q : G::KC;
        pcol[i] = (e - prow[i] * q) * 4;
        pidx[i] = 0;
    }
    auto fetch_idx = [&](int64_t c) {
        if (vec && c < nchunks) {
            const int rows = (int)min((int64_t)G::KC, n - c * G::KC);
#pragma unroll
            for (int i = 0; i < EPT; ++i)
                if (prow[i] < rows) pidx[i] = list[c * G::KC + prow[i]];
        }
    };
    auto issue = [&](int64_t c) {
        if (c < nchunks) {
            const int rows = (int)min((int64_t)G::KC, n - c * G::KC);
            const uint32_t buf = stage_s + (uint32_t)((c % G::STAGES) * G::KC * G::DP) * 4u;
            if (vec) {
#pragma unroll
                for (int i = 0; i < EPT; ++i)
                    if (prow[i] < rows) cp_async16(buf + (uint32_t)(prow[i] * G::DP + pcol[i]) * 4u, Y + (size_t)pidx[i] * d + pcol[i]);
            } else {
                for (int e = tid; e < rows * d; e += G::NT) {
                    const int r = e / d, cc = e - r * d;
                    cp_async4(buf + (uint32_t)(r * G::DP + cc) * 4u, Y + (size_t)list[c * G::KC + r] * d + cc);
                }
            }
        }
        cp_async_commit();
    };
    fetch_idx(0);
#pragma unroll
    for (int s = 0; s < G::STAGES - 1; ++s) { issue(s); fetch_idx(s + 1); }
    for (int64_t c = 0; c < nchunks; ++c) {
        cp_async_wait<G::STAGES - 2>();
        named_barrier(3, G::NT);                 // the NT tile threads (a blocked-variant block has one more warp, which is not here)
        issue(c + G::STAGES - 1);
        fetch_idx(c + G::STAGES);
        const float* buf = stage + (c % G::STAGES) * G::KC * G::DP;
        const int rows = (int)min((int64_t)G::KC, n - c * G::KC);
        // lower_only: a diagonal block whose upper-triangle quarter (rows 0..31 x columns 32..63 of the block) nobody reads
        // -- every solve; not the shared Gram, whose reduce kernel writes the full matrix.  Two of a scheduler's five warps
        // then issue 48 instead of 64 FFMA per gathered row.
        auto rank1 = [&](auto lo) {
            constexpr bool LO = decltype(lo)::value;
#pragma unroll 2
            for (int k = 0; k < rows; ++k) {
                const float* y = buf + k * G::DP;
                const float4 b0 = *reinterpret_cast<const float4*>(y + t.c0), b1 = *reinterpret_cast<const float4*>(y + t.c0 + 32);
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int gi = 0; gi < G::TR / 4; ++gi) {
                    const float4 a4 = *reinterpret_cast<const float4*>(y + t.r0 + gi * G::RS);
                    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < (LO && gi == 0 ? 4 : 8); ++j) acc[gi * 4 + i][j] = fmaf(av[i], bv[j], acc[gi * 4 + i][j]);
                }
                if (tid < G::DP) ssum += y[tid];
            }
        };
        if (lower_only) rank1(std::true_type{}); else rank1(std::false_type{});
    }
    cp_async_wait<0>();
    named_barrier(3, G::NT);
}

template <int NB>
__device__ __forceinline__ void store_partial(const float (&acc)[Geo<NB>::TR][8], float ssum, float* slot) {
    using G = Geo<NB>;
    const int tid = threadIdx.x, blk = tid / G::TPB, u = tid % G::TPB;
    float4* p = reinterpret_cast<float4*>(slot) + (size_t)blk * 1024 + u;
#pragma unroll
    for (int i = 0; i < G::TR; ++i) {
        p[(i * 2) * G::TPB] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        p[(i * 2 + 1) * G::TPB] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    if (tid < G::DP) slot[G::NBLK * 4096 + tid] = ssum;
}
template <int NB>
__device__ __forceinline__ void add_partial(float (&acc)[Geo<NB>::TR][8], float& ssum, const float* slot, int blk, int u, bool with_sum) {
    using G = Geo<NB>;
    const float4* p = reinterpret_cast<const float4*>(slot) + (size_t)blk * 1024 + u;
#pragma unroll
    for (int i = 0; i < G::TR; ++i) {
        const float4 lo = p[(i * 2) * G::TPB], hi = p[(i * 2 + 1) * G::TPB];
        acc[i][0] += lo.x; acc[i][1] += lo.y; acc[i][2] += lo.z; acc[i][3] += lo.w;
        acc[i][4] += hi.x; acc[i][5] += hi.y; acc[i][6] += hi.z; acc[i][7] += hi.w;
    }
    if (with_sum && (int)threadIdx.x < G::DP) ssum += slot[G::NBLK * 4096 + threadIdx.x];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Loss terms of one row (cer.py:46,58-63 / wmf.py:77,91-96), threads tid < DP hold x, rhs0, s of their column.
template <int NB>
__device__ __forceinline__ void row_loss(const RowArgs& p, int row, int64_t n, bool solved, float x, float rhs0, float ssum,
                                         float pr, float* red) {
    using G = Geo<NB>;
    const int tid = threadIdx.x;
    if (tid < G::DP) {
        const float dx = p.prior ? x - pr : x;
        float v0 = warp_sum(x * x), v1 = warp_sum(x * rhs0), v2 = warp_sum(ssum * x), v3 = warp_sum(dx * dx);
        if ((tid & 31) == 0) { float* o = red + (tid >> 5) * 4; o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3; }
    }
    __syncthreads();
    if (tid == 0) {
        double xx = 0, xr = 0, sx = 0, dd = 0;
        for (int w = 0; w < G::DP / 32; ++w) { xx += red[w * 4]; xr += red[w * 4 + 1]; sx += red[w * 4 + 2]; dd += red[w * 4 + 3]; }
        double loss;
        if (!p.item_loss) loss = 0.5 * (double)p.lreg * xx;
        else {
            loss = 0.5 * (double)p.lreg * dd;
            if (n > 0 && solved) loss += 0.5 * (xr - (double)p.ridge * xx) + 0.5 * (double)n * (double)p.a - (double)p.a * sx;
        }
        if (p.loss_rows) p.loss_rows[row] = loss;   // optional output (C callers may pass NULL)
    }
}

template <int NB> struct Smem {
    using G = Geo<NB>;
    float* M; float* stage; float* raw; float* zs; float* xs; float* d44; float* red; float* keep; float* dinv;
    __device__ explicit Smem(float* base) {
        M = base; stage = M + G::FACTOR;     // FACTOR * 4 bytes is a multiple of 16 for every DP
        raw = stage + G::STAGES * G::KC * G::DP; zs = raw + 4 * G::DP; xs = zs + G::DP; d44 = xs + G::DP; red = d44 + 16; keep = red + 64;
        dinv = keep + 3 * G::DP;              // [PW] 1 / D of the round's diagonal block
    }
};

// A = L D L^T four columns per round (see finish_row), then back substitution; z = right-hand side entry of thread
// tid < DP on entry, x = solution entry on exit.
template <int NB>
__device__ __forceinline__ void factor_solve_panels(float (&acc)[Geo<NB>::TR][8], const Tile& t, const Smem<NB>& sm, float z, float& x) {
    using G = Geo<NB>;
    const int tid = threadIdx.x;
    float4* const L4 = reinterpret_cast<float4*>(sm.M);
    float4* const R4 = reinterpret_cast<float4*>(sm.raw);
    const int rmax = row_pos<NB>(t.r0, G::TR - 1), cmax = t.c0 + 35;
    const bool on_diag_block = t.I == t.J;
    const int ty4 = t.r0 - t.I * 64;            // 4 * ty
    float pinv_mine = 1.f;                      // 1 / D[tid]
    int poff = 0;                               // float4 offset of the round's panel in the factor
    int j0 = 0;
    for (int Jb = 0; Jb < NB; ++Jb) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll 1
            for (int tq = 0; tq < 8; ++tq) {
                const bool owner = t.J == Jb && t.tx == tq;
                // ---- A
                if (owner && on_diag_block && ty4 == (G::TR == 8 ? tq : (tq & 3)) * 4) {
                    const bool th = G::TR == 16 && (tq >> 2);     // which of the two candidate row groups holds rows j0..j0+3
                    auto dg = [&](int qr, int qc) -> float {
                        if (G::TR == 8) return acc[g * 4 + qr][g * 4 + qc];
                        return th ? acc[((2 * g + 1) * 4 + qr) % G::TR][g * 4 + qc] : acc[((2 * g) * 4 + qr) % G::TR][g * 4 + qc];
                    };
                    float a00 = dg(0, 0);
                    float a10 = dg(1, 0), a11 = dg(1, 1);
                    float a20 = dg(2, 0), a21 = dg(2, 1), a22 = dg(2, 2);
                    float a30 = dg(3, 0), a31 = dg(3, 1), a32 = dg(3, 2), a33 = dg(3, 3);
                    const float i0 = __frcp_rn(a00);
                    const float l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
                    a11 = fmaf(-l10, a10, a11); a21 = fmaf(-l20, a10, a21); a31 = fmaf(-l30, a10, a31);
                    a22 = fmaf(-l20, a20, a22); a32 = fmaf(-l30, a20, a32); a33 = fmaf(-l30, a30, a33);
                    const float i1 = __frcp_rn(a11);
                    const float l21 = a21 * i1, l31 = a31 * i1;
                    a22 = fmaf(-l21, a21, a22); a32 = fmaf(-l31, a21, a32); a33 = fmaf(-l31, a31, a33);
                    const float i2 = __frcp_rn(a22);
                    const float l32 = a32 * i2;
                    a33 = fmaf(-l32, a32, a33);
                    const float i3 = __frcp_rn(a33);
                    float4* o = reinterpret_cast<float4*>(sm.d44);
                    o[0] = make_float4(i0, i1, i2, i3);
                    o[1] = make_float4(l10, l20, l30, l21);
                    o[2] = make_float4(l31, l32, 0.f, 0.f);
                    L4[poff + 0] = make_float4(0.f, 0.f, 0.f, 0.f);      // rows of the diagonal block: L entries left of the diagonal
                    L4[poff + 1] = make_float4(l10, 0.f, 0.f, 0.f);
                    L4[poff + 2] = make_float4(l20, l21, 0.f, 0.f);
                    L4[poff + 3] = make_float4(l30, l31, l32, 0.f);
                }
                if (tid >= j0 && tid < j0 + 4) sm.zs[tid] = z;
                __syncthreads();
                // ---- B
                const float4 di = reinterpret_cast<const float4*>(sm.d44)[0];
                const float4 la = reinterpret_cast<const float4*>(sm.d44)[1];
                const float4 lb = reinterpret_cast<const float4*>(sm.d44)[2];
                const float l10 = la.x, l20 = la.y, l30 = la.z, l21 = la.w, l31 = lb.x, l32 = lb.y;
                if (owner) {
#pragma unroll
                    for (int i = 0; i < G::TR; ++i) {
                        const int r = row_pos<NB>(t.r0, i);
                        if (r > j0 + 3) {
                            const float m0 = acc[i][g * 4 + 0];
                            const float m1 = fmaf(-m0, l10, acc[i][g * 4 + 1]);
                            const float m2 = fmaf(-m1, l21, fmaf(-m0, l20, acc[i][g * 4 + 2]));
                            const float m3 = fmaf(-m2, l32, fmaf(-m1, l31, fmaf(-m0, l30, acc[i][g * 4 + 3])));
                            R4[r] = make_float4(m0, m1, m2, m3);
                            L4[poff + r - j0] = make_float4(m0 * di.x, m1 * di.y, m2 * di.z, m3 * di.w);
                        }
                    }
                }
                float zq0 = 0.f, zq1 = 0.f, zq2 = 0.f, zq3 = 0.f;
                if (tid < G::DP) {
                    const float4 zz = *reinterpret_cast<const float4*>(sm.zs + j0);
                    zq0 = zz.x;
                    zq1 = fmaf(-l10, zq0, zz.y);
                    zq2 = fmaf(-l21, zq1, fmaf(-l20, zq0, zz.z));
                    zq3 = fmaf(-l32, zq2, fmaf(-l31, zq1, fmaf(-l30, zq0, zz.w)));
                    if (tid >= j0 && tid < j0 + 4) {
                        const int qq = tid - j0;
                        z = qq == 0 ? zq0 : qq == 1 ? zq1 : qq == 2 ? zq2 : zq3;
                        pinv_mine = qq == 0 ? di.x : qq == 1 ? di.y : qq == 2 ? di.z : di.w;
                    }
                }
                __syncthreads();
                // ---- C
                if (tid > j0 + 3 && tid < G::DP) {
                    const float4 lz = L4[poff + tid - j0];
                    z = fmaf(-lz.w, zq3, fmaf(-lz.z, zq2, fmaf(-lz.y, zq1, fmaf(-lz.x, zq0, z))));
                }
                if (rmax > j0 + 3 && cmax > j0 + 3) {
                    const float4* lrow = L4 + poff - j0 + t.r0;     // rows <= j0+3 read earlier panels (valid memory) and are zeroed
                    const bool edge = t.r0 <= j0 + 3 || t.c0 <= j0 + 3;
                    constexpr int CW = 4;                           // columns per pass
#pragma unroll
                    for (int h = 0; h < 8 / CW; ++h) {              // column registers CW*h .. CW*h + CW-1
                        float4 mc[CW];
#pragma unroll
                        for (int jj = 0; jj < CW; ++jj) {
                            const int cq = h * CW + jj, c = t.c0 + (cq & 3) + (cq >> 2) * 32;
                            mc[jj] = R4[c];
                            if (edge && c <= j0 + 3) mc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int i = 0; i < G::TR; ++i) {
                            float4 lr = lrow[(i & 3) + (i >> 2) * G::RS];
                            if (edge && t.r0 + (i & 3) + (i >> 2) * G::RS <= j0 + 3) lr = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int jj = 0; jj < CW; ++jj) {
                                float v = acc[i][h * CW + jj];
                                v = fmaf(-lr.x, mc[jj].x, v); v = fmaf(-lr.y, mc[jj].y, v);
                                v = fmaf(-lr.z, mc[jj].z, v); v = fmaf(-lr.w, mc[jj].w, v);
                                acc[i][h * CW + jj] = v;
                            }
                        }
                    }
                }
                poff += G::DP - j0;
                j0 += 4;
            }
        }
    }
    // back substitution, four rows per round from the bottom: x_j = z_j / D_j - sum_{r > j} L[r][j] x_r
    if (tid < G::DP) {
        float w = z * pinv_mine;
        const int myp = tid >> 2, myq = tid & 3;
        const float* mycol = sm.M + (size_t)(myp * G::DP - 2 * myp * (myp - 1) - 4 * myp) * 4 + myq;   // + 4*r: L[r][tid]
        for (int pp = G::NP - 1; pp >= 0; --pp) {
            const int b0 = pp * 4;
            if (myp == pp) sm.xs[tid] = w;
            named_barrier(1, G::DP);
            const int po = pp * G::DP - 2 * pp * (pp - 1);
            const float4 ww = *reinterpret_cast<const float4*>(sm.xs + b0);
            const float4 r1 = L4[po + 1], r2 = L4[po + 2], r3 = L4[po + 3];
            const float x3 = ww.w;
            const float x2 = fmaf(-r3.z, x3, ww.z);
            const float x1 = fmaf(-r3.y, x3, fmaf(-r2.y, x2, ww.y));
            const float x0 = fmaf(-r3.x, x3, fmaf(-r2.x, x2, fmaf(-r1.x, x1, ww.x)));
            if (myp == pp) x = myq == 0 ? x0 : myq == 1 ? x1 : myq == 2 ? x2 : x3;
            if (myp < pp) {
                const float* c = mycol + 4 * b0;
                w = fmaf(-c[12], x3, fmaf(-c[8], x2, fmaf(-c[4], x1, fmaf(-c[0], x0, w))));
            }
        }
    }
}

// The same factorisation one column per barrier: the variant for NB = 4.  The owners of column j publish it raw (M, packed
// column-major: column j at M + j*DP - j(j-1)/2, `col - j` = the column's virtual row 0), one barrier, every tile below /
// right applies the rank-1 update M_r M_c / D_j; the right-hand side rides along; back substitution walks the packed columns.
// Measured alternatives at this register budget, 24 009 rows of d=256 (this loop: 32.1 ms): the 4-column rounds 46.5 ms and a
// 2-column version 39.7 ms (both spill ~0.6 KB per thread); a blocked right-looking variant -- 64-column panels factored by
// the panel's own warps behind a named barrier, trailing blocks updated 64 columns at a time at the gather loop's density,
// next column published early -- 37.3 ms although it executes 2.6x fewer instructions: a column's critical path is the
// in-order latency of ONE warp's ~100-150 dependent instructions, not the instruction total.
template <int NB>
__device__ __forceinline__ void factor_solve_columns(float (&acc)[Geo<NB>::TR][8], const Tile& t, const Smem<NB>& sm, float z, float& x) {
    using G = Geo<NB>;
    const int tid = threadIdx.x;
    const int rmax = t.r0 + 35, cmax = t.c0 + 35;
    float* col = sm.M;
    int j = 0;
    for (int Jb = 0; Jb < NB; ++Jb) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll 1
            for (int tq = 0; tq < 8; ++tq) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cj = g * 4 + q;
                    if (t.J == Jb && t.tx == tq && rmax >= j) {
                        float* pw = col - j + t.r0;
                        if (t.r0 >= j) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) pw[(i & 3) + (i >> 2) * 32] = acc[i][cj];
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (t.r0 + (i & 3) + (i >> 2) * 32 >= j) pw[(i & 3) + (i >> 2) * 32] = acc[i][cj];
                        }
                    }
                    if (tid == j) sm.zs[j] = z;
                    __syncthreads();
                    const float inv = 1.0f / col[0];
                    if (rmax > j && cmax > j) {
                        const float* pr = col - j + t.r0;     // rows <= j read earlier columns (valid memory) and are zeroed
                        const float* pc = col - j + t.c0;
                        float cr[8], cc[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            cr[i] = pr[(i & 3) + (i >> 2) * 32] * inv;
                            cc[i] = pc[(i & 3) + (i >> 2) * 32];
                        }
                        if (t.r0 <= j || t.c0 <= j) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (t.r0 + (i & 3) + (i >> 2) * 32 <= j) cr[i] = 0.f;
                                if (t.c0 + (i & 3) + (i >> 2) * 32 <= j) cc[i] = 0.f;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(-cr[i], cc[jj], acc[i][jj]);
                    }
                    if (tid > j && tid < G::DP) z = fmaf(-col[tid - j] * inv, sm.zs[j], z);
                    col += G::DP - j;
                    ++j;
                }
            }
        }
    }
    // x_r = (z_r - sum_{q > r} M[q][r] x_q) / D_r, walking r downwards; thread c keeps z_c - (partial sum)
    if (tid < G::DP) {
        const float* mycol = sm.M + ((size_t)tid * G::DP - (size_t)tid * (tid - 1) / 2);
        const float pinv = 1.0f / mycol[0];
        for (int r = G::DP - 1; r >= 0; --r) {
            if (tid == r) { x = z * pinv; sm.xs[r] = x; }
            named_barrier(1, G::DP);
            if (tid < r) z = fmaf(-mycol[r - tid], sm.xs[r], z);
        }
    }
}

// Blocked right-looking A = L D L^T, PW = 16 columns per round (the default at NB = 4), the matrix staying in the threads'
// register tiles.  Round p, columns [16p, 16p + 16):
//   publish  the tile owners of the panel's columns store them k-major into the panel's slot of the factor   -- barrier --
//   panel    threads 16p..DP-1 take one matrix ROW each (16 registers).  The warp holding the diagonal block factors it with
//            shuffles (row i at lane i: pivot, L[i][j], the rank-1 update of a_i[c] by the value of lane c), forward-solves
//            the panel's right-hand-side entries and publishes L (straight into the factor), 1/D and y          -- named barrier --
//            every row below solves its 16 entries against that block on its own (120 FMA, no communication), stores
//            L (into the factor: this IS what back substitution walks) and raw M = L D (transient), takes its share of
//            the forward substitution                                                                  -- barrier --
//   update   every live tile applies the rank-16 update at the gather loop's density (64 FFMA per 4 LDS.128); 32-row / 32-column
//            halves of a tile that the factorisation has passed are skipped (warp-uniform), and so is the upper-triangle
//            quarter of the diagonal blocks.  Entries the factorisation has passed hold garbage from here on: never read.
// Three barriers per 16 columns instead of one per column, and the serial part of a column is a shuffle + an FMA of one warp
// instead of a store / barrier / load round trip of 20.  Back substitution runs panel by panel the same way (the diagonal
// block inside one warp, everybody else one 16-term dot product per panel).
template <int NB, int CG>
__device__ __forceinline__ void blocked_publish(const float (&acc)[Geo<NB>::TR][8], const Tile& t, const Smem<NB>& sm, int p) {
    using G = Geo<NB>;
    if (t.J == (p >> 2) && (t.tx >> 2) == (p & 1)) {
        const int pitch = G::pitch(p);
        float* base = sm.M + G::panel_off(p) + (t.tx & 3) * 4 * pitch - G::PW * p;
#pragma unroll
        for (int rg = 0; rg < 2; ++rg) {
            const int r = t.r0 + rg * 32;
            if (r >= G::PW * p) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    *reinterpret_cast<float4*>(base + kk * pitch + r) =
                        make_float4(acc[rg * 4 + 0][CG * 4 + kk], acc[rg * 4 + 1][CG * 4 + kk], acc[rg * 4 + 2][CG * 4 + kk], acc[rg * 4 + 3][CG * 4 + kk]);
            }
        }
    }
}

template <int NB, bool R0, bool C0, bool DIAG>
__device__ __forceinline__ void blocked_update(float (&acc)[Geo<NB>::TR][8], const float* __restrict__ lrow, int pitch, const float* __restrict__ mcol) {
    using G = Geo<NB>;
#pragma unroll 4
    for (int k = 0; k < G::PW; ++k) {
        const float4 l1 = *reinterpret_cast<const float4*>(lrow + k * pitch + 32);
        const float4 m1 = *reinterpret_cast<const float4*>(mcol + k * G::MT_PITCH + 32);
        const float lv1[4] = {l1.x, l1.y, l1.z, l1.w}, mv1[4] = {m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[4 + i][4 + j] = fmaf(-lv1[i], mv1[j], acc[4 + i][4 + j]);
        if (C0) {
            const float4 m0 = *reinterpret_cast<const float4*>(mcol + k * G::MT_PITCH);
            const float mv0[4] = {m0.x, m0.y, m0.z, m0.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[4 + i][j] = fmaf(-lv1[i], mv0[j], acc[4 + i][j]);
            if (R0) {
                const float4 l0 = *reinterpret_cast<const float4*>(lrow + k * pitch);
                const float lv0[4] = {l0.x, l0.y, l0.z, l0.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(-lv0[i], mv0[j], acc[i][j]);
                if (!DIAG) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][4 + j] = fmaf(-lv0[i], mv1[j], acc[i][4 + j]);
                }
            }
        } else if (R0) {
            const float4 l0 = *reinterpret_cast<const float4*>(lrow + k * pitch);
            const float lv0[4] = {l0.x, l0.y, l0.z, l0.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][4 + j] = fmaf(-lv0[i], mv1[j], acc[i][4 + j]);
        }
    }
}

// the 16 x 16 diagonal block of round p, factored inside the pivot warp (warp 0): row i at lane i, everything by shuffles
template <int NB>
__device__ __forceinline__ void blocked_pivot(const Smem<NB>& sm, int p) {
    using G = Geo<NB>;
    constexpr int PW = G::PW;
    constexpr unsigned FULL = 0xffffffffu;
    const int pitch = G::pitch(p);
    float* const Lp = sm.M + G::panel_off(p) - PW * p;       // Lp[k * pitch + r] = entry (r, 16p + k)
    const int i = threadIdx.x & 31;
    const bool dg = i < PW;
    const int r = PW * p + (dg ? i : 0);
    float a[PW];
#pragma unroll
    for (int k = 0; k < PW; ++k) a[k] = dg ? Lp[k * pitch + r] : 1.f;     // (lanes 16..31 only take part in the shuffles)
    float zi = dg ? sm.zs[r] : 0.f;
    static_for<0, PW>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const float dj = __shfl_sync(FULL, a[j], j);
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(dj));
        inv = fmaf(fmaf(-dj, inv, 1.f), inv, inv);            // one Newton step: within an ulp of 1 / D[j]
        const float lij = a[j] * inv;                         // L[i][j] (rows i > j; garbage in the others, never read)
        const float yj = __shfl_sync(FULL, zi, j);
        if (i == j) { sm.dinv[j] = inv; sm.raw[r] = inv; sm.zs[r] = zi; }   // lane j: 1 / D[j], and its y is final
        zi = fmaf(-lij, yj, zi);                              // (lanes i <= j: already published)
        // rank-1 update of the trailing block, all lanes unconditionally: entries right of the diagonal of a row are never
        // read (lane c's a[j] with c > j is a lower entry), so no predicates; the shuffles go first, then the FMAs
        float m[PW];
        static_for<j + 1, PW>([&](auto cc) { constexpr int c = decltype(cc)::value; m[c] = __shfl_sync(FULL, a[j], c); });   // raw M[c][j]
        static_for<j + 1, PW>([&](auto cc) { constexpr int c = decltype(cc)::value; a[c] = fmaf(-lij, m[c], a[c]); });
        if (dg) Lp[j * pitch + r] = lij;                      // the factor: column j, rows contiguous (read by the rows below and
    });                                                       // by back substitution, rows > j only)
}

// threads that meet at barrier 2 of round p: the warps with rows below the diagonal block and the pivot warp
template <int NB> __device__ __forceinline__ int blocked_b2_count(int p) {
    const int ws = (Geo<NB>::PW * (p + 1)) >> 5;
    return (ws < Geo<NB>::DP / 32 ? Geo<NB>::DP - 32 * ws : 0) + (ws > 0 ? 32 : 0);
}

// a round of the pivot warp once its own tiles (block (0, 0)) are finished: rounds 4 and later.  The register tiles are not
// live in this loop, so the factorisation of the diagonal block has the whole register budget.
template <int NB>
__device__ __forceinline__ void blocked_round_pivot(const Smem<NB>& sm, int p) {
    using G = Geo<NB>;
    named_barrier(4, G::NT);                                  // the panel's columns are published
    blocked_pivot<NB>(sm, p);
    named_barrier(2, blocked_b2_count<NB>(p));
    named_barrier(4, G::NT);                                  // L and M of the panel are in shared memory
}

// a round of the tile threads: rows below the diagonal block (threads 16p+16 .. DP-1, one row each), then the rank-16 update.
// PIVOT: warp 0 in rounds 0..3, when it still holds live tiles and factors the diagonal block as well.
template <int NB, bool PIVOT>
__device__ __forceinline__ void blocked_round(float (&acc)[Geo<NB>::TR][8], const Tile& t, const Smem<NB>& sm, int p, float& z) {
    using G = Geo<NB>;
    constexpr int PW = G::PW;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int pitch = G::pitch(p);
    float* const Mt = sm.stage;
    float* const Lp = sm.M + G::panel_off(p) - PW * p;       // Lp[k * pitch + r] = entry (r, 16p + k)
    named_barrier(4, G::NT);                                  // the panel's columns are published
    const int ws = (PW * (p + 1)) >> 5;
    if (PIVOT) blocked_pivot<NB>(sm, p);
    if (warp >= ws && tid < G::DP) {
        const int i = tid - PW * p;                           // >= 16: a row below the diagonal block
        float a[PW];
#pragma unroll
        for (int k = 0; k < PW; ++k) a[k] = i >= PW ? Lp[k * pitch + tid] : 0.f;     // (rows of the diagonal block are the pivot warp's)
        named_barrier(2, blocked_b2_count<NB>(p));
        if (i >= PW) {
            // M[r][j] = A[r][j] - sum_{q < j} M[r][q] L[j][q], right-looking: once M[r][q] is final it leaves all the later
            // entries (independent FMAs); column q of the diagonal block's L is read from the factor itself, uniformly
            static_for<0, PW - 1>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                static_for<(q + 1) / 4, PW / 4>([&](auto jb) {
                    constexpr int j0 = decltype(jb)::value * 4;
                    const float4 l = *reinterpret_cast<const float4*>(Lp + q * pitch + PW * p + j0);
                    if constexpr (j0 + 0 > q) a[j0 + 0] = fmaf(-a[q], l.x, a[j0 + 0]);
                    if constexpr (j0 + 1 > q) a[j0 + 1] = fmaf(-a[q], l.y, a[j0 + 1]);
                    if constexpr (j0 + 2 > q) a[j0 + 2] = fmaf(-a[q], l.z, a[j0 + 2]);
                    if constexpr (j0 + 3 > q) a[j0 + 3] = fmaf(-a[q], l.w, a[j0 + 3]);
                });
            });
            float zz = z;
#pragma unroll
            for (int k4 = 0; k4 < PW / 4; ++k4) {
                const float4 iv = *reinterpret_cast<const float4*>(sm.dinv + k4 * 4);
                const float4 yv = *reinterpret_cast<const float4*>(sm.zs + PW * p + k4 * 4);
                const float ivv[4] = {iv.x, iv.y, iv.z, iv.w}, yvv[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = k4 * 4 + e;
                    const float l = a[k] * ivv[e];
                    zz = fmaf(-l, yvv[e], zz);
                    Lp[k * pitch + tid] = l;
                    Mt[k * G::MT_PITCH + tid] = a[k];
                }
            }
            z = zz;
            if (i < 2 * PW) sm.zs[tid] = zz;                  // rows of the next diagonal block hand their right-hand side to the pivot warp
        }
    } else if (PIVOT) {
        named_barrier(2, blocked_b2_count<NB>(p));
    }
    named_barrier(4, G::NT);                                  // L and M of the panel are in shared memory
    const int pe = PW * p + PW - 1;
    if (t.I * 64 + 63 > pe && t.J * 64 + 63 > pe) {
        const bool r0live = t.I * 64 + 31 > pe, c0live = t.J * 64 + 31 > pe;     // c0live implies r0live (I >= J)
        const float* lrow = Lp + t.r0;
        const float* mcol = Mt + t.c0;
        if (c0live) {
            if (t.I == t.J) blocked_update<NB, true, true, true>(acc, lrow, pitch, mcol);
            else blocked_update<NB, true, true, false>(acc, lrow, pitch, mcol);
        } else if (r0live) blocked_update<NB, true, false, false>(acc, lrow, pitch, mcol);
        else blocked_update<NB, false, false, false>(acc, lrow, pitch, mcol);
    }
    if (p + 1 < G::NPAN) {                                    // CG of blocked_publish: which 32-column half of the block holds the panel
        if (((p + 1) >> 1) & 1) blocked_publish<NB, 1>(acc, t, sm, p + 1);
        else blocked_publish<NB, 0>(acc, t, sm, p + 1);
    }
}

template <int NB>
__device__ __forceinline__ void factor_solve_blocked(float (&acc)[Geo<NB>::TR][8], const Tile& t, const Smem<NB>& sm, float z, float& x) {
    using G = Geo<NB>;
    constexpr int PW = G::PW;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NP0 = G::NPAN < 4 ? G::NPAN : 4;            // rounds in which block (0, 0) -- the pivot warp's own tiles -- is live
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < G::DP) sm.zs[tid] = z;                          // the right-hand side: rows of a diagonal block are read by the pivot warp
    blocked_publish<NB, 0>(acc, t, sm, 0);
    if (warp == 0) {
#pragma unroll 1
        for (int p = 0; p < NP0; ++p) blocked_round<NB, true>(acc, t, sm, p, z);
#pragma unroll 1
        for (int p = NP0; p < G::NPAN; ++p) blocked_round_pivot<NB>(sm, p);
    } else {
#pragma unroll 1
        for (int p = 0; p < G::NPAN; ++p) blocked_round<NB, false>(acc, t, sm, p, z);
    }
    __syncthreads();                                          // y (zs) and 1/D (raw) of the last panel
    // back substitution: x = L^-T D^-1 y, panel by panel from the bottom
    if (tid < G::DP) {
        float w = sm.zs[tid] * sm.raw[tid];
#pragma unroll 1
        for (int p = G::NPAN - 1; p >= 0; --p) {
            const int wd = (PW * p) >> 5, lb = (PW * p) & 31;
            if (warp == wd) {
                const int i = tid - PW * p;
                const bool mine = i >= 0 && i < PW;
                const float* col = sm.M + G::panel_off(p) + (mine ? i : 0) * G::pitch(p);    // col[q] = L[16p + q][16p + i]
                float lc[PW];
#pragma unroll
                for (int q = 0; q < PW; ++q) lc[q] = col[q];
                static_for<0, PW - 1>([&](auto qc) {
                    constexpr int q = PW - 1 - decltype(qc)::value;     // 15 .. 1
                    const float xq = __shfl_sync(FULL, w, lb + q);
                    if (mine && i < q) w = fmaf(-lc[q], xq, w);
                });
                if (mine) sm.xs[tid] = w;
            }
            named_barrier(1, G::DP);
            if (tid < PW * p) {
                const int pc = tid >> 4;
                const float* col = sm.M + G::panel_off(pc) + (tid & 15) * G::pitch(pc) + PW * (p - pc);
#pragma unroll
                for (int q4 = 0; q4 < PW / 4; ++q4) {
                    const float4 l = *reinterpret_cast<const float4*>(col + q4 * 4);
                    const float4 xv = *reinterpret_cast<const float4*>(sm.xs + PW * p + q4 * 4);
                    w = fmaf(-l.w, xv.w, fmaf(-l.z, xv.z, fmaf(-l.y, xv.y, fmaf(-l.x, xv.x, w))));
                }
            }
        }
        x = w;
    }
}

// Build A from the accumulated Gram, factor it (A = L D L^T, in registers, four columns per round), solve, write the
// row and its loss.  One round for columns j0..j0+3:
//   A  the thread holding the 4x4 diagonal block factors it and publishes 1/D and the six L entries; threads j0..j0+3
//      publish their right-hand-side entries                                                            -- barrier --
//   B  the threads holding those columns eliminate the block from their rows (raw M and scaled L = M/D, a float4 per
//      row each: M into a transient buffer, L into the factor, which back substitution reads later)      -- barrier --
//   C  every tile to the right/below applies the rank-4 update from L (its rows) and M (its columns); the right-hand
//      side does the same (forward substitution rides along).
template <int NB, bool BLK>
__device__ __forceinline__ void finish_row(const RowArgs& p, int row, int64_t n, float (&acc)[Geo<NB>::TR][8], float ssum, const Tile& t,
                                           const Smem<NB>& sm) {
    using G = Geo<NB>;
    const int tid = threadIdx.x, d = p.d;
    if (tid < G::NT) {
        const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(p.base) & 15) == 0;
#pragma unroll
        for (int i = 0; i < G::TR; ++i) {
            const int r = row_pos<NB>(t.r0, i);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = t.c0 + h * 32;                  // four consecutive columns
                if (vec && r < d && c + 3 < d) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.base + (size_t)r * d + c));
                    acc[i][h * 4 + 0] = fmaf(p.amb, acc[i][h * 4 + 0], b4.x) + (r == c + 0 ? p.ridge : 0.f);
                    acc[i][h * 4 + 1] = fmaf(p.amb, acc[i][h * 4 + 1], b4.y) + (r == c + 1 ? p.ridge : 0.f);
                    acc[i][h * 4 + 2] = fmaf(p.amb, acc[i][h * 4 + 2], b4.z) + (r == c + 2 ? p.ridge : 0.f);
                    acc[i][h * 4 + 3] = fmaf(p.amb, acc[i][h * 4 + 3], b4.w) + (r == c + 3 ? p.ridge : 0.f);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int cc = c + j;
                        float v;
                        if (r < d && cc < d) v = fmaf(p.amb, acc[i][h * 4 + j], __ldg(p.base + (size_t)r * d + cc)) + (r == cc ? p.ridge : 0.f);
                        else v = (r == cc) ? 1.f : 0.f;
                        acc[i][h * 4 + j] = v;
                    }
                }
            }
        }
    }
    float pr = 0.f, z = 0.f;
    if (tid < d) {
        if (p.prior) pr = p.prior[(size_t)row * d + tid];
        z = fmaf(p.a, ssum, p.ridge * pr);      // a * sum_p y_p + ridge * prior  (cer.py:43,55)
    }
    if (tid < G::DP) { sm.keep[tid] = z; sm.keep[G::DP + tid] = pr; sm.keep[2 * G::DP + tid] = ssum; }   // read back for the loss
    float x = 0.f;
    if constexpr (BLK) factor_solve_blocked<NB>(acc, t, sm, z, x);
    else if constexpr (NB == 4) factor_solve_columns<NB>(acc, t, sm, z, x);
    else factor_solve_panels<NB>(acc, t, sm, z, x);
    if (tid < G::DP) {
        if (tid < d) p.X[(size_t)row * d + tid] = x;
    }
    __syncthreads();
    float rhs0 = 0.f;
    if (tid < G::DP) { rhs0 = sm.keep[tid]; pr = sm.keep[G::DP + tid]; ssum = sm.keep[2 * G::DP + tid]; }
    row_loss<NB>(p, row, n, true, x, rhs0, ssum, pr, sm.red);
}

template <int NB>
__device__ __forceinline__ void loss_only_row(const RowArgs& p, int row, float* red) {
    const int tid = threadIdx.x;
    float x = 0.f, pr = 0.f;
    if (tid < p.d) {
        x = p.X[(size_t)row * p.d + tid];
        if (p.prior) pr = p.prior[(size_t)row * p.d + tid];
    }
    row_loss<NB>(p, row, 0, false, x, 0.f, 0.f, pr, red);
}

// kernel 1: one block per segment
template <int NB, bool BLK>
__global__ void __launch_bounds__(BLK ? Geo<NB>::NTB : Geo<NB>::NT) __maxnreg__(BLK ? Geo<NB>::MAXREG_B : Geo<NB>::MAXREG) als_segment_kernel(const RowArgs p) {
    using G = Geo<NB>;
    extern __shared__ __align__(16) float smem_raw[];
    Smem<NB> sm(smem_raw);
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int row = p.seg_row[seg], len = p.seg_len[seg], slot = p.seg_slot[seg];
    if (slot < 0 && len == 0 && !p.solve_empty) {
        if (p.loss_rows) loss_only_row<NB>(p, row, sm.red);
        return;
    }
    for (int e = tid; e < G::STAGES * G::KC * G::DP; e += blockDim.x) sm.stage[e] = 0.f;   // pad columns stay zero
    __syncthreads();
    const bool tile_thread = tid < G::NT;                   // (a blocked-variant block has one more warp: the pivot warp)
    const Tile t = tile_of<NB>(tile_thread ? tid : 0);
    float acc[G::TR][8];
#pragma unroll
    for (int i = 0; i < G::TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float ssum = 0.f;
    if (tile_thread) gram_accumulate<NB>(acc, ssum, t, p.Y, p.d, p.idx + p.seg_off[seg], len, sm.stage, t.I == t.J && !p.full_diag);
    if (slot >= 0) {
        if (tile_thread) store_partial<NB>(acc, ssum, p.partial + (size_t)slot * G::PART);
        return;
    }
    if (BLK) __syncthreads();                               // the gather stages are free (the pivot warp was not in the gather's barriers)
    finish_row<NB, BLK>(p, row, len, acc, ssum, t, sm);
}

// kernel 2: one block per split row: ordered sum of its partial slots, then the same finish
template <int NB, bool BLK>
__global__ void __launch_bounds__(BLK ? Geo<NB>::NTB : Geo<NB>::NT) __maxnreg__(BLK ? Geo<NB>::MAXREG_B : Geo<NB>::MAXREG) als_multi_kernel(const RowArgs p) {
    using G = Geo<NB>;
    extern __shared__ __align__(16) float smem_raw[];
    Smem<NB> sm(smem_raw);
    const int m = blockIdx.x, tid = threadIdx.x;
    const int row = p.multi_row[m], slot0 = p.multi_slot0[m], ns = p.multi_nslots[m];
    const bool tile_thread = tid < G::NT;
    const Tile t = tile_of<NB>(tile_thread ? tid : 0);
    float acc[G::TR][8];
#pragma unroll
    for (int i = 0; i < G::TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float ssum = 0.f;
    if (tile_thread)
        for (int s = 0; s < ns; ++s) add_partial<NB>(acc, ssum, p.partial + (size_t)(slot0 + s) * G::PART, tid / G::TPB, tid % G::TPB, true);
    finish_row<NB, BLK>(p, row, p.multi_total[m], acc, ssum, t, sm);
}

// shared Gram: out[d,d] = scale * sum of n_slots partial matrices + ridge * I  (one 64-thread block per 64x64 block)
template <int NB>
__global__ void __launch_bounds__(Geo<NB>::TPB) als_gram_reduce_kernel(const float* partial, int n_slots, int d, float scale, float ridge,
                                                            float* out) {
    using G = Geo<NB>;
    const int blk = blockIdx.x, u = threadIdx.x;
    const Tile t = tile_of<NB>(blk * G::TPB + u);
    float acc[G::TR][8];
#pragma unroll
    for (int i = 0; i < G::TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float dummy = 0.f;
    for (int s = 0; s < n_slots; ++s) add_partial<NB>(acc, dummy, partial + (size_t)s * G::PART, blk, u, false);
#pragma unroll
    for (int i = 0; i < G::TR; ++i) {
        const int r = row_pos<NB>(t.r0, i);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = reg_pos(t.c0, j);
            if (r < d && c < d) {
                const float v = scale * acc[i][j] + (r == c ? ridge : 0.f);
                out[(size_t)r * d + c] = v;
                if (t.I != t.J) out[(size_t)c * d + r] = v;
            }
        }
    }
}

template <int NB> static int set_attr() {
    static thread_local int done_dev = -1;
    int dev = 0;
    TKR_CUDA(cudaGetDevice(&dev));
    if (done_dev != dev) {
        TKR_CUDA(cudaFuncSetAttribute(als_segment_kernel<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NB>::SMEM));
        TKR_CUDA(cudaFuncSetAttribute(als_multi_kernel<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NB>::SMEM));
        TKR_CUDA(cudaFuncSetAttribute(als_segment_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NB>::SMEM));
        TKR_CUDA(cudaFuncSetAttribute(als_multi_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<NB>::SMEM));
        done_dev = dev;
    }
    return TKR_OK;
}

// factorisation variant (tkr_debug_set_als_factor): 0 = default = 1 = the blocked 16-column rounds, 2 = the per-column (NB = 4) /
// four-column loops of round 1 (kept as the reference point the blocked rounds are measured against).  A process global like the other tkr_debug_* switches.
int g_als_factor = 0;

template <int NB> static int launch_rows(const RowArgs& p, int64_t n_segs, int64_t n_multi, cudaStream_t st) {
    using G = Geo<NB>;
    int rc = set_attr<NB>();
    if (rc) return rc;
    const bool blk = g_als_factor != 2;
    if (n_segs > 0) {
        if (blk) als_segment_kernel<NB, true><<<(unsigned)n_segs, G::NTB, G::SMEM, st>>>(p);
        else als_segment_kernel<NB, false><<<(unsigned)n_segs, G::NT, G::SMEM, st>>>(p);
        TKR_LAUNCH_CHECK();
    }
    if (n_multi > 0) {
        if (blk) als_multi_kernel<NB, true><<<(unsigned)n_multi, G::NTB, G::SMEM, st>>>(p);
        else als_multi_kernel<NB, false><<<(unsigned)n_multi, G::NT, G::SMEM, st>>>(p);
        TKR_LAUNCH_CHECK();
    }
    return TKR_OK;
}

template <int NB> static int launch_gram(const RowArgs& p, int64_t n_segs, int d, float scale, float ridge, float* out, cudaStream_t st) {
    using G = Geo<NB>;
    int rc = set_attr<NB>();
    if (rc) return rc;
    als_segment_kernel<NB, false><<<(unsigned)n_segs, G::NT, G::SMEM, st>>>(p);     // every segment is a partial slot: no solve
    TKR_LAUNCH_CHECK();
    als_gram_reduce_kernel<NB><<<G::NBLK, G::TPB, 0, st>>>(p.partial, (int)n_segs, d, scale, ridge, out);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

// segment table of the shared Gram: n_segs equal slices of the row list, every one a partial slot
__global__ void gram_plan_kernel(int64_t n, int n_segs, int64_t per, int32_t* seg_row, int64_t* seg_off, int32_t* seg_len, int32_t* seg_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const int64_t lo = min(n, (int64_t)s * per), hi = min(n, lo + per);
    seg_row[s] = 0; seg_off[s] = lo; seg_len[s] = (int32_t)(hi - lo); seg_slot[s] = s;
}

static inline int nb_of(int d) { return (d + 63) / 64; }
static inline size_t part_floats(int d) { const int nb = nb_of(d); return (size_t)nb * (nb + 1) / 2 * 4096 + 64 * nb; }
constexpr int kGramSegs = 2 * kNumSMs;

}  // namespace als
}  // namespace tkr

using namespace tkr;

extern "C" void tkr_debug_set_als_factor(int32_t mode) { als::g_als_factor = mode; }

extern "C" size_t tkr_als_partial_bytes(int32_t d, int64_t n_slots) {
    if (d < 1 || d > 256 || n_slots < 0) return 0;
    return align_up(als::part_floats(d) * sizeof(float) * (size_t)n_slots, 256);
}

extern "C" size_t tkr_als_gram_workspace_bytes(int32_t d) {
    if (d < 1 || d > 256) return 0;
    return tkr_als_partial_bytes(d, als::kGramSegs) + align_up((size_t)als::kGramSegs * 24, 256);
}

extern "C" int tkr_als_gram(const float* Y, int32_t d, const int32_t* rows, int64_t n_rows, float scale, float ridge,
                            float* out, void* ws, size_t ws_bytes, void* stream) {
    TKR_CHECK_ARG(Y && out && ws && (rows || n_rows == 0), "tkr_als_gram: null pointer");
    TKR_CHECK_ARG(d >= 1 && d <= 256, "tkr_als_gram: d=%d outside [1,256]", d);
    TKR_CHECK_ARG(n_rows >= 0, "tkr_als_gram: n_rows < 0");
    if (ws_bytes < tkr_als_gram_workspace_bytes(d) || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_error("tkr_als_gram: workspace of %zu bytes (256-aligned) needed, got %zu", tkr_als_gram_workspace_bytes(d), ws_bytes);
        return TKR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_segs = (int)std::min<int64_t>(als::kGramSegs, std::max<int64_t>(1, (n_rows + als::KC_ALIGN - 1) / als::KC_ALIGN));
    const int64_t per = ((n_rows + n_segs - 1) / n_segs + als::KC_ALIGN - 1) / als::KC_ALIGN * als::KC_ALIGN;
    char* w = (char*)ws;
    als::RowArgs p = {};
    p.partial = (float*)w; w += tkr_als_partial_bytes(d, als::kGramSegs);
    int64_t* seg_off = (int64_t*)w; w += (size_t)als::kGramSegs * 8;
    int32_t* seg_row = (int32_t*)w; w += (size_t)als::kGramSegs * 4;
    int32_t* seg_len = (int32_t*)w; w += (size_t)als::kGramSegs * 4;
    int32_t* seg_slot = (int32_t*)w;
    als::gram_plan_kernel<<<(n_segs + 127) / 128, 128, 0, st>>>(n_rows, n_segs, std::max<int64_t>(per, als::KC_ALIGN), seg_row, seg_off, seg_len, seg_slot);
    TKR_LAUNCH_CHECK();
    p.Y = Y; p.idx = rows; p.d = d; p.full_diag = 1;
    p.seg_row = seg_row; p.seg_off = seg_off; p.seg_len = seg_len; p.seg_slot = seg_slot;
    switch (als::nb_of(d)) {
        case 1: return als::launch_gram<1>(p, n_segs, d, scale, ridge, out, st);
        case 2: return als::launch_gram<2>(p, n_segs, d, scale, ridge, out, st);
        case 3: return als::launch_gram<3>(p, n_segs, d, scale, ridge, out, st);
        default: return als::launch_gram<4>(p, n_segs, d, scale, ridge, out, st);
    }
}

extern "C" int tkr_als_solve_rows(const tkr_als_cfg* cfg, const tkr_als_plan* plan, const float* Y, float* X, const int32_t* idx,
                                  const float* base, const float* prior, double* loss_rows, void* partial, size_t partial_bytes,
                                  void* stream) {
    TKR_CHECK_ARG(cfg && plan && Y && X && base, "tkr_als_solve_rows: null pointer");
    TKR_CHECK_ARG(cfg->d >= 1 && cfg->d <= 256, "tkr_als_solve_rows: d=%d outside [1,256]", cfg->d);
    TKR_CHECK_ARG(plan->n_segs >= 0 && plan->n_multi >= 0 && plan->n_slots >= 0, "tkr_als_solve_rows: negative plan counts");
    TKR_CHECK_ARG(plan->n_segs == 0 || (plan->seg_row && plan->seg_off && plan->seg_len && plan->seg_slot), "tkr_als_solve_rows: plan segment arrays missing");
    TKR_CHECK_ARG(plan->n_multi == 0 || (plan->multi_row && plan->multi_slot0 && plan->multi_nslots && plan->multi_total),
                  "tkr_als_solve_rows: plan split-row arrays missing");
    if (plan->n_slots > 0 && (!partial || partial_bytes < tkr_als_partial_bytes(cfg->d, plan->n_slots) || (reinterpret_cast<uintptr_t>(partial) & 15))) {
        set_error("tkr_als_solve_rows: partial workspace of %zu bytes needed, got %zu", tkr_als_partial_bytes(cfg->d, plan->n_slots), partial_bytes);
        return TKR_ERR_WORKSPACE;
    }
    als::RowArgs p = {};
    p.Y = Y; p.X = X; p.idx = idx; p.base = base; p.prior = prior;
    p.seg_row = plan->seg_row; p.seg_off = plan->seg_off; p.seg_len = plan->seg_len; p.seg_slot = plan->seg_slot;
    p.multi_row = plan->multi_row; p.multi_slot0 = plan->multi_slot0; p.multi_nslots = plan->multi_nslots; p.multi_total = plan->multi_total;
    p.partial = (float*)partial; p.loss_rows = loss_rows;
    p.a = cfg->a; p.amb = cfg->a - cfg->b; p.ridge = cfg->ridge; p.lreg = cfg->lreg;
    p.d = cfg->d; p.solve_empty = cfg->solve_empty; p.item_loss = cfg->item_loss;
    cudaStream_t st = (cudaStream_t)stream;
    switch (als::nb_of(cfg->d)) {
        case 1: return als::launch_rows<1>(p, plan->n_segs, plan->n_multi, st);
        case 2: return als::launch_rows<2>(p, plan->n_segs, plan->n_multi, st);
        case 3: return als::launch_rows<3>(p, plan->n_segs, plan->n_multi, st);
        default: return als::launch_rows<4>(p, plan->n_segs, plan->n_multi, st);
    }
}
