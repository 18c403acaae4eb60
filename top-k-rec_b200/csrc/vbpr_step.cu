// Hot path 1, VBPR variant (single/vbpr.py:29-74 run by sess.run at :114).
//
// The content model x = rb_i - rb_j + <ur_u, ir_i - ir_j> + <uc_u, (F_i - F_j) E> + (F_i - F_j) c is the BPR
// model on concatenated rows  U' = [ur | uc],  V' = [ir | F.E],  b' = rb + F.c  -- exactly the tables the
// reference exports (vbpr.py:124-126).  One step therefore is
//   vbpr_project_kernel   V'[:, h:] = F.E,  b' = rb + F.c          (fp32 SGEMM over the item table)
//   bpr_grad_kernel       the BPR gather/sigma/scatter step on U', V', b' (item regularisation only on the
//                         ir columns); the gradient landing on the F.E columns is W = d loss / d(F.E),
//                         the one landing on F.c is wq
//   vbpr_grad_dense_kernel dE += F^T W, dc += F^T wq over the touched item rows (split-K, red.add)
//   bpr_apply_kernel      sparse RMSProp on ur|uc rows, ir columns, rb; re-zeroes W / wq
//   vbpr_apply_dense_kernel  + lambda_e E, + lambda_b c, dense RMSProp on E and c (vbpr.py:63-73)
// The reference feeds feat[ib], feat[jb] from the host every step (2 x [B, d] fp32, vbpr.py:114); here F
// stays resident in HBM.
#include "bpr_internal.cuh"

namespace tkr {

// tensor-core route of the two content GEMMs (gemm_tf32x3.cu)
int gemm3_np(int h);
bool gemm3_legal(int M, int K, int h, int64_t a_pitch, const void* A);
int gemm3_build_b(const float* src, int ld, int off, const float* vec, int K, int Kp, int h, float* Bhi, float* Blo, cudaStream_t st);
int gemm3_transpose(const float* F, int M, int Kd, int Mp, float* Ft, cudaStream_t st);
int gemm3_run(int epi, const float* A, int M, int K, int64_t a_pitch, const float* Bhi, const float* Blo, int64_t b_pitch, int h,
              float* out, int ld, int off, float* vec_out, const float* vec_add, int splits, cudaStream_t st);
int g_vbpr_tc_mode = -1;   // -1 automatic, 0 never, 1 same as -1 (tkr_debug_set_vbpr_tc_mode: tests run both routes)

constexpr int GT = 64;   // GEMM tile edge (output tile GT x GT per block)
constexpr int GK = 32;   // K chunk staged per iteration; the four 64-thread groups of a block take 8 k's each

// fp32 SGEMM core shared by the two content GEMMs (both have a 64-wide N: the k/2 content dims).  A block of 256
// threads owns a 64 x 64 output tile; thread (g, ty, tx) of k-group g accumulates an 8 x 8 register tile over
// the k's [8g, 8g + 8) of every staged chunk (two LDS.128 per operand per k for 64 FMAs), plus the extra
// 64 x 1 product with the vector `vs` (the F.c / F^T wq column).  The four groups' partials are then summed
// through shared memory.  As[k][m], Bs[k][n]: the caller stages them.
// RT = output rows per thread: the tile is (8 * RT) x 64, so the row count of a tile can be chosen to make the number
// of tiles fit one wave of the 148 SMs (these kernels run one 256-thread block per SM: ~160 registers per thread).
template <int RT>
struct GemmSmem {
    float As[GK][8 * RT + 4];
    float Bs[GK][GT + 4];
    float vs[GK];
};

template <int RT>
__device__ __forceinline__ void gemm_chunk(const GemmSmem<RT>& sm, int g, int ty, int tx, float (&acc)[RT][8], float (&accq)[RT]) {
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {
        const int kk = g * 8 + k8;
        float a[RT], bv[8];
        if (RT == 8) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 8]), a1 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r) a[r] = sm.As[kk][ty * RT + r];
        }
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 8]), b1 = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 8 + 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        const float v = sm.vs[kk];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
#pragma unroll
            for (int n = 0; n < 8; ++n) acc[r][n] = fmaf(a[r], bv[n], acc[r][n]);
            if (tx == 0) accq[r] = fmaf(a[r], v, accq[r]);
        }
    }
}

// Sum the four k-groups' register tiles: afterwards red[m][n] (and redq[m]) hold the block's (8 RT) x 64 (+ x 1) result.
template <int RT>
__device__ __forceinline__ void gemm_reduce(float (*red)[GT + 1], float* redq, int g, int ty, int tx, const float (&acc)[RT][8], const float (&accq)[RT]) {
    for (int turn = 0; turn < 4; ++turn) {
        if (g == turn) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    float* dst = &red[ty * RT + r][tx * 8 + n];
                    *dst = turn == 0 ? acc[r][n] : *dst + acc[r][n];
                }
                if (tx == 0) redq[ty * RT + r] = turn == 0 ? accq[r] : redq[ty * RT + r] + accq[r];
            }
        }
        __syncthreads();
    }
}

// C[m, n] = sum_k F[m, k] * E[k, n]  (m < M items, n < h), written to Vp[m * ldv + h_off + n];
// q[m] = sum_k F[m, k] * c[k], bsum[m] = rb[m] + q[m].  One block per (8 RT)-row tile.
template <int RT>
__global__ void __launch_bounds__(256) vbpr_project_kernel(const float* __restrict__ F, const float* __restrict__ E,
                                                           const float* __restrict__ c, const float* __restrict__ rb, int M_all,
                                                           int h, int K, float* __restrict__ Vp, int ldv, int h_off,
                                                           float* __restrict__ bsum, const int32_t* __restrict__ rows,
                                                           const int32_t* __restrict__ n_rows_dev, int32_t* __restrict__ row_flag) {
    // rows != NULL: only the listed item rows (the <= 2B rows a small batch touches) are projected; their flags are re-armed
    const int M = rows ? *n_rows_dev : M_all;
    if ((int)blockIdx.x * 8 * RT >= M) return;
    constexpr int TM = 8 * RT, NA = TM * (GK / 4);          // rows per tile; float4 of an A chunk
    constexpr int AIT = (NA + 255) / 256;
    __shared__ GemmSmem<RT> sm;
    __shared__ float red[TM][GT + 1];
    __shared__ float redq[TM];
    const int tid = threadIdx.x, g = tid >> 6, ty = (tid & 63) >> 3, tx = tid & 7;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * GT;
    const bool vec = (K & 3) == 0 && (h & 3) == 0;
    float acc[RT][8] = {};
    float accq[RT] = {};
    // A chunk: TM rows x 32 k of F (row-major), transposed into As[k][row]; B chunk: 32 k x 64 n of E.  The next chunk's
    // global loads are issued before the current chunk is multiplied (register double buffer).
    // split-K over blockIdx.z (row-list mode only: a few hundred rows would otherwise be a few long serial loops); the
    // partial products are then added atomically onto rows the caller has zeroed / initialised to rb
    const int kper = (((K + GK - 1) / GK + gridDim.z - 1) / gridDim.z) * GK;
    const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
    if (kbeg >= kend) return;
    float4 pa[AIT], pb[2];
    float pc = 0.f;
    auto fetch = [&](int k0) {
#pragma unroll
        for (int it = 0; it < AIT; ++it) {
            const int idx = tid + it * 256;
            const int row = idx >> 3, k4 = (idx & 7) * 4;
            const int mi = m0 + row, k = k0 + k4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < NA && mi < M) {
                const int m = rows ? __ldg(rows + mi) : mi;
                const float* src = F + (int64_t)m * K + k;
                if (vec && k + 3 < kend) v = __ldg(reinterpret_cast<const float4*>(src));
                else { if (k < kend) v.x = __ldg(src); if (k + 1 < kend) v.y = __ldg(src + 1); if (k + 2 < kend) v.z = __ldg(src + 2); if (k + 3 < kend) v.w = __ldg(src + 3); }
            }
            pa[it] = v;
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * 256;
            const int kb = idx >> 4, n4 = (idx & 15) * 4;
            const int kk = k0 + kb, n = n0 + n4;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < kend) {
                const float* src = E + (int64_t)kk * h + n;
                if (vec && n + 3 < h) w = __ldg(reinterpret_cast<const float4*>(src));
                else { if (n < h) w.x = __ldg(src); if (n + 1 < h) w.y = __ldg(src + 1); if (n + 2 < h) w.z = __ldg(src + 2); if (n + 3 < h) w.w = __ldg(src + 3); }
            }
            pb[it] = w;
        }
        pc = (tid < GK && k0 + tid < kend) ? __ldg(c + k0 + tid) : 0.f;
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
        for (int it = 0; it < AIT; ++it) {
            const int idx = tid + it * 256;
            const int row = idx >> 3, k4 = (idx & 7) * 4;
            if (idx < NA) { sm.As[k4][row] = pa[it].x; sm.As[k4 + 1][row] = pa[it].y; sm.As[k4 + 2][row] = pa[it].z; sm.As[k4 + 3][row] = pa[it].w; }
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * 256;
            *reinterpret_cast<float4*>(&sm.Bs[idx >> 4][(idx & 15) * 4]) = pb[it];
        }
        if (tid < GK) sm.vs[tid] = pc;
        __syncthreads();
        if (k0 + GK < kend) fetch(k0 + GK);
        gemm_chunk<RT>(sm, g, ty, tx, acc, accq);
        __syncthreads();
    }
    gemm_reduce<RT>(red, redq, g, ty, tx, acc, accq);
    for (int e = tid; e < TM * GT; e += 256) {
        const int r = e >> 6, n = e & 63;
        const int mi = m0 + r, col = n0 + n;
        if (mi < M && col < h) {
            float* dst = Vp + (int64_t)(rows ? __ldg(rows + mi) : mi) * ldv + h_off + col;
            if (gridDim.z > 1) atomicAdd(dst, red[r][n]); else *dst = red[r][n];
        }
    }
    if (blockIdx.y == 0 && tid < TM && m0 + tid < M) {
        const int m = rows ? __ldg(rows + m0 + tid) : m0 + tid;
        if (gridDim.z > 1) atomicAdd(bsum + m, redq[tid]); else bsum[m] = rb[m] + redq[tid];
        if (rows && blockIdx.z == 0) row_flag[m] = 0;
    }
}

// GE[f, n] += sum_r F[row_r, f] * W[row_r, n],  Gc[f] += sum_r F[row_r, f] * wq[row_r]
// over the touched item rows (a list, or every row when rows == nullptr); split-K over blockIdx.z.
__global__ void __launch_bounds__(256) vbpr_grad_dense_kernel(const float* __restrict__ F, int dF, const float* __restrict__ GV,
                                                              int ldv, int h_off, int h, const float* __restrict__ wq,
                                                              const int32_t* __restrict__ rows, const int32_t* __restrict__ n_rows_dev,
                                                              int n_rows_all, float* __restrict__ GE, float* __restrict__ Gc) {
    __shared__ GemmSmem<8> sm;         // As[r][f0..f0+64) = F rows, Bs[r][n0..n0+64) = W rows, vs[r] = wq
    __shared__ float red[GT][GT + 1];
    __shared__ float redq[GT];
    const int tid = threadIdx.x, g = tid >> 6, ty = (tid & 63) >> 3, tx = tid & 7;
    const int f0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
    const int n_rows = rows ? *n_rows_dev : n_rows_all;
    const int chunks = (n_rows + GK - 1) / GK;
    const int per = (chunks + gridDim.z - 1) / gridDim.z;
    const int c_beg = blockIdx.z * per, c_end = min(chunks, c_beg + per);
    if (c_beg >= c_end) return;
    const bool vecF = (dF & 3) == 0, vecW = (ldv & 3) == 0 && (h_off & 3) == 0 && (h & 3) == 0;
    float acc[8][8] = {};
    float accq[8] = {};
    float4 pa[2], pb[2];
    float pq = 0.f;
    auto fetch = [&](int ch) {     // next chunk's rows straight into registers (row ids read from global)
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * 256;
            const int kk = idx >> 4, q4 = (idx & 15) * 4;
            const int rr = ch * GK + kk;
            const int r = rr < n_rows ? (rows ? __ldg(rows + rr) : rr) : -1;
            const int f = f0 + q4, n = n0 + q4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r >= 0) {
                const float* sf = F + (int64_t)r * dF + f;
                if (vecF && f + 3 < dF) v = __ldg(reinterpret_cast<const float4*>(sf));
                else { if (f < dF) v.x = __ldg(sf); if (f + 1 < dF) v.y = __ldg(sf + 1); if (f + 2 < dF) v.z = __ldg(sf + 2); if (f + 3 < dF) v.w = __ldg(sf + 3); }
                const float* sw = GV + (int64_t)r * ldv + h_off + n;
                if (vecW && n + 3 < h) w = *reinterpret_cast<const float4*>(sw);
                else { if (n < h) w.x = sw[0]; if (n + 1 < h) w.y = sw[1]; if (n + 2 < h) w.z = sw[2]; if (n + 3 < h) w.w = sw[3]; }
            }
            pa[it] = v; pb[it] = w;
        }
        if (tid < GK) {
            const int rr = ch * GK + tid;
            const int r = rr < n_rows ? (rows ? __ldg(rows + rr) : rr) : -1;
            pq = r >= 0 ? wq[r] : 0.f;
        }
    };
    fetch(c_beg);
    for (int ch = c_beg; ch < c_end; ++ch) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * 256;
            *reinterpret_cast<float4*>(&sm.As[idx >> 4][(idx & 15) * 4]) = pa[it];
            *reinterpret_cast<float4*>(&sm.Bs[idx >> 4][(idx & 15) * 4]) = pb[it];
        }
        if (tid < GK) sm.vs[tid] = pq;
        __syncthreads();
        if (ch + 1 < c_end) fetch(ch + 1);
        gemm_chunk<8>(sm, g, ty, tx, acc, accq);
        __syncthreads();
    }
    gemm_reduce<8>(red, redq, g, ty, tx, acc, accq);
    for (int e = tid; e < GT * GT; e += 256) {
        const int r = e >> 6, n = e & 63;
        const int f = f0 + r, col = n0 + n;
        if (f < dF && col < h && red[r][n] != 0.f) atomicAdd(GE + (int64_t)f * h + col, red[r][n]);
    }
    if (blockIdx.y == 0 && tid < GT && f0 + tid < dF && redq[tid] != 0.f) atomicAdd(Gc + f0 + tid, redq[tid]);
}

// Dense optimiser step on E[dF*h] and c[dF] (vbpr.py:63-73: the E and c regularisers are NOT per occurrence);
// adds their regularisation value to the step's loss; re-zeroes the gradient buffers.
__global__ void __launch_bounds__(256) vbpr_apply_dense_kernel(tkr_bpr_cfg cfg, float lambda_e, float* __restrict__ E,
                                                               float* __restrict__ msE, float* __restrict__ GE, int64_t nE,
                                                               float* __restrict__ c, float* __restrict__ msc,
                                                               float* __restrict__ Gc, int64_t nc, float* __restrict__ loss_out) {
    const bool l1 = cfg.l1 != 0;
    float reg = 0.f;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nE + nc; t += (int64_t)gridDim.x * blockDim.x) {
        const bool isE = t < nE;
        float* var = isE ? E + t : c + (t - nE);
        float* ms = isE ? msE + t : msc + (t - nE);
        float* G = isE ? GE + t : Gc + (t - nE);
        const float lam = isE ? lambda_e : cfg.lambda_b;
        const float v = *var;
        const float g = *G + (l1 ? lam * (float)((v > 0.f) - (v < 0.f)) : lam * v);
        reg += l1 ? lam * fabsf(v) : 0.5f * lam * v * v;
        if (cfg.optimizer == TKR_OPT_RMSPROP) {
            const float m = cfg.rms_decay * *ms + (1.0f - cfg.rms_decay) * g * g;
            *ms = m;
            *var = v - cfg.lr * g / sqrtf(m + cfg.rms_eps);
        } else {
            *var = v - cfg.lr * g;
        }
        *G = 0.f;
    }
    if (loss_out != nullptr) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) reg += __shfl_xor_sync(0xffffffffu, reg, s);
        if ((threadIdx.x & 31) == 0 && reg != 0.f) atomicAdd(loss_out, reg);
    }
}

// ---- VBPR with the graph as written (D-14): x[a, b] = r_a + y_b ------------------------------------------------------
// r_n = b'_i - b'_j (b' = rb + F.c), y_n = <U'_u, V'_i - V'_j> over all k columns (V' = [ir | F.E]); one warp per triple.
__global__ void __launch_bounds__(256) vbpr_xparts_kernel(const float* __restrict__ U, const float* __restrict__ V, const float* __restrict__ bsum,
                                                          const int32_t* __restrict__ ub, const int32_t* __restrict__ ib, const int32_t* __restrict__ jb,
                                                          int B, int d, float* __restrict__ r, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < B; n += (gridDim.x * blockDim.x) >> 5) {
        const int u = ub[n], i = ib[n], j = jb[n];
        float acc = 0.f;
        for (int c = lane; c < d; c += 32) acc = fmaf(U[(int64_t)u * d + c], V[(int64_t)i * d + c] - V[(int64_t)j * d + c], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) { y[n] = acc; r[n] = bsum[i] - bsum[j]; }
    }
}
// block n: s_emb[n] = sum_a sigma(-(r_a + y_n)), s_bias[n] = sum_b sigma(-(r_n + y_b)); the B*B softplus terms go to *loss
__global__ void __launch_bounds__(256) vbpr_pair_kernel(const float* __restrict__ r, const float* __restrict__ y, int B, float* __restrict__ s_emb,
                                                        float* __restrict__ s_bias, float* __restrict__ loss) {
    __shared__ float red[3][8];
    const int n = blockIdx.x;
    const float yn = y[n], rn = r[n];
    float se = 0.f, sb = 0.f, ls = 0.f;
    for (int a = threadIdx.x; a < B; a += blockDim.x) {
        const float x = r[a] + yn;                                        // column n of x
        se += __fdividef(1.0f, 1.0f + __expf(x));
        ls += fmaxf(-x, 0.f) + __logf(1.0f + __expf(-fabsf(x)));
        sb += __fdividef(1.0f, 1.0f + __expf(rn + y[a]));                 // row n of x
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); ls += __shfl_xor_sync(0xffffffffu, ls, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = se; red[1][threadIdx.x >> 5] = sb; red[2][threadIdx.x >> 5] = ls; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a0 += red[0][w]; a1 += red[1][w]; a2 += red[2][w]; }
        s_emb[n] = a0; s_bias[n] = a1;
        if (loss != nullptr) atomicAdd(loss, a2);
    }
}
constexpr int64_t kPairwiseMaxBatch = 4096;        // the literal graph is O(B^2) per step

struct VbprWs {
    StepWs s; float* wq; float* GE; float* Gc; size_t total;
    // tensor-core route (large batches, dense features): F^T built once per workspace, the pre-split B operands per step
    bool tc; int Mp; float* Ft; float* Bp_hi; float* Bp_lo; float* Bg_hi; float* Bg_lo; unsigned long long* ft_tag;
    int32_t* tflag; int32_t* tlist; int32_t* tcount; int32_t* trip;   // small batches: flags / list / count of the batch's item rows, triples drawn ahead
    float* pair;             // cfg.pairwise: [4][B] floats r, y, s_emb, s_bias  + the triples drawn ahead when the sampler is fused [3][B] int32
};

// The content GEMMs go to the tensor cores when every item row is (potentially) touched each step -- the dense mode of
// large batches -- and the shapes fit the TMA descriptors; small batches keep the list-driven CUDA-core kernels, which
// only read the <= 2B touched feature rows.
static bool vbpr_tc_shape(const tkr_vbpr_cfg* cfg, int64_t B) {
    const int h = cfg->base.d / 2;
    return g_vbpr_tc_mode != 0 && bpr_pick_mode(&cfg->base, B, 0) == MODE_DENSE && cfg->d_feat % 4 == 0 && cfg->d_feat >= 64 &&
           cfg->base.n_items >= 128 && gemm3_np(h) <= 256;
}

static size_t vbpr_ws_bytes(const tkr_vbpr_cfg* cfg, int64_t B) {
    const size_t h = cfg->base.d / 2;
    size_t n = align_up(bpr_ws_total(&cfg->base, B), 256) + align_up((size_t)cfg->base.n_items * 4, 256) +
               align_up((size_t)cfg->d_feat * h * 4, 256) + align_up((size_t)cfg->d_feat * 4, 256);
    if (vbpr_tc_shape(cfg, B)) {
        const size_t Mp = align_up((size_t)cfg->base.n_items, 4), NP = (size_t)gemm3_np((int)h);
        n += align_up((size_t)cfg->d_feat * Mp * 4, 1024) + 2 * align_up(NP * cfg->d_feat * 4, 1024) + 2 * align_up(NP * Mp * 4, 1024) + 1024 + 1024;
    }
    if (cfg->pairwise) n += align_up((size_t)7 * B * 4, 1024) + 1024;
    if (bpr_pick_mode(&cfg->base, B, 0) == MODE_LIST)
        n += align_up((size_t)cfg->base.n_items * 4, 1024) + align_up((size_t)2 * B * 4, 1024) + 1024 + align_up((size_t)3 * B * 4, 1024) + 1024;
    return n;
}

static int vbpr_carve(const tkr_vbpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes, VbprWs* out) {
    const size_t need = vbpr_ws_bytes(cfg, B);
    if (ws == nullptr || ws_bytes < need) { set_error("vbpr workspace too small: have %zu, need %zu", ws_bytes, need); return TKR_ERR_WORKSPACE; }
    if (int rc = bpr_carve(&cfg->base, B, ws, ws_bytes, &out->s)) return rc;
    const size_t h = cfg->base.d / 2;
    char* p = (char*)ws + align_up(bpr_ws_total(&cfg->base, B), 256);
    out->wq = (float*)p; p += align_up((size_t)cfg->base.n_items * 4, 256);
    out->GE = (float*)p; p += align_up((size_t)cfg->d_feat * h * 4, 256);
    out->Gc = (float*)p; p += align_up((size_t)cfg->d_feat * 4, 256);
    out->tc = vbpr_tc_shape(cfg, B);
    if (out->tc) {
        const size_t Mp = align_up((size_t)cfg->base.n_items, 4), NP = (size_t)gemm3_np((int)h);
        p = (char*)align_up((size_t)(uintptr_t)p, 1024);
        out->Mp = (int)Mp;
        out->ft_tag = (unsigned long long*)p; p += 1024;
        out->Ft = (float*)p; p += align_up((size_t)cfg->d_feat * Mp * 4, 1024);
        out->Bp_hi = (float*)p; p += align_up(NP * cfg->d_feat * 4, 1024);
        out->Bp_lo = (float*)p; p += align_up(NP * cfg->d_feat * 4, 1024);
        out->Bg_hi = (float*)p; p += align_up(NP * Mp * 4, 1024);
        out->Bg_lo = (float*)p; p += align_up(NP * Mp * 4, 1024);
    }
    out->pair = nullptr;
    if (cfg->pairwise) { p = (char*)align_up((size_t)(uintptr_t)p, 1024); out->pair = (float*)p; p += align_up((size_t)7 * B * 4, 1024); }
    out->tflag = nullptr;
    if (bpr_pick_mode(&cfg->base, B, 0) == MODE_LIST) {
        p = (char*)align_up((size_t)(uintptr_t)p, 1024);
        out->tflag = (int32_t*)p; p += align_up((size_t)cfg->base.n_items * 4, 1024);
        out->tlist = (int32_t*)p; p += align_up((size_t)2 * B * 4, 1024);
        out->tcount = (int32_t*)p; p += 1024;
        out->trip = (int32_t*)p;
    }
    out->total = need;
    return TKR_OK;
}

static int vbpr_check(const tkr_vbpr_cfg* cfg, int64_t B) {
    TKR_CHECK_ARG(cfg != nullptr, "cfg is NULL");
    if (int rc = bpr_check_cfg(&cfg->base, B)) return rc;
    TKR_CHECK_ARG(cfg->base.d % 2 == 0, "VBPR needs an even k (k/2 rating + k/2 content dims, vbpr.py:37-44), got %d", cfg->base.d);
    TKR_CHECK_ARG(cfg->d_feat > 0, "d_feat must be positive");
    TKR_CHECK_ARG(!cfg->pairwise || B <= kPairwiseMaxBatch, "pairwise (the [B,B] graph of vbpr.py:61 as written) is O(B^2): batch %lld > %lld",
                  (long long)B, (long long)kPairwiseMaxBatch);
    return TKR_OK;
}

// distinct item rows of a batch: first toucher of a row appends it (flag 0 -> 1)
__global__ void __launch_bounds__(256) vbpr_touch_kernel(const int32_t* __restrict__ ib, const int32_t* __restrict__ jb, int B,
                                                         int32_t* __restrict__ flag, int32_t* __restrict__ list, int32_t* __restrict__ n) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * B; t += gridDim.x * blockDim.x) {
        const int r = t < B ? ib[t] : jb[t - B];
        if (atomicExch(flag + r, 1) == 0) list[atomicAdd(n, 1)] = r;
    }
}
// prepares the listed rows for the split-K projection: content columns zeroed, bsum = rb (one warp per row)
__global__ void __launch_bounds__(256) vbpr_rows_init_kernel(const int32_t* __restrict__ list, const int32_t* __restrict__ n, float* __restrict__ V,
                                                             int ldv, int h, const float* __restrict__ rb, float* __restrict__ bsum) {
    const int lane = threadIdx.x & 31, cnt = *n;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < cnt; w += (gridDim.x * blockDim.x) >> 5) {
        const int r = list[w];
        for (int c = lane; c < h; c += 32) V[(int64_t)r * ldv + h + c] = 0.f;
        if (lane == 0) bsum[r] = rb[r];
    }
}

static void launch_project(const tkr_vbpr_cfg* cfg, const float* F, const float* E, const float* c, const float* rb, float* V,
                           float* bsum, cudaStream_t st, const int32_t* rows = nullptr, const int32_t* n_rows_dev = nullptr,
                           int32_t* row_flag = nullptr, int max_rows = 0) {
    const int h = cfg->base.d / 2, M = rows ? max_rows : cfg->base.n_items;
    // rows per tile: the smallest multiple of 8 in [64, 96] that fits the item table into one wave of blocks; a row list
    // (small batches) gets small tiles instead, so that a few hundred rows still spread over the chip
    int rt = ((M + kNumSMs - 1) / kNumSMs + 7) / 8;
    if (rows != nullptr) rt = rt <= 1 ? 1 : rt <= 2 ? 2 : rt <= 4 ? 4 : 8;
    else if (rt < 8 || rt > 12) rt = 8;
    const unsigned gy = (unsigned)((h + GT - 1) / GT);
    // row lists: split K so that tiles x splits cover the chip several times over (each block is a short latency-bound loop)
    unsigned gz = 1;
    if (rows != nullptr) {
        const int tiles = (M + 8 * rt - 1) / (8 * rt) * (int)gy, kchunks = (cfg->d_feat + GK - 1) / GK;
        int z = (8 * kNumSMs + tiles - 1) / tiles;
        gz = (unsigned)(z < 1 ? 1 : z > kchunks ? kchunks : z > 32 ? 32 : z);
    }
#define TKR_PROJ(RT) vbpr_project_kernel<RT><<<dim3((unsigned)((M + 8 * RT - 1) / (8 * RT)), gy, gz), 256, 0, st>>>(F, E, c, rb, cfg->base.n_items, h, cfg->d_feat, V, cfg->base.d, h, bsum, rows, n_rows_dev, row_flag)
    switch (rt) {
        case 1: TKR_PROJ(1); break;
        case 2: TKR_PROJ(2); break;
        case 4: TKR_PROJ(4); break;
        case 9: TKR_PROJ(9); break;
        case 10: TKR_PROJ(10); break;
        case 11: TKR_PROJ(11); break;
        case 12: TKR_PROJ(12); break;
        default: TKR_PROJ(8); break;
    }
#undef TKR_PROJ
}

}  // namespace tkr

using namespace tkr;

extern "C" void tkr_debug_set_vbpr_tc_mode(int32_t m) { g_vbpr_tc_mode = m < -1 || m > 1 ? -1 : m; }

extern "C" size_t tkr_vbpr_workspace_bytes(const tkr_vbpr_cfg* cfg, int64_t B) {
    if (cfg == nullptr || B <= 0 || cfg->base.n_users <= 0 || cfg->base.n_items <= 0 || cfg->base.d <= 0 || cfg->d_feat <= 0) return 0;
    return vbpr_ws_bytes(cfg, B);
}

extern "C" int tkr_vbpr_workspace_init(const tkr_vbpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = vbpr_check(cfg, B)) return rc;
    VbprWs w;
    if (int rc = vbpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    TKR_CUDA(cudaMemsetAsync(ws, 0, w.total, (cudaStream_t)stream));
    return TKR_OK;
}

extern "C" int tkr_vbpr_project(const tkr_vbpr_cfg* cfg, const float* F, const float* E, const float* c, const float* rb,
                                float* V, float* bsum, void* stream) {
    if (int rc = vbpr_check(cfg, 1)) return rc;
    TKR_CHECK_ARG(F && E && c && rb && V && bsum, "NULL pointer");
    launch_project(cfg, F, E, c, rb, V, bsum, (cudaStream_t)stream);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}

namespace tkr {
enum { VBPR_ALL = 0, VBPR_GRAD = 1, VBPR_APPLY = 2 };

// The step loop of tkr_vbpr_step; with phase = VBPR_GRAD / VBPR_APPLY (n_steps = 1) it runs the two halves of one step for
// data-parallel training: everything up to the gradients (projection, gather/scatter step, LOCAL dE / dc), then -- after the
// caller has summed [GV|Gb|tchV] and [GE|Gc] over the ranks -- the sparse and dense updates.
static int vbpr_run(int phase, int data_parallel, const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* bsum, float* E, float* c,
                    const float* F, float* msU, float* msV, float* msrb, float* msE, float* msc, const int32_t* u,
                    const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const tkr_sampler* smp,
                    uint64_t first_draw, float* loss_out, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = vbpr_check(cfg, B)) return rc;
    TKR_CHECK_ARG(U && V && rb && E && c, "U, V, rb, E, c must not be NULL");
    TKR_CHECK_ARG(phase == VBPR_APPLY || (bsum && F), "bsum and F must not be NULL");
    TKR_CHECK_ARG(phase == VBPR_GRAD || cfg->base.optimizer == TKR_OPT_SGD || (msU && msV && msrb && msE && msc), "RMSProp needs every rms slot");
    TKR_CHECK_ARG(n_steps >= 0, "n_steps < 0");
    TKR_CHECK_ARG(!(cfg->pairwise && data_parallel), "pairwise (x over all pairs of the batch) cannot be split over ranks");
    SamplerDev sd = {};
    if (phase != VBPR_APPLY) {
        if (u == nullptr) {
            if (int rc = bpr_make_sampler(smp, &sd)) return rc;
            TKR_CHECK_ARG(smp->n_items == cfg->base.n_items, "sampler n_items != cfg n_items");
        } else {
            TKR_CHECK_ARG(i && j, "i, j must not be NULL when u is given");
        }
    }
    VbprWs w;
    if (int rc = vbpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const tkr_bpr_cfg* bc = &cfg->base;
    const int h = bc->d / 2, dF = cfg->d_feat;
    const int mode = bpr_pick_mode(bc, B, data_parallel);
    const StepExtra ex{h, rb, w.wq, hot_rows_for(bc->d)};    // popular item rows privatised per block, as in the plain BPR step
    if (phase == VBPR_ALL && loss_out != nullptr && n_steps > 0) TKR_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float) * (size_t)n_steps, st));
    // split-K so that the dE GEMM covers the chip: (dF/64) x (h/64) output tiles
    const int tiles = ((dF + GT - 1) / GT) * ((h + GT - 1) / GT);
    int splitk = (2 * kNumSMs + tiles - 1) / tiles;
    if (splitk < 1) splitk = 1;
    if (splitk > 64) splitk = 64;
    // tensor-core route: F must be 16-byte aligned for the TMA descriptor; F^T is built once per (workspace, F)
    const bool tc = phase != VBPR_APPLY && w.tc && mode == MODE_DENSE && gemm3_legal(bc->n_items, dF, h, dF, F);
    if (tc) {
        unsigned long long tag = 0;
        TKR_CUDA(cudaMemcpyAsync(&tag, w.ft_tag, 8, cudaMemcpyDeviceToHost, st));
        TKR_CUDA(cudaStreamSynchronize(st));
        if (tag != (unsigned long long)(uintptr_t)F) {            // (a workspace is zero-initialised: tag 0 = not built)
            if (int rc = gemm3_transpose(F, bc->n_items, dF, w.Mp, w.Ft, st)) return rc;
            tag = (unsigned long long)(uintptr_t)F;
            TKR_CUDA(cudaMemcpyAsync(w.ft_tag, &tag, 8, cudaMemcpyHostToDevice, st));
            TKR_CUDA(cudaStreamSynchronize(st));
        }
    }
    // split-K of the gradient GEMM: as many K ranges as fit ONE wave of CTAs (160 CTAs on 148 SMs ran as two waves)
    const int tc_splits = tc ? (kNumSMs / ((dF + 127) / 128) > 0 ? kNumSMs / ((dF + 127) / 128) : 1) : 1;
    auto project = [&]() -> int {
        if (tc) {
            if (int rc = gemm3_build_b(E, h, 0, c, dF, dF, h, w.Bp_hi, w.Bp_lo, st)) return rc;
            return gemm3_run(0, F, bc->n_items, dF, dF, w.Bp_hi, w.Bp_lo, dF, h, V, bc->d, h, bsum, rb, 1, st);
        }
        launch_project(cfg, F, E, c, rb, V, bsum, st);
        TKR_LAUNCH_CHECK();
        return TKR_OK;
    };
    for (int64_t t = 0; t < n_steps; ++t) {
        float* lt = loss_out ? loss_out + t : nullptr;
        if (phase != VBPR_APPLY) {
            const int32_t *ut = u ? u + t * B : nullptr, *it = u ? i + t * B : nullptr, *jt = u ? j + t * B : nullptr;
            if (mode == MODE_LIST && !tc && w.tflag != nullptr) {
                // small batch (the reference's 256): project only the <= 2B item rows the batch touches -- F.[E|c] over all
                // 10 000 items was 5.3 GFLOP per step at C3, 20x what the batch reads.  The triples must be known first:
                // with the fused sampler they are drawn ahead (same draws).
                if (u == nullptr) {
                    if (int rc = tkr_bpr_sample(smp, first_draw + (uint64_t)t * (uint64_t)B, B, w.trip, w.trip + B, w.trip + 2 * B, stream)) return rc;
                    ut = w.trip; it = w.trip + B; jt = w.trip + 2 * B;
                }
                TKR_CUDA(cudaMemsetAsync(w.tcount, 0, 4, st));
                vbpr_touch_kernel<<<(unsigned)((2 * B + 255) / 256), 256, 0, st>>>(it, jt, (int)B, w.tflag, w.tlist, w.tcount);
                TKR_LAUNCH_CHECK();
                vbpr_rows_init_kernel<<<(unsigned)((2 * B + 7) / 8), 256, 0, st>>>(w.tlist, w.tcount, V, bc->d, h, rb, bsum);
                TKR_LAUNCH_CHECK();
                const int max_rows = (int)(2 * B < bc->n_items ? 2 * B : bc->n_items);
                launch_project(cfg, F, E, c, rb, V, bsum, st, w.tlist, w.tcount, w.tflag, max_rows);
                TKR_LAUNCH_CHECK();
            } else if (int rc = project()) return rc;
            StepExtra ext = ex;
            if (cfg->pairwise) {
                float* pr = w.pair;                                       // r | y | s_emb | s_bias, then the staged triples
                if (ut == nullptr) {                                      // fused sampler: the same draws, made ahead (x needs all of them)
                    int32_t* su = (int32_t*)(pr + 4 * B);
                    if (int rc = tkr_bpr_sample(smp, first_draw + (uint64_t)t * (uint64_t)B, B, su, su + B, su + 2 * B, stream)) return rc;
                    ut = su; it = su + B; jt = su + 2 * B;
                }
                int64_t xb = (B + 7) / 8;
                if (xb > kNumSMs * 8) xb = kNumSMs * 8;
                vbpr_xparts_kernel<<<(unsigned)xb, 256, 0, st>>>(U, V, bsum, ut, it, jt, (int)B, bc->d, pr, pr + B);
                TKR_LAUNCH_CHECK();
                vbpr_pair_kernel<<<(unsigned)B, 256, 0, st>>>(pr, pr + B, (int)B, pr + 2 * B, pr + 3 * B, lt);
                TKR_LAUNCH_CHECK();
                ext.s_emb = pr + 2 * B; ext.s_bias = pr + 3 * B;
            }
            if (int rc = bpr_dispatch_grad(bc, U, V, bsum, ut, it, jt, B, sd, first_draw + (uint64_t)t * (uint64_t)B, w.s, mode, ext, lt, st)) return rc;
            if (tc) {   // [dE | dc] += F^T . [W | wq] over all items (untouched rows of W are zero)
                if (int rc = gemm3_build_b(w.s.GV, bc->d, h, w.wq, bc->n_items, w.Mp, h, w.Bg_hi, w.Bg_lo, st)) return rc;
                if (int rc = gemm3_run(1, w.Ft, dF, bc->n_items, w.Mp, w.Bg_hi, w.Bg_lo, w.Mp, h, w.GE, h, 0, w.Gc, nullptr, tc_splits, st)) return rc;
            } else {
                dim3 grid((dF + GT - 1) / GT, (h + GT - 1) / GT, splitk);
                vbpr_grad_dense_kernel<<<grid, 256, 0, st>>>(F, dF, w.s.GV, bc->d, h, h, w.wq, mode == MODE_LIST ? w.s.listV : nullptr,
                                                             w.s.n_touched + 1, bc->n_items, w.GE, w.Gc);
                TKR_LAUNCH_CHECK();
            }
        }
        if (phase != VBPR_GRAD) {
            bpr_launch_apply(bc, U, V, rb, msU, msV, msrb, B, w.s, mode, ex, st);
            TKR_LAUNCH_CHECK();
            const int64_t nE = (int64_t)dF * h;
            int64_t blocks = (nE + dF + 255) / 256;
            if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
            vbpr_apply_dense_kernel<<<(unsigned)blocks, 256, 0, st>>>(*bc, cfg->lambda_e, E, msE, w.GE, nE, c, msc, w.Gc, dF, lt);
            TKR_LAUNCH_CHECK();
        }
    }
    // leave V[:, h:] = F.E and bsum = rb + F.c consistent with the final E, c: they ARE the export (vbpr.py:124-126)
    if (phase == VBPR_ALL)
        if (int rc = project()) return rc;
    return TKR_OK;
}
}  // namespace tkr

extern "C" int tkr_vbpr_step(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* bsum, float* E, float* c,
                             const float* F, float* msU, float* msV, float* msrb, float* msE, float* msc, const int32_t* u,
                             const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const tkr_sampler* smp,
                             uint64_t first_draw, float* loss_out, void* ws, size_t ws_bytes, void* stream) {
    return vbpr_run(VBPR_ALL, 0, cfg, U, V, rb, bsum, E, c, F, msU, msV, msrb, msE, msc, u, i, j, B, n_steps, smp, first_draw, loss_out, ws, ws_bytes, stream);
}

extern "C" int tkr_vbpr_grad(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* bsum, float* E, float* c, const float* F,
                             const int32_t* u, const int32_t* i, const int32_t* j, int64_t B, const tkr_sampler* smp, uint64_t first_draw,
                             float* loss_out, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream) {
    return vbpr_run(VBPR_GRAD, data_parallel, cfg, U, V, rb, bsum, E, c, F, nullptr, nullptr, nullptr, nullptr, nullptr, u, i, j, B, 1, smp, first_draw,
                    loss_out, ws, ws_bytes, stream);
}

extern "C" int tkr_vbpr_apply(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* E, float* c, float* msU, float* msV, float* msrb,
                              float* msE, float* msc, int64_t B, float* loss_out, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream) {
    return vbpr_run(VBPR_APPLY, data_parallel, cfg, U, V, rb, nullptr, E, c, nullptr, msU, msV, msrb, msE, msc, nullptr, nullptr, nullptr, B, 1, nullptr, 0,
                    loss_out, ws, ws_bytes, stream);
}

// byte offsets of the regions a data-parallel caller sums over the ranks: offsets[0..1] = [begin, end) of the fp32 region
// [GV | Gb | tchV] (the sparse item side), offsets[2..3] = [begin, end) of the fp32 region [GE | (pad) | Gc] (dense content side)
extern "C" int tkr_vbpr_workspace_layout(const tkr_vbpr_cfg* cfg, int64_t B, int64_t* offsets) {
    if (int rc = vbpr_check(cfg, B)) return rc;
    TKR_CHECK_ARG(offsets != nullptr, "offsets is NULL");
    int64_t o[TKR_WS_NFIELDS];
    if (int rc = tkr_bpr_workspace_layout(&cfg->base, B, o)) return rc;
    const size_t h = cfg->base.d / 2, ni = cfg->base.n_items;
    offsets[0] = o[TKR_WS_GV];
    offsets[1] = o[TKR_WS_GV] + (int64_t)(ni * cfg->base.d + 2 * ni) * 4;
    const size_t ge = align_up(bpr_ws_total(&cfg->base, B), 256) + align_up(ni * 4, 256);
    offsets[2] = (int64_t)ge;
    offsets[3] = (int64_t)(ge + align_up((size_t)cfg->d_feat * h * 4, 256) + (size_t)cfg->d_feat * 4);
    return TKR_OK;
}
