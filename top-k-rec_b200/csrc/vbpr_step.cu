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

constexpr int GT = 64;   // GEMM tile edge
constexpr int GK = 16;   // K chunk

// C[m, n] = sum_k F[m, k] * E[k, n]  (m < M items, n < h), written to Vp[m * ldv + h_off + n];
// q[m] = sum_k F[m, k] * c[k], bsum[m] = rb[m] + q[m].
__global__ void __launch_bounds__(256) vbpr_project_kernel(const float* __restrict__ F, const float* __restrict__ E,
                                                           const float* __restrict__ c, const float* __restrict__ rb, int M,
                                                           int h, int K, float* __restrict__ Vp, int ldv, int h_off,
                                                           float* __restrict__ bsum) {
    __shared__ float As[GK][GT + 4];
    __shared__ float Bs[GK][GT + 4];
    __shared__ float cs[GK];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
    float acc[4][4] = {};
    float accq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += GK) {
        {   // A chunk: 64 rows x 16 k, transposed into As[k][row]
            const int row = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m0 + row, k = k0 + kq + e;
                As[kq + e][row] = (m < M && k < K) ? __ldg(F + (int64_t)m * K + k) : 0.f;
            }
            // B chunk: 16 k x 64 n
            const int kk = tid >> 4, nq = (tid & 15) * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + kk, n = n0 + nq + e;
                Bs[kk][nq + e] = (k < K && n < h) ? __ldg(E + (int64_t)k * h + n) : 0.f;
            }
            if (tid < GK) cs[tid] = (k0 + tid < K) ? __ldg(c + k0 + tid) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
            for (int n = 0; n < 4; ++n) bv[n] = Bs[kk][tx * 4 + n];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(a[r], bv[n], acc[r][n]);
                if (tx == 0) accq[r] = fmaf(a[r], cs[kk], accq[r]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty * 4 + r;
        if (m >= M) continue;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int col = n0 + tx * 4 + n;
            if (col < h) Vp[(int64_t)m * ldv + h_off + col] = acc[r][n];
        }
        if (tx == 0 && blockIdx.y == 0) bsum[m] = rb[m] + accq[r];
    }
}

// GE[f, n] += sum_r F[row_r, f] * W[row_r, n],  Gc[f] += sum_r F[row_r, f] * wq[row_r]
// over the touched item rows (a list, or every row when rows == nullptr); split-K over blockIdx.z.
__global__ void __launch_bounds__(256) vbpr_grad_dense_kernel(const float* __restrict__ F, int dF, const float* __restrict__ GV,
                                                              int ldv, int h_off, int h, const float* __restrict__ wq,
                                                              const int32_t* __restrict__ rows, const int32_t* __restrict__ n_rows_dev,
                                                              int n_rows_all, float* __restrict__ GE, float* __restrict__ Gc) {
    __shared__ float As[GK][GT + 4];   // F[row, f0..f0+64)
    __shared__ float Bs[GK][GT + 4];   // W[row, n0..n0+64)
    __shared__ float ws_[GK];
    __shared__ int rs[GK];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int f0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
    const int n_rows = rows ? *n_rows_dev : n_rows_all;
    const int chunks = (n_rows + GK - 1) / GK;
    const int per = (chunks + gridDim.z - 1) / gridDim.z;
    const int c_beg = blockIdx.z * per, c_end = min(chunks, c_beg + per);
    float acc[4][4] = {};
    float accq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ch = c_beg; ch < c_end; ++ch) {
        if (tid < GK) {
            const int r = ch * GK + tid;
            rs[tid] = r < n_rows ? (rows ? rows[r] : r) : -1;
        }
        __syncthreads();
        {
            const int kk = tid >> 4, q4 = (tid & 15) * 4;
            const int r = rs[kk];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int f = f0 + q4 + e, n = n0 + q4 + e;
                As[kk][q4 + e] = (r >= 0 && f < dF) ? __ldg(F + (int64_t)r * dF + f) : 0.f;
                Bs[kk][q4 + e] = (r >= 0 && n < h) ? GV[(int64_t)r * ldv + h_off + n] : 0.f;
            }
            if (tid < GK) ws_[tid] = rs[tid] >= 0 ? wq[rs[tid]] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
            for (int n = 0; n < 4; ++n) bv[n] = Bs[kk][tx * 4 + n];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(a[r], bv[n], acc[r][n]);
                if (tx == 0) accq[r] = fmaf(a[r], ws_[kk], accq[r]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int f = f0 + ty * 4 + r;
        if (f >= dF) continue;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int col = n0 + tx * 4 + n;
            if (col < h && acc[r][n] != 0.f) atomicAdd(GE + (int64_t)f * h + col, acc[r][n]);
        }
        if (tx == 0 && blockIdx.y == 0 && accq[r] != 0.f) atomicAdd(Gc + f, accq[r]);
    }
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

struct VbprWs { StepWs s; float* wq; float* GE; float* Gc; size_t total; };

static size_t vbpr_ws_bytes(const tkr_vbpr_cfg* cfg, int64_t B) {
    const size_t h = cfg->base.d / 2;
    return align_up(bpr_ws_total(&cfg->base, B), 256) + align_up((size_t)cfg->base.n_items * 4, 256) +
           align_up((size_t)cfg->d_feat * h * 4, 256) + align_up((size_t)cfg->d_feat * 4, 256);
}

static int vbpr_carve(const tkr_vbpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes, VbprWs* out) {
    const size_t need = vbpr_ws_bytes(cfg, B);
    if (ws == nullptr || ws_bytes < need) { set_error("vbpr workspace too small: have %zu, need %zu", ws_bytes, need); return TKR_ERR_WORKSPACE; }
    if (int rc = bpr_carve(&cfg->base, B, ws, ws_bytes, &out->s)) return rc;
    const size_t h = cfg->base.d / 2;
    char* p = (char*)ws + align_up(bpr_ws_total(&cfg->base, B), 256);
    out->wq = (float*)p; p += align_up((size_t)cfg->base.n_items * 4, 256);
    out->GE = (float*)p; p += align_up((size_t)cfg->d_feat * h * 4, 256);
    out->Gc = (float*)p;
    out->total = need;
    return TKR_OK;
}

static int vbpr_check(const tkr_vbpr_cfg* cfg, int64_t B) {
    TKR_CHECK_ARG(cfg != nullptr, "cfg is NULL");
    if (int rc = bpr_check_cfg(&cfg->base, B)) return rc;
    TKR_CHECK_ARG(cfg->base.d % 2 == 0, "VBPR needs an even k (k/2 rating + k/2 content dims, vbpr.py:37-44), got %d", cfg->base.d);
    TKR_CHECK_ARG(cfg->d_feat > 0, "d_feat must be positive");
    return TKR_OK;
}

static void launch_project(const tkr_vbpr_cfg* cfg, const float* F, const float* E, const float* c, const float* rb, float* V,
                           float* bsum, cudaStream_t st) {
    const int h = cfg->base.d / 2;
    dim3 grid((cfg->base.n_items + GT - 1) / GT, (h + GT - 1) / GT);
    vbpr_project_kernel<<<grid, 256, 0, st>>>(F, E, c, rb, cfg->base.n_items, h, cfg->d_feat, V, cfg->base.d, h, bsum);
}

}  // namespace tkr

using namespace tkr;

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

extern "C" int tkr_vbpr_step(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* bsum, float* E, float* c,
                             const float* F, float* msU, float* msV, float* msrb, float* msE, float* msc, const int32_t* u,
                             const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const tkr_sampler* smp,
                             uint64_t first_draw, float* loss_out, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = vbpr_check(cfg, B)) return rc;
    TKR_CHECK_ARG(U && V && rb && bsum && E && c && F, "U, V, rb, bsum, E, c, F must not be NULL");
    TKR_CHECK_ARG(cfg->base.optimizer == TKR_OPT_SGD || (msU && msV && msrb && msE && msc), "RMSProp needs every rms slot");
    TKR_CHECK_ARG(n_steps >= 0, "n_steps < 0");
    SamplerDev sd = {};
    if (u == nullptr) {
        if (int rc = bpr_make_sampler(smp, &sd)) return rc;
        TKR_CHECK_ARG(smp->n_items == cfg->base.n_items, "sampler n_items != cfg n_items");
    } else {
        TKR_CHECK_ARG(i && j, "i, j must not be NULL when u is given");
    }
    VbprWs w;
    if (int rc = vbpr_carve(cfg, B, ws, ws_bytes, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const tkr_bpr_cfg* bc = &cfg->base;
    const int h = bc->d / 2, dF = cfg->d_feat;
    const int mode = bpr_pick_mode(bc, B, 0);
    const StepExtra ex{h, rb, w.wq};
    if (loss_out != nullptr && n_steps > 0) TKR_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float) * (size_t)n_steps, st));
    // split-K so that the dE GEMM covers the chip: (dF/64) x (h/64) output tiles
    const int tiles = ((dF + GT - 1) / GT) * ((h + GT - 1) / GT);
    int splitk = (2 * kNumSMs + tiles - 1) / tiles;
    if (splitk < 1) splitk = 1;
    if (splitk > 64) splitk = 64;
    for (int64_t t = 0; t < n_steps; ++t) {
        float* lt = loss_out ? loss_out + t : nullptr;
        launch_project(cfg, F, E, c, rb, V, bsum, st);
        TKR_LAUNCH_CHECK();
        if (int rc = bpr_dispatch_grad(bc, U, V, bsum, u ? u + t * B : nullptr, u ? i + t * B : nullptr, u ? j + t * B : nullptr, B, sd,
                                       first_draw + (uint64_t)t * (uint64_t)B, w.s, mode, ex, lt, st)) return rc;
        dim3 grid((dF + GT - 1) / GT, (h + GT - 1) / GT, splitk);
        vbpr_grad_dense_kernel<<<grid, 256, 0, st>>>(F, dF, w.s.GV, bc->d, h, h, w.wq, mode == MODE_LIST ? w.s.listV : nullptr,
                                                     w.s.n_touched + 1, bc->n_items, w.GE, w.Gc);
        TKR_LAUNCH_CHECK();
        bpr_launch_apply(bc, U, V, rb, msU, msV, msrb, B, w.s, mode, ex, st);
        TKR_LAUNCH_CHECK();
        const int64_t nE = (int64_t)dF * h;
        int64_t blocks = (nE + dF + 255) / 256;
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        vbpr_apply_dense_kernel<<<(unsigned)blocks, 256, 0, st>>>(*bc, cfg->lambda_e, E, msE, w.GE, nE, c, msc, w.Gc, dF, lt);
        TKR_LAUNCH_CHECK();
    }
    // leave V[:, h:] = F.E and bsum = rb + F.c consistent with the final E, c: they ARE the export (vbpr.py:124-126)
    launch_project(cfg, F, E, c, rb, V, bsum, st);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}
