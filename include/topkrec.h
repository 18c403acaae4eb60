/* topkrec.h -- C ABI of libtopkrec.so, the B200 (sm_100a) engine under the two
 * hot paths of domainxz/top-k-rec.
 *
 * The reference has NO FFI: the seams this ABI sits under are Python calls into
 * third-party numerics.  Each entry point cites the reference line it replaces:
 *
 *   tkr_bpr_step / _host   <- sess.run([solver, obj], feed_dict={u,i,j})   single/bpr.py:141
 *                             (graph = single/bpr.py:71-101, RMSProp :100; SGD = old/methods/bpr.py:57-61)
 *   tkr_bpr_sample         <- BPR._uniform_user_sampling                    single/bpr.py:155-165
 *   tkr_vbpr_step/_project <- sess.run(..., feed_dict={u,i,j,ic,jc})        single/vbpr.py:114 (graph :29-74, export :124-126)
 *   tkr_score_topk / _host <- np.dot + np.argsort + rated-filter walk       evaluate.py:78,81,96-105
 *   tkr_eval_hits          <- the hits[] accumulation of the evaluation walk       evaluate.py:84-112
 *   tkr_topk_merge         <- (no reference equivalent: merges item-sharded candidates so the
 *                              sharded result equals the single-process one)
 *
 * Conventions
 *   - Caller owns every buffer.  Pointers are DEVICE pointers unless the
 *     parameter name ends in _host.  The library never allocates persistent
 *     memory; scratch comes from the caller (`*_workspace_bytes`).
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered and
 *     asynchronous unless the name ends in _host (those synchronise the stream
 *     before returning because they write host memory).
 *   - Return 0 on success, a negative TKR_ERR_* otherwise; the message is in
 *     thread-local storage, `tkr_last_error()`.  No C++ exception crosses.
 *   - Re-entrant across streams/devices.  The compute entry points keep no global
 *     mutable state but the TLS error string and launch counter.  The tkr_debug_*
 *     setters (profiling / test aids: filter counters, filter mode, seed fraction,
 *     count mode) write PROCESS-WIDE variables without synchronisation: set them
 *     from one thread while no other thread is inside the library.
 */
#ifndef TOPKREC_H
#define TOPKREC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TKR_VERSION 10100 /* 1.1.0 */

#define TKR_OK 0
#define TKR_ERR_INVALID (-1)     /* bad argument */
#define TKR_ERR_WORKSPACE (-2)   /* workspace too small / misaligned */
#define TKR_ERR_CUDA (-3)        /* CUDA runtime error */
#define TKR_ERR_UNSUPPORTED (-4) /* shape outside what the kernels are built for */

int tkr_version(void);
const char* tkr_last_error(void);
/* number of kernels launched by this library on the calling thread since the
 * last tkr_reset_launch_count() -- bench.py reports it as gpu_launches */
int64_t tkr_launch_count(void);
void tkr_reset_launch_count(void);

/* ------------------------------------------------------------------ path 1 */

#define TKR_OPT_RMSPROP 0 /* tf.train.RMSPropOptimizer defaults, single/bpr.py:100 */
#define TKR_OPT_SGD 1     /* old/methods/bpr.py:57-61 */

typedef struct tkr_bpr_cfg {
    int32_t n_users, n_items, d;                  /* d = embedding width k of the reference */
    float lambda_u, lambda_i, lambda_j, lambda_b; /* single/bpr.py:20 */
    float lr;
    float rms_decay; /* 0.9  */
    float rms_eps;   /* 1e-10, inside the sqrt */
    int32_t l1;      /* 0: mode=='l2' (bpr.py:92-95); 1: L1 (bpr.py:96-99) */
    int32_t optimizer;
} tkr_bpr_cfg;

/* Device-side sampler tables (CSR of positives; users with >= 1 positive). */
typedef struct tkr_sampler {
    const int32_t* tr_users;   /* [n_tr_users] user rows that have positives (bpr.py:67) */
    int32_t n_tr_users;
    const int64_t* pos_indptr; /* [n_users + 1] */
    const int32_t* pos_idx;    /* positives of each user, ascending within a user */
    int32_t n_items;
    uint64_t seed;
} tkr_sampler;

/* Scratch for a step of `batch` triples.  Must be zero-filled once before the
 * first step (tkr_bpr_workspace_init); every step leaves it zeroed again. */
size_t tkr_bpr_workspace_bytes(const tkr_bpr_cfg* cfg, int64_t batch);
int tkr_bpr_workspace_init(const tkr_bpr_cfg* cfg, int64_t batch, void* ws, size_t ws_bytes, void* stream);
/* Optional: name up to TKR_MAX_HOT popular item rows (HOST array of distinct item ids, e.g. the items with the most
 * positives).  The gradient kernel then sums their gradients in shared memory per thread block and adds each block's
 * sum once, instead of one L2 atomic per occurrence -- popular items otherwise serialise on one L2 slice.  Results are
 * the same sums in another order.  n = 0 clears the set. */
int tkr_bpr_workspace_set_hot_items(const tkr_bpr_cfg* cfg, int64_t batch, void* ws, size_t ws_bytes,
                                    const int32_t* item_ids_host, int32_t n, void* stream);

/* Byte offsets of the workspace regions, for callers that exchange gradients
 * between devices (data-parallel training): offsets[TKR_WS_*]. The fp32 region
 * [GV | Gb | TCHV] is contiguous ((n_items*d + 2*n_items) floats from offsets[TKR_WS_GV]). */
#define TKR_WS_GU 0
#define TKR_WS_CNTU 1
#define TKR_WS_LISTU 2
#define TKR_WS_NTOUCHED 3
#define TKR_WS_GV 4
#define TKR_WS_GB 5
#define TKR_WS_TCHV 6
#define TKR_WS_CNTV 7
#define TKR_WS_LISTV 8
#define TKR_WS_HOTV 9      /* int32 hot_slot[n_items] (0 = cold, s+1 = privatised slot s) then int32 hot_ids[TKR_MAX_HOT] */
#define TKR_WS_STAGE 10    /* batches <= 1024: 256 B of barrier words + int32[3][65536] triples sampled ahead for the persistent multi-step kernel */
#define TKR_WS_TOTAL 11
#define TKR_WS_NFIELDS 12
#define TKR_MAX_HOT 32
int tkr_bpr_workspace_layout(const tkr_bpr_cfg* cfg, int64_t batch, int64_t* offsets);

/* The two halves of a step, for data-parallel training: users are
 * partitioned across ranks (U rows never exchanged), V/b replicated.
 *   tkr_bpr_grad   gathers + accumulates the summed per-row gradients of `batch` triples
 *                  into the workspace (loss added to *loss_out, which the caller zeroes);
 *   -- with data_parallel != 0 the caller now all-reduces (sum) the contiguous fp32
 *      region [GV | Gb | TCHV] across ranks --
 *   tkr_bpr_apply  one optimiser update per touched row; with data_parallel != 0 item
 *                  rows are those whose summed occurrence count TCHV is non-zero, so
 *                  every replica applies the identical update.
 * tkr_bpr_step == tkr_bpr_grad + tkr_bpr_apply with data_parallel = 0. */
int tkr_bpr_grad(const tkr_bpr_cfg* cfg, const float* U, const float* V, const float* b, const int32_t* u,
                 const int32_t* i, const int32_t* j, int64_t batch, const tkr_sampler* smp, uint64_t first_draw,
                 float* loss_out, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream);
int tkr_bpr_apply(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                  int64_t batch, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream);

/* ---- multi-GPU plumbing: peer-mapped exchange buffers (one process per GPU, one box) ----
 * Replaces nothing in the reference (it is single-process); this is how the sharded paths of SURVEY 8(e)
 * move data: every rank allocates an exchange buffer with tkr_peer_alloc (the one place the library allocates device
 * memory, explicitly and on request; zero-filled), exports its 64-byte CUDA-IPC handle, the handles travel over any
 * host channel (topkrec/peer.py uses torch.distributed), every rank imports the others' and hands the table of mapped
 * base addresses to the fused kernels, which then load from / store to peer HBM over NVLink.  All ranks use the same
 * buffer layout.  tkr_peer_release unmaps an imported buffer, tkr_peer_free frees an own one. */
#define TKR_MAX_PEERS 8
#define TKR_PEER_HANDLE_BYTES 64
typedef struct tkr_peers {
    int32_t rank, world;
    void* base[TKR_MAX_PEERS]; /* base[p] = rank p's exchange buffer as mapped here; base[rank] = the local allocation */
} tkr_peers;
int tkr_peer_alloc(size_t bytes, void** out);
int tkr_peer_free(void* p);
int tkr_peer_export(void* p, void* handle_out /* TKR_PEER_HANDLE_BYTES, host */);
int tkr_peer_import(const void* handle /* host */, void** out);
int tkr_peer_release(void* p);

/* Data-parallel step with the exchange fused into the update (replaces grad + NCCL all-reduce +
 * apply): users are partitioned over the ranks (U, msU rows never move), the item side lives in the exchange buffer:
 *     [ V[n_items,d] | b[n_items] | G0 | G1 | flags ]   G* = [GV[n_items,d] | Gb[n_items] | tch[n_items]] (fp32)
 * (byte offsets from tkr_bpr_dp_layout: TKR_DP_V, _B, _G0, _G1, _FLAGS, _TOTAL).  One call =
 *   bpr_grad_kernel on this rank's triples (item gradients into G[epoch & 1], user gradients into the workspace), then
 *   ONE kernel in which: a cross-GPU barrier waits for every rank's gradients; this rank, owner of the item rows
 *   r % world == rank, LOADS those rows of every peer's G over NVLink, sums them in rank order, applies the optimiser
 *   update once (slots msV/msb of owned rows are local) and STORES the new V row / bias into every rank's buffer;
 *   meanwhile other thread blocks apply this rank's user rows and re-zero G[(epoch+1) & 1]; a second barrier closes the
 *   step.  Every replica of V / b therefore holds identical bits, equal to one GPU stepping the union batch up to fp32
 *   summation order.  `epoch` must be the same on every rank and grow by 1 per call, starting at 1.  msV / msb rows are
 *   only meaningful on their owner.  A rank that never arrives trips a 20 s device-side timeout, reported by
 *   tkr_bpr_dp_status (< 0) instead of hanging the GPU. */
#define TKR_DP_V 0
#define TKR_DP_B 1
#define TKR_DP_G0 2
#define TKR_DP_G1 3
#define TKR_DP_FLAGS 4
#define TKR_DP_TOTAL 5
#define TKR_DP_NFIELDS 6
int tkr_bpr_dp_layout(const tkr_bpr_cfg* cfg, int64_t* offsets);
int tkr_bpr_dp_step(const tkr_bpr_cfg* cfg, float* U, float* msU, float* msV, float* msb, const int32_t* u,
                    const int32_t* i, const int32_t* j, int64_t batch, const tkr_sampler* smp, uint64_t first_draw,
                    float* loss_out, void* ws, size_t ws_bytes, const tkr_peers* peers, uint64_t epoch, void* stream);
int tkr_bpr_dp_status(const tkr_bpr_cfg* cfg, const tkr_peers* peers, void* stream); /* synchronises; 0 = no timeout so far */

/* n_steps consecutive synchronous mini-batch steps.  Step t uses triples
 * [t*batch, (t+1)*batch) of u/i/j and writes the batch objective evaluated
 * before the update to loss_out[t] (what sess.run returns for `obj`).
 * All B gradients are taken at the pre-step snapshot, duplicate rows are
 * summed, then one optimiser update per touched row.
 * State: U[n_users,d], V[n_items,d], b[n_items] and the RMSProp `rms` slots
 * msU/msV/msb of the same shapes (ignored for SGD; may be NULL then).
 * If u == NULL the triples are drawn on the device from `smp` with the
 * counter-based generator of tkr_bpr_sample (draw index = first_draw + t*batch + n)
 * fused into the gradient kernel; i/j may then be NULL too. */
int tkr_bpr_step(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                 const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, int64_t n_steps,
                 const tkr_sampler* smp, uint64_t first_draw, float* loss_out, void* ws, size_t ws_bytes,
                 void* stream);

/* Barrier-free ("Hogwild") plain-SGD steps (SURVEY 8(f) NEXT-4; update rule of old/methods/bpr.py:57-61 without its batch
 * synchrony): ONE kernel per step adds -lr * gradient of every occurrence straight onto the parameter rows while other warps
 * read them.  A throughput mode: not bit-reproducible, no parity claim beyond "equals tkr_bpr_step(optimizer = SGD) when no row
 * occurs twice in a batch".  No workspace, no slots; same triple / sampler / loss conventions as tkr_bpr_step. */
int tkr_bpr_hogwild(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, const int32_t* u, const int32_t* i, const int32_t* j,
                    int64_t batch, int64_t n_steps, const tkr_sampler* smp, uint64_t first_draw, float* loss_out, void* stream);

/* Same, with the triples and the losses in HOST memory (the feed_dict /
 * fetch of sess.run): copies u/i/j host->device into `staging` (device,
 * >= 3*batch*n_steps*4 bytes), runs the steps, copies loss back, synchronises. */
int tkr_bpr_step_host(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                      const int32_t* u_host, const int32_t* i_host, const int32_t* j_host, int64_t batch,
                      int64_t n_steps, float* loss_host, void* staging, size_t staging_bytes, void* ws,
                      size_t ws_bytes, void* stream);

/* VBPR (single/vbpr.py): content-aware BPR with item features F[n_items, d_feat] resident on the device.
 * State in the layout the reference exports (vbpr.py:124-126):
 *   U[n_users, k] = [ur | uc],  V[n_items, k] = [ir | F.E],  rb[n_items] (trainable rating bias),
 *   bsum[n_items] = rb + F.c,  E[d_feat, k/2],  c[d_feat];  rms slots msU, msV (ir columns used), msrb, msE, msc.
 * cfg.base.d = k (even).  tkr_vbpr_project refreshes V[:, k/2:] and bsum from E, c (call it once after
 * initialising / importing E, c, rb); tkr_vbpr_step runs n_steps synchronous steps exactly like
 * tkr_bpr_step (same triple / sampler / loss conventions) and leaves V, bsum projected with the final E, c.
 * Large batches (every item potentially touched per step) with dense features run the two content GEMMs on the tensor
 * cores (tcgen05 kind::tf32, 3-term split: fp32-level accuracy); the workspace then also holds F^T, built at the first
 * step and keyed by the address of F -- re-initialise the workspace if the CONTENTS of F change under the same address. */
typedef struct tkr_vbpr_cfg {
    tkr_bpr_cfg base;
    int32_t d_feat;
    float lambda_e; /* single/vbpr.py:18 */
    /* 0: per-triple objective x_n = r_n + y_n (what the code evidently means; any batch size).
     * 1: the graph exactly as written: the bias variables are [n_items,1] / [d,1], so vbpr.py:61 broadcasts x to [B,B],
     *    x[a,b] = r_a + y_b with r = rb_i - rb_j + (f_i - f_j).c and y = x_ui - x_uj, and the loss sums over all B*B entries
     *    (reference defect D-14).  O(B^2) per step: batch <= 4096. */
    int32_t pairwise;
} tkr_vbpr_cfg;
size_t tkr_vbpr_workspace_bytes(const tkr_vbpr_cfg* cfg, int64_t batch);
int tkr_vbpr_workspace_init(const tkr_vbpr_cfg* cfg, int64_t batch, void* ws, size_t ws_bytes, void* stream);
int tkr_vbpr_project(const tkr_vbpr_cfg* cfg, const float* F, const float* E, const float* c, const float* rb,
                     float* V, float* bsum, void* stream);
int tkr_vbpr_step(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* bsum, float* E, float* c,
                  const float* F, float* msU, float* msV, float* msrb, float* msE, float* msc, const int32_t* u,
                  const int32_t* i, const int32_t* j, int64_t batch, int64_t n_steps, const tkr_sampler* smp,
                  uint64_t first_draw, float* loss_out, void* ws, size_t ws_bytes, void* stream);

/* The two halves of one VBPR step for data-parallel training (users partitioned over the ranks, every
 * other table replicated):  tkr_vbpr_grad = projection + gather/scatter gradients + this rank's dE / dc; the caller then sums
 * over the ranks the two fp32 regions named by tkr_vbpr_workspace_layout -- offsets[0..1) = [GV|Gb|tchV] (item rows, incl. the
 * content gradient W), offsets[2..3) = [GE|Gc] (dE and dc: linear in W, so the sum of the local products is the product of the
 * sum) --; tkr_vbpr_apply = the sparse and dense optimiser updates, identical on every replica.  With data_parallel != 0 item
 * rows are flagged through tchV as in tkr_bpr_grad / tkr_bpr_apply.  Call tkr_vbpr_project after the last step before
 * exporting V / bsum. */
int tkr_vbpr_grad(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* bsum, float* E, float* c, const float* F,
                  const int32_t* u, const int32_t* i, const int32_t* j, int64_t batch, const tkr_sampler* smp, uint64_t first_draw,
                  float* loss_out, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream);
int tkr_vbpr_apply(const tkr_vbpr_cfg* cfg, float* U, float* V, float* rb, float* E, float* c, float* msU, float* msV, float* msrb,
                   float* msE, float* msc, int64_t batch, float* loss_out, void* ws, size_t ws_bytes, int32_t data_parallel, void* stream);
int tkr_vbpr_workspace_layout(const tkr_vbpr_cfg* cfg, int64_t batch, int64_t* offsets /* [4] */);

/* Draw `n` triples (draw indices first_draw .. first_draw+n-1) with the
 * semantics of single/bpr.py:155-165: user uniform over tr_users with
 * replacement, positive uniform over the user's positives, negative uniform
 * over [0, n_items) redrawn while it is a positive of the user.  Counter-based
 * (Philox4x32-10 keyed by seed): reproducible and order-independent. */
int tkr_bpr_sample(const tkr_sampler* smp, uint64_t first_draw, int64_t n, int32_t* u_out, int32_t* i_out,
                   int32_t* j_out, void* stream);

/* ------------------------------------------------------------------ path 2 */

/* For every user row r of U[nu,d] the first k columns c of V[ni,d], in the
 * order (score desc, column desc), that are not in the user's rated list:
 *   score(r,c) = fp32 fma chain over the d products in ascending index order,
 *                + bias[c] (if bias != NULL), + 0.0f
 * Columns are reported as c + col_offset (item-sharded callers pass the shard's
 * first global column; rated_idx holds global columns, ascending per user,
 * rated_indptr[nu+1] indexes it).  Unused slots: idx -1, score -inf.
 * k <= 64.  Output [nu,k] each.  The score matrix is never written to memory. */
size_t tkr_score_topk_workspace_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k);
int tkr_score_topk(const float* U, int64_t nu, const float* V, int64_t ni, int32_t d, const float* bias,
                   const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset,
                   int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, void* stream);

/* Same contract and bit-identical results, computed on the tensor cores: BF16 tcgen05 GEMM with a
 * fused in-SM candidate filter (TMA-fed, accumulators in TMEM, scores never written), exact fp32
 * re-scoring of the <= 64 survivors per row, a per-row error-bound certificate, and the exact kernel
 * above for the rows that cannot be certified (their count is written to *n_fallback_rows, a DEVICE
 * int32, if not NULL).  Shapes the filter does not cover (d + 3 > 256 after padding, k > 48) are
 * routed to tkr_score_topk entirely.
 * items_prepared != 0: the previous call on this workspace used the same V / bias / ni / d / k and the same
 * nu, so its BF16 item table is reused (an evaluator scores many user batches against one item table). */
size_t tkr_score_topk_tc_workspace_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k, int32_t has_bias);
int tkr_score_topk_tc(const float* U, int64_t nu, const float* V, int64_t ni, int32_t d, const float* bias,
                      const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset,
                      int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int32_t* n_fallback_rows,
                      int32_t items_prepared, void* stream);

/* Profiling aid: device buffer of [n_ctas][14 warps][4] int64 cycle counters filled by the filter kernel
 * (total / wait cycles per warp role); NULL (default) disables it.  While counters are set,
 * tkr_debug_set_filter_mode selects 1 = normal, 2 = epilogue never reads TMEM (TMA + MMA ceiling probe),
 * 3 = epilogue drains TMEM without scanning (TMEM read ceiling probe); 2 and 3 return after the filter
 * kernel without producing lists. */
void tkr_debug_set_filter_counters(long long* dev_buf);
/* tkr_bpr_step path choice: -1 automatic (default), 0 never / 1 always (when legal) take the counting path that
 * updates rows occurring once in a batch in place; both paths follow the same step semantics. */
void tkr_debug_set_count_mode(int32_t mode);
/* tkr_bpr_step route for small batches (<= 1024 triples, d <= 256), many steps per launch: -1 automatic (default: the dataflow
 * kernel -- row-level version words, no grid-wide barriers -- for batches <= 64 and 257..1024, the cluster kernel with two grid
 * barriers per step for 65..256), 0 never (two launches per step), 1 the cluster kernel whenever legal, 2 the dataflow kernel
 * whenever legal; all routes follow the same step semantics. */
void tkr_debug_set_persist_mode(int32_t mode);
/* profiling aid: device int64[8] receiving warp 0's cycles per phase of the persistent kernel (gather+gradient, slot
 * prefetch, barrier 1, update, barrier 2, steps); NULL (default) disables it */
void tkr_debug_set_persist_counters(long long* dev_buf);
void tkr_debug_set_filter_mode(int32_t mode);
/* filter tuning aid: seed_rank 3 | 4 (0 = automatic), cap_trigger in [72, 128] (0 = automatic); results do not depend on them */
void tkr_debug_set_filter_tuning(int32_t seed_rank, int32_t cap_trigger);
/* tkr_vbpr_step content GEMMs: -1 automatic (tcgen05 3xTF32 route for large batches with dense 16-byte-aligned features), 0 never
 * (fp32 CUDA-core GEMMs).  Read when the workspace is SIZED as well: keep it fixed between tkr_vbpr_workspace_bytes and the steps. */
void tkr_debug_set_vbpr_tc_mode(int32_t mode);
/* experiments on the 3xTF32 GEMM: bit 0 = overwrite the A tile with its TF32-exact part, bit 1 = three separate products */
void tkr_debug_set_gemm3_flags(int32_t flags);
/* ALS factorisation variant: 0 = default (blocked 16-column rounds for d > 192, four-column rounds below), 1 = blocked at every
 * width, 2 = the per-column / four-column loops at every width (als_solve.cu) */
void tkr_debug_set_als_factor(int32_t mode);
void tkr_debug_set_seed_div(int32_t div);          /* seed fraction of a sweep = 1/div (default 12); tuning aid */
int32_t tkr_debug_filter_max_pairs(int32_t d);   /* resident CTA pairs of the filter kernel on the current device */

/* Same with HOST inputs/outputs (the np.dot/np.argsort seam of evaluate.py):
 * U_host/V_host/bias_host/rated_* in host memory, results to host memory.
 * `dev` is device scratch of >= tkr_score_topk_host_device_bytes(). */
size_t tkr_score_topk_host_device_bytes(int64_t nu, int64_t ni, int32_t d, int32_t k, int64_t n_rated);
int tkr_score_topk_host(const float* U_host, int64_t nu, const float* V_host, int64_t ni, int32_t d,
                        const float* bias_host, const int64_t* rated_indptr_host, const int32_t* rated_idx_host,
                        int32_t k, int32_t* out_idx_host, float* out_score_host, void* dev, size_t dev_bytes,
                        void* stream);

/* Hit counting of evaluate.py:84-112 on the device.  lists[*][total] = filtered top-`total` columns per user row
 * (output of tkr_score_topk*, -1 padded); the test file is given as n_lines entries: line_rows[l] = user row of line l,
 * likes_idx[likes_indptr[l] .. likes_indptr[l+1]) = its liked test columns, ascending and distinct.
 * pos_hits[p] (uint64[total], caller-zeroed, accumulated) += number of lines whose p-th kept column is liked;
 * the reference's hits[q] (q < total/step) is the sum of pos_hits[p] over p < (q+1)*step. */
int tkr_eval_hits(const int32_t* lists, int32_t total, const int32_t* line_rows, const int64_t* likes_indptr,
                  const int32_t* likes_idx, int64_t n_lines, unsigned long long* pos_hits, void* stream);

/* ---- the reference's text formats at native speed (host memory; SURVEY 8(f) NEXT-2) ----
 * .dat (utils.py:28-55, evaluate.py:19-28): rows x cols text matrix, '%f ' per element, one row per line.
 * tkr_dat_write is byte-identical to the reference writer; tkr_dat_read makes the roundings of np.float32(token). */
int tkr_dat_shape(const char* path, int64_t* rows, int64_t* cols);
int tkr_dat_read(const char* path, float* out, int64_t rows, int64_t cols);
int tkr_dat_write(const char* path, const float* mat, int64_t rows, int64_t cols);
/* same for a float64 matrix (CER's final-E.dat, single/cer.py:81-85): the doubles are formatted directly, like "'%f ' % x" */
int tkr_dat_write_f64(const char* path, const double* mat, int64_t rows, int64_t cols);
/* Rating file "uid,iid:like,..." (utils.py:58-89, evaluate.py:30-45) -> flat arrays in file order.  Call once with
 * line_user == NULL to get *n_lines / *n_pairs, allocate, call again.  line_user[l] / pair_item[p] = row of the id in
 * uid_path / iid_path (one id per line), -1 when unknown; pair_like[p] = 1 iff the label text is exactly "1";
 * the pairs of line l are [line_indptr[l], line_indptr[l+1]). */
int tkr_ratings_parse(const char* ratings_path, const char* uid_path, const char* iid_path, int64_t* n_lines,
                      int64_t* n_pairs, int32_t* line_user, int64_t* line_indptr, int32_t* pair_item, int8_t* pair_like);

/* ------------------------------------------------------------------ ALS (WMF / CER), SURVEY 8(f) NEXT-1 */

/* One half-step of the alternating least squares of single/cer.py:36-63 (and the intended single/wmf.py:67-96):
 * for every row r of the solved side X[n_rows, d]
 *     X_r = solve( base + (a-b) * sum_{p in pos(r)} Y_p Y_p^T + ridge*I ,  a * sum_p Y_p + ridge * prior_r )
 * replacing the reference's Python loop of np.dot + np.linalg.solve.  fp32 throughout (the reference forms the
 * matrices in fp32 and solves them in fp64 LAPACK; see DESIGN.md for the measured difference).  d <= 256.
 *   base   [d,d]   b * Yr^T Yr (+ lambda_u I on the user side, cer.py:38) from tkr_als_gram
 *   prior  [n_rows,d] or NULL: the content prior F_j E of CER (cer.py:35,55,62)
 *   cfg.solve_empty  rows without positives are solved too (CER items, cer.py:61-62) or keep their value
 *   cfg.item_loss    loss_rows[r] = the row's terms of the reference's loss: 0 -> 1/2 lreg |x|^2 (cer.py:46),
 *                    1 -> cer.py:58-63 / wmf.py:91-96 (with x^T B x evaluated as x^T rhs - ridge |x|^2)
 * The work list (`plan`, DEVICE arrays built by the caller once per data set) cuts rows into segments of
 * positives: seg_slot < 0 = the row's only segment, solved by the block that accumulates it; otherwise the segment's
 * partial matrix is written to partial slot seg_slot and the row is finished by the second kernel from its
 * multi_nslots consecutive slots starting at multi_slot0 (fixed summation order: results are deterministic).
 * Segments should be listed longest first. */
typedef struct tkr_als_cfg {
    int32_t d;
    float a, b;          /* confidence of positives / of everything else (wmf.py:11) */
    float ridge;         /* added to the diagonal per row: 0 on the user side (lambda_u is in base), lambda_v for items */
    float lreg;          /* coefficient of the row's quadratic loss term (lambda_u / lambda_v) */
    int32_t solve_empty;
    int32_t item_loss;
} tkr_als_cfg;
typedef struct tkr_als_plan {
    int64_t n_segs;
    const int32_t* seg_row;     /* [n_segs] row of X */
    const int64_t* seg_off;     /* [n_segs] first position in idx */
    const int32_t* seg_len;     /* [n_segs] positives in the segment */
    const int32_t* seg_slot;    /* [n_segs] -1 or partial slot */
    int64_t n_multi;
    const int32_t* multi_row;   /* [n_multi] rows that were split */
    const int32_t* multi_slot0; /* [n_multi] */
    const int32_t* multi_nslots;/* [n_multi] */
    const int64_t* multi_total; /* [n_multi] positives of the row */
    int64_t n_slots;
} tkr_als_plan;
size_t tkr_als_partial_bytes(int32_t d, int64_t n_slots);
int tkr_als_solve_rows(const tkr_als_cfg* cfg, const tkr_als_plan* plan, const float* Y, float* X, const int32_t* idx,
                       const float* base, const float* prior, double* loss_rows, void* partial, size_t partial_bytes,
                       void* stream);
/* out[d,d] = scale * sum_{r in rows} Y_r Y_r^T + ridge * I   (XX of cer.py:37-38 / :47-48); rows = DEVICE int32[n_rows]. */
size_t tkr_als_gram_workspace_bytes(int32_t d);
int tkr_als_gram(const float* Y, int32_t d, const int32_t* rows, int64_t n_rows, float scale, float ridge, float* out,
                 void* ws, size_t ws_bytes, void* stream);

/* Merge n_lists candidate lists idx/score[n_lists][nu][k] (each in the order
 * above, padded with idx -1) into out[nu][k] in the same order. */
int tkr_topk_merge(const int32_t* idx, const float* score, int32_t n_lists, int64_t nu, int32_t k,
                   int32_t* out_idx, float* out_score, void* stream);

/* Item-sharded scoring on several GPUs: every rank holds the filtered top-k lists of all `nu`
 * users of a batch against its own item shard (tkr_score_topk* with col_offset = the shard's first column).
 * tkr_topk_exchange_push stores row u of this rank's lists into list `rank` of the merge buffer of the row's owner
 * (owner = u / ceil(nu / world)) over NVLink and raises a flag; tkr_topk_exchange_merge waits for every rank's flag of
 * the same epoch and merges the world lists of the owned rows into out[rows_owned, k] -- bit-identical to the unsharded
 * call, each user's final list on exactly one rank.  The exchange buffer (tkr_peer_alloc'ed, tkr_topk_exchange_bytes,
 * sized for nu_cap users per batch) is double-buffered by epoch parity: push(t+1) may run while merge(t) is pending
 * (call them on different streams to overlap the exchange with the next batch's scoring).  `epoch` = 1, 2, 3, ... and
 * the same on every rank.  Barriers time out after 20 s (tkr_topk_exchange_status < 0). */
size_t tkr_topk_exchange_bytes(int64_t nu_cap, int32_t k, int32_t world);
int tkr_topk_exchange_push(const int32_t* idx, const float* score, int64_t nu, int64_t nu_cap, int32_t k,
                           const tkr_peers* peers, uint64_t epoch, void* stream);
int tkr_topk_exchange_merge(int64_t nu, int64_t nu_cap, int32_t k, const tkr_peers* peers, uint64_t epoch,
                            int32_t* out_idx, float* out_score, void* stream);
int tkr_topk_exchange_status(int64_t nu_cap, int32_t k, const tkr_peers* peers, void* stream);

/* Item-sharded scoring as a RING: the sweep of a user batch over the whole item table is cut into one SEGMENT per GPU, and
 * the running state of the sweep -- every row's threshold, candidate buffer and count (tkr_score_topk_tc_state_bytes) --
 * travels from GPU to GPU; GPU g works on segment g of batch t - g while GPU g+1 works on batch t - g - 1.  Unlike independent
 * per-shard top-k lists (tkr_topk_exchange_*), the per-row selection work -- which hardly depends on the sweep length -- is then
 * paid once per batch instead of once per shard.  `first` = the sweep starts here: its thresholds are seeded on a sample of
 * the WHOLE table when V_full is given (tables of >= 65 536 items; a BF16 copy of V_full is kept in the workspace next to the
 * shard's, built while items_prepared == 0 -- so V_full / bias_full must be the same table on every call that shares a
 * workspace), else on the shard alone.  `last` = it ends here: only then are the lists sorted out, re-scored exactly against V_full (the keys carry global columns) with the error-bound certificate,
 * and the uncertified rows re-done by the exact engine.  Results are bit-identical to tkr_score_topk on the whole table.
 * The caller moves the state (a plain device-to-device copy into the next GPU's mapped buffer) and orders the segments with
 * tkr_peer_signal_to / tkr_peer_wait_from (flag block of tkr_peer_flag_bytes() at a caller-chosen offset of the exchange
 * buffer; slots 4..7 are free for callers); topkrec.dist.RingScorer does both. */
size_t tkr_score_topk_tc_state_bytes(int64_t nu);
size_t tkr_score_topk_tc_segment_workspace_bytes(int64_t nu, int64_t ni_shard, int64_t ni_full, int32_t d, int32_t k, int32_t has_bias);
int tkr_score_topk_tc_segment(const float* U, int64_t nu, const float* V_shard, int64_t ni_shard, int32_t d, const float* bias_shard,
                              const int64_t* rated_indptr, const int32_t* rated_idx, int32_t k, int64_t col_offset, void* state,
                              int32_t first, int32_t last, const float* V_full, int64_t ni_full, const float* bias_full,
                              int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int32_t* n_fallback_rows,
                              int32_t items_prepared, void* stream);
size_t tkr_peer_flag_bytes(void);
int tkr_peer_signal_to(const tkr_peers* peers, size_t flag_off, int32_t slot, int32_t target, uint64_t epoch, void* stream);
int tkr_peer_wait_from(const tkr_peers* peers, size_t flag_off, int32_t slot, int32_t source, uint64_t epoch, void* stream);
int tkr_peer_status(const tkr_peers* peers, size_t flag_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOPKREC_H */
