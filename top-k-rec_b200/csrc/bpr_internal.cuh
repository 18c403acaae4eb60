// Internal interface between bpr_step.cu and vbpr_step.cu (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace tkr {

struct SamplerDev {
    const int32_t* tr_users;
    const int64_t* pos_indptr;
    const int32_t* pos_idx;
    uint32_t n_tr_users, n_items, seed_lo, seed_hi;
};

struct StepWs {            // views into the caller's workspace
    float* GU; float* GV; float* Gb;
    float* tchV;           // dense / data-parallel mode: per-item "touched" flag as fp32 (all-reduced with GV|Gb)
    int32_t* cntU; int32_t* cntV;
    int32_t* listU; int32_t* listV;
    int32_t* n_touched;    // [0] touched user rows, [1] touched item rows, [2] apply blocks finished
    int32_t* hot_slot;     // [n_items] 0 = cold, s+1 = the row's gradients are privatised in slot s of every block
    int32_t* hot_ids;      // [TKR_MAX_HOT] item id of slot s (-1 = unused)
    int32_t* stage;        // batches <= kPersistMaxBatch: 3 x kStageTriples ids sampled ahead for the persistent kernel (else NULL)
    uint32_t* sync;        // ... and the persistent kernel's barrier words: [0] arrival counter, [1] its value at the end of the last launch
    char* flow;            // scratch of the dataflow multi-step kernel (bpr_flow_bytes; NULL for batches > kPersistMaxBatch)
};

constexpr int64_t kPersistMaxBatch = 1024;     // batches up to this size CAN take the persistent multi-step kernel (bpr_persist.cu)
constexpr int64_t kPersistAutoBatch = 256;     // ... and up to this size do by default (one triple per warp of the cluster)
constexpr int64_t kStageTriples = 1 << 16;
constexpr int64_t kFlowTriples = 1 << 18;      // triples per launch of the dataflow multi-step kernel (bpr_flow_steps)
// automatic choice between the two multi-step kernels (profiles/r02w_probe_b256.txt, C2 tables, us per step dataflow / cluster /
// two launches: B=64 5.3 / 6.2 / 13.7, B=256 7.7 / 6.1 / 14.7, B=512 9.6 / 17.7 / 15.0, B=1024 11.8 / 28.7 / 16.3): the cluster
// kernel's two grid barriers cost the same whatever the batch, the dataflow kernel's per-step chain grows with the number of
// occurrences of the most popular rows
constexpr int64_t kFlowSmallBatch = 64;        // batches up to this size ...
constexpr int64_t kFlowLargeBatch = 256;       // ... and above this one (up to kPersistMaxBatch) take the dataflow kernel

// VBPR rides on the same kernels with concatenated rows U' = [ur|uc], V' = [ir | F.E]:
//   item_cols  leading columns of an item row that are parameters (regularised, updated); the rest is the
//              content projection, whose accumulated gradient W feeds the dense dE GEMM instead;
//   b_reg      the trainable bias (rb) used for regularisation / update, while x reads b = rb + F.c;
//   wq         per-item sum of -/+ s (the bias gradient without regularisation), the dc GEMV input.
struct StepExtra {
    int item_cols;
    const float* b_reg;
    float* wq;
    int hot_rows = 0;      // privatised item rows per block (0 = off): shared memory holds hot_rows * (d + 3) floats
    // VBPR with the graph exactly as written (vbpr.py:61 broadcasts x to [B, B], defect D-14): the weight of triple n is
    // s_emb[n] = sum_a sigma(-x[a, n]) on everything reached through the embeddings and s_bias[n] = sum_b sigma(-x[n, b])
    // on the biases / wq, both computed beforehand (vbpr_pair_kernel); the data term of the loss comes from there as well.
    const float* s_emb = nullptr;
    const float* s_bias = nullptr;
};

// MODE_LIST / MODE_DENSE: every touched row goes through the gradient accumulators (touched rows listed / flagged).
// MODE_COUNT: a counting pre-pass tells how often each row occurs in the batch; rows that occur once are updated
//             in place by the gradient kernel (no accumulator round trip), the others take the accumulator path.
constexpr int MODE_LIST = 0, MODE_DENSE = 1, MODE_COUNT = 2;
// MODE_HOGWILD: no accumulators, no barrier, no apply pass -- every occurrence adds -lr * gradient straight onto the
//             parameter rows (ws.GU / GV / Gb point AT U / V / b); rows read while others update them (SURVEY 8(f) NEXT-4).
constexpr int MODE_HOGWILD = 3;

size_t bpr_ws_total(const tkr_bpr_cfg* cfg, int64_t B);
int bpr_carve(const tkr_bpr_cfg* cfg, int64_t B, void* ws, size_t ws_bytes, StepWs* out);
int bpr_check_cfg(const tkr_bpr_cfg* cfg, int64_t B);
int bpr_pick_mode(const tkr_bpr_cfg* cfg, int64_t B, int data_parallel);
int bpr_make_sampler(const tkr_sampler* smp, SamplerDev* out);
int bpr_dispatch_grad(const tkr_bpr_cfg* cfg, const float* U, const float* V, const float* b, const int32_t* u,
                      const int32_t* i, const int32_t* j, int64_t B, const SamplerDev& smp, uint64_t first_draw,
                      const StepWs& ws, int mode, const StepExtra& ex, float* loss, cudaStream_t st,
                      float* msU = nullptr, float* msV = nullptr);   // the slots are only needed by MODE_COUNT
extern int g_persist_mode;
bool bpr_persist_legal(const tkr_bpr_cfg* cfg, int64_t B);
size_t bpr_flow_bytes(const tkr_bpr_cfg* cfg);
int bpr_flow_steps(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                   const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const StepWs& ws, float* loss, cudaStream_t st);
int bpr_persist_steps(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb, const int32_t* u,
                      const int32_t* i, const int32_t* j, int64_t B, int64_t n_steps, const StepWs& ws, float* loss, cudaStream_t st);
void bpr_launch_apply(const tkr_bpr_cfg* cfg, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                      int64_t B, const StepWs& ws, int mode, const StepExtra& ex, cudaStream_t st);

// Plain BPR rows: every column is a parameter; up to 16 KB of shared memory per block for the privatised hot rows
// (costs nothing when the caller named none: the per-triple slot lookups read zeros).
inline int hot_rows_for(int d) {
    int rows = (16 * 1024) / ((d + 3) * (int)sizeof(float));
    return rows > TKR_MAX_HOT ? TKR_MAX_HOT : rows;
}
inline StepExtra plain_extra(const tkr_bpr_cfg* cfg, const float* b) { return StepExtra{cfg->d, b, nullptr, hot_rows_for(cfg->d)}; }


}  // namespace tkr
