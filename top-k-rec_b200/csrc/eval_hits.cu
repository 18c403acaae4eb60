// NEXT-3 of SURVEY 8(f): the hit counting of evaluate.py:84-112 on the device.
// The reference walks, per test-file line with >= 1 like, the user's filtered top-`total` columns and adds 1 to
// hits[q] for every q >= p / step when the p-th kept column is liked.  Here: one thread per (line, position)
// looks its column up in the line's sorted like list and bumps a per-position histogram; the caller turns it into
// the cumulative hits@{step, 2*step, ...} (hits[q] = sum of pos_hits[p] over p < (q+1)*step).
#include "common.cuh"

namespace tkr {

__global__ void __launch_bounds__(256) eval_hits_kernel(const int32_t* __restrict__ lists, int32_t total,
                                                        const int32_t* __restrict__ line_rows, const int64_t* __restrict__ likes_indptr,
                                                        const int32_t* __restrict__ likes_idx, int64_t n_lines,
                                                        unsigned long long* __restrict__ pos_hits) {
    extern __shared__ unsigned int hist[];                 // [total]
    for (int t = threadIdx.x; t < total; t += blockDim.x) hist[t] = 0u;
    __syncthreads();
    const int64_t n = n_lines * total;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t line = e / total;
        const int p = (int)(e - line * total);
        const int32_t col = __ldg(lists + (int64_t)__ldg(line_rows + line) * total + p);
        if (col < 0) continue;                             // list shorter than `total`
        int64_t lo = __ldg(likes_indptr + line), hi = __ldg(likes_indptr + line + 1);
        const int64_t end = hi;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(likes_idx + mid) < col) lo = mid + 1; else hi = mid;
        }
        if (lo < end && __ldg(likes_idx + lo) == col) atomicAdd(hist + p, 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < total; t += blockDim.x)
        if (hist[t] != 0u) atomicAdd(pos_hits + t, (unsigned long long)hist[t]);
}

}  // namespace tkr

using namespace tkr;

extern "C" int tkr_eval_hits(const int32_t* lists, int32_t total, const int32_t* line_rows, const int64_t* likes_indptr,
                             const int32_t* likes_idx, int64_t n_lines, unsigned long long* pos_hits, void* stream) {
    TKR_CHECK_ARG(total >= 1 && total <= 4096, "total must be in [1, 4096]");
    TKR_CHECK_ARG(n_lines >= 0, "n_lines < 0");
    if (n_lines == 0) return TKR_OK;
    TKR_CHECK_ARG(lists && line_rows && likes_indptr && likes_idx && pos_hits, "NULL argument");
    int64_t blocks = (n_lines * total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    eval_hits_kernel<<<(unsigned)blocks, 256, (size_t)total * sizeof(unsigned int), (cudaStream_t)stream>>>(
        lists, total, line_rows, likes_indptr, likes_idx, n_lines, pos_hits);
    TKR_LAUNCH_CHECK();
    return TKR_OK;
}
