/* CPU oracle for hot path 2: score + rated-filtered top-k.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the
 * product library.
 *
 * Restates evaluate.py:75-105 of the reference:
 *   :78  scores = np.dot(umat, temat.T)          -> fp32 dot per (user, column)
 *   :79  scores += bias                          -> intended per-column bias
 *                                                   (old/methods/bpr_test.py:18-32)
 *   :81  rlist = np.argsort(scores, axis=1)      -> ranking
 *   :96-105 walk the ranking backwards, skip rated items, keep `total`
 *
 * The reference leaves two things open (BLAS summation order, unstable sort);
 * SURVEY.md section 8(c) fixes them and this file is that definition:
 *   score(u,c) = fmaf chain over k ascending starting from +0.0f, then
 *                + bias[c] (one fp32 add), then +0.0f (canonicalises -0);
 *   order      = score descending, ties by column index descending
 *                (= a stable ascending argsort read backwards).
 * Output per user: the first `k` columns in that order that are not in the
 * user's rated list; unused slots get idx -1 / score -inf.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t make_key(float s, int32_t c) {
    uint32_t b;
    memcpy(&b, &s, 4);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   /* order-preserving */
    return ((uint64_t)b << 32) | (uint32_t)c;
}

float tkr_ref_score(const float* u, const float* v, int d, const float* bias_c) {
    float acc = 0.0f;
    for (int k = 0; k < d; ++k) acc = fmaf(u[k], v[k], acc);
    if (bias_c) acc = acc + *bias_c;
    return acc + 0.0f;
}

/* rated_idx[rated_indptr[u] .. rated_indptr[u+1]) = sorted te-columns rated by u */
static int is_rated(const int64_t* indptr, const int32_t* idx, int64_t u, int32_t c) {
    if (!indptr) return 0;
    int64_t lo = indptr[u], hi = indptr[u + 1];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (idx[mid] < c) lo = mid + 1; else hi = mid;
    }
    return lo < indptr[u + 1] && idx[lo] == c;
}

static void sift_down(uint64_t* h, int n, int p) {       /* min-heap */
    for (;;) {
        int l = 2 * p + 1, r = l + 1, m = p;
        if (l < n && h[l] < h[m]) m = l;
        if (r < n && h[r] < h[m]) m = r;
        if (m == p) return;
        uint64_t t = h[p]; h[p] = h[m]; h[m] = t; p = m;
    }
}

static int cmp_desc(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}

int tkr_ref_score_topk(const float* U, int64_t nu, const float* V, int64_t ni, int d,
                       const float* bias, const int64_t* rated_indptr, const int32_t* rated_idx,
                       int k, int64_t col_offset, int32_t* out_idx, float* out_score) {
    if (k <= 0 || d <= 0) return -1;
#pragma omp parallel
    {
        uint64_t* heap = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)k);
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < nu; ++r) {
            int n = 0;
            const float* u = U + r * d;
            for (int64_t c = 0; c < ni; ++c) {
                float s = tkr_ref_score(u, V + c * d, d, bias ? bias + c : 0);
                uint64_t key = make_key(s, (int32_t)(c + col_offset));
                if (n == k && key < heap[0]) continue;
                if (is_rated(rated_indptr, rated_idx, r, (int32_t)(c + col_offset))) continue;
                if (n < k) {
                    heap[n++] = key;
                    if (n == k) for (int p = k / 2 - 1; p >= 0; --p) sift_down(heap, k, p);
                } else {
                    heap[0] = key;
                    sift_down(heap, k, 0);
                }
            }
            qsort(heap, (size_t)n, sizeof(uint64_t), cmp_desc);
            for (int p = 0; p < k; ++p) {
                if (p < n) {
                    int32_t c = (int32_t)(heap[p] & 0xffffffffu);
                    out_idx[r * k + p] = c;
                    out_score[r * k + p] = tkr_ref_score(u, V + (c - col_offset) * d, d, bias ? bias + (c - col_offset) : 0);
                } else {
                    out_idx[r * k + p] = -1;
                    out_score[r * k + p] = -INFINITY;
                }
            }
        }
        free(heap);
    }
    return 0;
}

/* Merge G per-shard candidate lists [G][nu][k] (each sorted by the order above,
 * padded with idx -1) into the global top-k: the reference has no such step
 * (single process); it is defined by "same answer as the unsharded call". */
int tkr_ref_topk_merge(const int32_t* idx, const float* score, int G, int64_t nu, int k,
                       int32_t* out_idx, float* out_score) {
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)G * (size_t)k);
    for (int64_t r = 0; r < nu; ++r) {
        int n = 0;
        for (int g = 0; g < G; ++g)
            for (int p = 0; p < k; ++p) {
                int64_t o = ((int64_t)g * nu + r) * k + p;
                if (idx[o] >= 0) keys[n++] = make_key(score[o], idx[o]);
            }
        qsort(keys, (size_t)n, sizeof(uint64_t), cmp_desc);
        for (int p = 0; p < k; ++p) {
            if (p < n) {
                uint32_t b = (uint32_t)(keys[p] >> 32);
                b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
                float s; memcpy(&s, &b, 4);
                out_idx[r * k + p] = (int32_t)(keys[p] & 0xffffffffu);
                out_score[r * k + p] = s;
            } else {
                out_idx[r * k + p] = -1;
                out_score[r * k + p] = -INFINITY;
            }
        }
    }
    free(keys);
    return 0;
}
