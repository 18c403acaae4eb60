/* CPU oracle / CPU baseline for hot path 1: the synchronous BPR mini-batch step.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
 * reference executes this arithmetic inside TensorFlow 1.15 (not vendored);
 * this is the same restatement as oracle/bpr_ref.py, in C
 * with OpenMP so that bench.py's cpu_baseline / --impl reference leg can use
 * every host core.  tests/ pin it against oracle/bpr_ref.py.
 *
 * Follows single/bpr.py:81-100:
 *   x = b_i - b_j + <U_u,V_i> - <U_u,V_j>; loss = sum log(1+e^-x) + reg
 *   per-occurrence grads (App. A.2), duplicates summed (A.3),
 *   sparse RMSProp decay .9, eps 1e-10 inside sqrt (A.4) or plain SGD (A.9).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t n_users, n_items, d;
    float lu, li, lj, lb, lr;
    int32_t l1, sgd;
} ref_bpr_cfg;

static inline float sgnf(float x) { return (x > 0.f) - (x < 0.f); }
static inline float regg(float x, float lam, int l1) { return l1 ? lam * sgnf(x) : lam * x; }

static void apply_row(float* var, float* ms, const float* G, int d, float lr, int sgd) {
    for (int k = 0; k < d; ++k) {
        if (sgd) { var[k] -= lr * G[k]; continue; }
        float m = 0.9f * ms[k] + (1.0f - 0.9f) * G[k] * G[k];
        ms[k] = m;
        var[k] -= lr * G[k] / sqrtf(m + 1e-10f);
    }
}

/* counting sort of occurrence ids by row; occurrence o in [0,n_occ) has row key[o] */
static void build_lists(const int32_t* key, int64_t n_occ, int32_t n_rows, int64_t* indptr, int64_t* list) {
    memset(indptr, 0, sizeof(int64_t) * ((size_t)n_rows + 1));
    for (int64_t o = 0; o < n_occ; ++o) indptr[key[o] + 1]++;
    for (int32_t r = 0; r < n_rows; ++r) indptr[r + 1] += indptr[r];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_rows);
    memcpy(cur, indptr, sizeof(int64_t) * (size_t)n_rows);
    for (int64_t o = 0; o < n_occ; ++o) list[cur[key[o]]++] = o;
    free(cur);
}

int tkr_ref_bpr_step(const ref_bpr_cfg* c, float* U, float* V, float* b, float* msU, float* msV, float* msb,
                     const int32_t* u, const int32_t* i, const int32_t* j, int64_t B, double* loss_out) {
    const int d = c->d;
    float* s = (float*)malloc(sizeof(float) * (size_t)B);
    double loss = 0.0;
#pragma omp parallel for reduction(+ : loss) schedule(static)
    for (int64_t n = 0; n < B; ++n) {
        const float *pu = U + (int64_t)u[n] * d, *pi = V + (int64_t)i[n] * d, *pj = V + (int64_t)j[n] * d;
        float xi = 0.f, xj = 0.f, r = 0.f;
        for (int k = 0; k < d; ++k) {
            xi += pu[k] * pi[k];
            xj += pu[k] * pj[k];
            r += c->l1 ? c->lu * fabsf(pu[k]) + c->li * fabsf(pi[k]) + c->lj * fabsf(pj[k])
                       : 0.5f * (c->lu * pu[k] * pu[k] + c->li * pi[k] * pi[k] + c->lj * pj[k] * pj[k]);
        }
        float bi = b[i[n]], bj = b[j[n]];
        float x = bi - bj + xi - xj;
        r += c->l1 ? c->lb * (fabsf(bi) + fabsf(bj)) : 0.5f * c->lb * (bi * bi + bj * bj);
        s[n] = 1.0f / (1.0f + expf(x));
        loss += (double)log1pf(expf(-x)) + (double)r;
    }
    /* user rows: gradient rows are buffered, applied after the item pass (snapshot semantics) */
    int64_t* up = (int64_t*)malloc(sizeof(int64_t) * ((size_t)c->n_users + 1));
    int64_t* ul = (int64_t*)malloc(sizeof(int64_t) * (size_t)B);
    build_lists(u, B, c->n_users, up, ul);
    float* GU = (float*)calloc((size_t)B * (size_t)d, sizeof(float)); /* slot = first occurrence position */
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = 0; r < c->n_users; ++r) {
        if (up[r] == up[r + 1]) continue;
        float* G = GU + up[r] * d;
        const float* pu = U + (int64_t)r * d;
        for (int64_t q = up[r]; q < up[r + 1]; ++q) {
            int64_t n = ul[q];
            const float *pi = V + (int64_t)i[n] * d, *pj = V + (int64_t)j[n] * d;
            for (int k = 0; k < d; ++k) G[k] += -s[n] * (pi[k] - pj[k]) + regg(pu[k], c->lu, c->l1);
        }
    }
    /* item rows: occurrences 0..B-1 are the i-gather, B..2B-1 the j-gather (concat order) */
    int32_t* key = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)B);
    memcpy(key, i, sizeof(int32_t) * (size_t)B);
    memcpy(key + B, j, sizeof(int32_t) * (size_t)B);
    int64_t* vp = (int64_t*)malloc(sizeof(int64_t) * ((size_t)c->n_items + 1));
    int64_t* vl = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)B);
    build_lists(key, 2 * B, c->n_items, vp, vl);
#pragma omp parallel
    {
        float* G = (float*)malloc(sizeof(float) * (size_t)d);
#pragma omp for schedule(dynamic, 16)
        for (int32_t r = 0; r < c->n_items; ++r) {
            if (vp[r] == vp[r + 1]) continue;
            memset(G, 0, sizeof(float) * (size_t)d);
            float gb = 0.f;
            float* pv = V + (int64_t)r * d;
            for (int64_t q = vp[r]; q < vp[r + 1]; ++q) {
                int64_t o = vl[q];
                int neg = o >= B;
                int64_t n = neg ? o - B : o;
                const float* pu = U + (int64_t)u[n] * d;
                float sg = neg ? s[n] : -s[n];
                float lam = neg ? c->lj : c->li;
                for (int k = 0; k < d; ++k) G[k] += sg * pu[k] + regg(pv[k], lam, c->l1);
                gb += sg + regg(b[r], c->lb, c->l1);
            }
            apply_row(pv, msV + (int64_t)r * d, G, d, c->lr, c->sgd);
            apply_row(b + r, msb + r, &gb, 1, c->lr, c->sgd);
        }
        free(G);
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = 0; r < c->n_users; ++r)
        if (up[r] != up[r + 1]) apply_row(U + (int64_t)r * d, msU + (int64_t)r * d, GU + up[r] * d, d, c->lr, c->sgd);
    *loss_out = loss;
    free(s); free(up); free(ul); free(GU); free(key); free(vp); free(vl);
    return 0;
}

/* bench.py's CPU legs: torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, which would make the "all host
 * cores" baseline single-threaded; n > 0 sets the team size explicitly, the return value is the size in effect. */
#ifdef _OPENMP
#include <omp.h>
int tkr_ref_omp_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int tkr_ref_omp_threads(int n) { (void)n; return 1; }
#endif
