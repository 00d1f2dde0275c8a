/*
 * oracle/chain.c -- TEST INFRASTRUCTURE ONLY (the parity checker and the CPU
 * baseline).  Nothing in the product path may link or call this file.
 *
 * CPU restatement, in plain C, of the arithmetic that the reference delegates
 * to faiss-cpu (>=1.7.4, un-vendored; requirements.txt:15) on the retrieval
 * path:
 *
 *   faiss.normalize_L2(x)            src/trainer.py:240, app/recommendation_service.py:71
 *   faiss.IndexFlatIP(d).add(x)      src/trainer.py:241-242
 *   index.search(q, k) -> (D, I)     app/recommendation_service.py:72
 *
 * FAISS is not installable in this image and the reference ships no golden
 * vectors for this path, so this oracle is "parity unpinned" against FAISS; it
 * is pinned by the known-answer tests in tests/  and by
 * the committed fixtures under tests/golden/.
 *
 * Arithmetic stated for this build (so that index-exact parity is well
 * defined):
 *   score(q, x) = fmaf(q[d-1], x[d-1], ... fmaf(q[1], x[1], fmaf(q[0], x[0], 0)) ...)
 *   i.e. a single fp32 accumulator, fused multiply-add, ascending feature index.
 *   normalize:  nr = the same chain over x[j]*x[j];  if nr > 0:  x[j] *= 1.0f / sqrtf(nr)
 *   top-k order: score descending, then index ascending; rows short of k are
 *   padded with (score = -FLT_MAX, index = -1) as FAISS pads its IP heaps.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RB 16 /* rows scored together so the 16 independent chains pipeline */

typedef struct {
    float s;
    int64_t i;
} hit_t;

/* a is strictly better than b under (score desc, index asc) */
static inline int better(float sa, int64_t ia, float sb, int64_t ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

/* min-heap on "better": root = the worst kept hit */
static void heap_sift_down(hit_t *h, int n, int p) {
    for (;;) {
        int l = 2 * p + 1, r = l + 1, w = p;
        if (l < n && better(h[w].s, h[w].i, h[l].s, h[l].i)) w = l;
        if (r < n && better(h[w].s, h[w].i, h[r].s, h[r].i)) w = r;
        if (w == p) return;
        hit_t t = h[p];
        h[p] = h[w];
        h[w] = t;
        p = w;
    }
}

static void heap_sift_up(hit_t *h, int p) {
    while (p > 0) {
        int parent = (p - 1) / 2;
        if (better(h[parent].s, h[parent].i, h[p].s, h[p].i)) {
            hit_t t = h[p];
            h[p] = h[parent];
            h[parent] = t;
            p = parent;
        } else
            return;
    }
}

static inline void heap_offer(hit_t *h, int *n, int k, float s, int64_t i) {
    if (*n < k) {
        h[*n].s = s;
        h[*n].i = i;
        heap_sift_up(h, *n);
        (*n)++;
    } else if (better(s, i, h[0].s, h[0].i)) {
        h[0].s = s;
        h[0].i = i;
        heap_sift_down(h, k, 0);
    }
}

static int cmp_hit_desc(const void *a, const void *b) {
    const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
    if (better(x->s, x->i, y->s, y->i)) return -1;
    if (better(y->s, y->i, x->s, x->i)) return 1;
    return 0;
}

/* the stated score chain */
float oracle_dot_chain(const float *q, const float *x, int d) {
    float acc = 0.0f;
    for (int j = 0; j < d; ++j) acc = __builtin_fmaf(q[j], x[j], acc);
    return acc;
}

/* scores[r] for rows [r0, r0+nr) of x against one query (nr <= RB) */
static inline void score_block(const float *q, const float *x, int d, int nr, float *out) {
    float acc[RB];
    for (int r = 0; r < RB; ++r) acc[r] = 0.0f;
    if (nr == RB) {
        for (int j = 0; j < d; ++j) {
            const float qj = q[j];
            for (int r = 0; r < RB; ++r) acc[r] = __builtin_fmaf(qj, x[(size_t)r * d + j], acc[r]);
        }
    } else {
        for (int j = 0; j < d; ++j) {
            const float qj = q[j];
            for (int r = 0; r < nr; ++r) acc[r] = __builtin_fmaf(qj, x[(size_t)r * d + j], acc[r]);
        }
    }
    for (int r = 0; r < nr; ++r) out[r] = acc[r];
}

/* all scores, dense [nq, n] (small cases / matvec checks) */
void oracle_scores_chain(const float *x, int64_t n, int d, const float *q, int64_t nq, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        for (int64_t r0 = 0; r0 < n; r0 += RB) {
            int nr = (int)((n - r0) < RB ? (n - r0) : RB);
            score_block(q + qi * d, x + r0 * d, d, nr, out + qi * n + r0);
        }
    }
}

/*
 * IndexFlatIP.search restated: exact top-k of q . x over all n rows.
 * D: float32 [nq, k], I: int64 [nq, k].  id_base is added to the row numbers
 * (used when the caller scores one shard of a partitioned catalog).
 */
void oracle_search_chain(const float *x, int64_t n, int d, const float *q, int64_t nq, int k,
                         int64_t id_base, float *D, int64_t *I) {
    int nchunk = 1;
#ifdef _OPENMP
    {
        int t = omp_get_max_threads();
        if (nq < t) nchunk = (int)((t + nq - 1) / nq);
        if ((int64_t)nchunk * 4096 > n) nchunk = (int)(n / 4096 > 0 ? n / 4096 : 1);
    }
#endif
    const int64_t ntask = nq * nchunk;
    hit_t *heaps = (hit_t *)malloc(sizeof(hit_t) * (size_t)ntask * (size_t)k);
    int *cnts = (int *)calloc((size_t)ntask, sizeof(int));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < ntask; ++t) {
        const int64_t qi = t / nchunk;
        const int c = (int)(t % nchunk);
        const int64_t lo = n * c / nchunk, hi = n * (c + 1) / nchunk;
        hit_t *h = heaps + (size_t)t * k;
        int cnt = 0;
        float sc[RB];
        for (int64_t r0 = lo; r0 < hi; r0 += RB) {
            int nr = (int)((hi - r0) < RB ? (hi - r0) : RB);
            score_block(q + qi * d, x + r0 * d, d, nr, sc);
            for (int r = 0; r < nr; ++r) heap_offer(h, &cnt, k, sc[r], r0 + r);
        }
        cnts[t] = cnt;
    }
    hit_t *all = (hit_t *)malloc(sizeof(hit_t) * (size_t)nchunk * (size_t)k);
    for (int64_t qi = 0; qi < nq; ++qi) {
        int m = 0;
        for (int c = 0; c < nchunk; ++c) {
            const int64_t t = qi * nchunk + c;
            memcpy(all + m, heaps + (size_t)t * k, sizeof(hit_t) * (size_t)cnts[t]);
            m += cnts[t];
        }
        qsort(all, (size_t)m, sizeof(hit_t), cmp_hit_desc);
        for (int j = 0; j < k; ++j) {
            if (j < m) {
                D[qi * k + j] = all[j].s;
                I[qi * k + j] = all[j].i + id_base;
            } else {
                D[qi * k + j] = -FLT_MAX;
                I[qi * k + j] = -1;
            }
        }
    }
    free(all);
    free(cnts);
    free(heaps);
}

/*
 * FAISS-style blocked search bookkeeping for the CPU *baseline* (not the
 * parity oracle): the caller computes a block of scores S[nq, nb] with sgemm
 * (as faiss does above 20 queries) and this folds it into per-query heaps.
 * heaps: [nq, k] hit_t, cnts: [nq] int, both owned by the caller.
 */
void oracle_heap_addn(hit_t *heaps, int *cnts, int64_t nq, int k, const float *S, int64_t nb,
                      int64_t id_base) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        hit_t *h = heaps + (size_t)qi * k;
        int cnt = cnts[qi];
        const float *s = S + qi * nb;
        for (int64_t j = 0; j < nb; ++j) {
            if (cnt < k || s[j] >= h[0].s) heap_offer(h, &cnt, k, s[j], id_base + j);
        }
        cnts[qi] = cnt;
    }
}

void oracle_heap_finish(hit_t *heaps, const int *cnts, int64_t nq, int k, float *D, int64_t *I) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        hit_t *h = heaps + (size_t)qi * k;
        int m = cnts[qi];
        qsort(h, (size_t)m, sizeof(hit_t), cmp_hit_desc);
        for (int j = 0; j < k; ++j) {
            D[qi * k + j] = j < m ? h[j].s : -FLT_MAX;
            I[qi * k + j] = j < m ? h[j].i : -1;
        }
    }
}

int oracle_hit_size(void) { return (int)sizeof(hit_t); }

/* faiss.normalize_L2 restated with the stated chain; zero rows are left alone */
void oracle_normalize_l2(float *x, int64_t n, int d) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        float *row = x + r * d;
        float nr = 0.0f;
        for (int j = 0; j < d; ++j) nr = __builtin_fmaf(row[j], row[j], nr);
        if (nr > 0.0f) {
            const float inv = 1.0f / sqrtf(nr);
            for (int j = 0; j < d; ++j) row[j] *= inv;
        }
    }
}

/*
 * Merge of per-shard top-k lists (the multi-GPU exchange step): parts is
 * [nparts, nq, k] of (D, I) already carrying global ids; keeps (score desc,
 * index asc) and drops -1 padding.
 */
void oracle_merge_topk(const float *Dp, const int64_t *Ip, int nparts, int64_t nq, int k, float *D,
                       int64_t *I) {
    hit_t *all = (hit_t *)malloc(sizeof(hit_t) * (size_t)nparts * (size_t)k);
    for (int64_t qi = 0; qi < nq; ++qi) {
        int m = 0;
        for (int p = 0; p < nparts; ++p)
            for (int j = 0; j < k; ++j) {
                const size_t o = ((size_t)p * nq + qi) * k + j;
                if (Ip[o] >= 0) {
                    all[m].s = Dp[o];
                    all[m].i = Ip[o];
                    ++m;
                }
            }
        qsort(all, (size_t)m, sizeof(hit_t), cmp_hit_desc);
        for (int j = 0; j < k; ++j) {
            D[qi * k + j] = j < m ? all[j].s : -FLT_MAX;
            I[qi * k + j] = j < m ? all[j].i : -1;
        }
    }
    free(all);
}
