/*
 * TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (inclusivegan_b200/) never imports, links or executes anything under oracle/.
 *
 * Exact float64 brute-force k-nearest-neighbour oracle for the IMLE matching path.
 *
 * What it restates (reference = /root/reference, ningyu1991/InclusiveGAN):
 *   - distance semantics: dci_code/src/util.c:62-69 `compute_dist` — Euclidean distance,
 *     sqrt of a float64 sum of squared differences accumulated sequentially i = 0..dim-1.
 *     `oracle_dist` below follows that loop order exactly, so distances are bit-identical
 *     to what the reference reports for the same (query, point) pair.
 *   - result contract: dci_code/src/dci.c:788-828 `dci_query` — per query, neighbours in
 *     ascending distance, indices are row positions in the array passed to add()
 *     (dci_code/src/py_dci.c:185 adds data_idx_offset), at most k per query.
 *   - what it deliberately does NOT restate: DCI's approximate candidate selection
 *     (dci.c:385-764).  BASELINE.json `north_star` defines correctness against the exact
 *     answer; the reference with exhaustive settings (num_levels=1, prop_to_visit=1,
 *     prop_to_retrieve=1) visits every point and is itself exact — that is what the
 *     golden fixtures in tests/golden/ pin this oracle against (see tests/golden/make_golden.py).
 *
 * Parity status: the reference ships NO tests/golden vectors for this path (SURVEY.md §4,
 * §8c).  The oracle is pinned against outputs of the compiled reference run in this
 * container (oracle/_ref/_dci.so, exhaustive mode) and committed as tests/golden/*.npz.
 *
 * Ties: ascending (distance, index) — equal distances resolve to the lower index.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QBLOCK 16 /* queries evaluated per pass over the pool (keeps a pool row in L1) */

static inline double oracle_dist(const double *a, const double *b, int dim)
{
    double s = 0.0;
    for (int i = 0; i < dim; i++) s += (a[i] - b[i]) * (a[i] - b[i]);
    return sqrt(s);
}

/* insert (d, idx) into an ascending list of length *len (capacity k), ordered by (d, idx) */
static inline void topk_insert(double *bd, int32_t *bi, int *len, int k, double d, int32_t idx)
{
    int n = *len;
    if (n == k) {
        if (d > bd[k - 1] || (d == bd[k - 1] && idx > bi[k - 1])) return;
        n = k - 1;
    }
    int p = n;
    while (p > 0 && (bd[p - 1] > d || (bd[p - 1] == d && bi[p - 1] > idx))) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        p--;
    }
    bd[p] = d;
    bi[p] = idx;
    *len = n + 1;
}

/*
 * data  : n  x dim  row-major float64        query : nq x dim row-major float64
 * out_idx  : nq x kk int32,  out_dist : nq x kk float64,  kk = min(k, n); rows ascending.
 * squared != 0 returns squared distances (metrics/precision_recall.py:20-57 uses squared L2).
 * returns kk, or -1 on bad arguments.
 */
int knn_oracle_f64(const double *data, int64_t n, const double *query, int64_t nq, int dim, int k,
                   int squared, int32_t *out_idx, double *out_dist)
{
    if (!data || !query || !out_idx || !out_dist || n <= 0 || nq < 0 || dim <= 0 || k <= 0) return -1;
    const int kk = (int64_t)k < n ? k : (int)n;
    int64_t qb;
#pragma omp parallel for schedule(dynamic, 1)
    for (qb = 0; qb < nq; qb += QBLOCK) {
        const int nb = (int)((nq - qb) < QBLOCK ? (nq - qb) : QBLOCK);
        double *bd = (double *)malloc(sizeof(double) * (size_t)kk * QBLOCK);
        int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)kk * QBLOCK);
        int len[QBLOCK];
        memset(len, 0, sizeof(len));
        for (int64_t j = 0; j < n; j++) {
            const double *x = data + j * dim;
            for (int b = 0; b < nb; b++) {
                const double d = oracle_dist(x, query + (qb + b) * dim, dim);
                topk_insert(bd + (size_t)b * kk, bi + (size_t)b * kk, &len[b], kk, d, (int32_t)j);
            }
        }
        for (int b = 0; b < nb; b++) {
            for (int r = 0; r < kk; r++) {
                const double d = bd[(size_t)b * kk + r];
                out_dist[(qb + b) * kk + r] = squared ? d * d : d;
                out_idx[(qb + b) * kk + r] = bi[(size_t)b * kk + r];
            }
        }
        free(bd);
        free(bi);
    }
    return kk;
}

/* Distances of explicit (query row, data row) pairs — used by tests to re-evaluate candidates. */
void knn_oracle_pair_dist_f64(const double *data, const double *query, int dim, int64_t npairs,
                              const int64_t *qrow, const int64_t *xrow, double *out)
{
    int64_t p;
#pragma omp parallel for schedule(static)
    for (p = 0; p < npairs; p++) out[p] = oracle_dist(data + xrow[p] * dim, query + qrow[p] * dim, dim);
}
