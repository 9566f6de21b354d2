/*
 * ecs_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's element-centric-similarity hot path for flat disjoint
 * partitions.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported baseline.  Nothing under
 * clustassess_b200/ links, imports or calls it; the product path has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (1) the unmodified reference C++ (src/calculate_ecs.cpp compiled against oracle/rcpp_standin,
 *       oracle/_ref/libref_ecs.so) when /root/reference is present,
 *   (2) the committed fixtures in tests/golden/ generated from (1) by oracle/gen_golden.py, and
 *   (3) the only deterministic known answer the reference documents
 *       (docs/reference/merge_partitions.html:117-142, example at R/ECS.R:1164-1165).
 *
 * Each function cites the reference lines it follows.  Labels are int32 (R INTSXP); a partition is
 * a length-n vector; "labels" arguments are partition-major M x N (partition p at labels + p*n),
 * which is R's list-of-vectors layout.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static void minmax_i32(const int32_t *x, int64_t n, int32_t *mn, int32_t *mx) {
    int32_t lo = x[0], hi = x[0];
    for (int64_t i = 1; i < n; i++) {
        if (x[i] < lo) lo = x[i];
        if (x[i] > hi) hi = x[i];
    }
    *mn = lo;
    *mx = hi;
}

/* myContTable -- src/calculate_ecs.cpp:7-19.
 * rows = max(a)-min_a+1, cols = max(b)-min_b+1 (labels unused inside the range give all-zero
 * rows/cols); table is column-major like an R matrix: T[i + j*rows].  Returns a calloc'd table. */
int32_t *orc_cont_table(const int32_t *a, const int32_t *b, int64_t n, int32_t min_a, int32_t min_b,
                        int32_t *nrow, int32_t *ncol) {
    int32_t mn, max_a, max_b;
    minmax_i32(a, n, &mn, &max_a);
    minmax_i32(b, n, &mn, &max_b);
    int32_t ka = max_a - min_a + 1, kb = max_b - min_b + 1;
    int32_t *t = (int32_t *)calloc((size_t)ka * (size_t)kb, sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) t[(size_t)(a[i] - min_a) + (size_t)(b[i] - min_b) * ka]++;
    *nrow = ka;
    *ncol = kb;
    return t;
}

/* Same, writing into caller memory (ctypes-friendly). table_out may be NULL to query the shape. */
int orc_cont_table_into(const int32_t *a, const int32_t *b, int64_t n, int32_t min_a, int32_t min_b,
                        int32_t *table_out, int32_t *nrow, int32_t *ncol) {
    int32_t *t = orc_cont_table(a, b, n, min_a, min_b, nrow, ncol);
    if (table_out) memcpy(table_out, t, sizeof(int32_t) * (size_t)(*nrow) * (size_t)(*ncol));
    free(t);
    return 0;
}

static void row_col_sums(const int32_t *t, int32_t ka, int32_t kb, int32_t *rs, int32_t *cs) {
    memset(rs, 0, sizeof(int32_t) * (size_t)ka);
    memset(cs, 0, sizeof(int32_t) * (size_t)kb);
    for (int32_t j = 0; j < kb; j++)
        for (int32_t i = 0; i < ka; i++) {
            rs[i] += t[(size_t)i + (size_t)j * ka];
            cs[j] += t[(size_t)i + (size_t)j * ka];
        }
}

/* disjointECS -- src/calculate_ecs.cpp:22-46.
 * ecs[n] = T[a_n][b_n] / max(|A(a_n)|, |B(b_n)|); one IEEE fp64 division of two exactly
 * representable integers per non-empty table cell (calculate_ecs.cpp:35). */
int orc_disjoint_ecs(const int32_t *a, const int32_t *b, int64_t n, double *ecs) {
    if (n <= 0) return 0;
    int32_t min_a, min_b, mx;
    minmax_i32(a, n, &min_a, &mx);
    minmax_i32(b, n, &min_b, &mx);
    int32_t ka, kb;
    int32_t *t = orc_cont_table(a, b, n, min_a, min_b, &ka, &kb);
    int32_t *rs = (int32_t *)malloc(sizeof(int32_t) * (size_t)ka);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)kb);
    row_col_sums(t, ka, kb, rs, cs);
    double *u = (double *)malloc(sizeof(double) * (size_t)ka * (size_t)kb);
    for (int32_t i = 0; i < ka; i++)
        for (int32_t j = kb - 1; j >= 0; j--) {
            int32_t m = rs[i] > cs[j] ? rs[i] : cs[j];
            u[(size_t)i + (size_t)j * ka] = (double)t[(size_t)i + (size_t)j * ka] / m; /* 0/0 -> NaN, never gathered */
        }
    for (int64_t i = 0; i < n; i++) ecs[i] = u[(size_t)(a[i] - min_a) + (size_t)(b[i] - min_b) * ka];
    free(u);
    free(cs);
    free(rs);
    free(t);
    return 0;
}

/* disjointECSaverage -- src/calculate_ecs.cpp:49-70:  sum_{i,j} T^2 / (n * max(S1_i, S2_j)),
 * iterating i ascending, j descending, skipping nothing (empty cells add 0/(n*max); an empty row AND
 * column would add 0/0 -- cannot happen because every label in range... can: handled like the
 * reference only where max > 0).
 * DEVIATION (documented in SURVEY.md section 8 a3): the reference forms n*max as int*int which
 * overflows once n*max >= 2^31; here the product is formed in fp64.  Identical wherever the
 * reference is defined. */
double orc_disjoint_ecs_average(const int32_t *a, const int32_t *b, int64_t n) {
    if (n <= 0) return NAN;
    int32_t min_a, min_b, mx;
    minmax_i32(a, n, &min_a, &mx);
    minmax_i32(b, n, &min_b, &mx);
    int32_t ka, kb;
    int32_t *t = orc_cont_table(a, b, n, min_a, min_b, &ka, &kb);
    int32_t *rs = (int32_t *)malloc(sizeof(int32_t) * (size_t)ka);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)kb);
    row_col_sums(t, ka, kb, rs, cs);
    double avg = 0.0;
    for (int32_t i = 0; i < ka; i++)
        for (int32_t j = kb - 1; j >= 0; j--) {
            double inter = (double)t[(size_t)i + (size_t)j * ka];
            int32_t m = rs[i] > cs[j] ? rs[i] : cs[j];
            if (m == 0) continue; /* label unused in both partitions' ranges: reference would add NaN */
            avg += inter * inter / ((double)n * (double)m);
        }
    free(cs);
    free(rs);
    free(t);
    return avg;
}

/* corrected_l1_mb -- R/ECS.R:468-507 (the character-label route; formula at :491-496).  Kept as an
 * independent second statement of the closed form: 1 - 0.5*(|1/C1-1/C2|*I + (C1-I)/C1 + (C2-I)/C2).
 * alpha does not appear (R/ECS.R:325-328, 510-522: it cancels). */
int orc_corrected_l1_mb(const int32_t *a, const int32_t *b, int64_t n, double *ecs) {
    if (n <= 0) return 0;
    int32_t min_a, min_b, mx;
    minmax_i32(a, n, &min_a, &mx);
    minmax_i32(b, n, &min_b, &mx);
    int32_t ka, kb;
    int32_t *t = orc_cont_table(a, b, n, min_a, min_b, &ka, &kb);
    int32_t *rs = (int32_t *)malloc(sizeof(int32_t) * (size_t)ka);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)kb);
    row_col_sums(t, ka, kb, rs, cs);
    for (int64_t i = 0; i < n; i++) {
        int32_t ia = a[i] - min_a, ib = b[i] - min_b;
        double c1 = rs[ia], c2 = cs[ib], inter = t[(size_t)ia + (size_t)ib * ka];
        double tmp = (1.0 / c1 - 1.0 / c2) * inter;
        if (tmp < 0) tmp = -tmp;
        ecs[i] = 1.0 - 0.5 * (tmp + (c1 - inter) / c1 + (c2 - inter) / c2);
    }
    free(cs);
    free(rs);
    free(t);
    return 0;
}

/* R's mean() for doubles (summary.c real_mean): long-double sum, then one refinement pass. */
static double r_mean(const double *x, int64_t n) {
    long double s = 0.0L;
    for (int64_t i = 0; i < n; i++) s += x[i];
    s /= (long double)n;
    long double t = 0.0L;
    for (int64_t i = 0; i < n; i++) t += (x[i] - s);
    s += t / (long double)n;
    return (double)s;
}

/* weighted_element_consistency -- R/ECS.R:1446-1500.
 * norm = W(W-1)/2 (:1471-1472); for i: consistency += w_i/norm*(w_i-1)/2 (:1475); for j>i:
 * consistency += ((ecs*w_i)/norm)*w_j (:1484), sim[i,j]=sim[j,i]=mean(ecs) (:1479-1482), diag 1
 * (:1488,1497); finally the last partition's scalar term (:1491).  The scalar/vector additions are
 * performed in exactly that order for every element.  weights==NULL -> all 1 (:1466-1468).
 * sim_out (m x m, symmetric so layout-free) may be NULL.  Returns 1 if m < 2 (R stop(), :1451). */
int orc_weighted_element_consistency(const int32_t *labels, int64_t n, int32_t m, const double *weights,
                                     double *ecc_out, double *sim_out) {
    if (m <= 1) return 1;
    double *w = (double *)malloc(sizeof(double) * (size_t)m);
    for (int32_t i = 0; i < m; i++) w[i] = weights ? weights[i] : 1.0;
    double total = 0.0;
    for (int32_t i = 0; i < m; i++) total += w[i];
    double norm = total * (total - 1.0) / 2.0;
    double *cur = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t k = 0; k < n; k++) ecc_out[k] = 0.0;
    if (sim_out) memset(sim_out, 0, sizeof(double) * (size_t)m * (size_t)m);
    for (int32_t i = 0; i < m - 1; i++) {
        double sc = w[i] / norm * (w[i] - 1.0) / 2.0;
        for (int64_t k = 0; k < n; k++) ecc_out[k] += sc;
        for (int32_t j = i + 1; j < m; j++) {
            orc_disjoint_ecs(labels + (size_t)i * n, labels + (size_t)j * n, n, cur);
            if (sim_out) {
                double mu = r_mean(cur, n);
                sim_out[(size_t)i * m + j] = mu;
                sim_out[(size_t)j * m + i] = mu;
            }
            for (int64_t k = 0; k < n; k++) ecc_out[k] += cur[k] * w[i] / norm * w[j];
        }
        if (sim_out) sim_out[(size_t)i * m + i] = 1.0;
    }
    {
        double sc = w[m - 1] / norm * (w[m - 1] - 1.0) / 2.0;
        for (int64_t k = 0; k < n; k++) ecc_out[k] += sc;
    }
    if (sim_out) sim_out[(size_t)(m - 1) * m + (m - 1)] = 1.0;
    free(cur);
    free(w);
    return 0;
}

/* element_sim_matrix_flat_disjoint -- R/ECS.R:931-981: pair means in row-major upper-triangle order
 * (:942-958), mirrored, unit diagonal (:961-965); m==1 -> 1x1 matrix of 1 (:938-940). */
int orc_element_sim_matrix(const int32_t *labels, int64_t n, int32_t m, double *sim_out) {
    if (m == 1) {
        sim_out[0] = 1.0;
        return 0;
    }
    double *cur = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < m; i++) {
        sim_out[(size_t)i * m + i] = 1.0;
        for (int32_t j = i + 1; j < m; j++) {
            orc_disjoint_ecs(labels + (size_t)i * n, labels + (size_t)j * n, n, cur);
            double mu = r_mean(cur, n);
            sim_out[(size_t)i * m + j] = mu;
            sim_out[(size_t)j * m + i] = mu;
        }
    }
    free(cur);
    return 0;
}

/* are_identical_memberships -- R/ECS.R:984-1016 (integer route, :989): table != 0; FALSE unless
 * square (:994); TRUE if nrow == n (:999, all singletons); every column exactly one non-zero
 * (:1003-1009) and every row exactly one non-zero (:1011-1015).  A label unused inside its range
 * yields an all-zero row => never identical (the quirk in SURVEY.md section 8 a6). */
int orc_are_identical_memberships(const int32_t *a, const int32_t *b, int64_t n) {
    int32_t min_a, min_b, mx, ka, kb;
    minmax_i32(a, n, &min_a, &mx);
    minmax_i32(b, n, &min_b, &mx);
    int32_t *t = orc_cont_table(a, b, n, min_a, min_b, &ka, &kb);
    int res = 1;
    if (ka != kb) {
        res = 0;
    } else if ((int64_t)ka == n) {
        res = 1;
    } else {
        for (int32_t j = 0; j < kb && res; j++) {
            int nz = 0;
            for (int32_t i = 0; i < ka; i++) nz += (t[(size_t)i + (size_t)j * ka] != 0);
            if (nz != 1) res = 0;
        }
        for (int32_t i = 0; i < ka && res; i++) {
            int nz = 0;
            for (int32_t j = 0; j < kb; j++) nz += (t[(size_t)i + (size_t)j * ka] != 0);
            if (nz != 1) res = 0;
        }
    }
    free(t);
    return res;
}

/* merge_identical_partitions -- R/ECS.R:1021-1054: greedy; partition i joins the FIRST kept
 * partition it is identical to (:1036-1042), frequencies summed (:1046), else appended (:1049).
 * keep_idx_out[u] = index of the u-th kept partition, freq_out[u] its merged frequency.
 * freq_in may be NULL (all 1, as merge_partitions wraps plain vectors, :1182-1188).  Returns U. */
int32_t orc_merge_identical_partitions(const int32_t *labels, int64_t n, int32_t m, const double *freq_in,
                                       int32_t *keep_idx_out, double *freq_out) {
    int32_t u = 0;
    keep_idx_out[u] = 0;
    freq_out[u] = freq_in ? freq_in[0] : 1.0;
    u = 1;
    for (int32_t i = 1; i < m; i++) {
        int32_t assigned = -1;
        for (int32_t j = 0; j < u; j++) {
            if (orc_are_identical_memberships(labels + (size_t)keep_idx_out[j] * n, labels + (size_t)i * n, n)) {
                assigned = j;
                break;
            }
        }
        double f = freq_in ? freq_in[i] : 1.0;
        if (assigned >= 0) {
            freq_out[assigned] += f;
        } else {
            keep_idx_out[u] = i;
            freq_out[u] = f;
            u++;
        }
    }
    return u;
}

/* element_consistency -- R/ECS.R:1350-1386 via merge_partitions(ecs_thresh=1, order_logic="none")
 * R/ECS.R:1166-1226: dedup (:1204), U==1 -> rep(1,N) (:1210-1212), else weighted ECC with
 * weights = freq (:1214-1222).  Returns 1 if m < 2 (:1374-1376). */
int orc_element_consistency(const int32_t *labels, int64_t n, int32_t m, double *ecc_out) {
    if (m <= 1) return 1;
    int32_t *keep = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    double *freq = (double *)malloc(sizeof(double) * (size_t)m);
    int32_t u = orc_merge_identical_partitions(labels, n, m, NULL, keep, freq);
    if (u == 1) {
        for (int64_t k = 0; k < n; k++) ecc_out[k] = 1.0;
    } else {
        int32_t *sub = (int32_t *)malloc(sizeof(int32_t) * (size_t)u * (size_t)n);
        for (int32_t k = 0; k < u; k++) memcpy(sub + (size_t)k * n, labels + (size_t)keep[k] * n, sizeof(int32_t) * (size_t)n);
        orc_weighted_element_consistency(sub, n, u, freq, ecc_out, NULL);
        free(sub);
    }
    free(freq);
    free(keep);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Timing helper for bench.py's cpu_baseline / --impl reference legs: run the per-pair body of
 * weighted_element_consistency (disjointECS + mean + the three N-length vector ops of
 * R/ECS.R:1478-1484) over an explicit list of pairs, split over nthreads POSIX threads (the
 * analogue of %dopar% at R/ECS.R:953; nthreads=1 is the reference's sequential loop).
 * Returns wall seconds; ecc_out (may be NULL) receives the un-normalised sum for sanity checks.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const int32_t *labels;
    int64_t n;
    const int32_t *pi, *pj;
    int64_t first, last;
    double *acc;
    double msum;
    double *means;           /* [npairs] mean(ECS) of every pair (may be NULL) */
    const int64_t *cell_idx; /* [ncell] cells whose ECS is kept per pair (may be NULL) */
    int64_t ncell;
    double *cells;           /* [npairs][ncell] */
} pair_job_t;

static void *pair_worker(void *arg) {
    pair_job_t *jb = (pair_job_t *)arg;
    double *cur = (double *)malloc(sizeof(double) * (size_t)jb->n);
    for (int64_t p = jb->first; p < jb->last; p++) {
        orc_disjoint_ecs(jb->labels + (size_t)jb->pi[p] * jb->n, jb->labels + (size_t)jb->pj[p] * jb->n, jb->n, cur);
        const double mean = r_mean(cur, jb->n);
        jb->msum += mean;
        if (jb->means) jb->means[p] = mean;
        if (jb->cells)
            for (int64_t c = 0; c < jb->ncell; c++) jb->cells[(size_t)p * (size_t)jb->ncell + (size_t)c] = cur[jb->cell_idx[c]];
        for (int64_t k = 0; k < jb->n; k++) jb->acc[k] += cur[k] * 1.0 / 1.0 * 1.0;
    }
    free(cur);
    return NULL;
}

/* means_out[npairs], cells_out[npairs][ncell] (ECS of every pair at the cells cell_idx[]) may be NULL: they are what
 * bench.py's parity block and the full-size tests compare the GPU's ECS matrix and per-pair ECS with. */
double orc_time_pairs_ex(const int32_t *labels, int64_t n, const int32_t *pi, const int32_t *pj, int64_t npairs,
                         int32_t nthreads, double *ecc_out, double *means_out, const int64_t *cell_idx, int64_t ncell,
                         double *cells_out) {
    if (nthreads < 1) nthreads = 1;
    pair_job_t *jobs = (pair_job_t *)calloc((size_t)nthreads, sizeof(pair_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int32_t t = 0; t < nthreads; t++) {
        jobs[t].labels = labels;
        jobs[t].n = n;
        jobs[t].pi = pi;
        jobs[t].pj = pj;
        jobs[t].first = npairs * t / nthreads;
        jobs[t].last = npairs * (t + 1) / nthreads;
        jobs[t].acc = (double *)calloc((size_t)n, sizeof(double));
        jobs[t].means = means_out;
        jobs[t].cell_idx = cell_idx;
        jobs[t].ncell = ncell;
        jobs[t].cells = cell_idx ? cells_out : NULL;
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (nthreads == 1) {
        pair_worker(&jobs[0]);
    } else {
        for (int32_t t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pair_worker, &jobs[t]);
        for (int32_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (ecc_out) {
        for (int64_t k = 0; k < n; k++) ecc_out[k] = 0.0;
        for (int32_t t = 0; t < nthreads; t++)
            for (int64_t k = 0; k < n; k++) ecc_out[k] += jobs[t].acc[k];
    }
    for (int32_t t = 0; t < nthreads; t++) free(jobs[t].acc);
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double orc_time_pairs(const int32_t *labels, int64_t n, const int32_t *pi, const int32_t *pj, int64_t npairs,
                      int32_t nthreads, double *ecc_out) {
    return orc_time_pairs_ex(labels, n, pi, pj, npairs, nthreads, ecc_out, NULL, NULL, 0, NULL);
}
