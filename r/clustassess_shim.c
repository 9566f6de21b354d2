/* r/clustassess_shim.c -- route B of INTEGRATION.md: the same .Call entry points WITHOUT Rcpp, using only
 * R's C API (R.h, Rinternals.h).  It keeps the three symbol names the reference registers
 * (src/RcppExports.cpp:130-132: _ClustAssess_myContTable / _disjointECS / _disjointECSaverage, called from
 * R/RcppExports.R:4-14) and adds the batched ones.  Use it when the package's registration table is written
 * by hand (add ca_shim_entries[] to the R_CallMethodDef table handed to R_registerRoutines).
 *
 * Rules kept: inputs are never written; outputs are allocated here under PROTECT and
 * only filled by the library; Rf_error (a longjmp) is raised only after the library call has returned.
 * The build container has no R headers; tests/test_r_shim.py compiles this file against a miniature of R's C runtime
 * (tests/r_standin/) and drives every entry below the way .Call would, on the CPU (argument errors, no-device error,
 * registration table) and on the B200 (results against the oracle, PROTECT balance, inputs untouched). */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "clustassess_b200.h"

#define CA_R_CHECK(rc, nprot)                        \
    do {                                             \
        if ((rc) != CA_OK) {                         \
            UNPROTECT(nprot);                        \
            Rf_error("%s", ca_last_error());         \
        }                                            \
    } while (0)

/* Rcpp's input_parameter<IntegerVector>: INTSXP (and factors) as is, REALSXP truncated, else an error. */
static SEXP as_int(SEXP x) { return TYPEOF(x) == INTSXP ? x : Rf_coerceVector(x, INTSXP); }

SEXP _ClustAssess_myContTable(SEXP a, SEXP b, SEXP min_a, SEXP min_b) {
    SEXP ai = PROTECT(as_int(a)), bi = PROTECT(as_int(b));
    int32_t lo, hi_a, hi_b, mina = Rf_asInteger(min_a), minb = Rf_asInteger(min_b);
    int rc = ca_label_range(INTEGER(ai), XLENGTH(ai), &lo, &hi_a);
    if (rc == CA_OK) rc = ca_label_range(INTEGER(bi), XLENGTH(bi), &lo, &hi_b);
    CA_R_CHECK(rc, 2);
    int ka = hi_a - mina + 1, kb = hi_b - minb + 1;
    SEXP res = PROTECT(Rf_allocMatrix(INTSXP, ka, kb));
    rc = ca_contingency(INTEGER(ai), INTEGER(bi), XLENGTH(ai), mina, minb, ka, kb, INTEGER(res), NULL, NULL);
    CA_R_CHECK(rc, 3);
    UNPROTECT(3);
    return res;
}

SEXP _ClustAssess_disjointECS(SEXP mb1, SEXP mb2) {
    SEXP a = PROTECT(as_int(mb1)), b = PROTECT(as_int(mb2));
    if (XLENGTH(a) != XLENGTH(b)) { UNPROTECT(2); Rf_error("clustering1 and clustering2 do not have the same length."); }
    SEXP res = PROTECT(Rf_allocVector(REALSXP, XLENGTH(a)));
    int rc = ca_ecs_pair(INTEGER(a), INTEGER(b), XLENGTH(a), REAL(res), NULL);
    CA_R_CHECK(rc, 3);
    UNPROTECT(3);
    return res;
}

SEXP _ClustAssess_disjointECSaverage(SEXP mb1, SEXP mb2) {
    SEXP a = PROTECT(as_int(mb1)), b = PROTECT(as_int(mb2));
    if (XLENGTH(a) != XLENGTH(b)) { UNPROTECT(2); Rf_error("clustering1 and clustering2 do not have the same length."); }
    double avg = NA_REAL;
    int rc = ca_ecs_pair_mean(INTEGER(a), INTEGER(b), XLENGTH(a), &avg);
    CA_R_CHECK(rc, 2);
    UNPROTECT(2);
    return Rf_ScalarReal(avg);
}

/* list of M membership vectors -> M protected INTSXPs + an array of column pointers (no packing copy) */
static const int32_t **columns(SEXP lst, SEXP *holder, R_xlen_t *n) {
    int m = Rf_length(lst);
    *holder = PROTECT(Rf_allocVector(VECSXP, m));
    const int32_t **cols = (const int32_t **)R_alloc(m, sizeof(int32_t *));
    *n = -1;
    for (int p = 0; p < m; p++) {
        SET_VECTOR_ELT(*holder, p, as_int(VECTOR_ELT(lst, p)));
        SEXP v = VECTOR_ELT(*holder, p);
        if (*n < 0) *n = XLENGTH(v);
        if (XLENGTH(v) != *n) { UNPROTECT(1); Rf_error("The partitions do not have the same length"); }
        cols[p] = INTEGER(v);
    }
    return cols;
}

/* .Call("_ClustAssess_ecs_consistency", mb_list, weights_or_NULL, calculate_sim_matrix) -> list(ecc, ecs_matrix) */
SEXP _ClustAssess_ecs_consistency(SEXP lst, SEXP weights, SEXP want_matrix) {
    int m = Rf_length(lst), want = Rf_asLogical(want_matrix) == TRUE;
    if (m <= 1) Rf_error("At least two clusterings are required to calculate the consistency.");
    SEXP holder;
    R_xlen_t n;
    const int32_t **cols = columns(lst, &holder, &n);
    SEXP w = PROTECT(Rf_isNull(weights) ? R_NilValue : Rf_coerceVector(weights, REALSXP));
    if (!Rf_isNull(w) && XLENGTH(w) != m) { UNPROTECT(2); Rf_error("weights must have one entry per clustering"); }
    SEXP ecc = PROTECT(Rf_allocVector(REALSXP, n));
    SEXP sim = PROTECT(want ? Rf_allocMatrix(REALSXP, m, m) : R_NilValue);
    int rc = ca_consistency_cols(cols, n, m, Rf_isNull(w) ? NULL : REAL(w), REAL(ecc), want ? REAL(sim) : NULL);
    CA_R_CHECK(rc, 4);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, ecc); SET_VECTOR_ELT(out, 1, sim);
    SET_STRING_ELT(nm, 0, Rf_mkChar("ecc")); SET_STRING_ELT(nm, 1, Rf_mkChar("ecs_matrix"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(6);
    return out;
}

/* .Call("_ClustAssess_ecs_sim_matrix", mb_list) -> M x M numeric matrix */
SEXP _ClustAssess_ecs_sim_matrix(SEXP lst) {
    int m = Rf_length(lst);
    SEXP holder;
    R_xlen_t n;
    const int32_t **cols = columns(lst, &holder, &n);
    SEXP sim = PROTECT(Rf_allocMatrix(REALSXP, m, m));
    int rc = ca_sim_matrix_cols(cols, n, m, REAL(sim));
    CA_R_CHECK(rc, 2);
    UNPROTECT(2);
    return sim;
}

/* .Call("_ClustAssess_ecs_agreement", reference, mb_list) -> numeric vector: mean ECS of every clustering of the list against
 * the reference (element_agreement's flat branch, R/ECS.R:1639-1647) */
SEXP _ClustAssess_ecs_agreement(SEXP reference, SEXP lst) {
    int m = Rf_length(lst);
    if (m < 1) Rf_error("At least one clustering is required.");
    SEXP ref = PROTECT(as_int(reference));
    SEXP holder;
    R_xlen_t n;
    const int32_t **cols = columns(lst, &holder, &n);
    if (XLENGTH(ref) != n) { UNPROTECT(2); Rf_error("The partitions do not have the same length"); }
    SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
    int rc = ca_agreement_cols(INTEGER(ref), cols, n, m, REAL(out));
    CA_R_CHECK(rc, 3);
    UNPROTECT(3);
    return out;
}

/* .Call("_ClustAssess_ecs_merge_identical", mb_list, freq) -> list(keep = 1-based indices of the kept partitions, freq = their
 * summed frequencies): merge_identical_partitions (R/ECS.R:1021-1054) without one contingency table per comparison */
SEXP _ClustAssess_ecs_merge_identical(SEXP lst, SEXP freq) {
    int m = Rf_length(lst);
    if (m < 1) Rf_error("At least one clustering is required.");
    SEXP f = PROTECT(Rf_coerceVector(freq, REALSXP));
    if (XLENGTH(f) != m) { UNPROTECT(1); Rf_error("freq must have one entry per clustering"); }
    SEXP holder;
    R_xlen_t n;
    const int32_t **cols = columns(lst, &holder, &n);
    int32_t *keep = (int32_t *)R_alloc(m, sizeof(int32_t));
    double *fout = (double *)R_alloc(m, sizeof(double));
    int32_t u = 0;
    int rc = ca_merge_identical_cols(cols, n, m, REAL(f), keep, fout, &u);
    CA_R_CHECK(rc, 2);
    SEXP k1 = PROTECT(Rf_allocVector(INTSXP, u)), f1 = PROTECT(Rf_allocVector(REALSXP, u));
    for (int t = 0; t < u; t++) {
        INTEGER(k1)[t] = keep[t] + 1; /* 1-based for R */
        REAL(f1)[t] = fout[t];
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, k1); SET_VECTOR_ELT(out, 1, f1);
    SET_STRING_ELT(nm, 0, Rf_mkChar("keep")); SET_STRING_ELT(nm, 1, Rf_mkChar("freq"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(6);
    return out;
}

/* ---- pipeline glue: the runs of one configuration stay on the device (R/stability-3-graph-clustering.R:25-69, 271-343 compare
 * the same membership vectors up to three times).  The external pointer owns the device memory; its finalizer releases it. */
static void resident_finalizer(SEXP ptr) {
    ca_labels_t *h = (ca_labels_t *)R_ExternalPtrAddr(ptr);
    if (h) ca_labels_destroy(h);
    R_ClearExternalPtr(ptr);
}
static ca_labels_t *resident_handle(SEXP ptr) {
    if (TYPEOF(ptr) != EXTPTRSXP) Rf_error("not a resident label set");
    ca_labels_t *h = (ca_labels_t *)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("the resident label set has been released");
    return h;
}
/* 1-based R indices -> 0-based, range-checked against the set (an index outside it is an R error, not a library error code) */
static int32_t *resident_indices(SEXP idx, ca_labels_t *h, int *k_out) {
    SEXP ix = PROTECT(as_int(idx));
    int k = Rf_length(ix);
    int64_t n;
    int32_t m;
    ca_labels_shape(h, &n, &m);
    int32_t *out = (int32_t *)R_alloc(k > 0 ? k : 1, sizeof(int32_t));
    for (int t = 0; t < k; t++) {
        int v = INTEGER(ix)[t];
        if (v < 1 || v > m) { UNPROTECT(1); Rf_error("index %d outside the resident label set (1..%d)", v, (int)m); }
        out[t] = v - 1;
    }
    UNPROTECT(1);
    *k_out = k;
    return out;
}

/* .Call("_ClustAssess_ecs_resident", mb_list) -> external pointer */
SEXP _ClustAssess_ecs_resident(SEXP lst) {
    int m = Rf_length(lst);
    if (m < 1) Rf_error("At least one clustering is required.");
    SEXP holder;
    R_xlen_t n;
    const int32_t **cols = columns(lst, &holder, &n);
    ca_labels_t *h = NULL;
    int rc = ca_labels_create(n, m, &h);
    CA_R_CHECK(rc, 1);
    rc = ca_labels_upload_cols(h, cols); /* one staged upload of R's (pageable) vectors */
    if (rc != CA_OK) ca_labels_destroy(h);
    CA_R_CHECK(rc, 1);
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, resident_finalizer, TRUE);
    UNPROTECT(2);
    return ptr;
}

/* .Call("_ClustAssess_ecs_resident_consistency", ptr, idx (1-based), weights_or_NULL, calculate_sim_matrix) -> list(ecc, ecs_matrix).
 * The length of ecc comes from the label set itself (ca_labels_shape), never from the caller. */
SEXP _ClustAssess_ecs_resident_consistency(SEXP ptr, SEXP idx, SEXP weights, SEXP want_matrix) {
    ca_labels_t *h = resident_handle(ptr);
    int want = Rf_asLogical(want_matrix) == TRUE, k = 0;
    int32_t *ix = resident_indices(idx, h, &k);
    if (k <= 1) Rf_error("At least two clusterings are required to calculate the consistency.");
    SEXP w = PROTECT(Rf_isNull(weights) ? R_NilValue : Rf_coerceVector(weights, REALSXP));
    if (!Rf_isNull(w) && XLENGTH(w) != k) { UNPROTECT(1); Rf_error("weights must have one entry per clustering"); }
    int64_t n;
    int32_t m;
    ca_labels_shape(h, &n, &m);
    SEXP ecc = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
    SEXP sim = PROTECT(want ? Rf_allocMatrix(REALSXP, k, k) : R_NilValue);
    int rc = ca_consistency_resident(h, ix, k, Rf_isNull(w) ? NULL : REAL(w), REAL(ecc), want ? REAL(sim) : NULL);
    CA_R_CHECK(rc, 3);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, ecc); SET_VECTOR_ELT(out, 1, sim);
    SET_STRING_ELT(nm, 0, Rf_mkChar("ecc")); SET_STRING_ELT(nm, 1, Rf_mkChar("ecs_matrix"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

/* .Call("_ClustAssess_ecs_resident_sim_matrix", ptr, idx (1-based)) -> k x k numeric matrix */
SEXP _ClustAssess_ecs_resident_sim_matrix(SEXP ptr, SEXP idx) {
    ca_labels_t *h = resident_handle(ptr);
    int k = 0;
    int32_t *ix = resident_indices(idx, h, &k);
    if (k < 1) Rf_error("At least one clustering is required.");
    SEXP sim = PROTECT(Rf_allocMatrix(REALSXP, k, k));
    int rc = ca_sim_matrix_resident(h, ix, k, REAL(sim));
    CA_R_CHECK(rc, 1);
    UNPROTECT(1);
    return sim;
}

const R_CallMethodDef ca_shim_entries[] = {
    {"_ClustAssess_myContTable", (DL_FUNC)&_ClustAssess_myContTable, 4},
    {"_ClustAssess_disjointECS", (DL_FUNC)&_ClustAssess_disjointECS, 2},
    {"_ClustAssess_disjointECSaverage", (DL_FUNC)&_ClustAssess_disjointECSaverage, 2},
    {"_ClustAssess_ecs_consistency", (DL_FUNC)&_ClustAssess_ecs_consistency, 3},
    {"_ClustAssess_ecs_sim_matrix", (DL_FUNC)&_ClustAssess_ecs_sim_matrix, 1},
    {"_ClustAssess_ecs_agreement", (DL_FUNC)&_ClustAssess_ecs_agreement, 2},
    {"_ClustAssess_ecs_merge_identical", (DL_FUNC)&_ClustAssess_ecs_merge_identical, 2},
    {"_ClustAssess_ecs_resident", (DL_FUNC)&_ClustAssess_ecs_resident, 1},
    {"_ClustAssess_ecs_resident_consistency", (DL_FUNC)&_ClustAssess_ecs_resident_consistency, 4},
    {"_ClustAssess_ecs_resident_sim_matrix", (DL_FUNC)&_ClustAssess_ecs_resident_sim_matrix, 2},
    {NULL, NULL, 0}};
