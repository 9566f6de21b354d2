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

const R_CallMethodDef ca_shim_entries[] = {
    {"_ClustAssess_myContTable", (DL_FUNC)&_ClustAssess_myContTable, 4},
    {"_ClustAssess_disjointECS", (DL_FUNC)&_ClustAssess_disjointECS, 2},
    {"_ClustAssess_disjointECSaverage", (DL_FUNC)&_ClustAssess_disjointECSaverage, 2},
    {"_ClustAssess_ecs_consistency", (DL_FUNC)&_ClustAssess_ecs_consistency, 3},
    {"_ClustAssess_ecs_sim_matrix", (DL_FUNC)&_ClustAssess_ecs_sim_matrix, 1},
    {NULL, NULL, 0}};
