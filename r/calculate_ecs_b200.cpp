// r/calculate_ecs_b200.cpp -- drop-in replacement for ClustAssess's src/calculate_ecs.cpp (route A of
// INTEGRATION.md).  Same three exported routines, same signatures, same return types -- so
// Rcpp::compileAttributes() regenerates R/RcppExports.R:4-14 and the registration table of
// src/RcppExports.cpp:129-145 unchanged -- but the bodies forward to the B200 library's C ABI
// (include/clustassess_b200.h) instead of computing on the CPU.  Four batched routines are added for the
// set-level R functions.  Link with  PKG_LIBS = -L$(CLUSTASSESS_B200_HOME) -lclustassess_b200  in src/Makevars.
//
// The build container has neither R nor Rcpp: tests/test_r_glue.py compiles this file UNMODIFIED against a stand-in
// Rcpp.h (tests/rcpp_standin/) and calls every exported routine -- argument errors and the no-device error on the CPU,
// results against the oracle on the B200.  clustassess_b200/_capi.py is the Python twin over the same entry points.
#include <Rcpp.h>
#include "clustassess_b200.h"

using namespace Rcpp;

static inline void ca_check(int rc) {
    // The C ABI never throws and never calls R; turn its error code into an R condition here, after every
    // C++ frame of the library has unwound (END_RCPP converts the exception, src/RcppExports.cpp:26).
    if (rc != CA_OK) stop("%s", ca_last_error());
}

// [[Rcpp::export]]
IntegerMatrix myContTable(IntegerVector &a, IntegerVector &b, int minim_mb1, int minim_mb2) {
    int lo, hi_a, hi_b;
    ca_check(ca_label_range(a.begin(), a.size(), &lo, &hi_a));
    ca_check(ca_label_range(b.begin(), b.size(), &lo, &hi_b));
    const int ka = hi_a - minim_mb1 + 1, kb = hi_b - minim_mb2 + 1;
    IntegerMatrix result(ka, kb);  // column-major, exactly what ca_contingency fills
    ca_check(ca_contingency(a.begin(), b.begin(), a.size(), minim_mb1, minim_mb2, ka, kb, result.begin(), NULL, NULL));
    return result;
}

// [[Rcpp::export]]
NumericVector disjointECS(IntegerVector mb1, IntegerVector mb2) {
    if (mb1.size() != mb2.size()) stop("clustering1 and clustering2 do not have the same length.");
    NumericVector ecsScore(mb1.size());
    ca_check(ca_ecs_pair(mb1.begin(), mb2.begin(), mb1.size(), ecsScore.begin(), NULL));
    return ecsScore;
}

// [[Rcpp::export]]
double disjointECSaverage(IntegerVector mb1, IntegerVector mb2) {
    if (mb1.size() != mb2.size()) stop("clustering1 and clustering2 do not have the same length.");
    double avg = NA_REAL;
    ca_check(ca_ecs_pair_mean(mb1.begin(), mb2.begin(), mb1.size(), &avg));
    return avg;
}

// ---- batched routines: one upload, all pairs on the GPU -------------------------------------------------
static std::vector<const int32_t *> column_pointers(List mb_list, std::vector<IntegerVector> &keep, R_xlen_t &n) {
    // as<IntegerVector> coerces numeric (truncating) and factor (codes) inputs exactly like the reference's
    // input_parameter<IntegerVector> (src/RcppExports.cpp:34-35); integer vectors are NOT copied.
    const int m = mb_list.size();
    keep.reserve(m);
    std::vector<const int32_t *> cols(m);
    n = -1;
    for (int p = 0; p < m; p++) {
        keep.push_back(as<IntegerVector>(mb_list[p]));
        if (n < 0) n = keep[p].size();
        if (keep[p].size() != n) stop("The partitions do not have the same length");
        cols[p] = keep[p].begin();
    }
    return cols;
}

// [[Rcpp::export]]
List ecsConsistency(List mb_list, Nullable<NumericVector> weights, bool calculate_sim_matrix) {
    const int m = mb_list.size();
    if (m <= 1) stop("At least two clusterings are required to calculate the consistency.");
    std::vector<IntegerVector> keep;
    R_xlen_t n;
    std::vector<const int32_t *> cols = column_pointers(mb_list, keep, n);
    NumericVector ecc(n);
    NumericMatrix sim(calculate_sim_matrix ? m : 0, calculate_sim_matrix ? m : 0);
    const double *w = NULL;
    NumericVector wv;
    if (weights.isNotNull()) {
        wv = NumericVector(weights);
        if (wv.size() != m) stop("weights must have one entry per clustering");
        w = wv.begin();
    }
    ca_check(ca_consistency_cols(cols.data(), n, m, w, ecc.begin(), calculate_sim_matrix ? sim.begin() : NULL));
    return List::create(Named("ecc") = ecc, Named("ecs_matrix") = sim);
}

// [[Rcpp::export]]
NumericMatrix ecsSimMatrix(List mb_list) {
    const int m = mb_list.size();
    std::vector<IntegerVector> keep;
    R_xlen_t n;
    std::vector<const int32_t *> cols = column_pointers(mb_list, keep, n);
    NumericMatrix sim(m, m);  // symmetric, so R's column-major layout needs no transposition
    ca_check(ca_sim_matrix_cols(cols.data(), n, m, sim.begin()));
    return sim;
}

// [[Rcpp::export]]
NumericVector ecsAgreement(IntegerVector reference, List mb_list) {
    const int m = mb_list.size();
    std::vector<IntegerVector> keep;
    R_xlen_t n;
    std::vector<const int32_t *> cols = column_pointers(mb_list, keep, n);
    if (reference.size() != n) stop("The partitions do not have the same length");
    NumericVector out(n);
    ca_check(ca_agreement_cols(reference.begin(), cols.data(), n, m, out.begin()));
    return out;
}

// [[Rcpp::export]]
List ecsMergeIdentical(List mb_list, NumericVector freq) {
    const int m = mb_list.size();
    std::vector<IntegerVector> keep;
    R_xlen_t n;
    std::vector<const int32_t *> cols = column_pointers(mb_list, keep, n);
    if (freq.size() != m) stop("freq must have one entry per clustering");
    IntegerVector kept(m);
    NumericVector fout(m);
    int u = 0;
    ca_check(ca_merge_identical_cols(cols.data(), n, m, freq.begin(), kept.begin(), fout.begin(), &u));
    IntegerVector kept1(u);
    NumericVector f1(u);
    for (int t = 0; t < u; t++) {
        kept1[t] = kept[t] + 1;  // 1-based for R
        f1[t] = fout[t];
    }
    return List::create(Named("keep") = kept1, Named("freq") = f1);
}

// ---- pipeline glue: the runs of one configuration stay on the device -------------
// R/stability-3-graph-clustering.R compares the same membership vectors per k (:271-300), per resolution (:25-69)
// and per k across resolutions (:303-343).  ecsResident(mb_list) uploads them once and returns an external pointer;
// ecsResidentConsistency(ptr, idx, weights, calculate_sim_matrix) / ecsResidentSimMatrix(ptr, idx) run a comparison of
// the partitions idx (1-based); the length of the result comes from the label set itself (ca_labels_shape), never from
// the caller.  The finalizer releases the device memory with the R object.
static void resident_finalizer(ca_labels_t *h) { ca_labels_destroy(h); }
typedef XPtr<ca_labels_t, PreserveStorage, resident_finalizer, true> ResidentPtr;

// [[Rcpp::export]]
SEXP ecsResident(List mb_list) {
    const int m = mb_list.size();
    if (m < 1) stop("At least one clustering is required.");
    std::vector<IntegerVector> keep;
    R_xlen_t n;
    std::vector<const int32_t *> cols = column_pointers(mb_list, keep, n);
    ca_labels_t *h = NULL;
    ca_check(ca_labels_create(n, m, &h));
    int rc = ca_labels_upload_cols(h, cols.data());  // one staged upload of R's (pageable) vectors
    if (rc != CA_OK) {
        ca_labels_destroy(h);
        ca_check(rc);
    }
    return ResidentPtr(h, true);
}

// [[Rcpp::export]]
List ecsResidentConsistency(SEXP resident, IntegerVector idx, Nullable<NumericVector> weights, bool calculate_sim_matrix) {
    ResidentPtr h(resident);
    const int k = idx.size();
    if (k <= 1) stop("At least two clusterings are required to calculate the consistency.");
    int64_t n_cells = 0;
    int32_t m_set = 0;
    ca_check(ca_labels_shape(h.get(), &n_cells, &m_set));
    std::vector<int32_t> ix(idx.begin(), idx.end());
    for (size_t t = 0; t < ix.size(); t++) {
        if (ix[t] < 1 || ix[t] > m_set) stop("index %d outside the resident label set (1..%d)", ix[t], (int)m_set);
        ix[t] -= 1;  // R indices are 1-based
    }
    const double *w = NULL;
    NumericVector wv;
    if (weights.isNotNull()) {
        wv = NumericVector(weights);
        if (wv.size() != k) stop("weights must have one entry per clustering");
        w = wv.begin();
    }
    NumericVector ecc(n_cells);
    NumericMatrix sim(calculate_sim_matrix ? k : 0, calculate_sim_matrix ? k : 0);
    ca_check(ca_consistency_resident(h.get(), ix.data(), k, w, ecc.begin(), calculate_sim_matrix ? sim.begin() : NULL));
    return List::create(Named("ecc") = ecc, Named("ecs_matrix") = sim);
}

// [[Rcpp::export]]
NumericMatrix ecsResidentSimMatrix(SEXP resident, IntegerVector idx) {
    ResidentPtr h(resident);
    int64_t n_cells = 0;
    int32_t m_set = 0;
    ca_check(ca_labels_shape(h.get(), &n_cells, &m_set));
    std::vector<int32_t> ix(idx.begin(), idx.end());
    for (size_t t = 0; t < ix.size(); t++) {
        if (ix[t] < 1 || ix[t] > m_set) stop("index %d outside the resident label set (1..%d)", ix[t], (int)m_set);
        ix[t] -= 1;
    }
    NumericMatrix sim(ix.size(), ix.size());
    ca_check(ca_sim_matrix_resident(h.get(), ix.data(), (int)ix.size(), sim.begin()));
    return sim;
}
