// glue_harness.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the Rcpp glue r/calculate_ecs_b200.cpp UNMODIFIED against the
// stand-in Rcpp.h of this directory and exposes its exported functions through a C interface that tests/test_r_glue.py
// binds with ctypes.  Every call plays the part of BEGIN_RCPP / END_RCPP (src/RcppExports.cpp:17,26): a C++ exception
// raised by the glue (Rcpp::stop) becomes return code 1 + gh_error(), i.e. what R would see as a condition.
#include "../../r/calculate_ecs_b200.cpp"

#include <cstring>

static std::string g_msg;
template <class F> static int guarded(F f) {
    g_msg.clear();
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_msg = e.what();
        return 1;
    }
}
static List make_list(const int* const* cols, const R_xlen_t* lens, int m, int real_every) {
    // every `real_every`-th membership vector is handed over as a REALSXP (value + 0.5) to exercise as<IntegerVector>
    List l;
    for (int p = 0; p < m; p++) {
        if (real_every > 0 && p % real_every == real_every - 1) {
            NumericVector d(lens[p]);
            for (R_xlen_t i = 0; i < lens[p]; i++) d[i] = cols[p][i] + (cols[p][i] >= 0 ? 0.5 : -0.5);
            l.push_back(wrap(d));
        } else {
            l.push_back(wrap(IntegerVector(cols[p], cols[p] + lens[p])));
        }
    }
    return l;
}

extern "C" {
const char* gh_error(void) { return g_msg.c_str(); }
void gh_release(void) { standin_arena().clear(); }  // "garbage collection": runs the external pointers' finalizers

int gh_cont_table(const int* a, const int* b, R_xlen_t n, int min_a, int min_b, int* out, int cap, int* ka, int* kb) {
    return guarded([&] {
        IntegerVector va(a, a + n), vb(b, b + n);
        IntegerMatrix t = myContTable(va, vb, min_a, min_b);
        *ka = t.nrow();
        *kb = t.ncol();
        if ((long long)t.nrow() * t.ncol() > cap) stop("harness: table larger than the caller's buffer");
        std::memcpy(out, t.begin(), sizeof(int) * (size_t)t.nrow() * t.ncol());
    });
}
int gh_disjoint_ecs(const int* a, R_xlen_t na, const int* b, R_xlen_t nb, double* out) {
    return guarded([&] {
        NumericVector r = disjointECS(IntegerVector(a, a + na), IntegerVector(b, b + nb));
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)r.size());
    });
}
int gh_disjoint_ecs_average(const int* a, R_xlen_t na, const int* b, R_xlen_t nb, double* out) {
    return guarded([&] { *out = disjointECSaverage(IntegerVector(a, a + na), IntegerVector(b, b + nb)); });
}
int gh_consistency(const int* const* cols, const R_xlen_t* lens, int m, int real_every, const double* w, int nw, int want_matrix,
                   double* ecc, double* sim) {
    return guarded([&] {
        Nullable<NumericVector> weights;
        if (w) weights = Nullable<NumericVector>(wrap(NumericVector(w, w + nw)));
        List res = ecsConsistency(make_list(cols, lens, m, real_every), weights, want_matrix != 0);
        if (res.names() != std::vector<std::string>{"ecc", "ecs_matrix"}) stop("harness: unexpected names in the result list");
        NumericVector e = as<NumericVector>(res.get("ecc"));
        std::memcpy(ecc, e.begin(), sizeof(double) * (size_t)e.size());
        SEXP s = res.get("ecs_matrix");
        if (want_matrix) {
            if (s->dm->nrow() != m || s->dm->ncol() != m) stop("harness: ecs_matrix is not m x m");
            std::memcpy(sim, s->dm->begin(), sizeof(double) * (size_t)m * m);
        } else if (s->dm->nrow() != 0) {
            stop("harness: ecs_matrix should be 0 x 0 when it was not asked for");
        }
    });
}
int gh_sim_matrix(const int* const* cols, const R_xlen_t* lens, int m, int real_every, double* sim) {
    return guarded([&] {
        NumericMatrix s = ecsSimMatrix(make_list(cols, lens, m, real_every));
        std::memcpy(sim, s.begin(), sizeof(double) * (size_t)m * m);
    });
}
int gh_agreement(const int* ref, R_xlen_t nref, const int* const* cols, const R_xlen_t* lens, int m, double* out) {
    return guarded([&] {
        NumericVector r = ecsAgreement(IntegerVector(ref, ref + nref), make_list(cols, lens, m, 0));
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)r.size());
    });
}
int gh_merge_identical(const int* const* cols, const R_xlen_t* lens, int m, const double* freq, int nfreq, int* keep, double* fout, int* u) {
    return guarded([&] {
        List res = ecsMergeIdentical(make_list(cols, lens, m, 0), NumericVector(freq, freq + nfreq));
        IntegerVector k = as<IntegerVector>(res.get("keep"));
        NumericVector f = as<NumericVector>(res.get("freq"));
        *u = (int)k.size();
        for (int t = 0; t < *u; t++) {
            keep[t] = k[t];  // 1-based, as R gets it
            fout[t] = f[t];
        }
    });
}
// ecsResident once, then two comparisons on index subsets without another upload
int gh_resident(const int* const* cols, const R_xlen_t* lens, int m, const int* idx1, int k, const double* w, int nw, int want_matrix, double* ecc,
                double* sim, double* sim_only) {
    return guarded([&] {
        SEXP handle = ecsResident(make_list(cols, lens, m, 0));
        Nullable<NumericVector> weights;
        if (w) weights = Nullable<NumericVector>(wrap(NumericVector(w, w + nw)));
        List res = ecsResidentConsistency(handle, IntegerVector(idx1, idx1 + k), weights, want_matrix != 0);
        NumericVector e = as<NumericVector>(res.get("ecc"));
        std::memcpy(ecc, e.begin(), sizeof(double) * (size_t)e.size());
        if (want_matrix) std::memcpy(sim, res.get("ecs_matrix")->dm->begin(), sizeof(double) * (size_t)k * k);
        NumericMatrix s = ecsResidentSimMatrix(handle, IntegerVector(idx1, idx1 + k));
        std::memcpy(sim_only, s.begin(), sizeof(double) * (size_t)k * k);
    });
}
}
