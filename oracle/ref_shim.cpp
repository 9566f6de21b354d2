// TEST INFRASTRUCTURE ONLY.  extern "C" doorway onto the UNMODIFIED reference functions in
// /root/reference/src/calculate_ecs.cpp (compiled from where it lies, never copied), so that
// python can call them through ctypes.  Built into oracle/_ref/libref_ecs.so by oracle/Makefile.
//   myContTable        calculate_ecs.cpp:7-19
//   disjointECS        calculate_ecs.cpp:22-46
//   disjointECSaverage calculate_ecs.cpp:49-70
#include <Rcpp.h>
#include <cstdint>
#include <cstring>

Rcpp::IntegerMatrix myContTable(Rcpp::IntegerVector& a, Rcpp::IntegerVector& b, int minim_mb1, int minim_mb2);
Rcpp::NumericVector disjointECS(Rcpp::IntegerVector mb1, Rcpp::IntegerVector mb2);
double disjointECSaverage(Rcpp::IntegerVector mb1, Rcpp::IntegerVector mb2);

extern "C" {

// table_out must hold (max(a)-min_a+1)*(max(b)-min_b+1) ints; column-major like the R matrix.
int ref_myContTable(const int32_t* a, const int32_t* b, int64_t n, int32_t min_a, int32_t min_b,
                    int32_t* table_out, int32_t* nrow_out, int32_t* ncol_out) {
    Rcpp::IntegerVector va(a, a + n), vb(b, b + n);
    Rcpp::IntegerMatrix t = myContTable(va, vb, min_a, min_b);
    *nrow_out = t.nrow();
    *ncol_out = t.ncol();
    if (table_out) std::memcpy(table_out, t.begin(), sizeof(int32_t) * t.length());
    return 0;
}

int ref_disjointECS(const int32_t* a, const int32_t* b, int64_t n, double* ecs_out) {
    Rcpp::IntegerVector va(a, a + n), vb(b, b + n);
    Rcpp::NumericVector e = disjointECS(va, vb);
    std::memcpy(ecs_out, e.begin(), sizeof(double) * static_cast<size_t>(n));
    return 0;
}

double ref_disjointECSaverage(const int32_t* a, const int32_t* b, int64_t n) {
    Rcpp::IntegerVector va(a, a + n), vb(b, b + n);
    return disjointECSaverage(va, vb);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Timing helper for bench.py (--impl reference / cpu_baseline kind "reference"): the reference's own
// disjointECS (unmodified) inside a restatement of the per-pair body of
// weighted_element_consistency (R/ECS.R:1478-1484: disjointECS, mean, scaled accumulate), over an
// explicit pair list split across nthreads std::threads (nthreads=1 == the reference's sequential
// loop; >1 is the %dopar% analogue of R/ECS.R:953).  Returns wall seconds.
// ---------------------------------------------------------------------------------------------
#include <chrono>
#include <thread>
#include <vector>

extern "C" double ref_time_pairs_ex(const int32_t* labels, int64_t n, const int32_t* pi, const int32_t* pj,
                                    int64_t npairs, int32_t nthreads, double* ecc_out, double* means_out,
                                    const int64_t* cell_idx, int64_t ncell, double* cells_out) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::vector<double> > acc(nthreads, std::vector<double>(static_cast<size_t>(n), 0.0));
    std::vector<double> msum(nthreads, 0.0);
    auto body = [&](int t) {
        int64_t first = npairs * t / nthreads, last = npairs * (t + 1) / nthreads;
        for (int64_t p = first; p < last; p++) {
            const int32_t* a = labels + static_cast<size_t>(pi[p]) * n;
            const int32_t* b = labels + static_cast<size_t>(pj[p]) * n;
            Rcpp::IntegerVector va(a, a + n), vb(b, b + n);
            Rcpp::NumericVector e = disjointECS(va, vb);
            long double s = 0.0L;
            for (int64_t k = 0; k < n; k++) s += e(static_cast<int>(k));
            const double mean = static_cast<double>(s / n);
            msum[t] += mean;
            if (means_out) means_out[p] = mean;  // per-pair mean(ECS): what the GPU's ECS matrix is compared with
            if (cell_idx && cells_out)
                for (int64_t c = 0; c < ncell; c++)
                    cells_out[static_cast<size_t>(p) * static_cast<size_t>(ncell) + static_cast<size_t>(c)] = e(static_cast<int>(cell_idx[c]));
            double* ac = acc[t].data();
            for (int64_t k = 0; k < n; k++) ac[k] += e(static_cast<int>(k)) * 1.0 / 1.0 * 1.0;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(body, t);
        for (auto& x : th) x.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (ecc_out) {
        for (int64_t k = 0; k < n; k++) {
            double s = 0.0;
            for (int t = 0; t < nthreads; t++) s += acc[t][k];
            ecc_out[k] = s;
        }
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

extern "C" double ref_time_pairs(const int32_t* labels, int64_t n, const int32_t* pi, const int32_t* pj,
                                 int64_t npairs, int32_t nthreads, double* ecc_out) {
    return ref_time_pairs_ex(labels, n, pi, pj, npairs, nthreads, ecc_out, nullptr, nullptr, 0, nullptr);
}
