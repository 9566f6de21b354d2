// Stand-in for <Rcpp.h> -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// It provides just enough of the Rcpp surface for the reference's
// /root/reference/src/calculate_ecs.cpp to compile UNMODIFIED with g++ (see oracle/Makefile,
// target _ref/libref_ecs.so).  Surface used by that file (calculate_ecs.cpp:7-70):
//   IntegerVector / NumericVector : (n) ctor zero-filled, operator()(i), size()
//   IntegerMatrix / NumericMatrix : (nrow, ncol) ctor zero-filled, column-major operator()(i,j),
//                                    nrow(), ncol()
//   sugar                         : min(v), max(v), sum(m), rowSums(m), colSums(m)
// Semantics follow R/Rcpp: matrices are column-major, vectors are zero-initialised.
#pragma once
#include <algorithm>
#include <cstddef>
#include <memory>
#include <vector>

namespace Rcpp {

// Like an Rcpp vector, a VectorStandin is a cheap handle: copying it aliases the same storage
// (Rcpp vectors are proxies onto one SEXP), so pass-by-value in the reference costs no O(n) copy.
template <typename T>
class VectorStandin {
public:
    VectorStandin() : p_(0), n_(0) {}
    explicit VectorStandin(int n) : own_(new std::vector<T>(static_cast<size_t>(n), T(0))), p_(own_->data()), n_(n) {}
    // non-owning view onto caller memory (what Rcpp does with an INTSXP/REALSXP it is handed)
    VectorStandin(const T* first, const T* last) : p_(const_cast<T*>(first)), n_(static_cast<int>(last - first)) {}
    T& operator()(int i) { return p_[i]; }
    const T& operator()(int i) const { return p_[i]; }
    T& operator[](int i) { return p_[i]; }
    const T& operator[](int i) const { return p_[i]; }
    int size() const { return n_; }
    const T* begin() const { return p_; }
    const T* end() const { return p_ + n_; }
    T* begin() { return p_; }
private:
    std::shared_ptr<std::vector<T> > own_;
    T* p_;
    int n_;
};

template <typename T>
class MatrixStandin {
public:
    MatrixStandin(int nr, int nc) : nr_(nr), nc_(nc), d_(static_cast<size_t>(nr) * nc, T(0)) {}
    T& operator()(int i, int j) { return d_[static_cast<size_t>(j) * nr_ + i]; }
    const T& operator()(int i, int j) const { return d_[static_cast<size_t>(j) * nr_ + i]; }
    int nrow() const { return nr_; }
    int ncol() const { return nc_; }
    const T* begin() const { return d_.data(); }
    size_t length() const { return d_.size(); }
private:
    int nr_, nc_;
    std::vector<T> d_;
};

typedef VectorStandin<int> IntegerVector;
typedef VectorStandin<double> NumericVector;
typedef MatrixStandin<int> IntegerMatrix;
typedef MatrixStandin<double> NumericMatrix;

template <typename T> inline T max(const VectorStandin<T>& v) { return *std::max_element(v.begin(), v.end()); }
template <typename T> inline T min(const VectorStandin<T>& v) { return *std::min_element(v.begin(), v.end()); }

template <typename T> inline T sum(const MatrixStandin<T>& m) {
    T s = T(0);
    for (size_t i = 0; i < m.length(); i++) s += m.begin()[i];
    return s;
}
template <typename T> inline VectorStandin<T> rowSums(const MatrixStandin<T>& m) {
    VectorStandin<T> r(m.nrow());
    for (int j = 0; j < m.ncol(); j++)
        for (int i = 0; i < m.nrow(); i++) r(i) += m(i, j);
    return r;
}
template <typename T> inline VectorStandin<T> colSums(const MatrixStandin<T>& m) {
    VectorStandin<T> r(m.ncol());
    for (int j = 0; j < m.ncol(); j++)
        for (int i = 0; i < m.nrow(); i++) r(j) += m(i, j);
    return r;
}

}  // namespace Rcpp
