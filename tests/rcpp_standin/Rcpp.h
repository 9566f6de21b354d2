// Stand-in for <Rcpp.h> -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// The build container has neither R nor Rcpp.  This header is the slice of Rcpp's surface that the glue
// r/calculate_ecs_b200.cpp (route A of INTEGRATION.md) uses, so that the tests can compile the glue UNMODIFIED and
// call its exported functions: IntegerVector / NumericVector (handles that alias one storage, like Rcpp's proxies
// onto a SEXP), IntegerMatrix / NumericMatrix (column-major), List with Named() = value, Nullable<T>, as<T>() with
// Rcpp's coercions (REALSXP -> integer truncates), XPtr with a finalizer, and stop() as a C++ exception
// (Rcpp::exception, which END_RCPP turns into an R condition, src/RcppExports.cpp:26).
// (The oracle has its own, smaller stand-in -- oracle/rcpp_standin/ -- for compiling the reference's
// src/calculate_ecs.cpp; the two are kept apart so that the oracle's pin does not move with the glue.)
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;
struct RObj;
typedef RObj* SEXP;
#define NA_REAL (std::nan("1954"))

namespace Rcpp {

class exception : public std::runtime_error {
  public:
    explicit exception(const std::string& m) : std::runtime_error(m) {}
};
template <typename... A> [[noreturn]] inline void stop(const char* fmt, A... a) {
    char buf[1024];
    std::snprintf(buf, sizeof(buf), fmt, a...);
    throw exception(buf);
}
[[noreturn]] inline void stop(const char* msg) { throw exception(msg); }

template <typename T> class Vector {
  public:
    Vector() : d_(std::make_shared<std::vector<T> >()) {}
    explicit Vector(R_xlen_t n) : d_(std::make_shared<std::vector<T> >(static_cast<size_t>(n), T(0))) {}
    Vector(const T* first, const T* last) : d_(std::make_shared<std::vector<T> >(first, last)) {}
    R_xlen_t size() const { return static_cast<R_xlen_t>(d_->size()); }
    T* begin() { return d_->data(); }
    T* end() { return d_->data() + d_->size(); }
    const T* begin() const { return d_->data(); }
    const T* end() const { return d_->data() + d_->size(); }
    T& operator[](R_xlen_t i) { return (*d_)[static_cast<size_t>(i)]; }
    const T& operator[](R_xlen_t i) const { return (*d_)[static_cast<size_t>(i)]; }
    T& operator()(R_xlen_t i) { return (*this)[i]; }
    bool same_storage(const Vector& o) const { return d_ == o.d_; }

  private:
    std::shared_ptr<std::vector<T> > d_;
};
typedef Vector<int> IntegerVector;
typedef Vector<double> NumericVector;

template <typename T> class Matrix {
  public:
    Matrix(int nr, int nc) : nr_(nr), nc_(nc), d_(std::make_shared<std::vector<T> >(static_cast<size_t>(nr) * nc, T(0))) {}
    int nrow() const { return nr_; }
    int ncol() const { return nc_; }
    T* begin() { return d_->data(); }
    const T* begin() const { return d_->data(); }
    T& operator()(int i, int j) { return (*d_)[static_cast<size_t>(j) * nr_ + i]; }

  private:
    int nr_, nc_;
    std::shared_ptr<std::vector<T> > d_;
};
typedef Matrix<int> IntegerMatrix;
typedef Matrix<double> NumericMatrix;

}  // namespace Rcpp

// A value as R would hold it.  Objects live until the harness drops them (standin_arena().clear()).
struct RObj {
    enum Kind { INT, REAL, INTMAT, REALMAT, LIST, XPTR } kind;
    Rcpp::IntegerVector iv;
    Rcpp::NumericVector dv;
    std::shared_ptr<Rcpp::IntegerMatrix> im;
    std::shared_ptr<Rcpp::NumericMatrix> dm;
    std::vector<SEXP> list;
    std::vector<std::string> names;
    void* xp = nullptr;
    void (*finalizer)(void*) = nullptr;
    explicit RObj(Kind k) : kind(k) {}
    ~RObj() {
        if (kind == XPTR && xp && finalizer) finalizer(xp);  // R runs the finalizer when the pointer is collected
    }
};
inline std::vector<std::unique_ptr<RObj> >& standin_arena() {
    static std::vector<std::unique_ptr<RObj> > a;
    return a;
}
inline SEXP standin_new(RObj::Kind k) {
    standin_arena().emplace_back(new RObj(k));
    return standin_arena().back().get();
}

namespace Rcpp {

inline SEXP wrap(const IntegerVector& v) { SEXP x = standin_new(RObj::INT); x->iv = v; return x; }
inline SEXP wrap(const NumericVector& v) { SEXP x = standin_new(RObj::REAL); x->dv = v; return x; }
inline SEXP wrap(const IntegerMatrix& m) { SEXP x = standin_new(RObj::INTMAT); x->im = std::make_shared<IntegerMatrix>(m); return x; }
inline SEXP wrap(const NumericMatrix& m) { SEXP x = standin_new(RObj::REALMAT); x->dm = std::make_shared<NumericMatrix>(m); return x; }

template <typename T> T as(SEXP x);
template <> inline IntegerVector as<IntegerVector>(SEXP x) {
    if (x && x->kind == RObj::INT) return x->iv;  // an INTSXP is used as it is: no copy
    if (x && x->kind == RObj::REAL) {             // REALSXP -> integer truncates toward zero (R's coercion)
        IntegerVector r(x->dv.size());
        for (R_xlen_t i = 0; i < r.size(); i++) r[i] = static_cast<int>(x->dv[i]);
        return r;
    }
    stop("Not compatible with requested type: [type=list; target=integer].");
}
template <> inline NumericVector as<NumericVector>(SEXP x) {
    if (x && x->kind == RObj::REAL) return x->dv;
    if (x && x->kind == RObj::INT) {
        NumericVector r(x->iv.size());
        for (R_xlen_t i = 0; i < r.size(); i++) r[i] = x->iv[i];
        return r;
    }
    stop("Not compatible with requested type: [type=list; target=double].");
}

struct NamedValue {
    std::string name;
    SEXP value;
};
class Named {
  public:
    explicit Named(const char* n) : n_(n) {}
    template <typename T> NamedValue operator=(const T& v) const { return NamedValue{n_, wrap(v)}; }

  private:
    std::string n_;
};

class List {
  public:
    List() {}
    List(SEXP x) {
        if (!x || x->kind != RObj::LIST) stop("Not compatible with requested type: [target=list].");
        v_ = x->list;
        names_ = x->names;
    }
    int size() const { return static_cast<int>(v_.size()); }
    SEXP operator[](int i) const { return v_[static_cast<size_t>(i)]; }
    void push_back(SEXP x) { v_.push_back(x); names_.push_back(""); }
    static List create(const NamedValue& a, const NamedValue& b) {
        List l;
        l.v_ = {a.value, b.value};
        l.names_ = {a.name, b.name};
        return l;
    }
    SEXP get(const std::string& name) const {
        for (size_t i = 0; i < v_.size(); i++)
            if (names_[i] == name) return v_[i];
        return nullptr;
    }
    const std::vector<std::string>& names() const { return names_; }

  private:
    std::vector<SEXP> v_;
    std::vector<std::string> names_;
};

template <typename T> class Nullable {
  public:
    Nullable() : x_(nullptr) {}
    Nullable(SEXP x) : x_(x) {}
    bool isNotNull() const { return x_ != nullptr; }
    bool isNull() const { return x_ == nullptr; }
    operator T() const { return as<T>(x_); }

  private:
    SEXP x_;
};

template <class> class PreserveStorage {};
template <typename T, template <class> class Storage, void Finalizer(T*), bool finalizeOnExit> class XPtr {
  public:
    XPtr(SEXP x) : x_(x) {
        if (!x || x->kind != RObj::XPTR) stop("Expecting an external pointer: [type=%s].", "other");
    }
    XPtr(T* p, bool /*set_delete_finalizer*/) : x_(standin_new(RObj::XPTR)) {
        x_->xp = p;
        x_->finalizer = [](void* q) { Finalizer(static_cast<T*>(q)); };
    }
    T* get() const { return static_cast<T*>(x_->xp); }
    operator SEXP() const { return x_; }

  private:
    SEXP x_;
};

}  // namespace Rcpp
