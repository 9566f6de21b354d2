/* mini_r.c -- a miniature of R's C runtime for driving r/clustassess_shim.c in the tests.
 * TEST INFRASTRUCTURE ONLY (never shipped, never on the product path); see Rinternals.h in this directory.
 *
 * What it models: vectors with type/length/payload and the names, dim and class attributes; an allocation arena
 * (everything is freed by mr_reset()); the PROTECT stack depth; Rf_error as a longjmp back into mr_call(), like
 * R's own error handling unwinds a .Call.  The mr_* functions are the harness the Python tests bind with ctypes. */
#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"

struct SEXPREC {
    int type;
    R_xlen_t len;
    void *data; /* int* (INTSXP, LGLSXP), double* (REALSXP), SEXP* (VECSXP, STRSXP), char* (CHARSXP) */
    SEXP names, dim, klass;
    struct SEXPREC *next;
    R_CFinalizer_t fin; /* EXTPTRSXP: C finalizer, run once when the object is collected */
};

static struct SEXPREC nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC names_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC dim_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
double R_NaReal;

static SEXP arena = NULL;
struct blk {
    struct blk *next;
};
static struct blk *ralloc_list = NULL;
static int prot_depth = 0, prot_underflow = 0, in_call = 0;
static jmp_buf call_jb;
static char err_msg[1024];

static void init_once(void) {
    static int done = 0;
    if (done) return;
    done = 1;
    union {
        double d;
        uint64_t u;
    } na;
    na.u = 0x7FF00000000007A2ull; /* R's NA_real_: a NaN whose low word is 1954 */
    R_NaReal = na.d;
}

static SEXP new_obj(int type, R_xlen_t n) {
    init_once();
    SEXP x = (SEXP)calloc(1, sizeof(struct SEXPREC));
    size_t elt = type == REALSXP ? sizeof(double) : (type == VECSXP || type == STRSXP || type == EXTPTRSXP) ? sizeof(SEXP) : type == CHARSXP ? 1 : sizeof(int);
    x->type = type;
    x->len = n;
    x->data = calloc((size_t)(n > 0 ? n : 1) + (type == CHARSXP), elt);
    x->names = x->dim = x->klass = R_NilValue;
    if (type == VECSXP || type == STRSXP)
        for (R_xlen_t i = 0; i < n; i++) ((SEXP *)x->data)[i] = R_NilValue;
    x->next = arena;
    arena = x;
    return x;
}

/* ---- the API surface the shim uses ---------------------------------------------------------------------- */
int TYPEOF(SEXP x) { return x->type; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int Rf_length(SEXP x) { return (int)x->len; }
int *INTEGER(SEXP x) {
    if (x->type != INTSXP && x->type != LGLSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type %d", x->type);
    return (int *)x->data;
}
int *LOGICAL(SEXP x) { return INTEGER(x); }
double *REAL(SEXP x) {
    if (x->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type %d", x->type);
    return (double *)x->data;
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) {
    if (x->type != VECSXP || i < 0 || i >= x->len) Rf_error("VECTOR_ELT: not a list or subscript out of bounds");
    return ((SEXP *)x->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    if (x->type != VECSXP || i < 0 || i >= x->len) Rf_error("SET_VECTOR_ELT: not a list or subscript out of bounds");
    ((SEXP *)x->data)[i] = v;
    return v;
}
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) {
    if (x->type != STRSXP || v->type != CHARSXP || i < 0 || i >= x->len) Rf_error("SET_STRING_ELT: bad arguments");
    ((SEXP *)x->data)[i] = v;
}
SEXP Rf_mkChar(const char *s) {
    SEXP x = new_obj(CHARSXP, (R_xlen_t)strlen(s));
    memcpy(x->data, s, strlen(s) + 1);
    return x;
}
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) {
    if (n < 0) Rf_error("negative length vectors are not allowed");
    return new_obj((int)type, n);
}
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol) {
    if (nrow < 0 || ncol < 0) Rf_error("negative extents to matrix");
    SEXP x = new_obj((int)type, (R_xlen_t)nrow * ncol);
    x->dim = new_obj(INTSXP, 2);
    ((int *)x->dim->data)[0] = nrow;
    ((int *)x->dim->data)[1] = ncol;
    return x;
}
SEXP Rf_ScalarReal(double v) {
    SEXP x = new_obj(REALSXP, 1);
    ((double *)x->data)[0] = v;
    return x;
}
SEXP Rf_coerceVector(SEXP x, SEXPTYPE type) {
    if ((SEXPTYPE)x->type == type) return x;
    if (type == INTSXP && x->type == REALSXP) { /* R truncates toward zero; NaN and out-of-range become NA */
        SEXP r = new_obj(INTSXP, x->len);
        for (R_xlen_t i = 0; i < x->len; i++) {
            double d = ((double *)x->data)[i];
            ((int *)r->data)[i] = (isnan(d) || d >= 2147483648.0 || d <= -2147483649.0) ? NA_INTEGER : (int)d;
        }
        return r;
    }
    if (type == INTSXP && x->type == LGLSXP) {
        SEXP r = new_obj(INTSXP, x->len);
        memcpy(r->data, x->data, (size_t)x->len * sizeof(int));
        return r;
    }
    if (type == REALSXP && (x->type == INTSXP || x->type == LGLSXP)) {
        SEXP r = new_obj(REALSXP, x->len);
        for (R_xlen_t i = 0; i < x->len; i++) {
            int v = ((int *)x->data)[i];
            ((double *)r->data)[i] = v == NA_INTEGER ? R_NaReal : (double)v;
        }
        return r;
    }
    Rf_error("cannot coerce type %d to vector of type %u", x->type, type);
}
int Rf_asInteger(SEXP x) {
    if (x->len < 1) return NA_INTEGER;
    if (x->type == INTSXP || x->type == LGLSXP) return ((int *)x->data)[0];
    if (x->type == REALSXP) return isnan(((double *)x->data)[0]) ? NA_INTEGER : (int)((double *)x->data)[0];
    return NA_INTEGER;
}
int Rf_asLogical(SEXP x) {
    if (x->len < 1) return NA_LOGICAL;
    if (x->type == LGLSXP || x->type == INTSXP) {
        int v = ((int *)x->data)[0];
        return v == NA_INTEGER ? NA_LOGICAL : v != 0;
    }
    if (x->type == REALSXP) return isnan(((double *)x->data)[0]) ? NA_LOGICAL : ((double *)x->data)[0] != 0.0;
    return NA_LOGICAL;
}
int Rf_isNull(SEXP x) { return x->type == NILSXP; }
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) {
    (void)tag;
    (void)prot;
    SEXP x = new_obj(EXTPTRSXP, 1);
    ((void **)x->data)[0] = p;
    return x;
}
void *R_ExternalPtrAddr(SEXP s) {
    if (s->type != EXTPTRSXP) Rf_error("R_ExternalPtrAddr: argument of type %d is not an external pointer", s->type);
    return ((void **)s->data)[0];
}
void R_ClearExternalPtr(SEXP s) {
    if (s->type == EXTPTRSXP) ((void **)s->data)[0] = NULL;
}
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) {
    (void)onexit;
    s->fin = fun;
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value) {
    if (name == R_NamesSymbol) x->names = value;
    else if (name == R_DimSymbol) x->dim = value;
    else Rf_error("mini R models only the names and dim attributes");
    return value;
}
SEXP Rf_getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : name == R_DimSymbol ? x->dim : R_NilValue; }
SEXP Rf_protect(SEXP x) {
    prot_depth++;
    return x;
}
void Rf_unprotect(int n) {
    prot_depth -= n;
    if (prot_depth < 0) prot_underflow = 1; /* R: "unprotect(): only 0 protected items" */
}
char *R_alloc(size_t n, int size) {
    struct blk *b = (struct blk *)calloc(1, sizeof(struct blk) + n * (size_t)size + 16);
    b->next = ralloc_list;
    ralloc_list = b;
    return (char *)(b + 1);
}
void Rf_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof(err_msg), fmt, ap);
    va_end(ap);
    if (!in_call) {
        fprintf(stderr, "mini R: Rf_error outside mr_call: %s\n", err_msg);
        abort();
    }
    longjmp(call_jb, 1);
}

/* ---- harness (bound from Python with ctypes) ------------------------------------------------------------------ */
/* "garbage collection" of one external pointer: runs its finalizer (once) */
static int finalizers_run = 0;
void mr_finalize(SEXP x) {
    if (x->type == EXTPTRSXP && x->fin) {
        R_CFinalizer_t f = x->fin;
        x->fin = NULL;
        in_call = 1; /* a finalizer may not raise, but keep Rf_error from aborting the test process */
        if (setjmp(call_jb) == 0) f(x);
        in_call = 0;
        finalizers_run++;
    }
}
int mr_finalizers_run(void) { return finalizers_run; }
void mr_reset(void) {
    for (SEXP x = arena; x; x = x->next) mr_finalize(x);
    while (arena) {
        SEXP n = arena->next;
        free(arena->data);
        free(arena);
        arena = n;
    }
    while (ralloc_list) {
        struct blk *n = ralloc_list->next;
        free(ralloc_list);
        ralloc_list = n;
    }
    prot_depth = prot_underflow = 0;
    err_msg[0] = 0;
}
SEXP mr_null(void) { return R_NilValue; }
SEXP mr_int(const int *v, R_xlen_t n) {
    SEXP x = new_obj(INTSXP, n);
    if (n > 0) memcpy(x->data, v, (size_t)n * sizeof(int));
    return x;
}
SEXP mr_factor(const int *codes, R_xlen_t n) { /* a factor is an INTSXP with a class attribute */
    SEXP x = mr_int(codes, n);
    x->klass = Rf_mkChar("factor");
    return x;
}
SEXP mr_real(const double *v, R_xlen_t n) {
    SEXP x = new_obj(REALSXP, n);
    if (n > 0) memcpy(x->data, v, (size_t)n * sizeof(double));
    return x;
}
SEXP mr_lgl(int v) {
    SEXP x = new_obj(LGLSXP, 1);
    ((int *)x->data)[0] = v;
    return x;
}
SEXP mr_list(R_xlen_t n) { return new_obj(VECSXP, n); }
void mr_list_set(SEXP l, R_xlen_t i, SEXP v) { ((SEXP *)l->data)[i] = v; }

typedef SEXP (*fn1)(SEXP);
typedef SEXP (*fn2)(SEXP, SEXP);
typedef SEXP (*fn3)(SEXP, SEXP, SEXP);
typedef SEXP (*fn4)(SEXP, SEXP, SEXP, SEXP);
/* .Call(fn, ...): NULL when the entry raised an R error (message in mr_error()) */
SEXP mr_call(void *fn, int nargs, SEXP a0, SEXP a1, SEXP a2, SEXP a3) {
    SEXP volatile res = NULL;
    err_msg[0] = 0;
    prot_depth = prot_underflow = 0;
    in_call = 1;
    if (setjmp(call_jb) == 0) {
        switch (nargs) {
            case 1: res = ((fn1)fn)(a0); break;
            case 2: res = ((fn2)fn)(a0, a1); break;
            case 3: res = ((fn3)fn)(a0, a1, a2); break;
            case 4: res = ((fn4)fn)(a0, a1, a2, a3); break;
            default: snprintf(err_msg, sizeof(err_msg), "mr_call: %d arguments not supported", nargs);
        }
    } else {
        res = NULL;
        prot_depth = 0; /* R unwinds the protect stack on error */
    }
    in_call = 0;
    return res;
}
const char *mr_error(void) { return err_msg; }
int mr_protect_depth(void) { return prot_depth; }
int mr_protect_underflow(void) { return prot_underflow; }
int mr_type(SEXP x) { return x->type; }
R_xlen_t mr_len(SEXP x) { return x->len; }
void *mr_data(SEXP x) { return x->data; }
SEXP mr_elt(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
const char *mr_name(SEXP x, R_xlen_t i) {
    if (x->names == R_NilValue || i >= x->names->len) return "";
    return (const char *)((SEXP *)x->names->data)[i]->data;
}
int mr_dim(SEXP x, int which) { return x->dim == R_NilValue ? -1 : ((int *)x->dim->data)[which]; }
