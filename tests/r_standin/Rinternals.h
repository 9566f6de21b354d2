/* Stand-in for R's <Rinternals.h> -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * The build container has no R.  This header plus mini_r.c is a miniature of the part of R's C API that
 * r/clustassess_shim.c uses, faithful where the shim depends on it: SEXP types and lengths, INTEGER()/REAL()
 * payloads, Rf_coerceVector (REALSXP -> INTSXP truncates toward zero, like R), the PROTECT stack (its depth is
 * checked by the tests: it must be back to zero when a .Call entry returns) and Rf_error as a longjmp out of
 * the entry point (which is why the shim may only raise it after the library call has returned).
 * It lets tests/test_r_shim.py drive the shim's .Call entry points the way R would. */
#ifndef MINI_RINTERNALS_H
#define MINI_RINTERNALS_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef struct SEXPREC *SEXP;
typedef unsigned int SEXPTYPE;

#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define NA_INTEGER (-2147483647 - 1)
#define NA_LOGICAL NA_INTEGER
extern double R_NaReal;
#define NA_REAL R_NaReal
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;

int TYPEOF(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int Rf_length(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
double *REAL(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_mkChar(const char *s);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_ScalarReal(double v);
SEXP Rf_coerceVector(SEXP x, SEXPTYPE type);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
int Rf_isNull(SEXP x);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value);
SEXP Rf_getAttrib(SEXP x, SEXP name);
/* external pointers (the resident label sets): address, tag / protected slots are not modelled; the finalizer runs when the
 * harness "collects" the object (mr_reset or mr_finalize), once, and not at all after R_ClearExternalPtr */
typedef void (*R_CFinalizer_t)(SEXP);
typedef int Rboolean;
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
char *R_alloc(size_t n, int size);
#if defined(__GNUC__)
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char *fmt, ...);
#endif

#ifdef __cplusplus
}
#endif
#endif
