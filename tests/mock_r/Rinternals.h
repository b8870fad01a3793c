/* MOCK of the few declarations of R's C API that r-shim/src/bgge_shim.c uses.  TEST INFRASTRUCTURE ONLY: R is not
 * installed in the build image, so tests/test_cpu_host.py type-checks the shim against these prototypes (gcc
 * -fsyntax-only) to catch arity / type drift between the shim and include/bgge_b200.h.  Signatures follow
 * R 4.x's Rinternals.h. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define FALSE 0
#define TRUE 1
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue;
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
double Rf_asReal(SEXP x);
Rboolean Rf_isNull(SEXP x);
Rboolean Rf_isReal(SEXP x);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_mkNamed(SEXPTYPE type, const char** names);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_install(const char* name);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_ScalarReal(double v);
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
char* R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
void Rprintf(const char* fmt, ...);
int R_IsNA(double x);
int R_IsNaN(double x);
#define ISNA(x) R_IsNA(x)
#define ISNAN(x) ((x) != (x))
#ifdef __cplusplus
}
#endif
#endif
