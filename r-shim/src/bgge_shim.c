/*
 * .Call shim between the R package BGGE and libbgge_b200.so (include/bgge_b200.h).
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no R (no Rinternals.h, libR.so).
 * It contains no arithmetic: argument unpacking, one library call, result packing.
 *
 * Build inside the R package:  src/Makevars:  PKG_CPPFLAGS = -I$(BGGE_B200_HOME)/include
 *                                             PKG_LIBS     = -L$(BGGE_B200_HOME)/bgge_b200 -lbgge_b200 -Wl,-rpath,$(BGGE_B200_HOME)/bgge_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "bgge_b200.h"

static void check(int status) {
  if (status != BGGE_OK) Rf_error("libbgge_b200: %s", bgge_last_error());   /* no device resources are held here */
}

/* getK numeric body (R/getK.R:131-149, 138/148/173, 199-201, 207-218, 232).
 * X: L x p double matrix (already X[subjects,]); gid/env: 0-based integer codes; returns a list of
 * n x n matrices in getK()'s order; the caller attaches names and Type. */
SEXP bgge_R_getk(SEXP X, SEXP kind, SEXP bandwidth, SEXP quantil, SEXP gid, SEXP env, SEXP n_env, SEXP model,
                 SEXP intercept) {
  const int64_t L = Rf_nrows(X), p = Rf_ncols(X), n = XLENGTH(gid);
  const int nb = Rf_asInteger(kind) == BGGE_KERNEL_GB ? 1 : (int)XLENGTH(bandwidth);
  const int E = Rf_asInteger(n_env), mdl = Rf_asInteger(model), gi = Rf_asLogical(intercept);
  const int n_out = nb * (mdl == 0 ? 1 : mdl == 1 ? 2 : 1 + E) + (gi ? 1 : 0);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n_out));
  double** ptr = (double**)R_alloc(n_out, sizeof(double*));
  for (int k = 0; k < n_out; ++k) {
    SEXP m = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
    SET_VECTOR_ELT(out, k, m);
    UNPROTECT(1);
    ptr[k] = REAL(m);
  }
  check(bgge_getk(REAL(X), L, p, Rf_asInteger(kind), REAL(bandwidth), nb, Rf_asReal(quantil), INTEGER(gid),
                  INTEGER(env), n, E, mdl, gi, ptr, n_out, NULL));
  UNPROTECT(1);
  return out;
}

/* BGGE() numeric body (R/BGGE.R:189-433): K = list of n x n matrices, types = integer codes (0 D, 1 BD). */
SEXP bgge_R_fit(SEXP y, SEXP K, SEXP types, SEXP XF, SEXP ne, SEXP ite, SEXP burn, SEXP thin, SEXP tol, SEXP R2,
                SEXP seed) {
  const int64_t n = XLENGTH(y);
  const int nk = (int)XLENGTH(K);
  const int q = Rf_isNull(XF) ? 0 : Rf_ncols(XF);
  const int64_t it = (int64_t)Rf_asReal(ite), th = (int64_t)Rf_asReal(thin), ncum = it / th;
  const double** kp = (const double**)R_alloc(nk, sizeof(double*));
  for (int j = 0; j < nk; ++j) kp[j] = REAL(VECTOR_ELT(K, j));
  uint8_t* na = (uint8_t*)R_alloc(n, 1);
  for (int64_t i = 0; i < n; ++i) na[i] = ISNAN(REAL(y)[i]) ? 1 : 0;        /* NA_real_ is a NaN payload */
  int64_t* ne64 = (int64_t*)R_alloc(XLENGTH(ne) > 0 ? XLENGTH(ne) : 1, sizeof(int64_t));
  for (R_xlen_t k = 0; k < XLENGTH(ne); ++k) ne64[k] = (int64_t)REAL(ne)[k];
  const char* nm[] = {"yHat", "varE", "varE.sd", "u", "u.sd", "varu", "varu.sd", "y", "chain.mu", "chain.varE",
                      "chain.varU", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SEXP yHat = Rf_allocVector(REALSXP, n);           SET_VECTOR_ELT(out, 0, yHat);
  SEXP varE = Rf_allocVector(REALSXP, 1);           SET_VECTOR_ELT(out, 1, varE);
  SEXP varEsd = Rf_allocVector(REALSXP, 1);         SET_VECTOR_ELT(out, 2, varEsd);
  SEXP u = Rf_allocMatrix(REALSXP, (int)n, nk);     SET_VECTOR_ELT(out, 3, u);
  SEXP usd = Rf_allocMatrix(REALSXP, (int)n, nk);   SET_VECTOR_ELT(out, 4, usd);
  SEXP varu = Rf_allocVector(REALSXP, nk);          SET_VECTOR_ELT(out, 5, varu);
  SEXP varusd = Rf_allocVector(REALSXP, nk);        SET_VECTOR_ELT(out, 6, varusd);
  SEXP yo = Rf_allocVector(REALSXP, n);             SET_VECTOR_ELT(out, 7, yo);
  SEXP cmu = Rf_allocVector(REALSXP, ncum);         SET_VECTOR_ELT(out, 8, cmu);
  SEXP cve = Rf_allocVector(REALSXP, ncum);         SET_VECTOR_ELT(out, 9, cve);
  SEXP cvu = Rf_allocMatrix(REALSXP, (int)ncum, nk); SET_VECTOR_ELT(out, 10, cvu);
  check(bgge_fit(REAL(y), na, n, kp, INTEGER(types), nk, q ? REAL(XF) : NULL, q, ne64, (int)XLENGTH(ne), it,
                 (int64_t)Rf_asReal(burn), th, Rf_asReal(tol), Rf_asReal(R2), BGGE_RNG_PHILOX,
                 (uint64_t)Rf_asReal(seed), NULL, 0, NULL, 0, REAL(yHat), REAL(varE), REAL(varEsd), REAL(u), REAL(usd),
                 REAL(varu), REAL(varusd), REAL(yo), REAL(cmu), REAL(cve), REAL(cvu)));
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef call_methods[] = {
  {"bgge_R_getk", (DL_FUNC)&bgge_R_getk, 9},
  {"bgge_R_fit", (DL_FUNC)&bgge_R_fit, 11},
  {NULL, NULL, 0}
};

void R_init_BGGE(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
