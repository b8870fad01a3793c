/*
 * .Call shim between the R package BGGE and libbgge_b200.so (include/bgge_b200.h).
 * R is not installed in this repository's build image, so the shim is never linked there; it is type-checked against
 * a mock of the R API (tests/mock_r, tests/test_cpu_host.py::test_r_shim_type_checks_against_the_header) so that it
 * cannot drift from the header.  No arithmetic here: argument unpacking, one library call, result packing.
 *
 * Build inside the R package:  src/Makevars:  PKG_CPPFLAGS = -I$(BGGE_B200_HOME)/include
 *                                             PKG_LIBS     = -L$(BGGE_B200_HOME)/bgge_b200 -lbgge_b200 -Wl,-rpath,$(BGGE_B200_HOME)/bgge_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>
#include "bgge_b200.h"

static void check(int status) {
  if (status != BGGE_OK) Rf_error("libbgge_b200: %s", bgge_last_error());   /* no device resources are held here */
}

/* ---- getK ----------------------------------------------------------------------------------------------------------- */

/* Base kernels over the L subjects (R/getK.R:131-149).  X = X[subjects, ] (L x p double).  Returns a list of nb
 * L x L matrices; attribute-free, the R code names them. */
SEXP bgge_R_getk_base(SEXP X, SEXP kind, SEXP bandwidth, SEXP quantil) {
  const int64_t L = Rf_nrows(X), p = Rf_ncols(X);
  const int nb = Rf_asInteger(kind) == BGGE_KERNEL_GB ? 1 : (int)XLENGTH(bandwidth);
  double* buf = (double*)R_alloc((size_t)nb * (size_t)L * (size_t)L, sizeof(double));
  check(bgge_getk_base(REAL(X), L, p, Rf_asInteger(kind), REAL(bandwidth), nb, Rf_asReal(quantil), buf, NULL));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, nb));
  for (int b = 0; b < nb; ++b) {
    SEXP m = PROTECT(Rf_allocMatrix(REALSXP, (int)L, (int)L));
    memcpy(REAL(m), buf + (size_t)b * L * L, (size_t)L * L * sizeof(double));
    SET_VECTOR_ELT(out, b, m);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}

/* One expansion Zg K0 Zg' (o ZeZe' / o ZEE_k Ze') of an L x L base kernel to n x n (R/getK.R:138,148,173,199-201,
 * 207-218,232): used for setKernel matrices and whenever the dense matrices are wanted.  K0 may be NULL for Gi. */
SEXP bgge_R_expand(SEXP K0, SEXP L_, SEXP gid, SEXP env, SEXP mode, SEXP env_k) {
  const int64_t n = XLENGTH(gid), L = (int64_t)Rf_asInteger(L_);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  check(bgge_getk_expand(Rf_isNull(K0) ? NULL : REAL(K0), L, INTEGER(gid), INTEGER(env), n, Rf_asInteger(mode),
                         Rf_asInteger(env_k), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* Does `Kernel` (n x n) still equal the expansion its factor describes?  Exact comparison of every entry (a user may
 * have edited the matrix after getK(); the reference always decomposes the matrix it is given, R/BGGE.R:144,160);
 * the library runs the pass on all host cores. */
SEXP bgge_R_factor_matches(SEXP Kernel, SEXP K0, SEXP L_, SEXP gid, SEXP env, SEXP mode, SEXP env_k) {
  const int64_t n = XLENGTH(gid), L = (int64_t)Rf_asInteger(L_);
  const int md = Rf_asInteger(mode), ek = Rf_asInteger(env_k);
  if (!Rf_isReal(Kernel) || Rf_nrows(Kernel) != n || Rf_ncols(Kernel) != n) return Rf_ScalarLogical(0);
  int ok = 0;
  check(bgge_factor_matches(REAL(Kernel), n, n, Rf_isNull(K0) ? NULL : REAL(K0), L, INTEGER(gid), INTEGER(env), md, ek, &ok));
  return Rf_ScalarLogical(ok);
}

/* ---- BGGE ----------------------------------------------------------------------------------------------------------- */

struct progress_state {
  int verbose;
};
static void poll_interrupt(void* unused) {
  (void)unused;
  R_CheckUserInterrupt();
}
/* Called by the library between pieces of `verbose` iterations, on the R main thread (R/BGGE.R:398-401).  An
 * interrupt must not longjmp through the library (its device state would leak): R_ToplevelExec catches it, the
 * callback asks the library to stop, and the error is raised after the library call has returned. */
static int on_progress(int64_t done, double seconds_per_iteration, void* user) {
  const struct progress_state* ps = (const struct progress_state*)user;
  if (ps->verbose) Rprintf("Iter:  %lld time:  %.3f \n", (long long)done, seconds_per_iteration);
  return R_ToplevelExec(poll_interrupt, NULL) ? 0 : 1;
}

static SEXP attr_get(SEXP x, const char* name) { return Rf_getAttrib(x, Rf_install(name)); }

/*
 * BGGE() numeric body (R/BGGE.R:189-433).
 *   kernels  list, one entry per K[[j]]: either a plain n x n double matrix (dense route) or an L x L base kernel
 *            carrying the attributes "mode" (bgge_expand_mode), "env_k", "L" (factored route; NULL K0 with mode GI is
 *            passed as a length-0 double with the attributes)
 *   types    integer codes (0 = "D", 1 = "BD")
 *   gid/env  0-based codes of Y[,2] / Y[,1] (may be length 0 when every kernel is dense)
 *   verbose  0 or the print interval;  tapes: NULL or double vectors (injected-draw mode)
 */
SEXP bgge_R_fit(SEXP y, SEXP kernels, SEXP types, SEXP gid, SEXP env, SEXP XF, SEXP ne, SEXP ite, SEXP burn, SEXP thin,
                SEXP verbose, SEXP tol, SEXP R2, SEXP seed, SEXP normals, SEXP chisq) {
  const int64_t n = XLENGTH(y);
  const int nk = (int)XLENGTH(kernels);
  const int q = Rf_isNull(XF) ? 0 : Rf_ncols(XF);
  const int64_t it = (int64_t)Rf_asReal(ite), th = (int64_t)Rf_asReal(thin), ncum = it / th;
  bgge_kernel_desc* kd = (bgge_kernel_desc*)R_alloc((size_t)nk, sizeof(bgge_kernel_desc));
  for (int j = 0; j < nk; ++j) {
    SEXP kj = VECTOR_ELT(kernels, j);
    SEXP md = attr_get(kj, "mode");
    memset(&kd[j], 0, sizeof(kd[j]));
    kd[j].type = INTEGER(types)[j];
    if (Rf_isNull(md)) {
      kd[j].mode = BGGE_KERNEL_DENSE;
      kd[j].dense = REAL(kj);
      if (Rf_nrows(kj) != n || Rf_ncols(kj) != n) Rf_error("non-conformable arguments");
    } else {
      kd[j].mode = Rf_asInteger(md);
      kd[j].env_k = Rf_asInteger(attr_get(kj, "env_k"));
      kd[j].L = (int64_t)Rf_asInteger(attr_get(kj, "L"));
      kd[j].K0 = XLENGTH(kj) > 0 ? REAL(kj) : NULL;
      if (XLENGTH(gid) != n) Rf_error("non-conformable arguments");
    }
  }
  uint8_t* na = (uint8_t*)R_alloc((size_t)n, 1);
  for (int64_t i = 0; i < n; ++i) na[i] = ISNAN(REAL(y)[i]) ? 1 : 0;        /* NA_real_ is a NaN payload */
  const R_xlen_t nne = Rf_isNull(ne) ? 0 : XLENGTH(ne);
  int64_t* ne64 = (int64_t*)R_alloc(nne > 0 ? (size_t)nne : 1, sizeof(int64_t));
  for (R_xlen_t k = 0; k < nne; ++k) ne64[k] = (int64_t)REAL(ne)[k];
  const int tape = !Rf_isNull(normals) && !Rf_isNull(chisq);
  const char* nm[] = {"yHat", "varE", "varE.sd", "u", "u.sd", "varu", "varu.sd", "y", "chain.mu", "chain.varE",
                      "chain.varU", "chain.Bet", "factored", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SEXP yHat = Rf_allocVector(REALSXP, n);              SET_VECTOR_ELT(out, 0, yHat);
  SEXP varE = Rf_allocVector(REALSXP, 1);              SET_VECTOR_ELT(out, 1, varE);
  SEXP varEsd = Rf_allocVector(REALSXP, 1);            SET_VECTOR_ELT(out, 2, varEsd);
  SEXP u = Rf_allocMatrix(REALSXP, (int)n, nk);        SET_VECTOR_ELT(out, 3, u);
  SEXP usd = Rf_allocMatrix(REALSXP, (int)n, nk);      SET_VECTOR_ELT(out, 4, usd);
  SEXP varu = Rf_allocVector(REALSXP, nk);             SET_VECTOR_ELT(out, 5, varu);
  SEXP varusd = Rf_allocVector(REALSXP, nk);           SET_VECTOR_ELT(out, 6, varusd);
  SEXP yo = Rf_allocVector(REALSXP, n);                SET_VECTOR_ELT(out, 7, yo);
  SEXP cmu = Rf_allocVector(REALSXP, ncum);            SET_VECTOR_ELT(out, 8, cmu);
  SEXP cve = Rf_allocVector(REALSXP, ncum);            SET_VECTOR_ELT(out, 9, cve);
  SEXP cvu = Rf_allocMatrix(REALSXP, (int)ncum, nk);   SET_VECTOR_ELT(out, 10, cvu);
  SEXP cbt = Rf_allocMatrix(REALSXP, q > 0 ? q : 1, (int)ncum); SET_VECTOR_ELT(out, 11, cbt);   /* q x n_cum: the library writes n_cum rows of q */
  int nfact = 0;
  struct progress_state ps;
  ps.verbose = Rf_asInteger(verbose) > 0;
  /* without `verbose` the run is still cut into pieces so that Ctrl-C is honoured within a second or so */
  const int64_t chunk = Rf_asInteger(verbose) > 0 ? (int64_t)Rf_asInteger(verbose) : (it > 200 ? 200 : it);
  const int st = bgge_fit_kernels(
      REAL(y), na, n, kd, nk, XLENGTH(gid) > 0 ? INTEGER(gid) : NULL, XLENGTH(env) > 0 ? INTEGER(env) : NULL,
      q ? REAL(XF) : NULL, q, nne > 0 ? ne64 : NULL, (int)nne, it, (int64_t)Rf_asReal(burn), th, Rf_asReal(tol),
      Rf_asReal(R2), tape ? BGGE_RNG_TAPE : BGGE_RNG_PHILOX, (uint64_t)Rf_asReal(seed), tape ? REAL(normals) : NULL,
      tape ? (int64_t)XLENGTH(normals) : 0, tape ? REAL(chisq) : NULL, tape ? (int64_t)XLENGTH(chisq) : 0, chunk,
      on_progress, &ps, REAL(yHat), REAL(varE), REAL(varEsd), REAL(u), REAL(usd), REAL(varu), REAL(varusd), REAL(yo),
      REAL(cmu), REAL(cve), REAL(cvu), REAL(cbt), &nfact);
  if (st == BGGE_ERR_STATE) {          /* our callback asked to stop: the library has released everything */
    UNPROTECT(1);
    R_CheckUserInterrupt();            /* an interrupt that arrived since then surfaces as one */
    Rf_error("libbgge_b200: %s", bgge_last_error());   /* "interrupted after i of ite iterations" */
  }
  check(st);
  SET_VECTOR_ELT(out, 12, Rf_ScalarReal((double)nfact));
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef call_methods[] = {
  {"bgge_R_getk_base", (DL_FUNC)&bgge_R_getk_base, 4},
  {"bgge_R_expand", (DL_FUNC)&bgge_R_expand, 6},
  {"bgge_R_factor_matches", (DL_FUNC)&bgge_R_factor_matches, 7},
  {"bgge_R_fit", (DL_FUNC)&bgge_R_fit, 16},
  {NULL, NULL, 0}
};

void R_init_BGGE(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
