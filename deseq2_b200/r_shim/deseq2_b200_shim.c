/*
 * deseq2_b200_shim.c -- R .Call shim: re-creates DESeq2's three native entry points on top of libb200nb.so.
 *
 * It exports exactly the symbols DESeq2's R code resolves (src/RcppExports.cpp:16,41,64 and the registration
 * table :84-94): _DESeq2_fitDisp (15 SEXP), _DESeq2_fitBeta (13 SEXP), _DESeq2_fitDispGrid (11 SEXP) and
 * R_init_DESeq2.  Build it as DESeq2's package shared object in place of src/DESeq2.cpp + src/RcppExports.cpp:
 *
 *     R CMD SHLIB -o DESeq2.so deseq2_b200_shim.c -L<repo>/deseq2_b200 -lb200nb -Wl,-rpath,<repo>/deseq2_b200
 *
 * (needs R's headers; they are NOT available in the build image of this repository, so __graft_entry__.build()
 * does not compile this file; tests/test_r_shim.py compiles it unchanged against the mock R API in tests/mock_r/
 * and exercises every entry point).  The R code of the package is unchanged:
 * R/RcppExports.R:4-14 keeps calling .Call('_DESeq2_fitDisp', PACKAGE = 'DESeq2', ...).
 *
 * Semantics kept from the reference: inputs are never modified; outputs are fresh R objects in a named list with
 * the member names of src/DESeq2.cpp:268-276, :458-464, :512; `y` may be INTSXP or REALSXP; errors of the engine
 * become R errors (Rf_error), like BEGIN_RCPP/END_RCPP (src/RcppExports.cpp:17,37).  Call with parallel = FALSE:
 * CUDA must not be initialised before BiocParallel forks workers.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rallocators.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Utils.h>
#include <stdlib.h>

#include "../../include/b200nb.h"

static SEXP named_list(int n, const char **names) {
  SEXP l = PROTECT(Rf_allocVector(VECSXP, n));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
  for (int i = 0; i < n; i++) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  Rf_setAttrib(l, R_NamesSymbol, nm);
  UNPROTECT(2);
  return l;
}

/* Large result matrices (hat_diagonals: n x m doubles) are allocated THROUGH R's custom-allocator hook
 * (allocVector3, R >= 3.1.0) in the engine's pool of page-locked host memory: the device-to-host copy is then one DMA
 * into the R object itself -- no staging, no first-touch page faults (a malloc'ed 40 MB matrix costs ~3 ms of faults).
 * The block goes back to the pool when R's gc collects the object.  If page-locked memory cannot be had the block is
 * ordinary malloc memory (a 16-byte prefix remembers which). */
static void *pin_alloc(R_allocator_t *a, size_t size) {
  (void)a;
  char *p = (char *)b200nb_host_alloc(size + 16);
  const int pinned = p != NULL;
  if (!p) p = (char *)malloc(size + 16);
  if (!p) return NULL;
  *(int *)p = pinned;
  return p + 16;
}
static void pin_free(R_allocator_t *a, void *q) {
  (void)a;
  char *p = (char *)q - 16;
  if (*(int *)p) b200nb_host_free(p); else free(p);
}
static SEXP alloc_result_matrix(int n, int m) {
  if ((double)n * (double)m * sizeof(double) < (double)(1 << 20)) return Rf_allocMatrix(REALSXP, n, m);
  R_allocator_t al = {pin_alloc, pin_free, NULL, NULL};
  SEXP s = PROTECT(Rf_allocVector3(REALSXP, (R_xlen_t)n * m, &al));
  SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
  INTEGER(dim)[0] = n;
  INTEGER(dim)[1] = m;
  Rf_setAttrib(s, R_DimSymbol, dim);
  UNPROTECT(2);
  return s;
}

static const void *y_ptr(SEXP y, int *type) {
  if (TYPEOF(y) == INTSXP) { *type = B200NB_Y_INT32; return INTEGER(y); }
  if (TYPEOF(y) == REALSXP) { *type = B200NB_Y_F64; return REAL(y); }
  Rf_error("y must be an integer or numeric matrix");
  return NULL;
}

SEXP _DESeq2_fitDisp(SEXP ySEXP, SEXP xSEXP, SEXP mu_hatSEXP, SEXP log_alphaSEXP, SEXP log_alpha_prior_meanSEXP,
                     SEXP log_alpha_prior_sigmasqSEXP, SEXP min_log_alphaSEXP, SEXP kappa_0SEXP, SEXP tolSEXP,
                     SEXP maxitSEXP, SEXP usePriorSEXP, SEXP weightsSEXP, SEXP useWeightsSEXP,
                     SEXP weightThresholdSEXP, SEXP useCRSEXP) {
  R_CheckUserInterrupt();   /* the reference polls every 100 genes (src/DESeq2.cpp:195,320,493); a call here takes milliseconds */
  const int n = Rf_nrows(ySEXP), m = Rf_ncols(ySEXP), p = Rf_ncols(xSEXP);
  int yt;
  const void *y = y_ptr(ySEXP, &yt);
  SEXP x = PROTECT(Rf_coerceVector(xSEXP, REALSXP));
  SEXP mu = PROTECT(Rf_coerceVector(mu_hatSEXP, REALSXP));
  SEXP la = PROTECT(Rf_coerceVector(log_alphaSEXP, REALSXP));
  SEXP pm = PROTECT(Rf_coerceVector(log_alpha_prior_meanSEXP, REALSXP));
  SEXP w = PROTECT(Rf_coerceVector(weightsSEXP, REALSXP));
  const char *names[] = {"log_alpha", "iter", "iter_accept", "last_change", "initial_lp",
                         "initial_dlp", "last_lp", "last_dlp", "last_d2lp"};
  SEXP out = PROTECT(named_list(9, names));
  for (int k = 0; k < 9; k++)
    SET_VECTOR_ELT(out, k, Rf_allocVector((k == 1 || k == 2) ? INTSXP : REALSXP, n));
  const int useW = Rf_asLogical(useWeightsSEXP);
  int rc = b200nb_fit_disp(y, yt, REAL(x), REAL(mu), REAL(la), REAL(pm), Rf_asReal(log_alpha_prior_sigmasqSEXP),
                           Rf_asReal(min_log_alphaSEXP), Rf_asReal(kappa_0SEXP), Rf_asReal(tolSEXP),
                           Rf_asInteger(maxitSEXP), Rf_asLogical(usePriorSEXP), useW ? REAL(w) : NULL, useW,
                           Rf_asReal(weightThresholdSEXP), Rf_asLogical(useCRSEXP), n, m, p,
                           REAL(VECTOR_ELT(out, 0)), INTEGER(VECTOR_ELT(out, 1)), INTEGER(VECTOR_ELT(out, 2)),
                           REAL(VECTOR_ELT(out, 3)), REAL(VECTOR_ELT(out, 4)), REAL(VECTOR_ELT(out, 5)),
                           REAL(VECTOR_ELT(out, 6)), REAL(VECTOR_ELT(out, 7)), REAL(VECTOR_ELT(out, 8)));
  UNPROTECT(6);
  if (rc) Rf_error("fitDisp (b200nb): %s", b200nb_last_error());
  return out;
}

SEXP _DESeq2_fitDispGrid(SEXP ySEXP, SEXP xSEXP, SEXP mu_hatSEXP, SEXP disp_gridSEXP,
                         SEXP log_alpha_prior_meanSEXP, SEXP log_alpha_prior_sigmasqSEXP, SEXP usePriorSEXP,
                         SEXP weightsSEXP, SEXP useWeightsSEXP, SEXP weightThresholdSEXP, SEXP useCRSEXP) {
  R_CheckUserInterrupt();   /* the reference polls every 100 genes (src/DESeq2.cpp:195,320,493); a call here takes milliseconds */
  const int n = Rf_nrows(ySEXP), m = Rf_ncols(ySEXP), p = Rf_ncols(xSEXP);
  int yt;
  const void *y = y_ptr(ySEXP, &yt);
  SEXP x = PROTECT(Rf_coerceVector(xSEXP, REALSXP));
  SEXP mu = PROTECT(Rf_coerceVector(mu_hatSEXP, REALSXP));
  SEXP grid = PROTECT(Rf_coerceVector(disp_gridSEXP, REALSXP));
  SEXP pm = PROTECT(Rf_coerceVector(log_alpha_prior_meanSEXP, REALSXP));
  SEXP w = PROTECT(Rf_coerceVector(weightsSEXP, REALSXP));
  const char *names[] = {"log_alpha"};
  SEXP out = PROTECT(named_list(1, names));
  SET_VECTOR_ELT(out, 0, Rf_allocVector(REALSXP, n));
  const int useW = Rf_asLogical(useWeightsSEXP);
  int rc = b200nb_fit_disp_grid(y, yt, REAL(x), REAL(mu), REAL(grid), LENGTH(grid), REAL(pm),
                                Rf_asReal(log_alpha_prior_sigmasqSEXP), Rf_asLogical(usePriorSEXP),
                                useW ? REAL(w) : NULL, useW, Rf_asReal(weightThresholdSEXP),
                                Rf_asLogical(useCRSEXP), n, m, p, REAL(VECTOR_ELT(out, 0)));
  UNPROTECT(6);
  if (rc) Rf_error("fitDispGrid (b200nb): %s", b200nb_last_error());
  return out;
}

SEXP _DESeq2_fitBeta(SEXP ySEXP, SEXP xSEXP, SEXP nfSEXP, SEXP alpha_hatSEXP, SEXP contrastSEXP, SEXP beta_matSEXP,
                     SEXP lambdaSEXP, SEXP weightsSEXP, SEXP useWeightsSEXP, SEXP tolSEXP, SEXP maxitSEXP,
                     SEXP useQRSEXP, SEXP minmuSEXP) {
  R_CheckUserInterrupt();   /* the reference polls every 100 genes (src/DESeq2.cpp:195,320,493); a call here takes milliseconds */
  const int n = Rf_nrows(ySEXP), m = Rf_ncols(ySEXP), p = Rf_ncols(xSEXP);
  int yt;
  const void *y = y_ptr(ySEXP, &yt);
  SEXP x = PROTECT(Rf_coerceVector(xSEXP, REALSXP));
  SEXP nf = PROTECT(Rf_coerceVector(nfSEXP, REALSXP));
  SEXP alpha = PROTECT(Rf_coerceVector(alpha_hatSEXP, REALSXP));
  SEXP contrast = PROTECT(Rf_coerceVector(contrastSEXP, REALSXP));
  SEXP beta0 = PROTECT(Rf_coerceVector(beta_matSEXP, REALSXP));
  SEXP lambda = PROTECT(Rf_coerceVector(lambdaSEXP, REALSXP));
  SEXP w = PROTECT(Rf_coerceVector(weightsSEXP, REALSXP));
  const char *names[] = {"beta_mat", "beta_var_mat", "iter", "hat_diagonals",
                         "contrast_num", "contrast_denom", "deviance"};
  SEXP out = PROTECT(named_list(7, names));
  SET_VECTOR_ELT(out, 0, Rf_allocMatrix(REALSXP, n, p));
  SET_VECTOR_ELT(out, 1, Rf_allocMatrix(REALSXP, n, p));
  SET_VECTOR_ELT(out, 2, Rf_allocVector(REALSXP, n));      /* NumericVector in the reference (:317) */
  SET_VECTOR_ELT(out, 3, alloc_result_matrix(n, m));
  SET_VECTOR_ELT(out, 4, Rf_allocMatrix(REALSXP, n, 1));
  SET_VECTOR_ELT(out, 5, Rf_allocMatrix(REALSXP, n, 1));
  SET_VECTOR_ELT(out, 6, Rf_allocVector(REALSXP, n));
  const int useW = Rf_asLogical(useWeightsSEXP);
  int rc = b200nb_fit_beta(y, yt, REAL(x), REAL(nf), REAL(alpha), REAL(contrast), REAL(beta0), REAL(lambda),
                           useW ? REAL(w) : NULL, useW, Rf_asReal(tolSEXP), Rf_asInteger(maxitSEXP),
                           Rf_asLogical(useQRSEXP), Rf_asReal(minmuSEXP), n, m, p, REAL(VECTOR_ELT(out, 0)),
                           REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)), REAL(VECTOR_ELT(out, 3)),
                           REAL(VECTOR_ELT(out, 4)), REAL(VECTOR_ELT(out, 5)), REAL(VECTOR_ELT(out, 6)), NULL);
  UNPROTECT(8);
  if (rc) Rf_error("fitBeta (b200nb): %s", b200nb_last_error());
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"_DESeq2_fitDisp", (DL_FUNC)&_DESeq2_fitDisp, 15},
    {"_DESeq2_fitBeta", (DL_FUNC)&_DESeq2_fitBeta, 13},
    {"_DESeq2_fitDispGrid", (DL_FUNC)&_DESeq2_fitDispGrid, 11},
    {NULL, NULL, 0}};

void R_init_DESeq2(DllInfo *dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
