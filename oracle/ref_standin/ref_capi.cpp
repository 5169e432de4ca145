/*
 * ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.  Plain-C entry points around the REFERENCE's own fitDisp / fitBeta /
 * fitDispGrid (/root/reference/src/DESeq2.cpp:164,283,469, compiled unchanged from where it lies against the stand-in
 * headers of this directory by oracle/Makefile -> oracle/_ref/libdeseq2_ref.so).  Same argument lists as the oracle's
 * oracle_fit_* (oracle/nbglm_oracle.c) so that oracle/ref.py can drive either; used by tests/ (oracle == reference)
 * and by bench.py's --impl reference / cpu_baseline legs.  The product never loads it.
 *
 * What the wrapper adds around the reference call, and nothing else: building the R-shaped arguments (column-major
 * REALSXP / INTSXP matrices with dims, length-1 scalars, an all-ones weights matrix when the caller has none -- what
 * R/core.R:2749 passes), copying the named list members out, and -- for nthreads > 1 -- the BiocParallel emulation of
 * R/parallel.R:9-10: contiguous gene chunks, one independent reference call per chunk, results concatenated.
 */
#include "RcppArmadillo.h"

#include <omp.h>

#include <cstdint>
#include <string>

Rcpp::List fitDisp(SEXP ySEXP, SEXP xSEXP, SEXP mu_hatSEXP, SEXP log_alphaSEXP, SEXP log_alpha_prior_meanSEXP,
                   SEXP log_alpha_prior_sigmasqSEXP, SEXP min_log_alphaSEXP, SEXP kappa_0SEXP, SEXP tolSEXP,
                   SEXP maxitSEXP, SEXP usePriorSEXP, SEXP weightsSEXP, SEXP useWeightsSEXP, SEXP weightThresholdSEXP,
                   SEXP useCRSEXP);
Rcpp::List fitBeta(SEXP ySEXP, SEXP xSEXP, SEXP nfSEXP, SEXP alpha_hatSEXP, SEXP contrastSEXP, SEXP beta_matSEXP,
                   SEXP lambdaSEXP, SEXP weightsSEXP, SEXP useWeightsSEXP, SEXP tolSEXP, SEXP maxitSEXP, SEXP useQRSEXP,
                   SEXP minmuSEXP);
Rcpp::List fitDispGrid(SEXP ySEXP, SEXP xSEXP, SEXP mu_hatSEXP, SEXP disp_gridSEXP, SEXP log_alpha_prior_meanSEXP,
                       SEXP log_alpha_prior_sigmasqSEXP, SEXP usePriorSEXP, SEXP weightsSEXP, SEXP useWeightsSEXP,
                       SEXP weightThresholdSEXP, SEXP useCRSEXP);

std::vector<std::unique_ptr<standin_sexp>> &standin_arena() {
  static thread_local std::vector<std::unique_ptr<standin_sexp>> arena;
  return arena;
}

static thread_local std::string g_err;
static std::string g_last_err;

namespace {

SEXP scalar_real(double v) {
  SEXP s = standin_new(STANDIN_REALSXP, 1);
  (*s->real)[0] = v;
  return s;
}
SEXP scalar_int(int v) {
  SEXP s = standin_new(STANDIN_INTSXP, 1);
  (*s->ints)[0] = v;
  return s;
}
SEXP scalar_lgl(int v) {
  SEXP s = standin_new(STANDIN_LGLSXP, 1);
  (*s->ints)[0] = v ? 1 : 0;
  return s;
}
SEXP vec_real(const double *v, long n) {
  SEXP s = standin_new(STANDIN_REALSXP, n);
  std::copy(v, v + n, s->real->begin());
  return s;
}
/* rows [g0, g1) of a column-major n x m matrix, as a REALSXP (or INTSXP when as_int) with dims; src == NULL -> ones */
SEXP sub_matrix(const double *src, long n, long m, long g0, long g1, bool as_int = false) {
  long k = g1 - g0;
  SEXP s = standin_new(as_int ? STANDIN_INTSXP : STANDIN_REALSXP, k * m, (int)k, (int)m);
  for (long j = 0; j < m; j++)
    for (long i = 0; i < k; i++) {
      double v = src ? src[(g0 + i) + n * j] : 1.0;
      if (as_int) (*s->ints)[i + k * j] = (int)v; else (*s->real)[i + k * j] = v;
    }
  return s;
}
void copy_out(SEXP s, double *dst, long n, long g0, long k, long cols) { /* k x cols result -> rows g0.. of n x cols */
  if (!dst) return;
  auto r = standin_as_real(s);
  for (long j = 0; j < cols; j++)
    for (long i = 0; i < k; i++) dst[(g0 + i) + n * j] = (*r)[i + k * j];
}
void copy_out_int(SEXP s, int32_t *dst, long g0, long k) {
  if (!dst) return;
  for (long i = 0; i < k; i++) dst[g0 + i] = s->type == STANDIN_REALSXP ? (int32_t)(*s->real)[i] : (*s->ints)[i];
}

template <class F> int run_chunks(long n, int nthreads, F body) {
  int T = nthreads > 0 ? nthreads : 1;
  if (T > n) T = n > 0 ? (int)n : 1;
  int status = 0;
#pragma omp parallel for schedule(static, 1) num_threads(T)
  for (int c = 0; c < T; c++) {
    long g0 = n * c / T, g1 = n * (c + 1) / T;
    try {
      if (g1 > g0) body(g0, g1);
    } catch (const std::exception &e) {
#pragma omp critical
      { status = 1; g_last_err = e.what(); }
    }
    standin_arena().clear();
  }
  return status;
}

}  // namespace

extern "C" {

const char *ref_last_error(void) { return g_last_err.c_str(); }
const char *ref_source(void) { return "/root/reference/src/DESeq2.cpp compiled unchanged against oracle/ref_standin"; }

int ref_fit_disp(const double *y, int y_is_int, const double *x, const double *mu_hat, const double *log_alpha_in,
                 const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                 double kappa_0, double tol, int maxit, int usePrior, const double *weights, int useWeights,
                 double weightThreshold, int useCR, int n, int m, int p, int nthreads, double *log_alpha,
                 int32_t *iter, int32_t *iter_accept, double *last_change, double *initial_lp, double *initial_dlp,
                 double *last_lp, double *last_dlp, double *last_d2lp) {
  return run_chunks(n, nthreads, [&](long g0, long g1) {
    long k = g1 - g0;
    SEXP xs = standin_new(STANDIN_REALSXP, (long)m * p, m, p);
    std::copy(x, x + (long)m * p, xs->real->begin());
    Rcpp::List r = fitDisp(sub_matrix(y, n, m, g0, g1, y_is_int != 0), xs, sub_matrix(mu_hat, n, m, g0, g1),
                           vec_real(log_alpha_in + g0, k), vec_real(log_alpha_prior_mean + g0, k),
                           scalar_real(log_alpha_prior_sigmasq), scalar_real(min_log_alpha), scalar_real(kappa_0),
                           scalar_real(tol), scalar_int(maxit), scalar_lgl(usePrior),
                           sub_matrix(weights, n, m, g0, g1), scalar_lgl(useWeights), scalar_real(weightThreshold),
                           scalar_lgl(useCR));
    copy_out(r["log_alpha"], log_alpha, n, g0, k, 1);
    copy_out_int(r["iter"], iter, g0, k);
    copy_out_int(r["iter_accept"], iter_accept, g0, k);
    copy_out(r["last_change"], last_change, n, g0, k, 1);
    copy_out(r["initial_lp"], initial_lp, n, g0, k, 1);
    copy_out(r["initial_dlp"], initial_dlp, n, g0, k, 1);
    copy_out(r["last_lp"], last_lp, n, g0, k, 1);
    copy_out(r["last_dlp"], last_dlp, n, g0, k, 1);
    copy_out(r["last_d2lp"], last_d2lp, n, g0, k, 1);
  });
}

int ref_fit_disp_grid(const double *y, int y_is_int, const double *x, const double *mu_hat, const double *disp_grid,
                      int grid_n, const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq, int usePrior,
                      const double *weights, int useWeights, double weightThreshold, int useCR, int n, int m, int p,
                      int nthreads, double *log_alpha) {
  return run_chunks(n, nthreads, [&](long g0, long g1) {
    long k = g1 - g0;
    SEXP xs = standin_new(STANDIN_REALSXP, (long)m * p, m, p);
    std::copy(x, x + (long)m * p, xs->real->begin());
    Rcpp::List r = fitDispGrid(sub_matrix(y, n, m, g0, g1, y_is_int != 0), xs, sub_matrix(mu_hat, n, m, g0, g1),
                               vec_real(disp_grid, grid_n), vec_real(log_alpha_prior_mean + g0, k),
                               scalar_real(log_alpha_prior_sigmasq), scalar_lgl(usePrior),
                               sub_matrix(weights, n, m, g0, g1), scalar_lgl(useWeights), scalar_real(weightThreshold),
                               scalar_lgl(useCR));
    copy_out(r["log_alpha"], log_alpha, n, g0, k, 1);
  });
}

/* beta_mat: n x p column-major, start values in, fit out (the reference returns a fresh matrix; copied back here) */
int ref_fit_beta(const double *y, int y_is_int, const double *x, const double *nf, const double *alpha_hat,
                 const double *contrast, double *beta_mat, const double *lambda, const double *weights, int useWeights,
                 double tol, int maxit, int useQR, double minmu, int n, int m, int p, int nthreads,
                 double *beta_var_mat, double *iter, double *hat_diagonals, double *contrast_num,
                 double *contrast_denom, double *deviance) {
  return run_chunks(n, nthreads, [&](long g0, long g1) {
    long k = g1 - g0;
    SEXP xs = standin_new(STANDIN_REALSXP, (long)m * p, m, p);
    std::copy(x, x + (long)m * p, xs->real->begin());
    Rcpp::List r = fitBeta(sub_matrix(y, n, m, g0, g1, y_is_int != 0), xs, sub_matrix(nf, n, m, g0, g1),
                           vec_real(alpha_hat + g0, k), vec_real(contrast, p), sub_matrix(beta_mat, n, p, g0, g1),
                           vec_real(lambda, p), sub_matrix(weights, n, m, g0, g1), scalar_lgl(useWeights),
                           scalar_real(tol), scalar_int(maxit), scalar_lgl(useQR), scalar_real(minmu));
    copy_out(r["beta_mat"], beta_mat, n, g0, k, p);
    copy_out(r["beta_var_mat"], beta_var_mat, n, g0, k, p);
    copy_out(r["iter"], iter, n, g0, k, 1);
    copy_out(r["hat_diagonals"], hat_diagonals, n, g0, k, m);
    copy_out(r["contrast_num"], contrast_num, n, g0, k, 1);
    copy_out(r["contrast_denom"], contrast_denom, n, g0, k, 1);
    copy_out(r["deviance"], deviance, n, g0, k, 1);
  });
}

}  // extern "C"
