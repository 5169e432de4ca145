/*
 * engine_stub_oracle.c -- the three host entry points of include/b200nb.h implemented on the CPU oracle, so that
 * the R shim's argument marshalling can be tested on a machine without a GPU.  TEST INFRASTRUCTURE ONLY (lives
 * under tests/, links oracle/libnbglm_oracle.so); the product library is deseq2_b200/libb200nb.so and has no
 * CPU path.  tests/test_r_shim.py links the shim against this stub for the CPU test and against the real
 * library for the GPU test.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200nb.h"

int oracle_fit_disp(const double *y, const double *x, const double *mu_hat, const double *log_alpha_in,
                    const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                    double kappa_0, double tol, int maxit, int usePrior, const double *weights, int useWeights,
                    double weightThreshold, int useCR, int n, int m, int p, double *log_alpha, int32_t *iter,
                    int32_t *iter_accept, double *last_change, double *initial_lp, double *initial_dlp,
                    double *last_lp, double *last_dlp, double *last_d2lp, double *margin);
int oracle_fit_disp_grid(const double *y, const double *x, const double *mu_hat, const double *disp_grid, int grid_n,
                         const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq, int usePrior,
                         const double *weights, int useWeights, double weightThreshold, int useCR, int n, int m,
                         int p, double *log_alpha);
int oracle_fit_beta(const double *y, const double *x, const double *nf, const double *alpha_hat,
                    const double *contrast, double *beta_mat, const double *lambda, const double *weights,
                    int useWeights, double tol, int maxit, int useQR, double minmu, int n, int m, int p,
                    double *beta_var_mat, double *iter, double *hat_diagonals, double *contrast_num,
                    double *contrast_denom, double *deviance);

static const char *last_error = "";
const char *b200nb_last_error(void) { return last_error; }

/* the pool of page-locked result memory: plain heap here; the counters let tests/test_r_shim.py see that the shim's
 * custom allocator (allocVector3) obtains large result matrices from it and gives them back when R collects them */
#include <stdlib.h>
static int stub_allocs = 0, stub_frees = 0, stub_fail_alloc = 0;
void *b200nb_host_alloc(size_t bytes) {
  if (stub_fail_alloc) return NULL;
  stub_allocs++;
  return malloc(bytes);
}
void b200nb_host_free(void *p) {
  stub_frees++;
  free(p);
}
int stub_host_allocs(void) { return stub_allocs; }
int stub_host_frees(void) { return stub_frees; }
void stub_set_fail_alloc(int v) { stub_fail_alloc = v; }

static double *y_as_double(const void *y, int y_type, size_t len, int *owned) {
  *owned = 0;
  if (y_type == B200NB_Y_F64) return (double *)y;
  if (y_type != B200NB_Y_INT32) return NULL;
  double *d = (double *)malloc(sizeof(double) * (len ? len : 1));
  for (size_t i = 0; i < len; i++) d[i] = (double)((const int32_t *)y)[i];
  *owned = 1;
  return d;
}

int b200nb_fit_disp(const void *y, int y_type, const double *x, const double *mu_hat, const double *log_alpha,
                    const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                    double kappa_0, double tol, int maxit, int use_prior, const double *weights, int use_weights,
                    double weight_threshold, int use_cr, int n, int m, int p, double *out_log_alpha,
                    int32_t *out_iter, int32_t *out_iter_accept, double *out_last_change, double *out_initial_lp,
                    double *out_initial_dlp, double *out_last_lp, double *out_last_dlp, double *out_last_d2lp) {
  int owned;
  double *yd = y_as_double(y, y_type, (size_t)n * m, &owned);
  if (!yd) { last_error = "stub: bad y_type"; return 1; }
  if (p < 1) { if (owned) free(yd); last_error = "stub: p must be >= 1"; return 1; }
  int rc = oracle_fit_disp(yd, x, mu_hat, log_alpha, log_alpha_prior_mean, log_alpha_prior_sigmasq, min_log_alpha,
                           kappa_0, tol, maxit, use_prior, weights, use_weights, weight_threshold, use_cr, n, m, p,
                           out_log_alpha, out_iter, out_iter_accept, out_last_change, out_initial_lp,
                           out_initial_dlp, out_last_lp, out_last_dlp, out_last_d2lp, NULL);
  if (owned) free(yd);
  return rc;
}

int b200nb_fit_disp_grid(const void *y, int y_type, const double *x, const double *mu_hat, const double *disp_grid,
                         int disp_grid_n, const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq,
                         int use_prior, const double *weights, int use_weights, double weight_threshold, int use_cr,
                         int n, int m, int p, double *out_log_alpha) {
  int owned;
  double *yd = y_as_double(y, y_type, (size_t)n * m, &owned);
  if (!yd) { last_error = "stub: bad y_type"; return 1; }
  int rc = oracle_fit_disp_grid(yd, x, mu_hat, disp_grid, disp_grid_n, log_alpha_prior_mean,
                                log_alpha_prior_sigmasq, use_prior, weights, use_weights, weight_threshold, use_cr,
                                n, m, p, out_log_alpha);
  if (owned) free(yd);
  return rc;
}

int b200nb_fit_beta(const void *y, int y_type, const double *x, const double *nf, const double *alpha_hat,
                    const double *contrast, const double *beta_mat, const double *lambda, const double *weights,
                    int use_weights, double tol, int maxit, int use_qr, double minmu, int n, int m, int p,
                    double *out_beta_mat, double *out_beta_var_mat, double *out_iter, double *out_hat_diagonals,
                    double *out_contrast_num, double *out_contrast_denom, double *out_deviance, double *out_mu) {
  int owned;
  double *yd = y_as_double(y, y_type, (size_t)n * m, &owned);
  if (!yd) { last_error = "stub: bad y_type"; return 1; }
  if (out_mu) { if (owned) free(yd); last_error = "stub: out_mu not supported"; return 1; }
  memcpy(out_beta_mat, beta_mat, sizeof(double) * (size_t)n * p); /* the oracle fits in place */
  int rc = oracle_fit_beta(yd, x, nf, alpha_hat, contrast, out_beta_mat, lambda, weights, use_weights, tol, maxit,
                           use_qr, minmu, n, m, p, out_beta_var_mat, out_iter, out_hat_diagonals, out_contrast_num,
                           out_contrast_denom, out_deviance);
  if (owned) free(yd);
  return rc;
}
