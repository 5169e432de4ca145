/*
 * nbglm_oracle.c -- CPU fp64 ORACLE for the DESeq2 native hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA engine in
 * deseq2_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (deseq2_b200/) never links or calls it.
 *
 * PARITY PIN: R, Rcpp, RcppArmadillo and libRmath are absent from this image, but the reference's
 * translation unit /root/reference/src/DESeq2.cpp IS compiled here, unchanged, against stand-in
 * headers for the closed subset of Rcpp / Armadillo / Rmath it uses (oracle/ref_standin/, recipe
 * oracle/Makefile `ref` -> oracle/_ref/libdeseq2_ref.so), and tests/test_oracle_vs_reference.py holds
 * this restatement to it: identical line-search control flow on every non-knife-edge gene, 1e-9 on
 * every fitBeta output, identical grid points.  What stays unpinned is only what the reference itself
 * does not vendor: R's nmath (restated from the published algorithms, checked against mpmath) and
 * LAPACK (any backward-stable p x p solve).  This file is the restatement, function by function:
 *
 *   oracle_fit_beta       <- src/DESeq2.cpp:283-465  (fitBeta: QR branch :334-383, normal-eq :388-425,
 *                                                     post-loop hat diagonal / sandwich covariance :429-455)
 *   oracle_fit_disp       <- src/DESeq2.cpp:164-277  (fitDisp: Armijo back-tracking line search)
 *   oracle_fit_disp_grid  <- src/DESeq2.cpp:469-513  (fitDispGrid: coarse + fine grid argmax)
 *   log_posterior         <- src/DESeq2.cpp:31-64
 *   dlog_posterior        <- src/DESeq2.cpp:68-107
 *   d2log_posterior       <- src/DESeq2.cpp:111-158
 *
 * Third-party arithmetic the reference calls but does not vendor (R nmath, version unpinned by
 * DESCRIPTION): dnbinom_mu / dbinom_raw / bd0 / stirlerr are restated below from the published
 * algorithm (Loader 2000 saddle-point form as used by R <= 4.3); lgammafn -> glibc lgamma_r;
 * digamma / trigamma -> asymptotic series + upward recurrence (checked against mpmath in tests/).
 * Armadillo qr_econ / solve / inv / det -> Householder QR and partially pivoted LU below.
 *
 * All matrices are R layout: column-major, genes = rows (y[i + n*j] is gene i, sample j).
 * Genes are independent, so the outer loops carry an OpenMP pragma (emulating the reference's
 * BiocParallel gene chunking, R/parallel.R:9-10); with OMP_NUM_THREADS=1 this is the reference's
 * default serial loop.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sferr_halves.h"

#define LN_SQRT_2PI 0.918938533204672741780329736406
#define LN_2PI 1.837877066409345483560659472811

/* ------------------------------------------------------------------ special functions */

static double o_lgamma(double x) {
  int sg;
  return lgamma_r(x, &sg);
}

/* digamma for x > 0: recurrence up to x >= 20 then asymptotic series to B_14 */
static double o_digamma(double x) {
  double s = 0.0;
  while (x < 20.0) {
    s -= 1.0 / x;
    x += 1.0;
  }
  double xi = 1.0 / x, x2 = xi * xi;
  /* B_2k/(2k): 1/12, -1/120, 1/252, -1/240, 1/132, -691/32760, 1/12 */
  double ser = x2 * (1.0 / 12.0 -
               x2 * (1.0 / 120.0 -
               x2 * (1.0 / 252.0 -
               x2 * (1.0 / 240.0 -
               x2 * (1.0 / 132.0 -
               x2 * (691.0 / 32760.0 -
               x2 * (1.0 / 12.0)))))));
  return s + log(x) - 0.5 * xi - ser;
}

/* trigamma for x > 0 */
static double o_trigamma(double x) {
  double s = 0.0;
  while (x < 20.0) {
    s += 1.0 / (x * x);
    x += 1.0;
  }
  double xi = 1.0 / x, x2 = xi * xi;
  /* 1/x + 1/(2x^2) + sum B_2k / x^(2k+1) ; B2=1/6,B4=-1/30,B6=1/42,B8=-1/30,B10=5/66,B12=-691/2730,B14=7/6 */
  double ser = xi * x2 * (1.0 / 6.0 -
                    x2 * (1.0 / 30.0 -
                    x2 * (1.0 / 42.0 -
                    x2 * (1.0 / 30.0 -
                    x2 * (5.0 / 66.0 -
                    x2 * (691.0 / 2730.0 -
                    x2 * (7.0 / 6.0)))))));
  return s + xi + 0.5 * x2 + ser;
}

/* ---- R nmath restatement: stirlerr, bd0, dbinom_raw, dnbinom_mu (all log=TRUE) ---- */

static double o_stirlerr(double n) {
  const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
  if (n <= 15.0) {
    double nn = n + n;
    if (nn == (int)nn) return sferr_halves[(int)nn];
    return o_lgamma(n + 1.0) - (n + 0.5) * log(n) + n - LN_SQRT_2PI;
  }
  double nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

static double o_bd0(double x, double np) {
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    if (fabs(s) < DBL_MIN) return s;
    double ej = 2 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / ((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / np) + np - x;
}

static double o_dbinom_raw_log(double x, double n, double p, double q) {
  if (p == 0) return (x == 0) ? 0.0 : -INFINITY;
  if (q == 0) return (x == n) ? 0.0 : -INFINITY;
  if (x == 0) {
    if (n == 0) return 0.0;
    return (p < 0.1) ? -o_bd0(n, n * q) - n * p : n * log(q);
  }
  if (x == n) {
    return (q < 0.1) ? -o_bd0(n, n * p) - n * q : n * log(p);
  }
  if (x < 0 || x > n) return -INFINITY;
  double lc = o_stirlerr(n) - o_stirlerr(x) - o_stirlerr(n - x) - o_bd0(x, n * p) - o_bd0(n - x, n * q);
  double lf = LN_2PI + log(x) + log1p(-x / n);
  return lc - 0.5 * lf;
}

/* Rf_dnbinom_mu(x, size, mu, give_log=1) as called at src/DESeq2.cpp:369,371,411,413 */
double oracle_dnbinom_mu_log(double x, double size, double mu) {
  if (x == 0) return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
  if (x < 1e-10 * size) {
    double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
    return x * p - mu - o_lgamma(x + 1) + log1p(x * (x - 1) / (2 * size));
  }
  double p = size / (size + x);
  double ans = o_dbinom_raw_log(size, x + size, size / (size + mu), mu / (size + mu));
  return log(p) + ans;
}

double oracle_digamma(double x) { return o_digamma(x); }
double oracle_trigamma(double x) { return o_trigamma(x); }
double oracle_lgamma(double x) { return o_lgamma(x); }

/* ------------------------------------------------------------------ small dense linear algebra
 * square matrices are column-major k x k with leading dimension k                              */

/* LU with partial pivoting in place; returns sign of permutation (0 if exactly singular) */
static int lu_factor(double *a, int k, int *piv) {
  int sign = 1;
  for (int c = 0; c < k; c++) {
    int pr = c;
    double best = fabs(a[c + k * c]);
    for (int r = c + 1; r < k; r++)
      if (fabs(a[r + k * c]) > best) { best = fabs(a[r + k * c]); pr = r; }
    piv[c] = pr;
    if (pr != c) {
      sign = -sign;
      for (int cc = 0; cc < k; cc++) { double t = a[c + k * cc]; a[c + k * cc] = a[pr + k * cc]; a[pr + k * cc] = t; }
    }
    double d = a[c + k * c];
    if (d == 0.0) return 0;
    for (int r = c + 1; r < k; r++) {
      double f = a[r + k * c] / d;
      a[r + k * c] = f;
      for (int cc = c + 1; cc < k; cc++) a[r + k * cc] -= f * a[c + k * cc];
    }
  }
  return sign;
}

static double lu_det(const double *a, int k, double *work, int *piv) {
  memcpy(work, a, sizeof(double) * k * k);
  int s = lu_factor(work, k, piv);
  if (s == 0) return 0.0;
  double d = s;
  for (int c = 0; c < k; c++) d *= work[c + k * c];
  return d;
}

/* solve A X = B (B is k x nrhs col-major, overwritten) using LU; returns 0 if singular */
static int lu_solve(const double *a, int k, double *b, int nrhs, double *work, int *piv) {
  memcpy(work, a, sizeof(double) * k * k);
  if (lu_factor(work, k, piv) == 0) return 0;
  for (int q = 0; q < nrhs; q++) {
    double *v = b + (size_t)k * q;
    for (int c = 0; c < k; c++) {
      if (piv[c] != c) { double t = v[c]; v[c] = v[piv[c]]; v[piv[c]] = t; }
    }
    for (int c = 0; c < k; c++)
      for (int r = c + 1; r < k; r++) v[r] -= work[r + k * c] * v[c];
    for (int c = k - 1; c >= 0; c--) {
      v[c] /= work[c + k * c];
      for (int r = 0; r < c; r++) v[r] -= work[r + k * c] * v[c];
    }
  }
  return 1;
}

static int lu_inverse(const double *a, int k, double *inv, double *work, int *piv) {
  memset(inv, 0, sizeof(double) * k * k);
  for (int c = 0; c < k; c++) inv[c + k * c] = 1.0;
  int ok = lu_solve(a, k, inv, k, work, piv);
  if (!ok)
    for (int c = 0; c < k * k; c++) inv[c] = NAN;
  return ok;
}

/* B = X' diag(w) X over selected rows/cols.  xs: m x p col-major.  rows: index list (nr),
 * cols: index list (nc).  out: nc x nc col-major. */
static void xtwx_sub(const double *x, int m, const int *rows, int nr, const int *cols, int nc,
                     const double *w, double *out) {
  for (int a = 0; a < nc; a++)
    for (int b = 0; b <= a; b++) {
      const double *xa = x + (size_t)m * cols[a], *xb = x + (size_t)m * cols[b];
      double s = 0.0;
      for (int r = 0; r < nr; r++) { int j = rows[r]; s += xa[j] * w[j] * xb[j]; }
      out[a + nc * b] = s;
      out[b + nc * a] = s;
    }
}

static double mat_trace_prod(const double *a, const double *b, int k) { /* trace(A*B) */
  double t = 0.0;
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) t += a[i + k * j] * b[j + k * i];
  return t;
}

static void mat_mul(const double *a, const double *b, double *c, int k) {
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) {
      double s = 0.0;
      for (int l = 0; l < k; l++) s += a[i + k * l] * b[l + k * j];
      c[i + k * j] = s;
    }
}

/* ------------------------------------------------------------------ posterior of log alpha */

typedef struct {
  int m, p;
  const double *x;    /* m x p col-major */
  int *rows, *cols;   /* scratch: selected rows / columns for the Cox-Reid term */
  double *wd, *dwd, *d2wd; /* m */
  double *b, *db, *d2b, *bi, *t1, *t2, *work; /* p*p each */
  int *piv;
  double *yrow, *murow, *wrow; /* m: contiguous copies of the gene's rows */
  double abs_sum;              /* diagnostics only: sum of |terms| of the last log_posterior() call */
} disp_ws;

static disp_ws *disp_ws_new(int m, int p, const double *x) {
  disp_ws *w = (disp_ws *)calloc(1, sizeof(disp_ws));
  w->m = m; w->p = p; w->x = x;
  w->rows = (int *)malloc(sizeof(int) * m);
  w->cols = (int *)malloc(sizeof(int) * p);
  w->wd = (double *)malloc(sizeof(double) * m);
  w->dwd = (double *)malloc(sizeof(double) * m);
  w->d2wd = (double *)malloc(sizeof(double) * m);
  size_t pp = (size_t)p * p;
  w->b = (double *)malloc(sizeof(double) * pp); w->db = (double *)malloc(sizeof(double) * pp);
  w->d2b = (double *)malloc(sizeof(double) * pp); w->bi = (double *)malloc(sizeof(double) * pp);
  w->t1 = (double *)malloc(sizeof(double) * pp); w->t2 = (double *)malloc(sizeof(double) * pp);
  w->work = (double *)malloc(sizeof(double) * pp);
  w->piv = (int *)malloc(sizeof(int) * p);
  w->yrow = (double *)malloc(sizeof(double) * m);
  w->murow = (double *)malloc(sizeof(double) * m);
  w->wrow = (double *)malloc(sizeof(double) * m);
  return w;
}

static void disp_ws_free(disp_ws *w) {
  free(w->rows); free(w->cols); free(w->wd); free(w->dwd); free(w->d2wd);
  free(w->b); free(w->db); free(w->d2b); free(w->bi); free(w->t1); free(w->t2); free(w->work);
  free(w->piv); free(w->yrow); free(w->murow); free(w->wrow); free(w);
}

/* Cox-Reid row/column subsetting, src/DESeq2.cpp:38-44: keep rows with weight > threshold, then
 * keep columns whose sum of |x| over kept rows is > 0.  Returns nc. */
static int cr_select(disp_ws *s, int useWeights, double weightThreshold, int *nr_out) {
  int m = s->m, p = s->p, nr = 0, nc = 0;
  for (int j = 0; j < m; j++)
    if (!useWeights || s->wrow[j] > weightThreshold) s->rows[nr++] = j;
  for (int k = 0; k < p; k++) {
    if (!useWeights) { s->cols[nc++] = k; continue; }
    double a = 0.0;
    for (int r = 0; r < nr; r++) a += fabs(s->x[s->rows[r] + (size_t)m * k]);
    if (a > 0.0) s->cols[nc++] = k;
  }
  *nr_out = nr;
  return nc;
}

/* src/DESeq2.cpp:31-64 */
static double log_posterior(disp_ws *s, double log_alpha, double prior_mean, double prior_sigmasq,
                            int usePrior, int useWeights, double weightThreshold, int useCR) {
  int m = s->m;
  double alpha = exp(log_alpha);
  double cr_term = 0.0;
  if (useCR) {
    for (int j = 0; j < m; j++) s->wd[j] = 1.0 / (1.0 / s->murow[j] + alpha);
    int nr, nc = cr_select(s, useWeights, weightThreshold, &nr);
    xtwx_sub(s->x, m, s->rows, nr, s->cols, nc, s->wd, s->b);
    cr_term = -0.5 * log(lu_det(s->b, nc, s->work, s->piv));
  }
  double alpha_neg1 = 1.0 / alpha;
  double lg_an1 = o_lgamma(alpha_neg1);
  double ll = 0.0, asum = 0.0;
  for (int j = 0; j < m; j++) {
    double y = s->yrow[j], mu = s->murow[j];
    double lg1 = o_lgamma(y + alpha_neg1), t2 = y * log(mu + alpha_neg1), t3 = alpha_neg1 * log(1.0 + mu * alpha);
    double t = lg1 - lg_an1 - t2 - t3;
    ll += useWeights ? s->wrow[j] * t : t;
    asum += (useWeights ? s->wrow[j] : 1.0) * (fabs(lg1) + fabs(lg_an1) + fabs(t2) + fabs(t3));
  }
  double prior_part = usePrior ? -0.5 * (log_alpha - prior_mean) * (log_alpha - prior_mean) / prior_sigmasq : 0.0;
  s->abs_sum = asum + fabs(prior_part) + fabs(cr_term);
  return ll + prior_part + cr_term;
}

/* src/DESeq2.cpp:68-107 */
static double dlog_posterior(disp_ws *s, double log_alpha, double prior_mean, double prior_sigmasq,
                             int usePrior, int useWeights, double weightThreshold, int useCR) {
  int m = s->m;
  double alpha = exp(log_alpha);
  double cr_term = 0.0;
  if (useCR) {
    for (int j = 0; j < m; j++) {
      double u = 1.0 / s->murow[j] + alpha;
      s->wd[j] = 1.0 / u;
      s->dwd[j] = -1.0 / (u * u);
    }
    int nr, nc = cr_select(s, useWeights, weightThreshold, &nr);
    xtwx_sub(s->x, m, s->rows, nr, s->cols, nc, s->wd, s->b);
    xtwx_sub(s->x, m, s->rows, nr, s->cols, nc, s->dwd, s->db);
    double detb = lu_det(s->b, nc, s->work, s->piv);
    lu_inverse(s->b, nc, s->bi, s->work, s->piv);
    double ddetb = detb * mat_trace_prod(s->bi, s->db, nc);
    cr_term = -0.5 * ddetb / detb;
  }
  double alpha_neg1 = 1.0 / alpha, alpha_neg2 = 1.0 / (alpha * alpha);
  double dg_an1 = o_digamma(alpha_neg1);
  double ll = 0.0;
  for (int j = 0; j < m; j++) {
    double y = s->yrow[j], mu = s->murow[j];
    double t = dg_an1 + log(1 + mu * alpha) - mu * alpha / (1.0 + mu * alpha) - o_digamma(y + alpha_neg1) +
               y / (mu + alpha_neg1);
    ll += useWeights ? s->wrow[j] * t : t;
  }
  ll *= alpha_neg2;
  double prior_part = usePrior ? -1.0 * (log_alpha - prior_mean) / prior_sigmasq : 0.0;
  return (ll + cr_term) * alpha + prior_part;
}

/* src/DESeq2.cpp:111-158 */
static double d2log_posterior(disp_ws *s, double log_alpha, double prior_mean, double prior_sigmasq,
                              int usePrior, int useWeights, double weightThreshold, int useCR) {
  int m = s->m;
  double alpha = exp(log_alpha);
  double cr_term = 0.0;
  if (useCR) {
    for (int j = 0; j < m; j++) {
      double u = 1.0 / s->murow[j] + alpha;
      s->wd[j] = 1.0 / u;
      s->dwd[j] = -1.0 / (u * u);
      s->d2wd[j] = 2.0 / (u * u * u);
    }
    int nr, nc = cr_select(s, useWeights, weightThreshold, &nr);
    xtwx_sub(s->x, m, s->rows, nr, s->cols, nc, s->wd, s->b);
    xtwx_sub(s->x, m, s->rows, nr, s->cols, nc, s->dwd, s->db);
    xtwx_sub(s->x, m, s->rows, nr, s->cols, nc, s->d2wd, s->d2b);
    double detb = lu_det(s->b, nc, s->work, s->piv);
    lu_inverse(s->b, nc, s->bi, s->work, s->piv);
    double tr1 = mat_trace_prod(s->bi, s->db, nc);
    double ddetb = detb * tr1;
    mat_mul(s->bi, s->db, s->t1, nc);   /* b_i * db */
    mat_mul(s->t1, s->t1, s->t2, nc);   /* b_i db b_i db */
    double tr2 = 0.0;
    for (int c = 0; c < nc; c++) tr2 += s->t2[c + nc * c];
    double tr3 = mat_trace_prod(s->bi, s->d2b, nc);
    double d2detb = detb * (tr1 * tr1 - tr2 + tr3);
    cr_term = 0.5 * (ddetb / detb) * (ddetb / detb) - 0.5 * d2detb / detb;
  }
  double alpha_neg1 = 1.0 / alpha, alpha_neg2 = 1.0 / (alpha * alpha);
  double dg_an1 = o_digamma(alpha_neg1), tg_an1 = o_trigamma(alpha_neg1);
  double s1 = 0.0, s2 = 0.0;
  for (int j = 0; j < m; j++) {
    double y = s->yrow[j], mu = s->murow[j];
    double onema = 1 + mu * alpha;
    double t1 = dg_an1 + log(onema) - mu * alpha / onema - o_digamma(y + alpha_neg1) + y / (mu + alpha_neg1);
    double t2 = -1 * alpha_neg2 * tg_an1 + mu * mu * alpha / (onema * onema) +
                alpha_neg2 * o_trigamma(y + alpha_neg1) + alpha_neg2 * y / ((mu + alpha_neg1) * (mu + alpha_neg1));
    if (useWeights) { t1 *= s->wrow[j]; t2 *= s->wrow[j]; }
    s1 += t1;
    s2 += t2;
  }
  double ll = -2 * (1.0 / (alpha * alpha * alpha)) * s1 + alpha_neg2 * s2;
  double prior_part = usePrior ? -1.0 / prior_sigmasq : 0.0;
  return ((ll + cr_term) * alpha * alpha +
          dlog_posterior(s, log_alpha, prior_mean, prior_sigmasq, 0, useWeights, weightThreshold, useCR)) +
         prior_part;
}

static void load_rows(disp_ws *s, const double *y, const double *mu, const double *w, int n, int i) {
  for (int j = 0; j < s->m; j++) {
    s->yrow[j] = y[i + (size_t)n * j];
    s->murow[j] = mu[i + (size_t)n * j];
    s->wrow[j] = w ? w[i + (size_t)n * j] : 1.0;
  }
}

/* src/DESeq2.cpp:164-277.  Outputs are the nine named list members (:268-276).
 * `margin` (optional, may be NULL) is NOT part of the reference: it records, per gene, the smallest
 * distance |lhs-rhs| over every floating-point branch decision of the line search, in units of the
 * a-priori rounding-error bound of the log-posterior values entering that decision
 * (2^-52 * sum of |terms|).  margin >> 1 means no correct fp64 evaluation of the same formulas can take
 * the other branch; margin ~ 1 is a knife-edge decision.  Tests use it to separate the two. */
int oracle_fit_disp(const double *y, const double *x, const double *mu_hat, const double *log_alpha_in,
                    const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                    double kappa_0, double tol, int maxit, int usePrior, const double *weights, int useWeights,
                    double weightThreshold, int useCR, int n, int m, int p,
                    double *log_alpha, int32_t *iter, int32_t *iter_accept, double *last_change,
                    double *initial_lp, double *initial_dlp, double *last_lp, double *last_dlp,
                    double *last_d2lp, double *margin) {
  const double epsilon = 1.0e-4;
#pragma omp parallel
  {
    disp_ws *s = disp_ws_new(m, p, x);
#pragma omp for schedule(dynamic, 8)
    for (int i = 0; i < n; i++) {
      load_rows(s, y, mu_hat, weights, n, i);
      double pm = log_alpha_prior_mean[i];
      double a = log_alpha_in[i];
      double lp = log_posterior(s, a, pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
      double lp_abs = s->abs_sum;
      double dlp = dlog_posterior(s, a, pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
      double kappa = kappa_0;
      initial_lp[i] = lp;
      initial_dlp[i] = dlp;
      double change = -1.0;
      int it = 0, acc = 0;
      double mg = INFINITY;
      for (int t = 0; t < maxit; t++) {
        it++;
        double a_propose = a + kappa * dlp;
        if (a_propose < -30.0) kappa = (-30.0 - a) / dlp;
        if (a_propose > 10.0) kappa = (10.0 - a) / dlp;
        const int clamped_hi = a_propose > 10.0;
        double theta_kappa = -1.0 * log_posterior(s, a + kappa * dlp, pm, log_alpha_prior_sigmasq, usePrior,
                                                  useWeights, weightThreshold, useCR);
        double theta_hat_kappa = -1.0 * lp - kappa * epsilon * dlp * dlp;
        double prop_abs = s->abs_sum;
        if (a + kappa * dlp != a) { /* a zero-length step re-evaluates the same point: an exact, implementation-independent tie */
          double d = fabs(theta_kappa - theta_hat_kappa) / (DBL_EPSILON * (prop_abs + lp_abs));
          if (d < mg) mg = d;
        }
        if (theta_kappa <= theta_hat_kappa) {
          acc++;
          a = a + kappa * dlp;
          double lpnew = log_posterior(s, a, pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
          change = lpnew - lp;
          {
            double d = fabs(change - tol) / (DBL_EPSILON * (prop_abs + lp_abs));
            if (d < mg) mg = d;
          }
          if (change < tol) { lp = lpnew; break; }
          {
            double d = fabs(a - min_log_alpha) / (DBL_EPSILON * 64.0 * (fabs(a) + 1.0));
            if (d < mg) mg = d;
          }
          if (a < min_log_alpha) break;
          lp = lpnew;
          lp_abs = prop_abs;
          dlp = dlog_posterior(s, a, pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
          /* The search goes on from a point that was clamped to the upper bound: a + ((10 - a) / dlp) * dlp is 10 only up
           * to rounding (it can land a few ulps above), and with dlp > 0 every later proposal is clamped again with
           * kappa = (10 - a) / dlp of either sign or zero -- the reference itself then stops after one more step or
           * spins to maxit depending on the last bits.  Not a decision any implementation can be held to. */
          if (clamped_hi && dlp > 0.0) mg = 0.0;
          kappa = fmin(kappa * 1.1, kappa_0);
          if (acc % 5 == 0) kappa = kappa / 2.0;
        } else {
          kappa = kappa / 2.0;
        }
      }
      last_lp[i] = lp;
      last_dlp[i] = dlp;
      last_d2lp[i] = d2log_posterior(s, a, pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
      log_alpha[i] = a;
      last_change[i] = change;
      iter[i] = it;
      iter_accept[i] = acc;
      if (margin) margin[i] = mg;
    }
    disp_ws_free(s);
  }
  return 0;
}

/* src/DESeq2.cpp:469-513 */
int oracle_fit_disp_grid(const double *y, const double *x, const double *mu_hat, const double *disp_grid,
                         int grid_n, const double *log_alpha_prior_mean, double log_alpha_prior_sigmasq,
                         int usePrior, const double *weights, int useWeights, double weightThreshold, int useCR,
                         int n, int m, int p, double *log_alpha) {
  double delta = disp_grid[1] - disp_grid[0];
#pragma omp parallel
  {
    disp_ws *s = disp_ws_new(m, p, x);
    double *lpv = (double *)malloc(sizeof(double) * grid_n);
    double *fine = (double *)malloc(sizeof(double) * grid_n);
#pragma omp for schedule(dynamic, 4)
    for (int i = 0; i < n; i++) {
      load_rows(s, y, mu_hat, weights, n, i);
      double pm = log_alpha_prior_mean[i];
      int idx = 0;
      for (int t = 0; t < grid_n; t++) {
        lpv[t] = log_posterior(s, disp_grid[t], pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
        if (lpv[t] > lpv[idx]) idx = t; /* arma .max(): first maximal element */
      }
      double a_hat = disp_grid[idx];
      /* arma::linspace(a_hat - delta, a_hat + delta, N): start + i*step, last point forced to end */
      double start = a_hat - delta, end = a_hat + delta;
      double step = (end - start) / (double)(grid_n - 1);
      for (int t = 0; t < grid_n - 1; t++) fine[t] = start + t * step;
      fine[grid_n - 1] = end;
      idx = 0;
      for (int t = 0; t < grid_n; t++) {
        lpv[t] = log_posterior(s, fine[t], pm, log_alpha_prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
        if (lpv[t] > lpv[idx]) idx = t;
      }
      log_alpha[i] = fine[idx];
    }
    free(lpv); free(fine);
    disp_ws_free(s);
  }
  return 0;
}

/* ------------------------------------------------------------------ fitBeta */

/* economy Householder QR of a (rows x p) col-major matrix `a` (overwritten); applies Q' to rhs
 * (length rows, overwritten).  On exit the upper triangle of a[0:p,0:p] holds R and rhs[0:p] = Q'rhs. */
static void householder_qr_apply(double *a, int rows, int p, double *rhs) {
  for (int c = 0; c < p; c++) {
    double *col = a + (size_t)rows * c;
    double nrm = 0.0;
    for (int r = c; r < rows; r++) nrm += col[r] * col[r];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    double alpha = (col[c] > 0) ? -nrm : nrm;
    double v0 = col[c] - alpha;
    /* v = (v0, col[c+1..]) ; beta = 2 / (v'v) */
    double vtv = v0 * v0;
    for (int r = c + 1; r < rows; r++) vtv += col[r] * col[r];
    if (vtv == 0.0) { col[c] = alpha; continue; }
    double beta = 2.0 / vtv;
    for (int cc = c + 1; cc < p; cc++) {
      double *oc = a + (size_t)rows * cc;
      double d = v0 * oc[c];
      for (int r = c + 1; r < rows; r++) d += col[r] * oc[r];
      d *= beta;
      oc[c] -= d * v0;
      for (int r = c + 1; r < rows; r++) oc[r] -= d * col[r];
    }
    {
      double d = v0 * rhs[c];
      for (int r = c + 1; r < rows; r++) d += col[r] * rhs[r];
      d *= beta;
      rhs[c] -= d * v0;
      for (int r = c + 1; r < rows; r++) rhs[r] -= d * col[r];
    }
    col[c] = alpha;
  }
}

/* src/DESeq2.cpp:283-465.  beta_mat is n x p col-major: input = starting values, output = fit.
 * iter is REALSXP in the reference (NumericVector, :317) hence double here. */
int oracle_fit_beta(const double *y, const double *x, const double *nf, const double *alpha_hat,
                    const double *contrast, double *beta_mat, const double *lambda, const double *weights,
                    int useWeights, double tol, int maxit, int useQR, double minmu, int n, int m, int p,
                    double *beta_var_mat, double *iter, double *hat_diagonals, double *contrast_num,
                    double *contrast_denom, double *deviance) {
  const double large = 30.0;
#pragma omp parallel
  {
    size_t pp = (size_t)p * p;
    double *yrow = (double *)malloc(sizeof(double) * m), *nfrow = (double *)malloc(sizeof(double) * m);
    double *wrow = (double *)malloc(sizeof(double) * m), *mu = (double *)malloc(sizeof(double) * m);
    double *wv = (double *)malloc(sizeof(double) * m), *z = (double *)malloc(sizeof(double) * m);
    double *beta = (double *)malloc(sizeof(double) * p);
    double *stack = (double *)malloc(sizeof(double) * (size_t)(m + p) * p);
    double *rhs = (double *)malloc(sizeof(double) * (m + p));
    double *A = (double *)malloc(sizeof(double) * pp), *xtwx = (double *)malloc(sizeof(double) * pp);
    double *Ainv = (double *)malloc(sizeof(double) * pp), *work = (double *)malloc(sizeof(double) * pp);
    double *T1 = (double *)malloc(sizeof(double) * pp), *sigma = (double *)malloc(sizeof(double) * pp);
    int *piv = (int *)malloc(sizeof(int) * p);
#pragma omp for schedule(dynamic, 8)
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < m; j++) {
        yrow[j] = y[i + (size_t)n * j];
        nfrow[j] = nf[i + (size_t)n * j];
        wrow[j] = weights ? weights[i + (size_t)n * j] : 1.0;
      }
      for (int k = 0; k < p; k++) beta[k] = beta_mat[i + (size_t)n * k];
      double alpha = alpha_hat[i];
      for (int j = 0; j < m; j++) {
        double eta = 0.0;
        for (int k = 0; k < p; k++) eta += x[j + (size_t)m * k] * beta[k];
        mu[j] = fmax(nfrow[j] * exp(eta), minmu);
      }
      double dev = 0.0, dev_old = 0.0;
      double it = 0.0;
      for (int t = 0; t < maxit; t++) {
        it += 1.0;
        for (int j = 0; j < m; j++) {
          wv[j] = useWeights ? wrow[j] * mu[j] / (1.0 + alpha * mu[j]) : mu[j] / (1.0 + alpha * mu[j]);
          z[j] = log(mu[j] / nfrow[j]) + (yrow[j] - mu[j]) / mu[j];
        }
        if (useQR) {
          int rows = m + p;
          for (int k = 0; k < p; k++) {
            for (int j = 0; j < m; j++) stack[j + (size_t)rows * k] = x[j + (size_t)m * k] * sqrt(wv[j]);
            for (int r = 0; r < p; r++) stack[m + r + (size_t)rows * k] = (r == k) ? sqrt(lambda[k]) : 0.0;
          }
          for (int j = 0; j < m; j++) rhs[j] = z[j] * sqrt(wv[j]);
          for (int r = 0; r < p; r++) rhs[m + r] = 0.0;
          householder_qr_apply(stack, rows, p, rhs);
          for (int c = p - 1; c >= 0; c--) { /* back substitution R beta = gamma */
            double v = rhs[c];
            for (int cc = c + 1; cc < p; cc++) v -= stack[c + (size_t)rows * cc] * beta[cc];
            beta[c] = v / stack[c + (size_t)rows * c];
          }
        } else {
          for (int a = 0; a < p; a++) {
            for (int b = 0; b < p; b++) {
              double s = 0.0;
              for (int j = 0; j < m; j++) s += x[j + (size_t)m * a] * wv[j] * x[j + (size_t)m * b];
              A[a + p * b] = s + ((a == b) ? lambda[a] : 0.0);
            }
            double s = 0.0;
            for (int j = 0; j < m; j++) s += x[j + (size_t)m * a] * z[j] * wv[j];
            rhs[a] = s;
          }
          if (!lu_solve(A, p, rhs, 1, work, piv))
            for (int k = 0; k < p; k++) rhs[k] = NAN;
          for (int k = 0; k < p; k++) beta[k] = rhs[k];
        }
        int big = 0;
        for (int k = 0; k < p; k++) big += (fabs(beta[k]) > large);
        if (big > 0) { it = maxit; break; }
        for (int j = 0; j < m; j++) {
          double eta = 0.0;
          for (int k = 0; k < p; k++) eta += x[j + (size_t)m * k] * beta[k];
          mu[j] = fmax(nfrow[j] * exp(eta), minmu);
        }
        dev = 0.0;
        for (int j = 0; j < m; j++) {
          double l = oracle_dnbinom_mu_log(yrow[j], 1.0 / alpha, mu[j]);
          dev = dev + -2.0 * (useWeights ? wrow[j] * l : l);
        }
        double conv_test = fabs(dev - dev_old) / (fabs(dev) + 0.1);
        if (isnan(conv_test)) { it = maxit; break; }
        if ((t > 0) & (conv_test < tol)) break;
        dev_old = dev;
      }
      deviance[i] = dev;
      iter[i] = it;
      for (int k = 0; k < p; k++) beta_mat[i + (size_t)n * k] = beta[k];
      /* post-loop block :429-455 */
      for (int j = 0; j < m; j++)
        wv[j] = useWeights ? wrow[j] * mu[j] / (1.0 + alpha * mu[j]) : mu[j] / (1.0 + alpha * mu[j]);
      for (int a = 0; a < p; a++)
        for (int b = 0; b < p; b++) {
          double s = 0.0;
          for (int j = 0; j < m; j++) s += x[j + (size_t)m * a] * wv[j] * x[j + (size_t)m * b];
          xtwx[a + p * b] = s;
          A[a + p * b] = s + ((a == b) ? lambda[a] : 0.0);
        }
      lu_inverse(A, p, Ainv, work, piv);
      for (int j = 0; j < m; j++) {
        double h = 0.0;
        for (int a = 0; a < p; a++)
          for (int b = 0; b < p; b++)
            h += (x[j + (size_t)m * a] * sqrt(wv[j])) * ((x[j + (size_t)m * b] * sqrt(wv[j])) * Ainv[b + p * a]);
        hat_diagonals[i + (size_t)n * j] = h;
      }
      mat_mul(Ainv, xtwx, T1, p);
      mat_mul(T1, Ainv, sigma, p);
      double cn = 0.0, cd = 0.0;
      for (int a = 0; a < p; a++) {
        cn += contrast[a] * beta[a];
        for (int b = 0; b < p; b++) cd += contrast[a] * sigma[a + p * b] * contrast[b];
        beta_var_mat[i + (size_t)n * a] = sigma[a + p * a];
      }
      contrast_num[i] = cn;
      contrast_denom[i] = sqrt(cd);
    }
    free(yrow); free(nfrow); free(wrow); free(mu); free(wv); free(z); free(beta); free(stack); free(rhs);
    free(A); free(xtwx); free(Ainv); free(work); free(T1); free(sigma); free(piv);
  }
  return 0;
}

/* scalar helpers exposed for unit tests of the posterior and its derivatives on one gene
 * (mirrors tests/testthat/test_dispersions.R:98-111 in the reference) */
double oracle_log_posterior_row(const double *yrow, const double *murow, const double *wrow, const double *x,
                                int m, int p, double log_alpha, double prior_mean, double prior_sigmasq,
                                int usePrior, int useWeights, double weightThreshold, int useCR, int deriv) {
  disp_ws *s = disp_ws_new(m, p, x);
  for (int j = 0; j < m; j++) { s->yrow[j] = yrow[j]; s->murow[j] = murow[j]; s->wrow[j] = wrow ? wrow[j] : 1.0; }
  double r;
  if (deriv == 0) r = log_posterior(s, log_alpha, prior_mean, prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
  else if (deriv == 1) r = dlog_posterior(s, log_alpha, prior_mean, prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
  else r = d2log_posterior(s, log_alpha, prior_mean, prior_sigmasq, usePrior, useWeights, weightThreshold, useCR);
  disp_ws_free(s);
  return r;
}
