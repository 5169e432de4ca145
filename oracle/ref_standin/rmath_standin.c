/*
 * rmath_standin.c -- TEST INFRASTRUCTURE ONLY (parity checker, see RcppArmadillo.h in this directory).
 *
 * Stand-in for the five entry points of R's nmath / arithmetic that /root/reference/src/DESeq2.cpp calls
 * (Rf_lgammafn :50, Rf_digamma :90, Rf_trigamma :139, Rf_dnbinom_mu :369, R_pow_di :48).  R's sources are not part
 * of /root/reference and R is not in this image; DESeq2's DESCRIPTION pins no R version.  The algorithms below are the
 * published ones:
 *   dnbinom_mu   R nmath/dnbinom.c (R >= 2.x ... 4.3): x == 0 closed form, x < 1e-10 size series form, otherwise
 *                size/(size+x) * dbinom_raw(size, x+size, size/(size+mu), mu/(size+mu)); size = Inf -> Poisson
 *   dbinom_raw   Catherine Loader (2000), "Fast and accurate computation of binomial probabilities": saddle-point
 *                expansion with stirlerr() and the deviance part bd0()
 *   stirlerr     exact table for 2n integer, n <= 15 (regenerated with mpmath: ../sferr_halves.h), asymptotic series else
 *   lgammafn     log|Gamma(x)|  -> glibc lgamma_r (agrees with R's gammafn/lgammacor construction to ~1 ulp)
 *   digamma, trigamma  (R: Amos' dpsifn) -> independent formulation: upward recurrence to x >= 16, then the Bernoulli
 *                asymptotic series with enough terms for < 1e-16 relative truncation error
 *   R_pow_di     R arithmetic.c: repeated squaring on |n|, reciprocal of the RESULT for negative n
 * Deliberately written separately from the oracle's own nmath restatement (oracle/nbglm_oracle.c) so that
 * "oracle == _ref" also cross-checks the two.
 */
#include <float.h>
#include <math.h>

#include "../sferr_halves.h"

#define LN_SQRT_2PI 0.918938533204672741780329736406
#define LN_2PI 1.837877066409345483560659472811

double Rf_lgammafn(double x) {
  int sign;
  return lgamma_r(x, &sign);
}

double R_pow_di(double x, int n) {
  double xn = 1.0;
  if (isnan(x)) return x;
  if (n != 0) {
    if (!isfinite(x)) return pow(x, (double)n);
    int is_neg = n < 0;
    if (is_neg) n = -n;
    for (;;) {
      if (n & 1) xn *= x;
      if (n >>= 1) x *= x; else break;
    }
    if (is_neg) xn = 1.0 / xn;
  }
  return xn;
}

/* B_{2k} for k = 1..10 */
static const double BERN[10] = {1.0 / 6, -1.0 / 30, 1.0 / 42, -1.0 / 30, 5.0 / 66, -691.0 / 2730, 7.0 / 6,
                                -3617.0 / 510, 43867.0 / 798, -174611.0 / 330};

double Rf_digamma(double x) {
  if (isnan(x)) return x;
  if (x <= 0.0) { /* reflection; poles at the non-positive integers */
    if (x == floor(x)) return NAN;
    return Rf_digamma(1.0 - x) - M_PI / tan(M_PI * x);
  }
  double shift = 0.0;
  while (x < 16.0) { shift += 1.0 / x; x += 1.0; }
  double r = 1.0 / x, r2 = r * r, term = r2, tail = 0.0;
  for (int k = 0; k < 10; k++) { /* sum B_2k / (2k x^2k) */
    tail += BERN[k] / (2.0 * (k + 1)) * term;
    term *= r2;
  }
  return log(x) - 0.5 * r - tail - shift;
}

double Rf_trigamma(double x) {
  if (isnan(x)) return x;
  if (x <= 0.0) {
    if (x == floor(x)) return NAN;
    double s = sin(M_PI * x);
    return -Rf_trigamma(1.0 - x) + (M_PI / s) * (M_PI / s);
  }
  double shift = 0.0;
  while (x < 16.0) { shift += 1.0 / (x * x); x += 1.0; }
  double r = 1.0 / x, r2 = r * r, term = r * r2, tail = 0.0;
  for (int k = 0; k < 10; k++) { /* sum B_2k / x^(2k+1) */
    tail += BERN[k] * term;
    term *= r2;
  }
  return r + 0.5 * r2 + tail + shift;
}

static double stirlerr(double n) {
  static const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778, S2 = 0.00079365079365079365079365,
                      S3 = 0.000595238095238095238095238, S4 = 0.0008417508417508417508417508;
  if (n <= 15.0) {
    double nn = n + n;
    if (nn == (int)nn) return sferr_halves[(int)nn];
    return Rf_lgammafn(n + 1.0) - (n + 0.5) * log(n) + n - LN_SQRT_2PI;
  }
  double nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

static double bd0(double x, double np) {
  if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np), s = (x - np) * v;
    if (fabs(s) < DBL_MIN) return s;
    double ej = 2 * x * v;
    v *= v;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / ((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / np) + np - x;
}

static double dbinom_raw(double x, double n, double p, double q, int give_log) {
  double lc, lf;
  if (p == 0) return (x == 0) ? (give_log ? 0.0 : 1.0) : (give_log ? -INFINITY : 0.0);
  if (q == 0) return (x == n) ? (give_log ? 0.0 : 1.0) : (give_log ? -INFINITY : 0.0);
  if (x == 0) {
    if (n == 0) return give_log ? 0.0 : 1.0;
    lc = (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
    return give_log ? lc : exp(lc);
  }
  if (x == n) {
    lc = (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
    return give_log ? lc : exp(lc);
  }
  if (x < 0 || x > n) return give_log ? -INFINITY : 0.0;
  lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
  lf = LN_2PI + log(x) + log1p(-x / n);
  return give_log ? lc - 0.5 * lf : exp(lc - 0.5 * lf);
}

static double dpois_raw(double x, double lambda, int give_log) {
  if (lambda == 0) return (x == 0) ? (give_log ? 0.0 : 1.0) : (give_log ? -INFINITY : 0.0);
  if (!isfinite(lambda)) return give_log ? -INFINITY : 0.0;
  if (x < 0) return give_log ? -INFINITY : 0.0;
  if (x <= lambda * DBL_MIN) return give_log ? -lambda : exp(-lambda);
  if (lambda < x * DBL_MIN) {
    if (!isfinite(x)) return give_log ? -INFINITY : 0.0;
    double v = -lambda + x * log(lambda) - Rf_lgammafn(x + 1);
    return give_log ? v : exp(v);
  }
  double v = -0.5 * log(2 * M_PI * x) + (-stirlerr(x) - bd0(x, lambda));
  return give_log ? v : exp(v);
}

double Rf_dnbinom_mu(double x, double size, double mu, int give_log) {
  if (isnan(x) || isnan(size) || isnan(mu)) return x + size + mu;
  if (mu < 0 || size < 0) return NAN;
  if (fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x))) return give_log ? -INFINITY : 0.0; /* non-integer x */
  if (x < 0 || !isfinite(x)) return give_log ? -INFINITY : 0.0;
  if (x == 0 && size == 0) return give_log ? 0.0 : 1.0;
  x = nearbyint(x);
  if (!isfinite(size)) return dpois_raw(x, mu, give_log);
  if (x == 0) {
    double v = size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
    return give_log ? v : exp(v);
  }
  if (x < 1e-10 * size) {
    double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
    double v = x * p - mu - Rf_lgammafn(x + 1) + log1p(x * (x - 1) / (2 * size));
    return give_log ? v : exp(v);
  }
  double p = size / (size + x), ans = dbinom_raw(size, x + size, size / (size + mu), mu / (size + mu), give_log);
  return give_log ? log(p) + ans : p * ans;
}
