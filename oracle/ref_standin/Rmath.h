/* Rmath.h (STAND-IN, test infrastructure -- see RcppArmadillo.h in this directory).
 * The four nmath entry points src/DESeq2.cpp calls plus R_pow_di, implemented in rmath_standin.c from the published
 * algorithms of R's nmath (R is not in this image; DESeq2's DESCRIPTION pins no R version). */
#ifndef STANDIN_RMATH_H
#define STANDIN_RMATH_H
#ifdef __cplusplus
extern "C" {
#endif
double Rf_lgammafn(double x);
double Rf_digamma(double x);
double Rf_trigamma(double x);
double Rf_dnbinom_mu(double x, double size, double mu, int give_log);
double R_pow_di(double x, int n);
#ifdef __cplusplus
}
#endif
#endif
