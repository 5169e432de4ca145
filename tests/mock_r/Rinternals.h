/*
 * Rinternals.h (MOCK) -- the slice of R's C API that deseq2_b200/r_shim/deseq2_b200_shim.c uses, implemented by
 * tests/mock_r/mock_r.c on plain heap objects.  TEST INFRASTRUCTURE: R is not installed in this image, so the
 * shim cannot be loaded by a real R; this mock lets tests/test_r_shim.py compile the shim unchanged and drive
 * its .Call entry points with R-shaped arguments (column-major matrices with a dim attribute, length-1 scalars,
 * INTSXP / REALSXP / LGLSXP) and inspect the named list it returns.  Names and semantics follow "Writing R
 * Extensions" section 5.9; nothing here comes from /root/reference.
 */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define SYMSXP 1
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19

typedef struct mock_sexp {
  SEXPTYPE type;
  long length;
  void *data;              /* int[] (LGLSXP, INTSXP), double[] (REALSXP), SEXP[] (STRSXP, VECSXP), char[] (CHARSXP) */
  struct mock_sexp *names; /* STRSXP or NULL */
  struct mock_sexp *dim;   /* INTSXP of length 2 or NULL */
} *SEXP;

extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern int R_NaInt;
extern double R_NaReal;
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
#define NA_REAL R_NaReal

SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)

typedef long R_xlen_t;
SEXP Rf_allocVector(SEXPTYPE, long);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_mkChar(const char *);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asLogical(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));

int TYPEOF(SEXP);
int LENGTH(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
double *REAL(SEXP);
SEXP VECTOR_ELT(SEXP, long);
SEXP SET_VECTOR_ELT(SEXP, long, SEXP);
void SET_STRING_ELT(SEXP, long, SEXP);
SEXP STRING_ELT(SEXP, long);
const char *R_CHAR(SEXP);
#define CHAR(x) R_CHAR(x)

#ifdef __cplusplus
}
#endif
#endif
