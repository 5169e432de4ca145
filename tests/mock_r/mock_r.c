/*
 * mock_r.c -- heap implementation of the mock R API in Rinternals.h / R_ext/Rdynload.h plus the helpers
 * tests/test_r_shim.py needs to build arguments and to call a .Call entry point with R's error semantics
 * (Rf_error long-jumps out of the native routine; here back into mock_dot_call, which returns NULL).
 * TEST INFRASTRUCTURE ONLY.  Objects are never freed individually; mock_reset() drops everything.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

static struct mock_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL};
static struct mock_sexp names_sym = {SYMSXP, 0, NULL, NULL, NULL};
static struct mock_sexp dim_sym = {SYMSXP, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
int R_NaInt = INT_MIN;
double R_NaReal;

static int protect_depth = 0, protect_underflow = 0, n_alloc = 0;
static void **allocs = NULL;
static int allocs_cap = 0;
static jmp_buf *err_jmp = NULL;
static char err_msg[1024];

__attribute__((constructor)) static void mock_init(void) {
  /* R's NA_real_ is a NaN with payload 1954; any NaN is enough for the shim tests. */
  union { double d; unsigned long long u; } v;
  v.u = 0x7FF00000000007A2ULL;
  R_NaReal = v.d;
}

static void *track(void *p) {
  if (n_alloc == allocs_cap) {
    allocs_cap = allocs_cap ? 2 * allocs_cap : 256;
    allocs = (void **)realloc(allocs, sizeof(void *) * allocs_cap);
  }
  allocs[n_alloc++] = p;
  return p;
}

/* vectors whose data came from a custom allocator (Rf_allocVector3): released through it, like R's gc does */
#include <R_ext/Rallocators.h>
#define MOCK_HEADER 48   /* R places its vector header in front of the data inside the allocator's block */
static struct { R_allocator_t al; void *block; } custom[256];
static int n_custom = 0;
int mock_custom_allocations(void) { return n_custom; }

void mock_reset(void) {
  for (int i = 0; i < n_custom; i++) custom[i].al.mem_free(&custom[i].al, custom[i].block);
  n_custom = 0;
  for (int i = 0; i < n_alloc; i++) free(allocs[i]);
  n_alloc = 0;
  protect_depth = 0;
  protect_underflow = 0;
  err_msg[0] = 0;
}
static int n_interrupt_checks = 0;
void R_CheckUserInterrupt(void) { n_interrupt_checks++; }
int mock_interrupt_checks(void) { return n_interrupt_checks; }
int mock_protect_depth(void) { return protect_depth; }
int mock_protect_underflow(void) { return protect_underflow; }
const char *mock_last_error(void) { return err_msg; }

SEXP Rf_protect(SEXP s) { protect_depth++; return s; }
void Rf_unprotect(int n) {
  protect_depth -= n;
  if (protect_depth < 0) protect_underflow = 1;
}

void Rf_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_msg, sizeof err_msg, fmt, ap);
  va_end(ap);
  if (err_jmp) longjmp(*err_jmp, 1);
  fprintf(stderr, "mock R: error outside mock_dot_call: %s\n", err_msg);
  abort();
}

static size_t elt_size(SEXPTYPE t) {
  switch (t) {
    case LGLSXP: case INTSXP: return sizeof(int);
    case REALSXP: return sizeof(double);
    case STRSXP: case VECSXP: return sizeof(SEXP);
    case CHARSXP: return 1;
    default: Rf_error("mock R: allocVector of unsupported type %u", t);
  }
}

SEXP Rf_allocVector(SEXPTYPE t, long n) {
  if (n < 0) Rf_error("mock R: negative length");
  SEXP s = (SEXP)track(calloc(1, sizeof *s));
  s->type = t;
  s->length = n;
  s->data = track(calloc((size_t)n + 1, elt_size(t)));
  if (t == REALSXP) /* R leaves fresh vectors uninitialised: poison so that unwritten outputs show up */
    for (long i = 0; i < n; i++) ((double *)s->data)[i] = -7.7e77;
  if (t == INTSXP)
    for (long i = 0; i < n; i++) ((int *)s->data)[i] = -777777;
  if (t == STRSXP || t == VECSXP)
    for (long i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
  return s;
}

SEXP Rf_allocVector3(SEXPTYPE t, R_xlen_t n, R_allocator_t *al) {
  if (!al) return Rf_allocVector(t, n);
  if (t != REALSXP && t != INTSXP) Rf_error("mock R: allocVector3 of unsupported type %u", t);
  if (n_custom == 256) Rf_error("mock R: too many custom allocations");
  SEXP s = (SEXP)track(calloc(1, sizeof *s));
  s->type = t;
  s->length = n;
  char *block = (char *)al->mem_alloc(al, MOCK_HEADER + (size_t)(n + 1) * elt_size(t));
  if (!block) Rf_error("mock R: custom allocator returned NULL");
  custom[n_custom].al = *al;
  custom[n_custom].block = block;
  n_custom++;
  s->data = block + MOCK_HEADER;
  if (t == REALSXP)
    for (long i = 0; i < n; i++) ((double *)s->data)[i] = -7.7e77;
  else
    for (long i = 0; i < n; i++) ((int *)s->data)[i] = -777777;
  return s;
}

SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) {
  SEXP s = Rf_allocVector(t, (long)nr * nc);
  s->dim = Rf_allocVector(INTSXP, 2);
  INTEGER(s->dim)[0] = nr;
  INTEGER(s->dim)[1] = nc;
  return s;
}

SEXP Rf_mkChar(const char *c) {
  SEXP s = Rf_allocVector(CHARSXP, (long)strlen(c));
  memcpy(s->data, c, strlen(c) + 1);
  return s;
}

SEXP Rf_setAttrib(SEXP x, SEXP sym, SEXP v) {
  if (sym == R_NamesSymbol) {
    if (v->type != STRSXP || v->length != x->length) Rf_error("mock R: bad names attribute");
    x->names = v;
  } else if (sym == R_DimSymbol) {
    if (v->type != INTSXP || v->length != 2 || (long)INTEGER(v)[0] * INTEGER(v)[1] != x->length)
      Rf_error("mock R: dims do not match the length of object");
    x->dim = v;
  } else {
    Rf_error("mock R: unsupported attribute");
  }
  return v;
}

int TYPEOF(SEXP s) { return (int)s->type; }
int LENGTH(SEXP s) { return (int)s->length; }
int *INTEGER(SEXP s) {
  if (s->type != INTSXP && s->type != LGLSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type %u", s->type);
  return (int *)s->data;
}
int *LOGICAL(SEXP s) {
  if (s->type != LGLSXP) Rf_error("LOGICAL() can only be applied to a 'logical', not a type %u", s->type);
  return (int *)s->data;
}
double *REAL(SEXP s) {
  if (s->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type %u", s->type);
  return (double *)s->data;
}
SEXP VECTOR_ELT(SEXP s, long i) {
  if (s->type != VECSXP || i < 0 || i >= s->length) Rf_error("mock R: bad VECTOR_ELT");
  return ((SEXP *)s->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP s, long i, SEXP v) {
  if (s->type != VECSXP || i < 0 || i >= s->length) Rf_error("mock R: bad SET_VECTOR_ELT");
  ((SEXP *)s->data)[i] = v;
  return v;
}
SEXP STRING_ELT(SEXP s, long i) {
  if (s->type != STRSXP || i < 0 || i >= s->length) Rf_error("mock R: bad STRING_ELT");
  return ((SEXP *)s->data)[i];
}
void SET_STRING_ELT(SEXP s, long i, SEXP v) {
  if (s->type != STRSXP || i < 0 || i >= s->length || v->type != CHARSXP) Rf_error("mock R: bad SET_STRING_ELT");
  ((SEXP *)s->data)[i] = v;
}
const char *R_CHAR(SEXP s) {
  if (s->type != CHARSXP) Rf_error("mock R: CHAR() of a non-CHARSXP");
  return (const char *)s->data;
}

int Rf_nrows(SEXP s) {
  if (s->dim) return INTEGER(s->dim)[0];
  if (s->type == LGLSXP || s->type == INTSXP || s->type == REALSXP || s->type == VECSXP) return (int)s->length;
  Rf_error("object is not a matrix");
}
int Rf_ncols(SEXP s) {
  if (s->dim) return INTEGER(s->dim)[1];
  if (s->type == LGLSXP || s->type == INTSXP || s->type == REALSXP || s->type == VECSXP) return 1;
  Rf_error("object is not a matrix");
}

SEXP Rf_coerceVector(SEXP s, SEXPTYPE t) {
  if (s->type == t) return s; /* R returns the object itself: no copy, callers must not write to it */
  if (s->type == NILSXP) return Rf_allocVector(t, 0);
  if (t != REALSXP || (s->type != INTSXP && s->type != LGLSXP)) Rf_error("mock R: unsupported coercion %u -> %u", s->type, t);
  SEXP r = Rf_allocVector(REALSXP, s->length);
  for (long i = 0; i < s->length; i++) {
    int v = ((int *)s->data)[i];
    ((double *)r->data)[i] = v == R_NaInt ? R_NaReal : (double)v;
  }
  r->dim = s->dim;
  r->names = s->names;
  return r;
}

int Rf_asInteger(SEXP s) {
  if (s->length < 1) return R_NaInt;
  if (s->type == INTSXP || s->type == LGLSXP) return ((int *)s->data)[0];
  if (s->type == REALSXP) {
    double d = ((double *)s->data)[0];
    if (isnan(d) || d >= 2147483648.0 || d <= -2147483649.0) return R_NaInt;
    return (int)d;
  }
  return R_NaInt;
}
int Rf_asLogical(SEXP s) {
  if (s->length < 1) return R_NaInt;
  if (s->type == LGLSXP) return ((int *)s->data)[0];
  if (s->type == INTSXP) { int v = ((int *)s->data)[0]; return v == R_NaInt ? R_NaInt : v != 0; }
  if (s->type == REALSXP) { double d = ((double *)s->data)[0]; return isnan(d) ? R_NaInt : d != 0; }
  return R_NaInt;
}
double Rf_asReal(SEXP s) {
  if (s->length < 1) return R_NaReal;
  if (s->type == REALSXP) return ((double *)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) { int v = ((int *)s->data)[0]; return v == R_NaInt ? R_NaReal : (double)v; }
  return R_NaReal;
}

int R_registerRoutines(DllInfo *dll, const void *c_methods, const R_CallMethodDef *call_methods, const void *f_methods,
                       const void *ext_methods) {
  (void)c_methods; (void)f_methods; (void)ext_methods;
  dll->call_methods = call_methods;
  dll->n_call_methods = 0;
  if (call_methods)
    while (call_methods[dll->n_call_methods].name) dll->n_call_methods++;
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *dll, Rboolean v) {
  Rboolean old = dll->dynamic_symbols;
  dll->dynamic_symbols = v;
  return old;
}

/* ------------------------------------------------------------------ helpers for the Python side */
SEXP mock_real(const double *src, long n, int nrow, int ncol) {
  SEXP s = nrow >= 0 ? Rf_allocMatrix(REALSXP, nrow, ncol) : Rf_allocVector(REALSXP, n);
  memcpy(s->data, src, sizeof(double) * (size_t)s->length);
  return s;
}
SEXP mock_int(const int *src, long n, int nrow, int ncol, int logical) {
  SEXP s = nrow >= 0 ? Rf_allocMatrix(INTSXP, nrow, ncol) : Rf_allocVector(INTSXP, n);
  memcpy(s->data, src, sizeof(int) * (size_t)s->length);
  if (logical) s->type = LGLSXP;
  return s;
}
SEXP mock_string(const char *c) {
  SEXP s = Rf_allocVector(STRSXP, 1);
  SET_STRING_ELT(s, 0, Rf_mkChar(c));
  return s;
}
SEXP mock_nil(void) { return R_NilValue; }
int mock_typeof(SEXP s) { return (int)s->type; }
long mock_length(SEXP s) { return s->length; }
void *mock_data(SEXP s) { return s->data; }
int mock_has_dim(SEXP s) { return s->dim != NULL; }
int mock_dim(SEXP s, int k) { return s->dim ? INTEGER(s->dim)[k] : -1; }
SEXP mock_list_elt(SEXP s, long i) { return ((SEXP *)s->data)[i]; }
const char *mock_name(SEXP s, long i) {
  if (!s->names) return NULL;
  return (const char *)((SEXP *)s->names->data)[i]->data;
}

/* .Call(): invoke a native routine of `nargs` SEXP arguments; an Rf_error inside it returns NULL with the
 * message in mock_last_error() (R would unwind to the top level and reset the protect stack). */
typedef SEXP (*fn11)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn13)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn15)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

SEXP mock_dot_call(void *fn, int nargs, SEXP *a) {
  jmp_buf jb;
  SEXP volatile res = NULL;
  err_msg[0] = 0;
  err_jmp = &jb;
  if (setjmp(jb) == 0) {
    if (nargs == 11) res = ((fn11)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
    else if (nargs == 13) res = ((fn13)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
    else if (nargs == 15) res = ((fn15)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14]);
    else snprintf(err_msg, sizeof err_msg, "mock R: unsupported .Call arity %d", nargs);
  } else {
    res = NULL;
    protect_depth = 0; /* R resets the pointer-protection stack when unwinding to top level */
  }
  err_jmp = NULL;
  return res;
}

/* Runs R_init_<pkg> and reports what it registered. */
static DllInfo the_dll = {NULL, 0, -1};
int mock_run_init(void (*init)(DllInfo *)) {
  the_dll.call_methods = NULL;
  the_dll.n_call_methods = 0;
  the_dll.dynamic_symbols = -1;
  init(&the_dll);
  return the_dll.n_call_methods;
}
const char *mock_registered_name(int i) { return the_dll.call_methods[i].name; }
int mock_registered_nargs(int i) { return the_dll.call_methods[i].numArgs; }
void *mock_registered_fun(int i) { return (void *)the_dll.call_methods[i].fun; }
int mock_dynamic_symbols(void) { return the_dll.dynamic_symbols; }
