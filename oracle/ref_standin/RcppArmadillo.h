/*
 * RcppArmadillo.h (STAND-IN) -- TEST INFRASTRUCTURE ONLY, part of the parity checker under oracle/.
 *
 * Purpose: let g++ compile /root/reference/src/DESeq2.cpp UNCHANGED (where it lies, never copied) into
 * oracle/_ref/libdeseq2_ref.so although R, Rcpp, RcppArmadillo, Armadillo, LAPACK and libRmath are absent from this
 * image.  The reference translation unit uses a small closed subset of those libraries; this header (plus R.h, Rmath.h,
 * R_ext/Utils.h next to it and rmath_standin.c) provides exactly that subset with the documented semantics:
 *
 *   Rcpp      SEXP (plain heap object here), NumericVector / IntegerVector / NumericMatrix (reference-semantics
 *             wrappers that share storage with a REALSXP and coerce an INTSXP by copy), NumericMatrix::Row,
 *             clone, as<T>, List::create / Named, checkUserInterrupt, `_`, and the sugar used by the posterior
 *             functions: element-wise + - * / between vectors and scalars, pow, log, lgamma, digamma, trigamma, sum.
 *             Sugar is evaluated EAGERLY here; per element the same IEEE operations run in the same order as Rcpp's
 *             lazy expression templates, and sum() accumulates sequentially in double exactly like
 *             Rcpp::sugar::Sum, so results are bit-identical to a lazy evaluation.
 *             pow(v, n) is std::pow(double, double) as in Rcpp/sugar/functions/pow.h; lgamma / digamma / trigamma
 *             map to Rf_lgammafn / Rf_digamma / Rf_trigamma as in Rcpp/sugar/functions/math.h.
 *   Armadillo mat / vec / colvec (ONE dynamic column-major class here; the reference never relies on the static
 *             row/column distinction except where noted at sum()), umat / uvec, .t() .i() .rows() .cols() .row()
 *             .each_col() .max(idx) .n_rows .n_cols .n_elem, element-wise % / + - and scalar forms, * (matrix
 *             product), abs sqrt exp log sum find det trace diagmat diagvec join_cols zeros ones linspace span,
 *             qr_econ (Householder, thin Q), solve (partially pivoted LU; if rcond < DBL_EPSILON the minimum-norm
 *             least-squares solution via a one-sided Jacobi SVD, as Armadillo's default solve() does after its
 *             "system is singular; attempting approx solution" warning), inv (Gauss-Jordan with partial pivoting,
 *             throws std::runtime_error("inv(): matrix is singular") like Armadillo), det (LU).
 *             Any backward-stable p x p algorithm agrees with LAPACK's to << 1e-6 on these p <= 32 systems.
 *
 * Nothing in this directory is derived from the reference's sources; it restates public library interfaces.
 * Nothing under deseq2_b200/ may include or link it.
 */
#ifndef STANDIN_RCPPARMADILLO_H
#define STANDIN_RCPPARMADILLO_H

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

/* ------------------------------------------------------------------------------------------------ SEXP */
#define STANDIN_LGLSXP 10
#define STANDIN_INTSXP 13
#define STANDIN_REALSXP 14

struct standin_sexp {
  int type = STANDIN_REALSXP;
  long length = 0;
  int nrow = -1, ncol = -1; /* dim attribute, -1 = none */
  std::shared_ptr<std::vector<double>> real;
  std::shared_ptr<std::vector<int>> ints;
};
typedef standin_sexp *SEXP;

/* objects created during one .Call (clone(), coercions, results) live in a per-thread arena that the C wrapper
 * (ref_capi.cpp) releases after copying the results out -- the stand-in for R's garbage collector */
std::vector<std::unique_ptr<standin_sexp>> &standin_arena();
inline SEXP standin_new(int type, long length, int nrow = -1, int ncol = -1) {
  std::unique_ptr<standin_sexp> s(new standin_sexp);
  s->type = type; s->length = length; s->nrow = nrow; s->ncol = ncol;
  if (type == STANDIN_REALSXP) s->real = std::make_shared<std::vector<double>>(length, 0.0);
  else s->ints = std::make_shared<std::vector<int>>(length, 0);
  SEXP r = s.get();
  standin_arena().push_back(std::move(s));
  return r;
}
inline std::shared_ptr<std::vector<double>> standin_as_real(SEXP s) { /* REALSXP: shared; others: coerced copy */
  if (s->type == STANDIN_REALSXP) return s->real;
  auto v = std::make_shared<std::vector<double>>(s->length);
  for (long i = 0; i < s->length; i++) (*v)[i] = (double)(*s->ints)[i];
  return v;
}

/* ------------------------------------------------------------------------------------------------ arma */
namespace Rcpp { struct SugarVec; class NumericVector; }

namespace arma {
typedef unsigned long long uword;

struct span { uword a, b; span(uword a_, uword b_) : a(a_), b(b_) {} };

class umat { /* result of comparisons / find(): column-major uword matrix */
 public:
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<uword> mem;
  umat() {}
  umat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, 0) {}
  uword operator()(uword i) const { return mem[i]; }
};
typedef umat uvec;

class mat;
struct subview_row;
struct subview_span;
struct each_col_proxy;

class mat {
 public:
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<double> mem;
  mat() {}
  mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, 0.0) {}
  mat(const Rcpp::SugarVec &s);        /* RcppArmadillo: arma::vec from a sugar expression */
  mat(const subview_row &r);
  double &operator()(uword i) { return mem[i]; }
  double operator()(uword i) const { return mem[i]; }
  double &operator()(uword r, uword c) { return mem[r + c * n_rows]; }
  double operator()(uword r, uword c) const { return mem[r + c * n_rows]; }
  mat operator()(const uvec &idx) const { /* elem subview */
    mat o(idx.n_elem, 1);
    for (uword k = 0; k < idx.n_elem; k++) o.mem[k] = mem[idx.mem[k]];
    return o;
  }
  subview_span operator()(const span &s);
  mat t() const {
    mat o(n_cols, n_rows);
    for (uword c = 0; c < n_cols; c++)
      for (uword r = 0; r < n_rows; r++) o.mem[c + r * n_cols] = mem[r + c * n_rows];
    return o;
  }
  mat i() const;
  mat rows(const uvec &idx) const {
    mat o(idx.n_elem, n_cols);
    for (uword c = 0; c < n_cols; c++)
      for (uword k = 0; k < idx.n_elem; k++) o(k, c) = (*this)(idx.mem[k], c);
    return o;
  }
  mat cols(const uvec &idx) const {
    mat o(n_rows, idx.n_elem);
    for (uword k = 0; k < idx.n_elem; k++)
      for (uword r = 0; r < n_rows; r++) o(r, k) = (*this)(r, idx.mem[k]);
    return o;
  }
  subview_row row(uword r);
  mat row_copy(uword r) const {
    mat o(1, n_cols);
    for (uword c = 0; c < n_cols; c++) o.mem[c] = (*this)(r, c);
    return o;
  }
  each_col_proxy each_col() const;
  double max(uword &idx) const { /* first maximum; NaN never wins (Armadillo's op_max::direct_max) */
    idx = 0;
    double best = -INFINITY;
    for (uword k = 0; k < n_elem; k++)
      if (mem[k] > best) { best = mem[k]; idx = k; }
    return best;
  }
};
typedef mat vec;
typedef mat colvec;
typedef mat rowvec;

struct subview_row {
  mat *m;
  uword r;
  mat t() const {
    mat o(m->n_cols, 1);
    for (uword c = 0; c < m->n_cols; c++) o.mem[c] = (*m)(r, c);
    return o;
  }
  subview_row &operator=(const mat &v) {
    if (v.n_elem != m->n_cols) throw std::logic_error("copy into submatrix: incompatible matrix dimensions");
    for (uword c = 0; c < m->n_cols; c++) (*m)(r, c) = v.mem[c];
    return *this;
  }
};
inline mat::mat(const subview_row &s) : n_rows(1), n_cols(s.m->n_cols), n_elem(s.m->n_cols), mem(s.m->n_cols) {
  for (uword c = 0; c < n_cols; c++) mem[c] = (*s.m)(s.r, c);
}
inline subview_row mat::row(uword r) { return subview_row{this, r}; }

struct subview_span {
  mat *m;
  uword a, b;
  subview_span &operator=(const mat &v) {
    if (v.n_elem != b - a + 1) throw std::logic_error("copy into subvector: incompatible dimensions");
    for (uword k = a; k <= b; k++) m->mem[k] = v.mem[k - a];
    return *this;
  }
};
inline subview_span mat::operator()(const span &s) { return subview_span{this, s.a, s.b}; }

/* element-wise helpers ---------------------------------------------------------------------------- */
template <class F> inline mat ew2(const mat &a, const mat &b, F f, const char *what) {
  if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::logic_error(std::string(what) + ": incompatible matrix dimensions");
  mat o(a.n_rows, a.n_cols);
  for (uword k = 0; k < a.n_elem; k++) o.mem[k] = f(a.mem[k], b.mem[k]);
  return o;
}
template <class F> inline mat ew1(const mat &a, F f) {
  mat o(a.n_rows, a.n_cols);
  for (uword k = 0; k < a.n_elem; k++) o.mem[k] = f(a.mem[k]);
  return o;
}
inline mat operator%(const mat &a, const mat &b) { return ew2(a, b, [](double x, double y) { return x * y; }, "element-wise multiplication"); }
inline mat operator/(const mat &a, const mat &b) { return ew2(a, b, [](double x, double y) { return x / y; }, "element-wise division"); }
inline mat operator+(const mat &a, const mat &b) { return ew2(a, b, [](double x, double y) { return x + y; }, "addition"); }
inline mat operator-(const mat &a, const mat &b) { return ew2(a, b, [](double x, double y) { return x - y; }, "subtraction"); }
inline mat operator+(double s, const mat &a) { return ew1(a, [s](double x) { return s + x; }); }
inline mat operator+(const mat &a, double s) { return ew1(a, [s](double x) { return x + s; }); }
inline mat operator-(const mat &a, double s) { return ew1(a, [s](double x) { return x - s; }); }
inline mat operator*(double s, const mat &a) { return ew1(a, [s](double x) { return s * x; }); }
inline mat operator*(const mat &a, double s) { return ew1(a, [s](double x) { return x * s; }); }
inline mat operator/(const mat &a, double s) { return ew1(a, [s](double x) { return x / s; }); }
inline mat operator*(const mat &a, const mat &b) {
  if (a.n_cols != b.n_rows) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
  mat o(a.n_rows, b.n_cols);
  for (uword c = 0; c < b.n_cols; c++)
    for (uword r = 0; r < a.n_rows; r++) {
      double s = 0.0;
      for (uword k = 0; k < a.n_cols; k++) s += a(r, k) * b(k, c);
      o(r, c) = s;
    }
  return o;
}
inline mat abs(const mat &a) { return ew1(a, [](double x) { return std::fabs(x); }); }
inline mat sqrt(const mat &a) { return ew1(a, [](double x) { return std::sqrt(x); }); }
inline mat exp(const mat &a) { return ew1(a, [](double x) { return std::exp(x); }); }
inline mat log(const mat &a) { return ew1(a, [](double x) { return std::log(x); }); }

inline umat operator>(const mat &a, double s) {
  umat o(a.n_rows, a.n_cols);
  for (uword k = 0; k < a.n_elem; k++) o.mem[k] = a.mem[k] > s ? 1 : 0;
  return o;
}
inline uvec find(const umat &a) {
  uvec o;
  for (uword k = 0; k < a.n_elem; k++)
    if (a.mem[k]) o.mem.push_back(k);
  o.n_elem = o.n_rows = o.mem.size();
  o.n_cols = 1;
  return o;
}
/* sum(mat): column sums as a 1 x n_cols row (Armadillo's dim = 0 default for a Mat; the reference applies it to the
 * design matrix only).  sum(umat) is applied by the reference to the comparison of a column VECTOR, for which
 * Armadillo returns the scalar total. */
inline mat sum(const mat &a) {
  mat o(1, a.n_cols);
  for (uword c = 0; c < a.n_cols; c++) {
    double s = 0.0;
    for (uword r = 0; r < a.n_rows; r++) s += a(r, c);
    o.mem[c] = s;
  }
  return o;
}
inline uword sum(const umat &a) {
  uword s = 0;
  for (uword k = 0; k < a.n_elem; k++) s += a.mem[k];
  return s;
}

struct each_col_proxy {
  const mat *m;
  mat operator%(const mat &v) const {
    if (v.n_elem != m->n_rows) throw std::logic_error("each_col(): incompatible size");
    mat o(m->n_rows, m->n_cols);
    for (uword c = 0; c < m->n_cols; c++)
      for (uword r = 0; r < m->n_rows; r++) o(r, c) = (*m)(r, c) * v.mem[r];
    return o;
  }
};
inline each_col_proxy mat::each_col() const { return each_col_proxy{this}; }

inline mat zeros(uword n) { return mat(n, 1); }
inline mat zeros(uword r, uword c) { return mat(r, c); }
inline mat ones(uword n) {
  mat o(n, 1);
  std::fill(o.mem.begin(), o.mem.end(), 1.0);
  return o;
}
template <class T> inline T linspace(double a, double b, uword n) {
  T o(n, 1);
  if (n == 1) { o.mem[0] = b; return o; }
  double delta = (b - a) / double(n - 1);
  for (uword k = 0; k + 1 < n; k++) o.mem[k] = a + double(k) * delta;
  o.mem[n - 1] = b;
  return o;
}
inline mat diagmat(const mat &v) { /* the reference passes a vector (lambda) */
  mat o(v.n_elem, v.n_elem);
  for (uword k = 0; k < v.n_elem; k++) o(k, k) = v.mem[k];
  return o;
}
inline mat diagvec(const mat &a) {
  uword n = std::min(a.n_rows, a.n_cols);
  mat o(n, 1);
  for (uword k = 0; k < n; k++) o.mem[k] = a(k, k);
  return o;
}
inline mat join_cols(const mat &a, const mat &b) {
  if (a.n_cols != b.n_cols) throw std::logic_error("join_cols(): number of columns must be the same");
  mat o(a.n_rows + b.n_rows, a.n_cols);
  for (uword c = 0; c < a.n_cols; c++) {
    for (uword r = 0; r < a.n_rows; r++) o(r, c) = a(r, c);
    for (uword r = 0; r < b.n_rows; r++) o(a.n_rows + r, c) = b(r, c);
  }
  return o;
}
inline double trace(const mat &a) {
  double s = 0.0;
  for (uword k = 0; k < std::min(a.n_rows, a.n_cols); k++) s += a(k, k);
  return s;
}

/* partially pivoted LU in place; returns sign of the permutation, 0 if exactly singular */
inline int lu_inplace(mat &a, std::vector<uword> &piv) {
  uword n = a.n_rows;
  int sign = 1;
  piv.resize(n);
  for (uword k = 0; k < n; k++) {
    uword pr = k;
    double best = std::fabs(a(k, k));
    for (uword r = k + 1; r < n; r++)
      if (std::fabs(a(r, k)) > best) { best = std::fabs(a(r, k)); pr = r; }
    piv[k] = pr;
    if (!(best > 0.0)) return 0; /* zero or NaN pivot */
    if (pr != k) {
      sign = -sign;
      for (uword c = 0; c < n; c++) std::swap(a(k, c), a(pr, c));
    }
    for (uword r = k + 1; r < n; r++) {
      double f = a(r, k) / a(k, k);
      a(r, k) = f;
      for (uword c = k + 1; c < n; c++) a(r, c) -= f * a(k, c);
    }
  }
  return sign;
}
inline double det(const mat &a0) {
  if (a0.n_rows != a0.n_cols) throw std::logic_error("det(): given matrix must be square sized");
  mat a = a0;
  std::vector<uword> piv;
  for (uword k = 0; k < a.n_elem; k++)
    if (a.mem[k] != a.mem[k]) return NAN;
  int sign = lu_inplace(a, piv);
  if (sign == 0) return 0.0;
  double d = double(sign);
  for (uword k = 0; k < a.n_rows; k++) d *= a(k, k);
  return d;
}
inline bool inv_gj(const mat &a0, mat &out) {
  uword n = a0.n_rows;
  mat a = a0;
  out = mat(n, n);
  for (uword k = 0; k < n; k++) out(k, k) = 1.0;
  for (uword k = 0; k < n; k++) {
    uword pr = k;
    double best = std::fabs(a(k, k));
    for (uword r = k + 1; r < n; r++)
      if (std::fabs(a(r, k)) > best) { best = std::fabs(a(r, k)); pr = r; }
    if (!(best > 0.0)) return false;
    if (pr != k)
      for (uword c = 0; c < n; c++) { std::swap(a(k, c), a(pr, c)); std::swap(out(k, c), out(pr, c)); }
    double d = a(k, k);
    for (uword c = 0; c < n; c++) { a(k, c) /= d; out(k, c) /= d; }
    for (uword r = 0; r < n; r++)
      if (r != k) {
        double f = a(r, k);
        if (f != 0.0)
          for (uword c = 0; c < n; c++) { a(r, c) -= f * a(k, c); out(r, c) -= f * out(k, c); }
      }
  }
  return true;
}
inline mat mat::i() const {
  if (n_rows != n_cols) throw std::logic_error("inv(): given matrix must be square sized");
  mat o;
  if (!inv_gj(*this, o)) throw std::runtime_error("inv(): matrix is singular");
  return o;
}
inline double norm1(const mat &a) {
  double best = 0.0;
  for (uword c = 0; c < a.n_cols; c++) {
    double s = 0.0;
    for (uword r = 0; r < a.n_rows; r++) s += std::fabs(a(r, c));
    if (s > best || s != s) best = s;
  }
  return best;
}
/* minimum-norm least-squares solution through a one-sided Jacobi SVD (Armadillo's approx-solve fallback is
 * LAPACK dgelsd; any SVD gives the same minimum-norm solution up to rounding) */
inline mat solve_minnorm(const mat &a, const mat &b) {
  uword n = a.n_rows, p = a.n_cols;
  mat u = a, v(p, p);
  for (uword k = 0; k < p; k++) v(k, k) = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0;
    for (uword i = 0; i + 1 < p; i++)
      for (uword j = i + 1; j < p; j++) {
        double al = 0, be = 0, ga = 0;
        for (uword r = 0; r < n; r++) { al += u(r, i) * u(r, i); be += u(r, j) * u(r, j); ga += u(r, i) * u(r, j); }
        if (ga == 0.0) continue;
        off = std::max(off, std::fabs(ga) / std::sqrt(al * be + DBL_MIN));
        double zeta = (be - al) / (2.0 * ga);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (uword r = 0; r < n; r++) { double x = u(r, i), y = u(r, j); u(r, i) = c * x - s * y; u(r, j) = s * x + c * y; }
        for (uword r = 0; r < p; r++) { double x = v(r, i), y = v(r, j); v(r, i) = c * x - s * y; v(r, j) = s * x + c * y; }
      }
    if (off < 1e-15) break;
  }
  std::vector<double> sv(p);
  double smax = 0.0;
  for (uword k = 0; k < p; k++) {
    double s = 0.0;
    for (uword r = 0; r < n; r++) s += u(r, k) * u(r, k);
    sv[k] = std::sqrt(s);
    smax = std::max(smax, sv[k]);
  }
  double cut = smax * DBL_EPSILON * double(std::max(n, p));
  mat out(p, b.n_cols);
  for (uword c = 0; c < b.n_cols; c++)
    for (uword k = 0; k < p; k++) {
      if (!(sv[k] > cut)) continue;
      double proj = 0.0;
      for (uword r = 0; r < n; r++) proj += u(r, k) * b(r, c);
      proj /= sv[k] * sv[k];
      for (uword r = 0; r < p; r++) out(r, c) += v(r, k) * proj;
    }
  return out;
}
inline bool solve(mat &out, const mat &a0, const mat &b0) {
  if (a0.n_rows != b0.n_rows) throw std::logic_error("solve(): number of rows in given matrices must be the same");
  if (a0.n_rows != a0.n_cols) { out = solve_minnorm(a0, b0); return true; }
  uword n = a0.n_rows;
  mat a = a0, b = b0;
  std::vector<uword> piv;
  bool finite = true;
  for (uword k = 0; k < a.n_elem; k++) finite = finite && std::isfinite(a.mem[k]);
  for (uword k = 0; k < b.n_elem; k++) finite = finite && std::isfinite(b.mem[k]);
  if (!finite) { /* LAPACK propagates NaNs; the reference then flags the row through its |beta| / NaN tests */
    out = mat(n, b.n_cols);
    std::fill(out.mem.begin(), out.mem.end(), NAN);
    return false;
  }
  int sign = lu_inplace(a, piv);
  bool ok = sign != 0;
  if (ok) {
    mat ainv;
    double rc = inv_gj(a0, ainv) ? 1.0 / (norm1(a0) * norm1(ainv)) : 0.0;
    ok = rc >= DBL_EPSILON;
  }
  if (!ok) { out = solve_minnorm(a0, b0); return true; }
  for (uword c = 0; c < b.n_cols; c++) {
    for (uword k = 0; k < n; k++)
      if (piv[k] != k) std::swap(b(k, c), b(piv[k], c));
    for (uword r = 1; r < n; r++)
      for (uword k = 0; k < r; k++) b(r, c) -= a(r, k) * b(k, c);
    for (uword r = n; r-- > 0;) {
      for (uword k = r + 1; k < n; k++) b(r, c) -= a(r, k) * b(k, c);
      b(r, c) /= a(r, r);
    }
  }
  out = b;
  return true;
}
/* thin Householder QR: X (rows x p, rows >= p) = Q (rows x p) R (p x p) */
inline bool qr_econ(mat &q, mat &r, const mat &x) {
  uword rows = x.n_rows, p = x.n_cols;
  mat a = x;
  std::vector<std::vector<double>> vs(p);
  std::vector<double> betas(p, 0.0);
  for (uword k = 0; k < p && k < rows; k++) {
    double nrm = 0.0;
    for (uword i = k; i < rows; i++) nrm += a(i, k) * a(i, k);
    nrm = std::sqrt(nrm);
    std::vector<double> v(rows, 0.0);
    if (nrm > 0.0) {
      double alpha = a(k, k) >= 0 ? -nrm : nrm;
      for (uword i = k; i < rows; i++) v[i] = a(i, k);
      v[k] -= alpha;
      double vn = 0.0;
      for (uword i = k; i < rows; i++) vn += v[i] * v[i];
      if (vn > 0.0) {
        betas[k] = 2.0 / vn;
        for (uword c = k; c < p; c++) {
          double s = 0.0;
          for (uword i = k; i < rows; i++) s += v[i] * a(i, c);
          s *= betas[k];
          for (uword i = k; i < rows; i++) a(i, c) -= s * v[i];
        }
      }
    }
    vs[k] = v;
  }
  r = mat(p, p);
  for (uword c = 0; c < p; c++)
    for (uword i = 0; i <= c && i < rows; i++) r(i, c) = a(i, c);
  q = mat(rows, p);
  for (uword c = 0; c < p; c++) q(c, c) = 1.0;
  for (uword k = p; k-- > 0;) {
    if (betas[k] == 0.0) continue;
    for (uword c = 0; c < p; c++) {
      double s = 0.0;
      for (uword i = k; i < rows; i++) s += vs[k][i] * q(i, c);
      s *= betas[k];
      for (uword i = k; i < rows; i++) q(i, c) -= s * vs[k][i];
    }
  }
  return true;
}
}  // namespace arma

/* ------------------------------------------------------------------------------------------------ Rcpp */
extern "C" {
double Rf_lgammafn(double);
double Rf_digamma(double);
double Rf_trigamma(double);
}

namespace Rcpp {

struct NamedPlaceHolder {};
static const NamedPlaceHolder _ = NamedPlaceHolder();

inline void checkUserInterrupt() {}

/* an evaluated sugar expression */
struct SugarVec {
  std::vector<double> v;
  explicit SugarVec(size_t n = 0) : v(n) {}
  size_t size() const { return v.size(); }
  double operator[](size_t i) const { return v[i]; }
};

template <class T> struct is_sugar : std::false_type {};
template <> struct is_sugar<SugarVec> : std::true_type {};

class NumericVector {
 public:
  std::shared_ptr<std::vector<double>> d;
  NumericVector() : d(std::make_shared<std::vector<double>>()) {}
  NumericVector(SEXP s) : d(standin_as_real(s)) {}
  explicit NumericVector(int n) : d(std::make_shared<std::vector<double>>(n, 0.0)) {}
  template <class S, class = typename std::enable_if<is_sugar<S>::value>::type> NumericVector(const S &s)
      : d(std::make_shared<std::vector<double>>(s.size())) {
    for (size_t i = 0; i < s.size(); i++) (*d)[i] = s[i];
  }
  double &operator()(int i) { return (*d)[i]; }
  double operator()(int i) const { return (*d)[i]; }
  double operator[](size_t i) const { return (*d)[i]; }
  size_t size() const { return d->size(); }
  operator SEXP() const {
    SEXP s = standin_new(STANDIN_REALSXP, 0);
    s->real = d;
    s->length = (long)d->size();
    return s;
  }
};
template <> struct is_sugar<NumericVector> : std::true_type {};

class IntegerVector {
 public:
  std::shared_ptr<std::vector<int>> d;
  explicit IntegerVector(int n) : d(std::make_shared<std::vector<int>>(n, 0)) {}
  int &operator()(int i) { return (*d)[i]; }
  operator SEXP() const {
    SEXP s = standin_new(STANDIN_INTSXP, 0);
    s->ints = d;
    s->length = (long)d->size();
    return s;
  }
};

class NumericMatrix {
 public:
  std::shared_ptr<std::vector<double>> d;
  int nr = 0, nc = 0;
  NumericMatrix(SEXP s) : d(standin_as_real(s)), nr(s->nrow), nc(s->ncol) {
    if (nr < 0 || nc < 0) throw std::runtime_error("not a matrix");
  }
  int nrow() const { return nr; }
  int ncol() const { return nc; }
  class Row { /* a strided view of one matrix row (MatrixRow) */
   public:
    std::shared_ptr<std::vector<double>> d;
    int r, nr, nc;
    size_t size() const { return (size_t)nc; }
    double operator[](size_t j) const { return (*d)[(size_t)r + j * (size_t)nr]; }
  };
  Row operator()(int i, NamedPlaceHolder) const { return Row{d, i, nr, nc}; }
  Row row(int i) const { return Row{d, i, nr, nc}; }
};
template <> struct is_sugar<NumericMatrix::Row> : std::true_type {};

#define STANDIN_SUGAR_BINOP(OP)                                                                                     \
  template <class A, class B>                                                                                       \
  inline typename std::enable_if<is_sugar<A>::value && is_sugar<B>::value, SugarVec>::type operator OP(const A &a,  \
                                                                                                       const B &b) { \
    if (a.size() != b.size()) throw std::runtime_error("sugar: incompatible sizes");                               \
    SugarVec o(a.size());                                                                                           \
    for (size_t i = 0; i < a.size(); i++) o.v[i] = a[i] OP b[i];                                                    \
    return o;                                                                                                       \
  }                                                                                                                 \
  template <class A, class S>                                                                                       \
  inline typename std::enable_if<is_sugar<A>::value && std::is_arithmetic<S>::value, SugarVec>::type operator OP(   \
      const A &a, S s) {                                                                                            \
    SugarVec o(a.size());                                                                                           \
    for (size_t i = 0; i < a.size(); i++) o.v[i] = a[i] OP (double)s;                                               \
    return o;                                                                                                       \
  }                                                                                                                 \
  template <class S, class A>                                                                                       \
  inline typename std::enable_if<is_sugar<A>::value && std::is_arithmetic<S>::value, SugarVec>::type operator OP(   \
      S s, const A &a) {                                                                                            \
    SugarVec o(a.size());                                                                                           \
    for (size_t i = 0; i < a.size(); i++) o.v[i] = (double)s OP a[i];                                               \
    return o;                                                                                                       \
  }
STANDIN_SUGAR_BINOP(+)
STANDIN_SUGAR_BINOP(-)
STANDIN_SUGAR_BINOP(*)
STANDIN_SUGAR_BINOP(/)
#undef STANDIN_SUGAR_BINOP

template <class A, class F> inline SugarVec sugar_map(const A &a, F f) {
  SugarVec o(a.size());
  for (size_t i = 0; i < a.size(); i++) o.v[i] = f(a[i]);
  return o;
}
template <class A, class E>
inline typename std::enable_if<is_sugar<A>::value && std::is_arithmetic<E>::value, SugarVec>::type pow(const A &a, E e) {
  return sugar_map(a, [e](double x) { return std::pow(x, (double)e); });
}
template <class A> inline typename std::enable_if<is_sugar<A>::value, SugarVec>::type log(const A &a) {
  return sugar_map(a, [](double x) { return std::log(x); });
}
template <class A> inline typename std::enable_if<is_sugar<A>::value, SugarVec>::type lgamma(const A &a) {
  return sugar_map(a, [](double x) { return ::Rf_lgammafn(x); });
}
template <class A> inline typename std::enable_if<is_sugar<A>::value, SugarVec>::type digamma(const A &a) {
  return sugar_map(a, [](double x) { return ::Rf_digamma(x); });
}
template <class A> inline typename std::enable_if<is_sugar<A>::value, SugarVec>::type trigamma(const A &a) {
  return sugar_map(a, [](double x) { return ::Rf_trigamma(x); });
}
template <class A> inline typename std::enable_if<is_sugar<A>::value, double>::type sum(const A &a) {
  double s = 0.0; /* Rcpp::sugar::Sum for REALSXP: plain sequential double accumulation */
  for (size_t i = 0; i < a.size(); i++) s += a[i];
  return s;
}

inline SEXP clone(SEXP s) {
  SEXP o = standin_new(s->type, 0, s->nrow, s->ncol);
  o->length = s->length;
  if (s->type == STANDIN_REALSXP) o->real = std::make_shared<std::vector<double>>(*s->real);
  else o->ints = std::make_shared<std::vector<int>>(*s->ints);
  return o;
}

template <class T> struct as_impl;
template <> struct as_impl<double> {
  static double get(SEXP s) {
    if (s->length != 1) throw std::runtime_error("Expecting a single value");
    return s->type == STANDIN_REALSXP ? (*s->real)[0] : (double)(*s->ints)[0];
  }
};
template <> struct as_impl<int> {
  static int get(SEXP s) {
    if (s->length != 1) throw std::runtime_error("Expecting a single value");
    return s->type == STANDIN_REALSXP ? (int)(*s->real)[0] : (*s->ints)[0];
  }
};
template <> struct as_impl<bool> {
  static bool get(SEXP s) {
    if (s->length != 1) throw std::runtime_error("Expecting a single value");
    return s->type == STANDIN_REALSXP ? (*s->real)[0] != 0.0 : (*s->ints)[0] != 0;
  }
};
template <> struct as_impl<arma::mat> { /* also arma::vec / arma::colvec (one class here) */
  static arma::mat get(SEXP s) {
    auto r = standin_as_real(s);
    arma::mat o(s->nrow >= 0 ? s->nrow : s->length, s->ncol >= 0 ? s->ncol : 1);
    std::copy(r->begin(), r->end(), o.mem.begin());
    return o;
  }
};
template <class T> inline T as(SEXP s) { return as_impl<T>::get(s); }
template <class T> inline T as(const NumericVector &v) { return as_impl<T>::get((SEXP)v); }

inline SEXP wrap(SEXP s) { return s; }
inline SEXP wrap(const NumericVector &v) { return (SEXP)v; }
inline SEXP wrap(const IntegerVector &v) { return (SEXP)v; }
inline SEXP wrap(const arma::mat &m) {
  SEXP s = standin_new(STANDIN_REALSXP, (long)m.n_elem, (int)m.n_rows, (int)m.n_cols);
  std::copy(m.mem.begin(), m.mem.end(), s->real->begin());
  return s;
}

struct NamedObject {
  std::string name;
  SEXP value;
};
template <class T> inline NamedObject Named(const char *name, const T &obj) { return NamedObject{name, wrap(obj)}; }

class List {
 public:
  std::vector<NamedObject> items;
  template <class... Args> static List create(const Args &... args) {
    List l;
    l.items = {args...};
    return l;
  }
  SEXP operator[](const std::string &name) const {
    for (const auto &it : items)
      if (it.name == name) return it.value;
    throw std::runtime_error("no such list member: " + name);
  }
};

}  // namespace Rcpp

inline arma::mat::mat(const Rcpp::SugarVec &s) : n_rows(s.size()), n_cols(1), n_elem(s.size()), mem(s.v) {}

#endif
