// fit_disp.cu -- gene-wise Cox-Reid adjusted NB dispersion MLE/MAP on sm_100a, one warp per gene.
//
// Behavioural contract (WHAT): /root/reference/src/DESeq2.cpp:164-277 (fitDisp: Armijo back-tracking
// ascent on log alpha, accept / reject / kappa schedule replicated decision for decision), :31-158
// (log posterior and its first two derivatives, Cox-Reid term with weight-threshold row/column
// subsetting) and :469-513 (fitDispGrid).  HOW is new:
//   * the gene's row (counts, mu, 1/mu, weights) is staged once into shared memory with 128-bit loads;
//     every lane owns samples lane, lane+32, ...; all per-gene scalars are warp-uniform;
//   * one fused pass per proposal evaluates the log posterior AND its derivative (they share
//     log(1+mu*alpha), 1/(1/mu+alpha), log(y+1/alpha) and 1/(y+1/alpha)), so an accepted step costs
//     one pass instead of the reference's three (theta(kappa), lpnew, dlp) -- the values are the same
//     because the reference re-evaluates the same function at the same point (:225 vs :233);
//   * lgamma/digamma come from one log + one reciprocal (nbmath.cuh); the identities
//     log(mu+1/alpha) = log(1+mu*alpha) - log(alpha),  mu*alpha/(1+mu*alpha) = alpha*wd,
//     y/(mu+1/alpha) = y*alpha*wd/mu  (wd = 1/(1/mu+alpha)) remove two logs/divisions per sample;
//   * X'WX, X'dWX are accumulated per lane and combined with xor-butterfly shuffles (bitwise identical in
//     every lane, so control flow stays warp-uniform); the p x p Cholesky runs in registers;
//   * genes are pulled from an atomic work queue by persistent warps (trip counts vary 1..maxit).
#include "engine.h"
#include "smallp.cuh"

namespace nb {

namespace {

struct DispRow {
  const double* y;
  const double* mu;
  const double* imu;
  const double* w;   // only read when USE_W
  const double* x;   // shared, column-major with stride mpad
  int m, mpad;
};

struct DispScal {
  double prior_sigmasq, weight_threshold;
  int use_prior, use_cr;
};

// One fused pass: lp (and dlp when WANT_D) at log-alpha `a`.
template <int P, bool USE_W, bool WANT_D>
__device__ __forceinline__ void disp_eval(const DispRow& rv, const DispScal& sc, double a, double pm, double sum_wy,
                                          int lane, double& lp, double& dlp) {
  constexpr int NS = SymP<P>::N;
  const double alpha = exp(a);
  const double r = 1.0 / alpha;
  double lg_r, dg_r;
  lgamma_digamma_pos(r, lg_r, dg_r);

  double acc[2 + 2 * NS];
#pragma unroll
  for (int i = 0; i < 2 + 2 * NS; i++) acc[i] = 0.0;

#pragma unroll 2
  for (int j = lane; j < rv.m; j += 32) {
    const double y = rv.y[j], mu = rv.mu[j], imu = rv.imu[j];
    const double wd = rcp_fast(imu + alpha);
    const double onema = fma(mu, alpha, 1.0);
    const double l2 = log(onema);
    const double xr = y + r;
    double lg, dg;
    lgamma_digamma_pos(xr, lg, dg);
    double t = (lg - lg_r) - xr * l2;
    double wt = 1.0;
    if (USE_W) {
      wt = rv.w[j];
      t *= wt;
    }
    acc[0] += t;
    if (WANT_D) {
      double d = (dg_r - dg) + l2 + alpha * wd * fma(y, imu, -1.0);
      if (USE_W) d *= wt;
      acc[1] += d;
    }
    if (sc.use_cr) {
      double wdm = wd;
      if (USE_W) wdm = (wt > sc.weight_threshold) ? wd : 0.0;
      const double dwd = -wdm * wd;
      double xv[P];
#pragma unroll
      for (int k = 0; k < P; k++) xv[k] = rv.x[k * rv.mpad + j];
#pragma unroll
      for (int aa = 0; aa < P; aa++)
#pragma unroll
        for (int bb = 0; bb <= aa; bb++) {
          const double xx = xv[aa] * xv[bb];
          acc[2 + aa * (aa + 1) / 2 + bb] = fma(wdm, xx, acc[2 + aa * (aa + 1) / 2 + bb]);
          if (WANT_D) acc[2 + NS + aa * (aa + 1) / 2 + bb] = fma(dwd, xx, acc[2 + NS + aa * (aa + 1) / 2 + bb]);
        }
    }
  }
  warp_allreduce_sum_n(acc);

  double cr = 0.0, dcr = 0.0;
  if (sc.use_cr) {
    SymP<P> B, dB;
#pragma unroll
    for (int i = 0; i < NS; i++) {
      B.v[i] = acc[2 + i];
      dB.v[i] = acc[2 + NS + i];
    }
    if (USE_W) {
      // a column whose kept rows are all zero is dropped by the reference (:42): make it inert
#pragma unroll
      for (int k = 0; k < P; k++)
        if (B.get(k, k) == 0.0) B.at(k, k) = 1.0;
    }
    chol_factor<P>(B);
    cr = -0.5 * log(chol_det<P>(B));
    if (WANT_D) {
      SymP<P> Bi;
      chol_inverse<P>(B, Bi);
      dcr = -0.5 * sym_trace_prod<P>(Bi, dB);
    }
  }
  double prior = 0.0, dprior = 0.0;
  if (sc.use_prior) {
    const double dd = a - pm;
    prior = -0.5 * dd * dd / sc.prior_sigmasq;
    dprior = -1.0 * dd / sc.prior_sigmasq;
  }
  lp = (acc[0] + a * sum_wy) + prior + cr;
  if (WANT_D) dlp = (r * r * acc[1] + dcr) * alpha + dprior;
}

// second derivative at `a` (once per gene): src/DESeq2.cpp:111-158
template <int P, bool USE_W>
__device__ __forceinline__ double disp_d2(const DispRow& rv, const DispScal& sc, double a, int lane) {
  constexpr int NS = SymP<P>::N;
  const double alpha = exp(a);
  const double r = 1.0 / alpha;
  const double r2 = r * r;
  const double dg_r = digamma_pos(r), tg_r = trigamma_pos(r);
  double acc[2 + 3 * NS];
#pragma unroll
  for (int i = 0; i < 2 + 3 * NS; i++) acc[i] = 0.0;
  for (int j = lane; j < rv.m; j += 32) {
    const double y = rv.y[j], mu = rv.mu[j], imu = rv.imu[j];
    const double wd = 1.0 / (imu + alpha);
    const double onema = fma(mu, alpha, 1.0);
    const double l2 = log(onema);
    const double xr = y + r;
    const double mpr = mu + r;
    double t1 = dg_r + l2 - mu * alpha / onema - digamma_pos(xr) + y / mpr;
    double t2 = -r2 * tg_r + mu * mu * alpha / (onema * onema) + r2 * trigamma_pos(xr) + r2 * y / (mpr * mpr);
    double wt = 1.0;
    if (USE_W) {
      wt = rv.w[j];
      t1 *= wt;
      t2 *= wt;
    }
    acc[0] += t1;
    acc[1] += t2;
    if (sc.use_cr) {
      double wdm = wd;
      if (USE_W) wdm = (wt > sc.weight_threshold) ? wd : 0.0;
      const double dwd = -wdm * wd;
      const double d2wd = 2.0 * wdm * wd * wd;
      double xv[P];
#pragma unroll
      for (int k = 0; k < P; k++) xv[k] = rv.x[k * rv.mpad + j];
#pragma unroll
      for (int aa = 0; aa < P; aa++)
#pragma unroll
        for (int bb = 0; bb <= aa; bb++) {
          const double xx = xv[aa] * xv[bb];
          const int q = aa * (aa + 1) / 2 + bb;
          acc[2 + q] = fma(wdm, xx, acc[2 + q]);
          acc[2 + NS + q] = fma(dwd, xx, acc[2 + NS + q]);
          acc[2 + 2 * NS + q] = fma(d2wd, xx, acc[2 + 2 * NS + q]);
        }
    }
  }
  warp_allreduce_sum_n(acc);
  double cr2 = 0.0, dcr = 0.0;
  if (sc.use_cr) {
    SymP<P> B, dB, d2B, Bi;
#pragma unroll
    for (int i = 0; i < NS; i++) {
      B.v[i] = acc[2 + i];
      dB.v[i] = acc[2 + NS + i];
      d2B.v[i] = acc[2 + 2 * NS + i];
    }
    if (USE_W) {
#pragma unroll
      for (int k = 0; k < P; k++)
        if (B.get(k, k) == 0.0) B.at(k, k) = 1.0;
    }
    chol_factor<P>(B);
    chol_inverse<P>(B, Bi);
    double M[P][P];
    sym_mul_full<P>(Bi, dB, M);
    double tr1 = 0.0, tr2 = 0.0;
#pragma unroll
    for (int i = 0; i < P; i++) {
      tr1 += M[i][i];
#pragma unroll
      for (int k = 0; k < P; k++) tr2 = fma(M[i][k], M[k][i], tr2);
    }
    const double tr3 = sym_trace_prod<P>(Bi, d2B);
    cr2 = 0.5 * tr1 * tr1 - 0.5 * (tr1 * tr1 - tr2 + tr3);
    dcr = -0.5 * tr1;
  }
  const double ll2 = -2.0 * r2 * r * acc[0] + r2 * acc[1];
  const double dlp_noprior = (r2 * acc[0] + dcr) * alpha;
  const double prior2 = sc.use_prior ? -1.0 / sc.prior_sigmasq : 0.0;
  return ((ll2 + cr2) * alpha * alpha + dlp_noprior) + prior2;
}

template <int P, bool USE_W>
__global__ void __launch_bounds__(256) fit_disp_kernel(const DispArgs A, int warps_per_cta, int mpad) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int NROW = USE_W ? 4 : 3;
  double* xs = smem;                                   // P * mpad
  double* rowbase = smem + (size_t)P * mpad + (size_t)warp * NROW * mpad;
  double* ys = rowbase;
  double* mus = rowbase + mpad;
  double* imus = rowbase + 2 * mpad;
  double* wsm = USE_W ? rowbase + 3 * mpad : nullptr;

  // stage the design matrix (column-major m x p -> column-major with padded stride)
  for (int idx = threadIdx.x; idx < P * A.m; idx += blockDim.x) {
    const int k = idx / A.m, j = idx - k * A.m;
    xs[k * mpad + j] = A.x[idx];
  }
  __syncthreads();

  DispRow rv{ys, mus, imus, wsm, xs, A.m, mpad};
  DispScal sc{A.prior_sigmasq, A.weight_threshold, A.use_prior, A.use_cr};
  const double epsilon = 1.0e-4;

  for (;;) {
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(A.counter, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= (unsigned int)A.n) break;

    // ---- stage the gene row: 128-bit loads along the sample axis
    double sum_wy_l = 0.0;
    {
      const size_t off = (size_t)g * A.ld;
      for (int j4 = lane * 4; j4 < mpad; j4 += 128) {
        double yv[4];
        if (A.y_is_f64) {
          const double2* p2 = reinterpret_cast<const double2*>(static_cast<const double*>(A.y) + off + j4);
          const double2 a0 = __ldg(p2), a1 = __ldg(p2 + 1);
          yv[0] = a0.x; yv[1] = a0.y; yv[2] = a1.x; yv[3] = a1.y;
        } else {
          const int4 v = __ldg(reinterpret_cast<const int4*>(static_cast<const int32_t*>(A.y) + off + j4));
          yv[0] = v.x; yv[1] = v.y; yv[2] = v.z; yv[3] = v.w;
        }
        const double2* m2 = reinterpret_cast<const double2*>(A.mu + off + j4);
        const double2 m0 = __ldg(m2), m1 = __ldg(m2 + 1);
        const double mv[4] = {m0.x, m0.y, m1.x, m1.y};
        double wv[4] = {1.0, 1.0, 1.0, 1.0};
        if (USE_W) {
          const double2* w2 = reinterpret_cast<const double2*>(A.w + off + j4);
          const double2 w0 = __ldg(w2), w1 = __ldg(w2 + 1);
          wv[0] = w0.x; wv[1] = w0.y; wv[2] = w1.x; wv[3] = w1.y;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int j = j4 + q;
          ys[j] = yv[q];
          mus[j] = mv[q];
          imus[j] = 1.0 / mv[q];
          if (USE_W) wsm[j] = wv[q];
          if (j < A.m) sum_wy_l += wv[q] * yv[q];
        }
      }
    }
    __syncwarp();
    const double sum_wy = warp_allreduce_sum(sum_wy_l);
    const double pm = A.prior_mean[g];

    if (A.grid != nullptr) {
      // ---- fitDispGrid (src/DESeq2.cpp:492-510)
      const int gn = A.grid_n;
      const double delta = A.grid[1] - A.grid[0];
      double best = 0.0, a_hat = 0.0, dummy;
      for (int t = 0; t < gn; t++) {
        const double a = A.grid[t];
        double lp;
        disp_eval<P, USE_W, false>(rv, sc, a, pm, sum_wy, lane, lp, dummy);
        if (t == 0 || lp > best) { best = lp; a_hat = a; }
      }
      const double start = a_hat - delta, end = a_hat + delta;
      const double step = (end - start) / (double)(gn - 1);
      double a_best = 0.0;
      for (int t = 0; t < gn; t++) {
        const double a = (t == gn - 1) ? end : start + t * step;
        double lp;
        disp_eval<P, USE_W, false>(rv, sc, a, pm, sum_wy, lane, lp, dummy);
        if (t == 0 || lp > best) { best = lp; a_best = a; }
      }
      if (lane == 0) A.log_alpha[g] = a_best;
      __syncwarp();
      continue;
    }

    // ---- fitDisp line search (src/DESeq2.cpp:201-265)
    double a = A.log_alpha_in[g];
    double lp, dlp;
    disp_eval<P, USE_W, true>(rv, sc, a, pm, sum_wy, lane, lp, dlp);
    const double initial_lp = lp, initial_dlp = dlp;
    double kappa = A.kappa_0;
    double change = -1.0;
    int it = 0, acc_n = 0;
    for (int t = 0; t < A.maxit; t++) {
      it++;
      const double a_propose = a + kappa * dlp;
      if (a_propose < -30.0) kappa = (-30.0 - a) / dlp;
      if (a_propose > 10.0) kappa = (10.0 - a) / dlp;
      const double a_new = a + kappa * dlp;
      double lp_new, dlp_new;
      disp_eval<P, USE_W, true>(rv, sc, a_new, pm, sum_wy, lane, lp_new, dlp_new);
      const double theta_kappa = -1.0 * lp_new;
      const double theta_hat_kappa = -1.0 * lp - kappa * epsilon * dlp * dlp;
      if (theta_kappa <= theta_hat_kappa) {
        acc_n++;
        a = a_new;
        change = lp_new - lp;
        if (change < A.tol) { lp = lp_new; break; }
        if (a < A.min_log_alpha) break;
        lp = lp_new;
        dlp = dlp_new;
        kappa = fmin(kappa * 1.1, A.kappa_0);
        if (acc_n % 5 == 0) kappa = kappa / 2.0;
      } else {
        kappa = kappa / 2.0;
      }
    }
    const double d2 = disp_d2<P, USE_W>(rv, sc, a, lane);
    if (lane == 0) {
      A.log_alpha[g] = a;
      A.iter[g] = it;
      A.iter_accept[g] = acc_n;
      A.last_change[g] = change;
      A.initial_lp[g] = initial_lp;
      A.initial_dlp[g] = initial_dlp;
      A.last_lp[g] = lp;
      A.last_dlp[g] = dlp;
      A.last_d2lp[g] = d2;
    }
    __syncwarp();
  }
}

template <int P, bool USE_W>
cudaError_t launch_disp_t(const DispArgs& a, cudaStream_t stream) {
  const int mpad = (a.m + 3) & ~3;
  constexpr int NROW = USE_W ? 4 : 3;
  const size_t xbytes = (size_t)P * mpad * sizeof(double);
  const size_t rowbytes = (size_t)NROW * mpad * sizeof(double);
  const size_t smem_cap = 227 * 1024;
  int warps = 8;
  while (warps > 1 && xbytes + warps * rowbytes > smem_cap / 2) warps >>= 1;   // aim for >= 2 CTAs/SM
  if (xbytes + warps * rowbytes > smem_cap) return cudaErrorInvalidValue;
  const size_t smem = xbytes + warps * rowbytes;
  auto kern = fit_disp_kernel<P, USE_W>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, warps * 32, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
  const int sms = device_sm_count();
  long long want = ((long long)a.n + warps - 1) / warps;
  long long grid = (long long)sms * ctas_per_sm;   // persistent: one resident wave, genes come from the queue
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)grid, warps * 32, smem, stream>>>(a, warps, mpad);
  return cudaGetLastError();
}

template <bool USE_W>
cudaError_t launch_disp_p(const DispArgs& a, cudaStream_t stream) {
  switch (a.p) {
    case 1: return launch_disp_t<1, USE_W>(a, stream);
    case 2: return launch_disp_t<2, USE_W>(a, stream);
    case 3: return launch_disp_t<3, USE_W>(a, stream);
    case 4: return launch_disp_t<4, USE_W>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t launch_fit_disp(const DispArgs& a, cudaStream_t stream) {
  if (a.n == 0) return cudaSuccess;
  return a.use_weights ? launch_disp_p<true>(a, stream) : launch_disp_p<false>(a, stream);
}

}  // namespace nb
