// pipeline_kernels.cu -- the per-gene pre-steps that feed the hot path, and the global dispersion-trend fit,
// as device kernels ("host pre-steps on device").  In the reference these are R:
//   getBaseMeansAndVariances  R/core.R:2138-2157      baseMean, baseVar, allZero
//   linearModelMu             R/core.R:2454-2459      (y Q)(x R^-1)'  ==  X (X'X)^-1 X' y
//   roughDispEstimate         R/core.R:2422-2437
//   momentsDispEstimate       R/core.R:2439-2448
//   alpha_hat bounds          R/core.R:716,727-728    min(rough, moments) clamped to [minDisp, max(10, m)]
//   linearModelMuNormalized   R/core.R:2461-2467      mu = linear-model fit * size factor, clamped at minmu
//   beta start values         R/fitNbinomGLMs.R:139-155  QR least squares of log(K/s + 0.1) on X
//   parametricDispersionFit   R/core.R:2166-2189      Gamma GLM (identity link) disp ~ a0 + a1/mean, iterated trimming
// prep_kernel: one warp per gene; the projection P = (X'X)^-1 X' (p x m) is computed once on the host.
// trend_fit_kernel: ONE CTA runs the whole iteratively re-trimmed IRLS on the device (no host round trips).
#include "engine.h"
#include "nbmath.cuh"

namespace nb {
namespace {

__global__ void __launch_bounds__(256) prep_kernel(const PrepArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int mpad = (A.m + 3) & ~3;
  double* sf = smem;                                   // mpad
  double* wrow = smem + mpad + (size_t)warp * (2 * mpad + 64);
  double* vn = wrow;                                   // normalised counts
  double* vl = wrow + mpad;                            // log(norm + 0.1)
  double* coef = wrow + 2 * mpad;                      // 32: P * norm
  double* coefl = coef + 32;                           // 32: P * log(norm + .1)
  for (int j = threadIdx.x; j < A.m; j += blockDim.x) sf[j] = A.size_factors[j];
  __syncthreads();
  const int g = blockIdx.x * (blockDim.x >> 5) + warp;
  if (g >= A.n) return;
  const size_t off = (size_t)g * A.ld;
  double s = 0.0, s2 = 0.0, raw = 0.0;
  for (int j = lane; j < A.m; j += 32) {
    const double y = A.y_is_f64 ? static_cast<const double*>(A.y)[off + j] : (double)static_cast<const int32_t*>(A.y)[off + j];
    const double v = y / sf[j];
    vn[j] = v;
    vl[j] = log(v + 0.1);
    s += v;
    raw += y;
  }
  s = warp_allreduce_sum(s);
  raw = warp_allreduce_sum(raw);
  const double bm = s / (double)A.m;
  for (int j = lane; j < A.m; j += 32) {
    const double d = vn[j] - bm;
    s2 = fma(d, d, s2);
  }
  s2 = warp_allreduce_sum(s2);
  const double bv = s2 / (double)(A.m - 1);
  __syncwarp();
  // coef = P v  (lane k owns coefficient k)
  if (lane < A.p) {
    double c = 0.0, cl = 0.0;
    const double* Pk = A.proj + (size_t)lane * A.m;
    for (int j = 0; j < A.m; j++) {
      const double pk = __ldg(Pk + j);
      c = fma(pk, vn[j], c);
      cl = fma(pk, vl[j], cl);
    }
    coef[lane] = c;
    coefl[lane] = cl;
    if (A.beta0 != nullptr) A.beta0[(size_t)g + (size_t)A.n * lane] = cl;
  }
  __syncwarp();
  // fitted = X coef; rough dispersion; linear mu
  double rs = 0.0;
  for (int j = lane; j < A.m; j += 32) {
    double f = 0.0;
    for (int k = 0; k < A.p; k++) f = fma(__ldg(A.x + (size_t)k * A.m + j), coef[k], f);
    const double mu1 = fmax(1.0, f);
    const double d = vn[j] - mu1;
    rs += (d * d - mu1) / (mu1 * mu1);
    if (A.mu_lin != nullptr) A.mu_lin[off + j] = fmax(f * sf[j], A.minmu);
  }
  rs = warp_allreduce_sum(rs);
  if (lane == 0) {
    const double rough = fmax(rs / (double)(A.m - A.p), 0.0);
    const double moments = (bv - A.xim * bm) / (bm * bm);
    double a0 = fmin(rough, moments);
    a0 = fmin(fmax(A.min_disp, a0), A.max_disp);   // pmin(pmax(minDisp, alpha_hat), maxDisp): NaN -> minDisp like pmax
    A.base_mean[g] = bm;
    A.base_var[g] = bv;
    A.alpha0[g] = a0;
    A.all_zero[g] = (raw == 0.0) ? 1 : 0;
  }
}

// ---------------------------------------------------------------- dispersion trend (single CTA)
__device__ __forceinline__ double block_sum(double v, double* red, int tid, int nthreads) {
  v = warp_allreduce_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (nthreads >> 5); w++) t += red[w];
  return t;
}

// out[0..1] = coefficients (asymptDisp, extraPois); out[2] = status (0 ok, 1 not converged, 2 non-positive
// coefficients, 3 no usable genes); out[3] = outer iterations used.
__global__ void __launch_bounds__(1024) trend_fit_kernel(const double* __restrict__ means,
                                                         const double* __restrict__ disps, int n, double min_disp,
                                                         double* __restrict__ out) {
  __shared__ double red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  double c0 = 0.1, c1 = 1.0;
  int status = 1, outer = 0;
  for (; outer <= 10; outer++) {   // R: iter > 10 => "dispersion fit did not converge"
    // glm(disps[good] ~ I(1/means[good]), Gamma(identity), start = coefs): IRLS, weights 1/mu^2, response y
    const double t0 = c0, t1 = c1;   // trimming uses the coefficients at loop entry
    double b0 = c0, b1 = c1;
    double dev_old;
    {
      double dv = 0.0, cnt = 0.0;
      for (int i = tid; i < n; i += nt) {
        const double d = disps[i], mn = means[i];
        if (!(d > 100.0 * min_disp)) continue;
        const double resid = d / (t0 + t1 / mn);
        if (!(resid > 1e-4 && resid < 15.0)) continue;
        const double mu = b0 + b1 / mn;
        dv += -log(d / mu) + (d - mu) / mu;
        cnt += 1.0;
      }
      dev_old = 2.0 * block_sum(dv, red, tid, nt);
      cnt = block_sum(cnt, red, tid, nt);
      if (cnt < 2.0) { status = 3; break; }
    }
    bool converged = false, bad = false;
    for (int it = 0; it < 25; it++) {
      double sw = 0.0, swx = 0.0, swxx = 0.0, swy = 0.0, swxy = 0.0;
      for (int i = tid; i < n; i += nt) {
        const double d = disps[i], mn = means[i];
        if (!(d > 100.0 * min_disp)) continue;
        const double resid = d / (t0 + t1 / mn);
        if (!(resid > 1e-4 && resid < 15.0)) continue;
        const double xi = 1.0 / mn;
        const double mu = b0 + b1 * xi;
        const double w = 1.0 / (mu * mu);
        sw += w; swx += w * xi; swxx += w * xi * xi; swy += w * d; swxy += w * xi * d;
      }
      sw = block_sum(sw, red, tid, nt);
      swx = block_sum(swx, red, tid, nt);
      swxx = block_sum(swxx, red, tid, nt);
      swy = block_sum(swy, red, tid, nt);
      swxy = block_sum(swxy, red, tid, nt);
      const double det = sw * swxx - swx * swx;
      b0 = (swxx * swy - swx * swxy) / det;
      b1 = (sw * swxy - swx * swy) / det;
      double dv = 0.0, neg = 0.0;
      for (int i = tid; i < n; i += nt) {
        const double d = disps[i], mn = means[i];
        if (!(d > 100.0 * min_disp)) continue;
        const double resid = d / (t0 + t1 / mn);
        if (!(resid > 1e-4 && resid < 15.0)) continue;
        const double mu = b0 + b1 / mn;
        if (!(mu > 0.0)) neg += 1.0;
        dv += -log(d / mu) + (d - mu) / mu;
      }
      const double dev = 2.0 * block_sum(dv, red, tid, nt);
      neg = block_sum(neg, red, tid, nt);
      if (neg > 0.0) { bad = true; break; }
      if (fabs(dev - dev_old) / (fabs(dev) + 0.1) < 1e-8) { converged = true; break; }
      dev_old = dev;
    }
    if (bad || !(b0 > 0.0) || !(b1 > 0.0)) { status = 2; c0 = b0; c1 = b1; break; }
    const double l0 = log(b0 / c0), l1 = log(b1 / c1);
    c0 = b0;
    c1 = b1;
    if ((l0 * l0 + l1 * l1 < 1e-6) && converged) { status = 0; break; }
  }
  if (tid == 0) {
    out[0] = c0;
    out[1] = c1;
    out[2] = (double)status;
    out[3] = (double)outer;
  }
}

}  // namespace

cudaError_t launch_prep(const PrepArgs& a, cudaStream_t stream) {
  if (a.n == 0) return cudaSuccess;
  const int mpad = (a.m + 3) & ~3;
  int warps = 8;
  size_t smem = ((size_t)mpad + (size_t)warps * (2 * mpad + 64)) * sizeof(double);
  while (warps > 1 && smem > 200 * 1024) {
    warps >>= 1;
    smem = ((size_t)mpad + (size_t)warps * (2 * mpad + 64)) * sizeof(double);
  }
  cudaError_t e = cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  prep_kernel<<<(a.n + warps - 1) / warps, warps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_trend_fit(const double* means, const double* disps, int n, double min_disp, double* out,
                             cudaStream_t stream) {
  trend_fit_kernel<<<1, 1024, 0, stream>>>(means, disps, n, min_disp, out);
  return cudaGetLastError();
}

}  // namespace nb
