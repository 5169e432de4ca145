// nbmath.cuh -- fp64 device math for the NB-GLM kernels (sm_100a).
//
// Own implementations of the special functions the reference pulls from R's nmath
// (Rf_lgammafn / Rf_digamma / Rf_trigamma, call sites /root/reference/src/DESeq2.cpp:50-58,90-96,139-145).
// They are written for the GPU's FP64 pipe: one log + one reciprocal per fused lgamma/digamma pair,
// the rest FMAs, no tables, no divergent slow paths beyond a short predicated shift loop.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace nb {

constexpr double kHalfLog2Pi = 0.918938533204672741780329736406;  // log(sqrt(2*pi))
constexpr double kShift = 10.0;  // arguments below this are shifted up before the asymptotic series

// fast fp64 reciprocal: MUFU.RCP64H seed (~20 bits) + 2 Newton steps (<= 1 ulp for normal inputs).
// Callers use it only where the operand is known positive, finite and normal.
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Stirling tail  sum_{k>=1} B_2k / (2k (2k-1) x^(2k-1))  evaluated as xi * poly(xi^2), x >= kShift
__device__ __forceinline__ double stirling_tail(double xi, double xi2) {
  double p = 1.0 / 156.0;                       // k=7
  p = fma(p, xi2, -691.0 / 360360.0);           // k=6
  p = fma(p, xi2, 1.0 / 1188.0);                // k=5
  p = fma(p, xi2, -1.0 / 1680.0);               // k=4
  p = fma(p, xi2, 1.0 / 1260.0);                // k=3
  p = fma(p, xi2, -1.0 / 360.0);                // k=2
  p = fma(p, xi2, 1.0 / 12.0);                  // k=1
  return p * xi;
}

// sum_{k>=1} B_2k / (2k x^2k), x >= kShift
__device__ __forceinline__ double digamma_tail(double xi2) {
  double p = 1.0 / 12.0;                        // k=7
  p = fma(p, xi2, -691.0 / 32760.0);            // k=6
  p = fma(p, xi2, 1.0 / 132.0);                 // k=5
  p = fma(p, xi2, -1.0 / 240.0);                // k=4
  p = fma(p, xi2, 1.0 / 252.0);                 // k=3
  p = fma(p, xi2, -1.0 / 120.0);                // k=2
  p = fma(p, xi2, 1.0 / 12.0);                  // k=1
  return p * xi2;
}

// lgamma(x) for x > 0.
__device__ __forceinline__ double lgamma_pos(double x) {
  double prod = 1.0;
  bool shifted = x < kShift;
  while (x < kShift) {
    prod *= x;
    x += 1.0;
  }
  double xi = rcp_fast(x);
  double lx = log(x);
  double r = fma(x - 0.5, lx, -x) + kHalfLog2Pi + stirling_tail(xi, xi * xi);
  if (shifted) r -= log(prod);
  return r;
}

// fused lgamma(x), digamma(x) for x > 0: shares log(x) and 1/x.
__device__ __forceinline__ void lgamma_digamma_pos(double x, double& lg, double& dg) {
  // shift: prod = x (x+1) ... (x+n-1), dprod = d prod / dx
  double prod = 1.0, dprod = 0.0;
  bool shifted = x < kShift;
  while (x < kShift) {
    dprod = fma(dprod, x, prod);
    prod *= x;
    x += 1.0;
  }
  double xi = rcp_fast(x);
  double xi2 = xi * xi;
  double lx = log(x);
  lg = fma(x - 0.5, lx, -x) + kHalfLog2Pi + stirling_tail(xi, xi2);
  dg = lx - 0.5 * xi - digamma_tail(xi2);
  if (shifted) {
    lg -= log(prod);
    dg -= dprod / prod;
  }
}

__device__ __forceinline__ double digamma_pos(double x) {
  double lg, dg;
  lgamma_digamma_pos(x, lg, dg);
  return dg;
}

// trigamma(x), x > 0.  Used once per gene (d2lp), so the shift loop may divide.
__device__ __forceinline__ double trigamma_pos(double x) {
  double s = 0.0;
  while (x < kShift) {
    double xi = 1.0 / x;
    s = fma(xi, xi, s);
    x += 1.0;
  }
  double xi = 1.0 / x, xi2 = xi * xi;
  double p = 7.0 / 6.0;                         // B14
  p = fma(p, xi2, -691.0 / 2730.0);             // B12
  p = fma(p, xi2, 5.0 / 66.0);                  // B10
  p = fma(p, xi2, -1.0 / 30.0);                 // B8
  p = fma(p, xi2, 1.0 / 42.0);                  // B6
  p = fma(p, xi2, -1.0 / 30.0);                 // B4
  p = fma(p, xi2, 1.0 / 6.0);                   // B2
  return s + xi + 0.5 * xi2 + p * xi2 * xi;
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double warp_allreduce_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int N>
__device__ __forceinline__ void warp_allreduce_sum_n(double (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}

}  // namespace nb
