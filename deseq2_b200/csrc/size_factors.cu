// size_factors.cu -- median-of-ratios size factors on the device.
// Reference: estimateSizeFactorsForMatrix, /root/reference/R/core.R:535-578 (locfunc = median, no geoMeans /
// controlGenes):
//   type "ratio":     loggeomean_i = mean_j log K_ij                 (-Inf as soon as one count of the gene is 0)
//   type "poscounts": loggeomean_i = mean_j [K_ij > 0] log K_ij      (-Inf only for all-zero genes)
//   s_j = exp(median_i { log K_ij - loggeomean_i : loggeomean_i finite and K_ij > 0 })
// Three launches, no host round trip:
//   sf_ratio_kernel   one warp per gene: log ratios into a gene-major n x ld matrix, +Inf where the entry does not
//                     take part; counts the genes with a finite log geometric mean (R stops if there are none)
//   to_col_major      (layout.cu) makes the ratios of one sample contiguous
//   sf_median_kernel  one CTA per sample: exact median by a most-significant-byte-first radix select on the
//                     order-preserving 64-bit image of the doubles (8 histogram passes + one pass for the upper
//                     middle element when the count is even) -- no sort, no scratch beyond 256 shared counters.
#include "engine.h"
#include "nbmath.cuh"

#include <math.h>

namespace nb {
namespace {

__device__ __forceinline__ unsigned long long key_of(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);      // ascending doubles <-> ascending unsigned keys
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) sf_ratio_kernel(const SizeFactorArgs A) {
  init_log_table();
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= A.n) return;                                      // whole warps leave together
  const size_t off = (size_t)g * A.ld;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double s = 0.0;
  int zeros = 0, pos = 0;
  for (int j = lane; j < A.m; j += 32) {
    const double y = A.y_is_f64 ? static_cast<const double*>(A.y)[off + j]
                                : (double)static_cast<const int32_t*>(A.y)[off + j];
    if (y > 0.0) {
      const double l = log_pos(y);
      A.ratios[off + j] = l;                                 // log K_ij for now; the mean is subtracted below
      s += l;
      pos++;
    } else {
      zeros++;
    }
  }
  s = warp_allreduce_sum(s);
  zeros = (int)warp_allreduce_sum((double)zeros);
  pos = (int)warp_allreduce_sum((double)pos);
  const bool finite = A.poscounts ? (pos > 0) : (zeros == 0);
  const double lg = finite ? s / (double)A.m : -inf;
  __syncwarp();
  for (int j = lane; j < A.m; j += 32) {
    const double y = A.y_is_f64 ? static_cast<const double*>(A.y)[off + j]
                                : (double)static_cast<const int32_t*>(A.y)[off + j];
    A.ratios[off + j] = (finite && y > 0.0) ? A.ratios[off + j] - lg : inf;
  }
  if (lane == 0) {
    A.loggeomeans[g] = lg;
    if (finite) atomicAdd(A.n_finite, 1);
  }
}

// col: the n ratios of one sample (contiguous); +Inf marks entries that do not take part.
__global__ void __launch_bounds__(512) sf_median_kernel(const double* __restrict__ ratios_colmajor, int n,
                                                        double* __restrict__ size_factors) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned int s_k, s_nv, s_le;
  __shared__ unsigned long long s_next;
  const double* col = ratios_colmajor + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  const unsigned long long kinf = key_of(__longlong_as_double(0x7ff0000000000000ll));
  if (tid == 0) { s_nv = 0; s_prefix = 0; }
  __syncthreads();
  unsigned int mine = 0;
  for (int i = tid; i < n; i += blockDim.x) mine += key_of(col[i]) < kinf;
  if (mine) atomicAdd(&s_nv, mine);
  __syncthreads();
  const unsigned int nv = s_nv;
  if (nv == 0) {                                             // median of nothing: NA in R
    if (tid == 0) size_factors[blockIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  if (tid == 0) s_k = (nv - 1) / 2;                          // lower middle element (0-based rank)
  for (int pass = 7; pass >= 0; pass--) {
    for (int b = tid; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const unsigned long long mask = (pass == 7) ? 0ull : (~0ull << (8 * (pass + 1)));
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long k = key_of(col[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(unsigned int)(k >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int k = s_k, b = 0;
      while (b < 255u && k >= hist[b]) { k -= hist[b]; b++; }   // k < number of keys under this prefix; b is bounded anyway
      s_k = k;
      s_prefix = prefix | ((unsigned long long)b << (8 * pass));
    }
    __syncthreads();
  }
  const unsigned long long k1 = s_prefix;
  double med = value_of(k1);
  if ((nv & 1u) == 0) {
    // upper middle element: k1 again if enough copies of it exist, otherwise the smallest key above k1
    if (tid == 0) { s_le = 0; s_next = ~0ull; }
    __syncthreads();
    unsigned int le = 0;
    unsigned long long nxt = ~0ull;
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long k = key_of(col[i]);
      le += k <= k1;
      if (k > k1 && k < nxt) nxt = k;
    }
    if (le) atomicAdd(&s_le, le);
    atomicMin(&s_next, nxt);
    __syncthreads();
    const unsigned long long k2 = (s_le > nv / 2) ? k1 : s_next;
    med = 0.5 * (med + value_of(k2));
  }
  if (tid == 0) size_factors[blockIdx.x] = exp(med);
}

}  // namespace

cudaError_t launch_size_factors(const SizeFactorArgs& a, cudaStream_t stream) {
  if (a.n == 0 || a.m == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(a.n_finite, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  sf_ratio_kernel<<<(a.n + 7) / 8, 256, 0, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = launch_to_col_major(a.ratios, a.ratios_colmajor, a.n, a.m, a.ld, stream);
  if (e != cudaSuccess) return e;
  sf_median_kernel<<<a.m, 512, 0, stream>>>(a.ratios_colmajor, a.n, a.size_factors);
  return cudaGetLastError();
}

}  // namespace nb
