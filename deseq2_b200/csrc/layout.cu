// layout.cu -- layout conversion between R's column-major n x m matrices (genes fastest; what
// .Call hands over, /root/reference/src/DESeq2.cpp:165-169,285-287) and the engine's gene-major
// rows (sample axis contiguous, row stride ld) that the warp-per-gene kernels read with 128-bit loads.
// Plain tiled shared-memory transposes: both global sides are fully coalesced.
#include "engine.h"
#include "nbmath.cuh"

namespace nb {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) to_gene_major_kernel(const T* __restrict__ src, T* __restrict__ dst, int n,
                                                            int m, long long ld) {
  __shared__ T tile[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int jj = ty; jj < 32; jj += 8) {
    const int i = i0 + tx, j = j0 + jj;
    if (i < n && j < m) tile[jj][tx] = src[(size_t)i + (size_t)n * j];
  }
  __syncthreads();
  for (int ii = ty; ii < 32; ii += 8) {
    const int i = i0 + ii, j = j0 + tx;
    if (i < n && j < m) dst[(size_t)i * ld + j] = tile[tx][ii];
  }
}

__global__ void __launch_bounds__(256) to_col_major_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                           int n, int m, long long ld) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int ii = ty; ii < 32; ii += 8) {
    const int i = i0 + ii, j = j0 + tx;
    if (i < n && j < m) tile[ii][tx] = src[(size_t)i * ld + j];
  }
  __syncthreads();
  for (int jj = ty; jj < 32; jj += 8) {
    const int i = i0 + tx, j = j0 + jj;
    if (i < n && j < m) dst[(size_t)i + (size_t)n * j] = tile[tx][jj];
  }
}

__global__ void special_test_kernel(const double* x, int n, double* lg, double* dg, double* tg) {
  init_log_table();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a, b;
  lgamma_digamma_pos(x[i], a, b);
  lg[i] = a;
  dg[i] = b;
  tg[i] = trigamma_pos(x[i]);
}

}  // namespace

cudaError_t launch_to_gene_major(const void* src, void* dst, int n, int m, long long ld, int elem_size,
                                 cudaStream_t stream) {
  if (n == 0 || m == 0) return cudaSuccess;
  dim3 grid((n + 31) / 32, (m + 31) / 32), block(32, 8);
  if (elem_size == 4)
    to_gene_major_kernel<int32_t><<<grid, block, 0, stream>>>(static_cast<const int32_t*>(src),
                                                              static_cast<int32_t*>(dst), n, m, ld);
  else if (elem_size == 8)
    to_gene_major_kernel<double><<<grid, block, 0, stream>>>(static_cast<const double*>(src),
                                                             static_cast<double*>(dst), n, m, ld);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_to_col_major(const double* src, double* dst, int n, int m, long long ld, cudaStream_t stream) {
  if (n == 0 || m == 0) return cudaSuccess;
  dim3 grid((n + 31) / 32, (m + 31) / 32), block(32, 8);
  to_col_major_kernel<<<grid, block, 0, stream>>>(src, dst, n, m, ld);
  return cudaGetLastError();
}

cudaError_t launch_special_test(const double* x, int n, double* lg, double* dg, double* tg, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  special_test_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x, n, lg, dg, tg);
  return cudaGetLastError();
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace nb
