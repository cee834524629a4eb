// One-CTA dense Cholesky helpers shared by the mode-statistics and the mixture kernels.
#pragma once
#include "tb_common.cuh"

namespace tb {

// in-place lower Cholesky of the d x d matrix in `A` (shared memory); returns false on a
// non-positive / non-finite pivot (LAPACK potrf info > 0  <=>  np.linalg.LinAlgError)
__device__ inline bool chol_lower(double* A, int d) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  for (int j = 0; j < d; ++j) {
    if (threadIdx.x == 0) {
      double s = A[j * d + j];
      for (int k = 0; k < j; ++k) s -= A[j * d + k] * A[j * d + k];
      if (!(s > 0.0) || !isfinite(s)) ok = 0; else A[j * d + j] = sqrt(s);
    }
    __syncthreads();
    if (!ok) return false;
    const double piv = A[j * d + j];
    for (int i = j + 1 + threadIdx.x; i < d; i += blockDim.x) {
      double s = A[i * d + j];
      for (int k = 0; k < j; ++k) s -= A[i * d + k] * A[j * d + k];
      A[i * d + j] = s / piv;
    }
    __syncthreads();
  }
  return true;
}

// L^{-1} of the lower factor in `A` (d x d, shared memory) into `Li`: column c by forward substitution
__device__ inline void lower_inverse(const double* A, double* Li, int d) {
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    for (int i = 0; i < d; ++i) {
      if (i < c) { Li[i * d + c] = 0.0; continue; }
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= A[i * d + k] * Li[k * d + c];
      Li[i * d + c] = s / A[i * d + i];
    }
  }
  __syncthreads();
}

}  // namespace tb
