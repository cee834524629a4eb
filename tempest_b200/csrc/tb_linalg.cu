// Small dense fp64 linear algebra on one CTA per matrix (d <= 128): Cholesky factor and inverse
// of the mode covariances with the reference's regularise-on-failure rule.
//   ref: tempest/modes.py:105-119 (ModeStatistics.__init__), tempest/student.py:75-79,
//        tempest/tools.py:101-110 (volume_variation regularisation / inverse)
#include "tb_common.cuh"
#include "tb_chol.cuh"

namespace {
using namespace tb;

__global__ void __launch_bounds__(128)
chol_inv_kernel(double* __restrict__ a, int d, double* __restrict__ chol, double* __restrict__ inv,
                int* __restrict__ info, double* __restrict__ norms) {
  extern __shared__ double sm[];
  double* A = sm;            // working copy / L
  double* Li = sm + d * d;   // L^{-1}
  double* mat = a + (size_t)blockIdx.x * d * d;
  __shared__ double red[40];
  int status = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) A[e] = mat[e];
    __syncthreads();
    if (chol_lower(A, d)) break;
    if (attempt == 1) { status = 2; break; }
    // reg = max(1e-6, 1e-6 * |trace|) on the diagonal (modes.py:115-117), written back to `a`
    double tr = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) tr += mat[i * d + i];
    tr = block_sum(tr, red);
    const double reg = fmax(1e-6, 1e-6 * fabs(tr));
    for (int i = threadIdx.x; i < d; i += blockDim.x) mat[i * d + i] += reg;
    status = 1;
    __syncthreads();
  }
  if (threadIdx.x == 0 && info) info[blockIdx.x] = status;
  if (status == 2) return;
  // zero the strict upper triangle (np.linalg.cholesky returns a lower-triangular matrix)
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) { int i = e / d, j = e - i * d; if (j > i) A[e] = 0.0; }
  __syncthreads();
  if (chol) for (int e = threadIdx.x; e < d * d; e += blockDim.x) chol[(size_t)blockIdx.x * d * d + e] = A[e];
  lower_inverse(A, Li, d);
  // inv = L^{-T} L^{-1}
  double fro_inv = 0.0, fro_a = 0.0;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e - i * d;
    double s = 0.0;
    const int k0 = i > j ? i : j;
    for (int k = k0; k < d; ++k) s += Li[k * d + i] * Li[k * d + j];
    if (inv) inv[(size_t)blockIdx.x * d * d + e] = s;
    fro_inv += s * s;
    const double av = mat[e];
    fro_a += av * av;
  }
  fro_inv = block_sum(fro_inv, red);
  fro_a = block_sum(fro_a, red);
  double tr = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) tr += mat[i * d + i];
  tr = block_sum(tr, red);
  if (threadIdx.x == 0 && norms) {
    norms[blockIdx.x * 3 + 0] = sqrt(fro_a);
    norms[blockIdx.x * 3 + 1] = sqrt(fro_inv);
    norms[blockIdx.x * 3 + 2] = tr;
  }
}

// Sigma = scatter/(M-1) * (M-1)/M + diag(scatter/M)/M   (student.py:63 with np.cov ddof=1, np.var ddof=0)
__global__ void student_sigma_kernel(const double* __restrict__ scatter, int d, double m_total, double* __restrict__ sigma) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d * d; e += gridDim.x * blockDim.x) {
    const int i = e / d, j = e - i * d;
    const double s = scatter[e];
    double v = ((s * (1.0 / (m_total - 1.0))) * (m_total - 1.0)) / m_total;   // np.cov multiplies by 1/(M-1)
    if (i == j) v += (1.0 / m_total) * (s / m_total);
    sigma[e] = v;
  }
}

// mean of the two middle order statistics per column (np.median on an even count, student.py:62)
__global__ void median_pair_kernel(const double* __restrict__ pair, int d, double* __restrict__ med) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) med[c] = (pair[2 * c] + pair[2 * c + 1]) * 0.5;
}

// cov += reg * I with reg = 1e-6 * trace(cov)   (tools.py:101-104)
__global__ void add_trace_reg_kernel(double* __restrict__ cov, int d, double factor) {
  __shared__ double tr;
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < d; ++i) t += cov[i * d + i]; tr = t; }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) cov[i * d + i] += factor * tr;
}

// volume_variation's rank test and regularisation without a host round trip (tools.py:101-110).  After a first
// chol_inv of a copy of cov (info / norms): flags[0] = "matrix_rank(cov) < d" (Cholesky failure, or
// |A|_F |A^-1|_F >= 1 / (d eps), an upper bound of cond_2); if set, cov += 1e-6 trace(cov) I.  work = cov.
__global__ void __launch_bounds__(128)
vv_regularise_kernel(double* __restrict__ cov, double* __restrict__ work, int d, const int* __restrict__ info,
                     const double* __restrict__ norms, int* __restrict__ flags) {
  __shared__ double tr;
  __shared__ int singular;
  if (threadIdx.x == 0) {
    const double n0 = norms[0], n1 = norms[1];
    singular = (info[0] != 0 || !isfinite(n0) || !isfinite(n1) || n0 * n1 >= 1.0 / ((double)d * 2.220446049250313e-16)) ? 1 : 0;
    flags[0] = singular;
    double t = 0.0;
    for (int i = 0; i < d; ++i) t += cov[i * d + i];
    tr = t;
  }
  __syncthreads();
  if (singular) for (int i = threadIdx.x; i < d; i += blockDim.x) cov[i * d + i] += 1e-6 * tr;
  __syncthreads();
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) work[e] = cov[e];
}

// cv = 0.5 sqrt(raw), or 1e10 when the regularised matrix could not be inverted either (tools.py:108-110)
__global__ void vv_finish_kernel(const double* __restrict__ raw, const int* __restrict__ info2, const double* __restrict__ norms2,
                                 const int* __restrict__ flags, double* __restrict__ result) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const bool failed = flags[0] && (info2[0] == 2 || !isfinite(norms2[0]) || !isfinite(norms2[1]));
    result[0] = (failed || info2[0] == 2) ? 1e10 : 0.5 * sqrt(raw[0]);
  }
}

// fp64 FMA peak probe for the roofline of the mutation kernel: eight independent register-resident DFMA chains
// per thread, nothing else in the loop (the denominator MEASURED_PEAKS.json does not carry)
__global__ void __launch_bounds__(256)
fp64_peak_kernel(int iters, double seed, double* __restrict__ out) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
  double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 1.2345e300) out[0] = r;     // keeps the chains alive without a store on the timed path
}

}  // namespace

extern "C" {

int tb_vv_regularise(double* cov, double* work, int32_t d, const int32_t* info, const double* norms, int32_t* flags,
                     tb_stream_t stream) {
  if (!cov || !work || d <= 0 || d > 128 || !info || !norms || !flags) return TB_ERR_ARG;
  vv_regularise_kernel<<<1, 128, 0, as_stream(stream)>>>(cov, work, d, info, norms, flags);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_vv_finish(const double* raw, const int32_t* info2, const double* norms2, const int32_t* flags, double* result,
                 tb_stream_t stream) {
  if (!raw || !info2 || !norms2 || !flags || !result) return TB_ERR_ARG;
  vv_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(raw, info2, norms2, flags, result);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int64_t tb_fp64_peak_flops(int32_t iters) {
  return (int64_t)2 * 8 * (int64_t)iters * 256 * (int64_t)tb::sm_count() * 8;
}

int tb_fp64_peak_run(int32_t iters, double* out, tb_stream_t stream) {
  if (iters <= 0 || !out) return TB_ERR_ARG;
  fp64_peak_kernel<<<tb::sm_count() * 8, 256, 0, as_stream(stream)>>>(iters, 1.0, out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_chol_inv(double* a, int32_t d, int32_t batch, double* chol, double* inv, int32_t* info, double* norms3,
                tb_stream_t stream) {
  if (!a || d <= 0 || d > 128 || batch <= 0) return TB_ERR_ARG;
  const size_t smem = sizeof(double) * 2 * (size_t)d * d;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  chol_inv_kernel<<<batch, 128, smem, as_stream(stream)>>>(a, d, chol, inv, info, norms3);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_student_sigma(const double* scatter, int32_t d, double m_total, double* sigma, tb_stream_t stream) {
  if (!scatter || !sigma || d <= 0 || m_total < 2.0) return TB_ERR_ARG;
  student_sigma_kernel<<<(d * d + 255) / 256, 256, 0, as_stream(stream)>>>(scatter, d, m_total, sigma);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_median_pairs(const double* pair, int32_t d, double* med, tb_stream_t stream) {
  if (!pair || !med || d <= 0) return TB_ERR_ARG;
  median_pair_kernel<<<(d + 127) / 128, 128, 0, as_stream(stream)>>>(pair, d, med);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_add_trace_reg(double* cov, int32_t d, double factor, tb_stream_t stream) {
  if (!cov || d <= 0) return TB_ERR_ARG;
  add_trace_reg_kernel<<<1, 128, 0, as_stream(stream)>>>(cov, d, factor);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // extern "C"
