// Weighted Gaussian-mixture EM, BIC bound and label prediction for the hierarchical clustering row
// of the hot path.
//   ref: tempest/cluster.py:56-133 (GaussianMixture.fit), :135-170 (weighted k-means++ init),
//        :172-193 (_e_step), :195-250 (_m_step / full covariances), :264-283 (_compute_lower_bound),
//        :285-308 (predict), :310-340 (bic), :377-383 (min-max normalisation), :633-696 (hierarchy predict)
//
// One fit lives in a device-resident PARAMETER BLOCK (doubles, layout below); an EM pass is
//   estep (densities -> responsibilities r*sw, masses R_k, bound of the current parameters)
//   -> column sums (tb_moments_partial, weights r*sw) -> means -> centred scatter -> commit
//   (covariance, Cholesky of Sigma + reg I, log normaliser).
// The reference evaluates the bound of the NEW parameters after each M-step and the next E-step
// re-evaluates the same densities; here pass p computes both from one evaluation at theta_p, so the
// convergence test of reference iteration p-1 happens inside pass p and a converged pass leaves
// theta_p untouched.  Passes enqueued after convergence return immediately (hdr[1] = done).
//
// Densities use the Cholesky factor of Sigma + reg I (scipy uses an eigendecomposition; same value up
// to rounding); no log-sum-exp shift -- the reference has none (cluster.py:182-191).
#include "tb_common.cuh"
#include "tb_chol.cuh"

namespace {
using namespace tb;

constexpr double kLog2Pi = 1.8378770664093454835606594728112353;
constexpr int kGmmHdr = 16;
constexpr int kEmMaxK = 2;        // the hierarchy only fits 1- and 2-component mixtures (cluster.py:462-477)
constexpr int kPredictMaxK = 64;

// hdr: [0] lower bound kept by the reference loop, [1] done, [2] completed passes (= n_iter when done),
//      [3] bound of the last evaluated parameters, [4] components whose Cholesky failed,
//      [7] result of a bound-only evaluation
struct GmmLayout {
  int d, K;
  __host__ __device__ GmmLayout(int d_, int K_) : d(d_), K(K_) {}
  __host__ __device__ int w() const { return kGmmHdr; }
  __host__ __device__ int mean() const { return w() + K; }
  __host__ __device__ int cov() const { return mean() + K * d; }
  __host__ __device__ int linv() const { return cov() + K * d * d; }
  __host__ __device__ int lognorm() const { return linv() + K * d * d; }
  __host__ __device__ int ok() const { return lognorm() + K; }
  __host__ __device__ int nw() const { return ok() + K; }
  __host__ __device__ int nmean() const { return nw() + K; }
  __host__ __device__ int R() const { return nmean() + K * d; }
  __host__ __device__ int M() const { return R() + K; }
  __host__ __device__ int S() const { return M() + K * d; }
  __host__ __device__ int total() const { return S() + K * d * d; }
};

struct GmmWs {
  unsigned int ticket;
  unsigned int pad[3];
  double partial[kMaxPartials][4];
};

enum { FLAG_INIT = 1, FLAG_BOUND = 2, FLAG_SKIP_BAD = 4 };

// |L^{-1} (x - mu)|^2 with the row either in registers (DT > 0) or re-read through L1 (DT == 0)
template <int DT>
__device__ __forceinline__ double maha_lower(const double* __restrict__ xr, const double* __restrict__ mu,
                                             const double* __restrict__ li, int d) {
  double acc = 0.0;
  if (DT > 0) {
    double df[DT > 0 ? DT : 1];
#pragma unroll
    for (int j = 0; j < DT; ++j) df[j] = xr[j] - __ldg(mu + j);
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      double y = 0.0;
#pragma unroll
      for (int j = 0; j <= i; ++j) y += __ldg(li + i * DT + j) * df[j];
      acc += y * y;
    }
  } else {
    for (int i = 0; i < d; ++i) {
      double y = 0.0;
      for (int j = 0; j <= i; ++j) y += __ldg(li + i * d + j) * (xr[j] - __ldg(mu + j));
      acc += y * y;
    }
  }
  return acc;
}

template <int DT>
__device__ __forceinline__ void load_row(const double* __restrict__ x, int64_t r, int d, const double* __restrict__ lo,
                                         const double* __restrict__ hi, double* row) {
  const int dd = DT > 0 ? DT : d;
  for (int j = 0; j < dd; ++j) {
    double v = __ldg(x + r * dd + j);
    if (lo) v = (v - __ldg(lo + j)) / ((__ldg(hi + j) - __ldg(lo + j)) + 1e-10);   // cluster.py:382
    row[j] = v;
  }
}

// ---- E-step / bound ------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kBlock)
gmm_estep_kernel(const double* __restrict__ x, const int64_t* __restrict__ rows, const double* __restrict__ sw,
                 int64_t n, int d, int K, double* __restrict__ blk, int flags, double tol, int max_iter,
                 double* __restrict__ wr, GmmWs* ws) {
  __shared__ double smem[40];
  const GmmLayout L(d, K);
  const bool bound_only = (flags & FLAG_BOUND) != 0, init = (flags & FLAG_INIT) != 0;
  if (!bound_only && !init) {
    // converged or out of iterations: nothing to do (every CTA takes the same branch)
    if (blk[1] != 0.0 || blk[2] >= (double)max_iter) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && blk[1] == 0.0) blk[1] = 1.0;
      return;
    }
  }
  const double eps = init ? 0.0 : 1e-10;
  const double unit = 1.0 / (double)n;
  double lb = 0.0, r_acc[kEmMaxK] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = rows ? __ldg(rows + i) : i;
    double rowbuf[DT > 0 ? DT : 1];
    const double* xr;
    if (DT > 0) { load_row<DT>(x, r, d, nullptr, nullptr, rowbuf); xr = rowbuf; } else xr = x + r * d;
    double dens[kEmMaxK] = {0.0, 0.0};
    double tot = 0.0, tot_lb = 0.0;
    for (int k = 0; k < K; ++k) {
      const double m2 = maha_lower<DT>(xr, blk + L.mean() + k * d, blk + L.linv() + k * d * d, d);
      const double dk = blk[L.w() + k] * exp(blk[L.lognorm() + k] - 0.5 * m2);
      dens[k] = dk;
      tot += dk;
      if (blk[L.ok() + k] != 0.0) tot_lb += dk;            // cluster.py:278-279: a rejected covariance is skipped
    }
    const double swi = sw ? __ldg(sw + i) : unit;
    lb += swi * log(tot_lb + 1e-10);                       // :282-283
    if (!bound_only) {
      const double den = tot + eps;                        // :191 (+1e-10) / :165 (initial, none)
      for (int k = 0; k < K; ++k) {
        const double a = (dens[k] / den) * swi;            // :200
        wr[(int64_t)k * n + i] = a;
        r_acc[k] += a;
      }
    }
  }
  lb = block_sum(lb, smem);
  const double r0 = block_sum(r_acc[0], smem);
  const double r1 = block_sum(r_acc[1], smem);
  if (threadIdx.x == 0) { double* p = ws->partial[blockIdx.x]; p[0] = lb; p[1] = r0; p[2] = r1; }
  if (last_block_arrives(&ws->ticket)) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      a += __ldcg(&ws->partial[i][0]);
      b += __ldcg(&ws->partial[i][1]);
      c += __ldcg(&ws->partial[i][2]);
    }
    a = block_sum(a, smem);
    b = block_sum(b, smem);
    c = block_sum(c, smem);
    if (threadIdx.x == 0) {
      if (bound_only) { blk[7] = a; return; }
      const double mass[kEmMaxK] = {b, c};
      if (!init) {
        const double passes = blk[2];
        blk[3] = a;
        if (passes >= 1.0) {
          if (a - blk[0] < tol) { blk[1] = 1.0; return; }   // :118-119, parameters stay theta_p
          blk[0] = a;                                       // :121
        }
      }
      double tot = 0.0;
      for (int k = 0; k < K; ++k) tot += mass[k];
      for (int k = 0; k < K; ++k) { blk[L.R() + k] = mass[k]; blk[L.nw() + k] = mass[k] / tot; }   // :203-204
    }
  }
}

// means = (sum_i r_ik sw_i x_i) / (mass_k + 1e-10)   (cluster.py:207-209)
__global__ void gmm_means_kernel(double* __restrict__ blk, int d, int K) {
  const GmmLayout L(d, K);
  if (blk[1] != 0.0) return;
  for (int e = threadIdx.x; e < K * d; e += blockDim.x) {
    const int k = e / d;
    blk[L.nmean() + e] = blk[L.M() + e] / (blk[L.R() + k] + 1e-10);
  }
}

// Cholesky of cov + reg I for component `k` of the block -> L^{-1}, log normaliser, ok flag.
// `fallback`: covariance used when the factorisation fails (cluster.py:185-188 reg I; :667-670 I).
__device__ void factor_component(double* blk, const GmmLayout& L, int k, double reg, double fallback, double* sm) {
  const int d = L.d;
  double* A = sm;
  double* Li = sm + d * d;
  const double* cov = blk + L.cov() + k * d * d;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) A[e] = cov[e] + ((e / d == e % d) ? reg : 0.0);
  __syncthreads();
  const bool ok = chol_lower(A, d);
  if (!ok) {
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) A[e] = (e / d == e % d) ? sqrt(fallback) : 0.0;
    __syncthreads();
  }
  lower_inverse(A, Li, d);
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) blk[L.linv() + k * d * d + e] = Li[e];
  if (threadIdx.x == 0) {
    double logdet = 0.0;
    for (int i = 0; i < d; ++i) logdet += log(A[i * d + i]);
    blk[L.lognorm() + k] = -0.5 * (d * kLog2Pi + 2.0 * logdet);
    blk[L.ok() + k] = ok ? 1.0 : 0.0;
  }
  __syncthreads();
}

// one CTA per component: adopt the M-step result and factor it
__global__ void __launch_bounds__(128)
gmm_commit_kernel(double* __restrict__ blk, int d, int K, double reg, int init) {
  extern __shared__ double sm[];
  const GmmLayout L(d, K);
  if (!init && blk[1] != 0.0) return;
  const int k = blockIdx.x;
  const double mass = blk[L.R() + k] + 1e-10;               // cluster.py:225
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) blk[L.cov() + k * d * d + e] = blk[L.S() + k * d * d + e] / mass;
  for (int e = threadIdx.x; e < d; e += blockDim.x) blk[L.mean() + k * d + e] = blk[L.nmean() + k * d + e];
  if (threadIdx.x == 0) blk[L.w() + k] = blk[L.nw() + k];
  __syncthreads();
  factor_component(blk, L, k, reg, reg, sm);
  if (k == 0 && threadIdx.x == 0) {
    if (init) { blk[0] = -INFINITY; blk[1] = 0.0; blk[2] = 0.0; blk[3] = -INFINITY; }
    else blk[2] += 1.0;
  }
}

__global__ void __launch_bounds__(128)
gmm_prepare_kernel(double* __restrict__ blk, int d, int K, double reg, double fallback) {
  extern __shared__ double sm[];
  const GmmLayout L(d, K);
  factor_component(blk, L, blockIdx.x, reg, fallback, sm);
}

// initial responsibilities exp(-0.5 |x - c_k|^2) (cluster.py:161-165) expressed as a unit mixture
__global__ void gmm_seed_kernel(double* __restrict__ blk, const double* __restrict__ centres, int d, int K) {
  const GmmLayout L(d, K);
  for (int e = threadIdx.x; e < K * d * d; e += blockDim.x) {
    const int q = e % (d * d);
    const double v = (q / d == q % d) ? 1.0 : 0.0;
    blk[L.linv() + e] = v;
    blk[L.cov() + e] = v;
  }
  for (int e = threadIdx.x; e < K * d; e += blockDim.x) blk[L.mean() + e] = centres[e];
  for (int e = threadIdx.x; e < K; e += blockDim.x) { blk[L.w() + e] = 1.0; blk[L.lognorm() + e] = 0.0; blk[L.ok() + e] = 1.0; }
  if (threadIdx.x == 0) { blk[0] = -INFINITY; blk[1] = 0.0; blk[2] = 0.0; }
}

// ---- weighted k-means++ pieces (cluster.py:139-158) ------------------------------------------
// p_i = min_c |x_i - c|^2 * sw_i over the first `k` centres
__global__ void __launch_bounds__(kBlock)
kpp_prob_kernel(const double* __restrict__ x, const int64_t* __restrict__ rows, const double* __restrict__ sw,
                int64_t n, int d, const double* __restrict__ centres, int k, double* __restrict__ p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = rows ? __ldg(rows + i) : i;
    double best = INFINITY;
    for (int c = 0; c < k; ++c) {
      double s = 0.0;
      for (int j = 0; j < d; ++j) { const double t = __ldg(x + r * d + j) - __ldg(centres + c * d + j); s += t * t; }
      best = fmin(best, s);
    }
    p[i] = best * __ldg(sw + i);
  }
}

// j = searchsorted(run, frac * run[n-1]) (left) ; centre = x[rows[j]]
__global__ void kpp_pick_kernel(const double* __restrict__ run, int64_t n, double frac, const double* __restrict__ x,
                                const int64_t* __restrict__ rows, int d, double* __restrict__ centre,
                                int64_t* __restrict__ picked) {
  __shared__ int64_t js;
  if (threadIdx.x == 0) {
    const double r = __dmul_rn(frac, run[n - 1]);
    int64_t lo = 0, hi = n;                     // first j with run[j] >= r
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (run[mid] < r) lo = mid + 1; else hi = mid; }
    if (lo >= n) lo = n - 1;                     // numpy would index out of range; cannot happen for frac < 1
    js = lo;
    if (picked) *picked = lo;
  }
  __syncthreads();
  const int64_t r = rows ? rows[js] : js;
  for (int j = threadIdx.x; j < d; j += blockDim.x) centre[j] = x[r * d + j];
}

// ---- prediction ----------------------------------------------------------------------------------
// label_i = argmax_k log(w_k + 1e-10) + log N(x_i; mu_k, Sigma_k + reg I)   (first maximum wins, np.argmax)
template <int DT>
__global__ void __launch_bounds__(kBlock)
gmm_predict_kernel(const double* __restrict__ x, const int64_t* __restrict__ rows, int64_t n, int d, int K,
                   const double* __restrict__ blk, const double* __restrict__ lo, const double* __restrict__ hi,
                   int flags, int32_t* __restrict__ labels) {
  const GmmLayout L(d, K);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = rows ? __ldg(rows + i) : i;
    double rowbuf[DT > 0 ? DT : 128];
    load_row<DT>(x, r, d, lo, hi, rowbuf);
    double best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
      double s;
      if ((flags & FLAG_SKIP_BAD) && blk[L.ok() + k] == 0.0) s = -INFINITY;    // cluster.py:305-306
      else {
        const double m2 = maha_lower<DT>(rowbuf, blk + L.mean() + k * d, blk + L.linv() + k * d * d, d);
        s = log(blk[L.w() + k] + 1e-10) + (blk[L.lognorm() + k] - 0.5 * m2);
      }
      if (s > best || (k == 0 && !(s < best))) { best = s; arg = k; }
    }
    labels[i] = arg;
  }
}

// ---- column minima / maxima and the normalised gather ---------------------------------------------
__global__ void __launch_bounds__(kBlock)
col_minmax_kernel(const double* __restrict__ x, const int64_t* __restrict__ rows, int64_t n, int d, double* __restrict__ part,
                  unsigned int* ticket, double* __restrict__ out_lo, double* __restrict__ out_hi) {
  __shared__ double slo[kBlock], shi[kBlock];
  const int rpb = kBlock / d, act = rpb * d;
  const int col = threadIdx.x % d, r0 = threadIdx.x / d;
  double lo = INFINITY, hi = -INFINITY;
  if (threadIdx.x < act) {
    for (int64_t j = (int64_t)blockIdx.x * rpb + r0; j < n; j += (int64_t)gridDim.x * rpb) {
      const int64_t r = rows ? __ldg(rows + j) : j;
      const double v = __ldg(x + r * d + col);
      lo = fmin(lo, v);
      hi = fmax(hi, v);
    }
  }
  slo[threadIdx.x] = lo;
  shi[threadIdx.x] = hi;
  __syncthreads();
  if (threadIdx.x < d) {
    for (int q = 1; q < rpb; ++q) { lo = fmin(lo, slo[q * d + threadIdx.x]); hi = fmax(hi, shi[q * d + threadIdx.x]); }
    part[((size_t)blockIdx.x * 2) * d + threadIdx.x] = lo;
    part[((size_t)blockIdx.x * 2 + 1) * d + threadIdx.x] = hi;
  }
  if (last_block_arrives(ticket)) {
    if (threadIdx.x < d) {
      lo = INFINITY; hi = -INFINITY;
      for (int b = 0; b < (int)gridDim.x; ++b) {
        lo = fmin(lo, __ldcg(part + ((size_t)b * 2) * d + threadIdx.x));
        hi = fmax(hi, __ldcg(part + ((size_t)b * 2 + 1) * d + threadIdx.x));
      }
      out_lo[threadIdx.x] = lo;
      out_hi[threadIdx.x] = hi;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
gather_norm_kernel(const double* __restrict__ x, const int64_t* __restrict__ rows, int64_t n, int d,
                   const double* __restrict__ lo, const double* __restrict__ hi, double* __restrict__ out) {
  const int64_t total = n * d, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / d;
    const int j = (int)(e - i * d);
    const int64_t r = rows ? __ldg(rows + i) : i;
    double v = __ldg(x + r * d + j);
    if (lo) v = (v - __ldg(lo + j)) / ((__ldg(hi + j) - __ldg(lo + j)) + 1e-10);
    out[e] = v;
  }
}

__global__ void __launch_bounds__(kBlock)
take_kernel(const double* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __ldg(src + __ldg(idx + i));
}

// ---- ordered split of a member list by label (cluster.py:495-496) ------------------------------------
constexpr int kSplitItems = 2048;   // members per CTA
__global__ void __launch_bounds__(kBlock)
split_count_kernel(const int32_t* __restrict__ labels, int64_t n, int64_t* __restrict__ block_ones) {
  __shared__ double smem[40];
  const int64_t base = (int64_t)blockIdx.x * kSplitItems;
  double c = 0.0;
  for (int t = threadIdx.x; t < kSplitItems; t += kBlock) { const int64_t i = base + t; if (i < n && labels[i] != 0) c += 1.0; }
  c = block_sum(c, smem);
  if (threadIdx.x == 0) block_ones[blockIdx.x] = (int64_t)c;
}
__global__ void split_offsets_kernel(int64_t* __restrict__ block_ones, int nb, int64_t* __restrict__ total_ones) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t run = 0;
    for (int b = 0; b < nb; ++b) { const int64_t c = block_ones[b]; block_ones[b] = run; run += c; }
    *total_ones = run;
  }
}
__global__ void __launch_bounds__(kBlock)
split_emit_kernel(const int32_t* __restrict__ labels, const int64_t* __restrict__ members, int64_t n,
                  const int64_t* __restrict__ block_ones, int64_t* __restrict__ out_zero, int64_t* __restrict__ out_one) {
  __shared__ int warp_ones[kBlock / 32 + 1];
  __shared__ int64_t run_one;
  const int64_t base = (int64_t)blockIdx.x * kSplitItems;
  if (threadIdx.x == 0) run_one = block_ones[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t0 = 0; t0 < kSplitItems; t0 += kBlock) {
    const int64_t i = base + t0 + threadIdx.x;
    const bool live = i < n;
    const bool one = live && labels[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, one);
    if (lane == 0) warp_ones[wid] = __popc(m);
    __syncthreads();
    int before = 0, all = 0;
    for (int q = 0; q < kBlock / 32; ++q) { if (q < wid) before += warp_ones[q]; all += warp_ones[q]; }
    const int64_t ones_before = run_one + before + __popc(m & ((1u << lane) - 1u));
    if (live) {
      const int64_t v = members ? members[i] : i;
      if (one) out_one[ones_before] = v; else out_zero[i - ones_before] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) run_one += all;
    __syncthreads();
  }
}

template <int DT>
int launch_estep(const double* x, const int64_t* rows, const double* sw, int64_t n, int d, int K, double* blk, int flags,
                 double tol, int max_iter, double* wr, GmmWs* ws, cudaStream_t st) {
  gmm_estep_kernel<DT><<<stream_grid(n, kBlock, 4), kBlock, 0, st>>>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, ws);
  return 0;
}
int estep_dispatch(const double* x, const int64_t* rows, const double* sw, int64_t n, int d, int K, double* blk, int flags,
                   double tol, int max_iter, double* wr, void* ws, cudaStream_t st) {
  GmmWs* w = (GmmWs*)ws;
  switch (d) {
    case 1: return launch_estep<1>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 2: return launch_estep<2>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 3: return launch_estep<3>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 4: return launch_estep<4>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 5: return launch_estep<5>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 6: return launch_estep<6>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 8: return launch_estep<8>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    case 10: return launch_estep<10>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
    default: return launch_estep<0>(x, rows, sw, n, d, K, blk, flags, tol, max_iter, wr, w, st);
  }
}

size_t factor_smem(int d) { return sizeof(double) * 2 * (size_t)d * d; }
template <typename Kern>
int allow_smem(Kern kern, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace

extern "C" {

int64_t tb_gmm_block_doubles(int32_t d, int32_t k) { return GmmLayout(d, k).total(); }

int tb_gmm_offsets(int32_t d, int32_t k, int64_t* out8) {
  if (!out8 || d <= 0 || k <= 0) return TB_ERR_ARG;
  const GmmLayout L(d, k);
  out8[0] = L.w(); out8[1] = L.mean(); out8[2] = L.cov(); out8[3] = L.linv();
  out8[4] = L.lognorm(); out8[5] = L.ok(); out8[6] = L.R(); out8[7] = L.total();
  return TB_OK;
}

size_t tb_gmm_workspace_bytes(void) { return sizeof(GmmWs); }

int tb_gmm_prepare(double* block, int32_t d, int32_t k, double reg, double fallback, tb_stream_t stream) {
  if (!block || d <= 0 || d > 128 || k <= 0 || k > kPredictMaxK) return TB_ERR_ARG;
  if (int e = allow_smem(gmm_prepare_kernel, factor_smem(d))) return e;
  gmm_prepare_kernel<<<k, 128, factor_smem(d), as_stream(stream)>>>(block, d, k, reg, fallback);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_gmm_bound(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, int32_t k,
                 double* block, void* workspace, tb_stream_t stream) {
  if (!x || !block || !workspace || n <= 0 || d <= 0 || d > 128 || k <= 0 || k > kEmMaxK) return TB_ERR_ARG;
  estep_dispatch(x, rows, sw, n, d, k, block, FLAG_BOUND, 0.0, 0, nullptr, workspace, as_stream(stream));
  TB_CHECK_LAUNCH();
  return TB_OK;
}

// M-step after an E-step whose r*sw columns are in `wr`: column sums, means, centred scatter, commit
static int gmm_mstep(const double* x, const int64_t* rows, int64_t n, int d, int k, double* block, const double* wr,
                     void* mom_ws, double reg, int init, tb_stream_t stream) {
  const GmmLayout L(d, k);
  for (int c = 0; c < k; ++c) {
    int rc = tb_moments_partial(x, rows, wr + (int64_t)c * n, nullptr, n, d, 1.0, 1, 0, mom_ws, block + L.M() + c * d,
                                nullptr, stream);
    if (rc) return rc;
  }
  gmm_means_kernel<<<1, 128, 0, as_stream(stream)>>>(block, d, k);
  for (int c = 0; c < k; ++c) {
    int rc = tb_moments_partial(x, rows, wr + (int64_t)c * n, nullptr, n, d, 1.0, 0, 1, mom_ws, block + L.nmean() + c * d,
                                block + L.S() + c * d * d, stream);
    if (rc) return rc;
  }
  if (int e = allow_smem(gmm_commit_kernel, factor_smem(d))) return e;
  gmm_commit_kernel<<<k, 128, factor_smem(d), as_stream(stream)>>>(block, d, k, reg, init);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_gmm_init(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, int32_t k,
                const double* centres, double* block, double* wr, void* workspace, void* mom_workspace, double reg,
                tb_stream_t stream) {
  if (!x || !sw || !centres || !block || !wr || !workspace || !mom_workspace || n <= 0 || d <= 0 || d > 128 || k <= 0 ||
      k > kEmMaxK)
    return TB_ERR_ARG;
  gmm_seed_kernel<<<1, 128, 0, as_stream(stream)>>>(block, centres, d, k);
  estep_dispatch(x, rows, sw, n, d, k, block, FLAG_INIT, 0.0, 0, wr, workspace, as_stream(stream));
  TB_CHECK_LAUNCH();
  return gmm_mstep(x, rows, n, d, k, block, wr, mom_workspace, reg, 1, stream);
}

int tb_gmm_em(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, int32_t k, double* block,
              double* wr, void* workspace, void* mom_workspace, double reg, double tol, int32_t max_iter,
              int32_t passes, tb_stream_t stream) {
  if (!x || !sw || !block || !wr || !workspace || !mom_workspace || n <= 0 || d <= 0 || d > 128 || k <= 0 ||
      k > kEmMaxK || passes < 0)
    return TB_ERR_ARG;
  for (int p = 0; p < passes; ++p) {
    estep_dispatch(x, rows, sw, n, d, k, block, 0, tol, max_iter, wr, workspace, as_stream(stream));
    TB_CHECK_LAUNCH();
    int rc = gmm_mstep(x, rows, n, d, k, block, wr, mom_workspace, reg, 0, stream);
    if (rc) return rc;
  }
  return TB_OK;
}

int tb_kpp_prob(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, const double* centres,
                int32_t k, double* p, tb_stream_t stream) {
  if (!x || !sw || !centres || !p || n <= 0 || d <= 0 || k <= 0) return TB_ERR_ARG;
  kpp_prob_kernel<<<stream_grid(n, kBlock, 8), kBlock, 0, as_stream(stream)>>>(x, rows, sw, n, d, centres, k, p);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_kpp_pick(const double* run, int64_t n, double frac, const double* x, const int64_t* rows, int32_t d,
                double* centre, int64_t* picked, tb_stream_t stream) {
  if (!run || !x || !centre || n <= 0 || d <= 0) return TB_ERR_ARG;
  kpp_pick_kernel<<<1, 128, 0, as_stream(stream)>>>(run, n, frac, x, rows, d, centre, picked);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_gmm_predict(const double* x, const int64_t* rows, int64_t n, int32_t d, int32_t k, const double* block,
                   const double* lo, const double* hi, int32_t skip_bad, int32_t* labels, tb_stream_t stream) {
  if (!x || !block || !labels || n <= 0 || d <= 0 || d > 128 || k <= 0 || k > kPredictMaxK || ((lo == nullptr) != (hi == nullptr)))
    return TB_ERR_ARG;
  const int flags = skip_bad ? FLAG_SKIP_BAD : 0;
  const int grid = stream_grid(n, kBlock, 4);
  cudaStream_t st = as_stream(stream);
  switch (d) {
    case 1: gmm_predict_kernel<1><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 2: gmm_predict_kernel<2><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 3: gmm_predict_kernel<3><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 4: gmm_predict_kernel<4><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 5: gmm_predict_kernel<5><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 6: gmm_predict_kernel<6><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 8: gmm_predict_kernel<8><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    case 10: gmm_predict_kernel<10><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
    default: gmm_predict_kernel<0><<<grid, kBlock, 0, st>>>(x, rows, n, d, k, block, lo, hi, flags, labels); break;
  }
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_col_minmax_workspace_bytes(int32_t d) { return 256 + sizeof(double) * 2 * (size_t)d * tb::sm_count() * 4; }

int tb_col_minmax(const double* x, const int64_t* rows, int64_t n, int32_t d, void* workspace, double* lo, double* hi,
                  tb_stream_t stream) {
  if (!x || !workspace || !lo || !hi || n <= 0 || d <= 0 || d > 128) return TB_ERR_ARG;
  const int rpb = kBlock / d;
  const int grid = stream_grid(n, rpb * 8, 4);
  col_minmax_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(x, rows, n, d, (double*)((char*)workspace + 256),
                                                            (unsigned int*)workspace, lo, hi);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_gather_normalised(const double* x, const int64_t* rows, int64_t n, int32_t d, const double* lo, const double* hi,
                         double* out, tb_stream_t stream) {
  if (!x || !out || n <= 0 || d <= 0 || ((lo == nullptr) != (hi == nullptr))) return TB_ERR_ARG;
  gather_norm_kernel<<<stream_grid(n * d, kBlock, 8), kBlock, 0, as_stream(stream)>>>(x, rows, n, d, lo, hi, out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_take(const double* src, const int64_t* idx, int64_t n, double* out, tb_stream_t stream) {
  if (!src || !idx || !out || n <= 0) return TB_ERR_ARG;
  take_kernel<<<stream_grid(n, kBlock, 8), kBlock, 0, as_stream(stream)>>>(src, idx, n, out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_split_workspace_bytes(int64_t n) { return sizeof(int64_t) * (size_t)((n + kSplitItems - 1) / kSplitItems + 1); }

int tb_split_by_label(const int32_t* labels, const int64_t* members, int64_t n, void* workspace, int64_t* out_zero,
                      int64_t* out_one, int64_t* n_one, tb_stream_t stream) {
  if (!labels || !workspace || !out_zero || !out_one || !n_one || n <= 0) return TB_ERR_ARG;
  const int nb = (int)((n + kSplitItems - 1) / kSplitItems);
  int64_t* block_ones = (int64_t*)workspace;
  cudaStream_t st = as_stream(stream);
  split_count_kernel<<<nb, kBlock, 0, st>>>(labels, n, block_ones);
  split_offsets_kernel<<<1, 32, 0, st>>>(block_ones, nb, n_one);
  split_emit_kernel<<<nb, kBlock, 0, st>>>(labels, members, n, block_ones, out_zero, out_one);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // extern "C"
