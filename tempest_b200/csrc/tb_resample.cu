// Resampling kernels (a8, a9).
//   ref: tempest/steps/resample.py:52-99, tempest/tools.py:178-228,
//        numpy legacy RandomState.choice (cumsum -> /= last -> searchsorted 'right').
//
// The exact cumulative sum itself lives in tb_cdf.cu; this file holds the searches and the row gather.
#include "tb_common.cuh"

namespace {
using namespace tb;

__global__ void cdf_sequential_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ cdf) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double s = p[0];
  cdf[0] = s;
  for (int64_t i = 1; i < n; ++i) { s = __dadd_rn(s, p[i]); cdf[i] = s; }
}

// ---- searches / gather -----------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
search_right_kernel(const double* __restrict__ cdf, int64_t n, const double* __restrict__ draws, int64_t m,
                    int64_t* __restrict__ idx) {
  const double last = __ldg(cdf + n - 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double u = __ldg(draws + k);
    int64_t lo = 0, hi = n;  // first j with cdf_j/last > u
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (__ddiv_rn(__ldg(cdf + mid), last) <= u) lo = mid + 1; else hi = mid;
    }
    idx[k] = lo;
  }
}

// Guided search for many draws against one cdf: guide[g] = searchsorted(cdf/last, g/G, 'right') for the
// G+1 grid points g/G (G a power of two, so u*G and g/G are exact).  For g/G <= u < (g+1)/G the answer
// lies in [guide[g], guide[g+1]] by monotonicity, so the same comparison loop restricted to that bracket
// returns exactly the index of the full binary search, in ~2-3 probes instead of log2(n).
__global__ void __launch_bounds__(kBlock)
search_guide_kernel(const double* __restrict__ cdf, int64_t n, int64_t G, int64_t* __restrict__ guide) {
  const double last = __ldg(cdf + n - 1);
  const double invG = 1.0 / (double)G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= G; g += stride) {
    if (g == G) { guide[g] = n; continue; }
    const double u = (double)g * invG;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ddiv_rn(__ldg(cdf + mid), last) <= u) lo = mid + 1; else hi = mid;
    }
    guide[g] = lo;
  }
}

__global__ void __launch_bounds__(kBlock)
search_right_guided_kernel(const double* __restrict__ cdf, int64_t n, const double* __restrict__ draws, int64_t m,
                           const int64_t* __restrict__ guide, int64_t G, int64_t* __restrict__ idx) {
  const double last = __ldg(cdf + n - 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double u = __ldg(draws + k);
    int64_t lo = 0, hi = n;
    if (u >= 0.0 && u < 1.0) {
      int64_t g = (int64_t)(u * (double)G);
      if (g > G - 1) g = G - 1;
      lo = __ldg(guide + g);
      hi = __ldg(guide + g + 1);
    }
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ddiv_rn(__ldg(cdf + mid), last) <= u) lo = mid + 1; else hi = mid;
    }
    idx[k] = lo;
  }
}

// Sharded multinomial search.  The global cdf runs over generations, and inside a generation over
// ranks; this rank holds one contiguous SEGMENT per generation.  cdf[] is the rank-local cumulative
// sum; global value of local element j in segment s:  g(j) = cdf[j] + seg_shift[s]  (seg_shift folds
// "global offset of the segment" minus "local sum before it").  idx = local ancestor index, or -1
// when the draw's ancestor lives on another rank (seg_start[s] = global cdf value just before
// segment s decides ownership of a segment's first element).
__device__ __forceinline__ int seg_of(const int64_t* __restrict__ seg_begin, int S, int64_t j) {
  int lo = 0, hi = S;                  // last s with seg_begin[s] <= j
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(seg_begin + mid) <= j) lo = mid; else hi = mid; }
  return lo;
}

__global__ void __launch_bounds__(kBlock)
search_right_sharded_kernel(const double* __restrict__ cdf, int64_t n, const int64_t* __restrict__ seg_begin,
                            const double* __restrict__ seg_shift, const double* __restrict__ seg_start, int S,
                            double total, const double* __restrict__ draws, int64_t m, int64_t* __restrict__ idx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double u = __ldg(draws + k);
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int s = seg_of(seg_begin, S, mid);
      if (__ddiv_rn(__dadd_rn(__ldg(cdf + mid), __ldg(seg_shift + s)), total) <= u) lo = mid + 1; else hi = mid;
    }
    bool mine = lo < n;
    if (mine) {
      const int s = seg_of(seg_begin, S, lo);
      if (lo == __ldg(seg_begin + s)) mine = __ddiv_rn(__ldg(seg_start + s), total) <= u;
    }
    idx[k] = mine ? lo : -1;
  }
}

// guide table for the sharded search: guide[g] = first local j whose GLOBAL cdf value / total exceeds g/G
__global__ void __launch_bounds__(kBlock)
search_sharded_guide_kernel(const double* __restrict__ cdf, int64_t n, const int64_t* __restrict__ seg_begin,
                            const double* __restrict__ seg_shift, int S, double total, int64_t G,
                            int64_t* __restrict__ guide) {
  const double invG = 1.0 / (double)G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= G; g += stride) {
    if (g == G) { guide[g] = n; continue; }
    const double u = (double)g * invG;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int s = seg_of(seg_begin, S, mid);
      if (__ddiv_rn(__dadd_rn(__ldg(cdf + mid), __ldg(seg_shift + s)), total) <= u) lo = mid + 1; else hi = mid;
    }
    guide[g] = lo;
  }
}

__global__ void __launch_bounds__(kBlock)
search_right_sharded_guided_kernel(const double* __restrict__ cdf, int64_t n, const int64_t* __restrict__ seg_begin,
                                   const double* __restrict__ seg_shift, const double* __restrict__ seg_start, int S,
                                   double total, const double* __restrict__ draws, int64_t m,
                                   const int64_t* __restrict__ guide, int64_t G, int64_t* __restrict__ idx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double u = __ldg(draws + k);
    int64_t lo = 0, hi = n;
    if (u >= 0.0 && u < 1.0) {
      int64_t g = (int64_t)(u * (double)G);
      if (g > G - 1) g = G - 1;
      lo = __ldg(guide + g);
      hi = __ldg(guide + g + 1);
    }
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int s = seg_of(seg_begin, S, mid);
      if (__ddiv_rn(__dadd_rn(__ldg(cdf + mid), __ldg(seg_shift + s)), total) <= u) lo = mid + 1; else hi = mid;
    }
    bool mine = lo < n;
    if (mine) {
      const int s = seg_of(seg_begin, S, lo);
      if (lo == __ldg(seg_begin + s)) mine = __ddiv_rn(__ldg(seg_start + s), total) <= u;
    }
    idx[k] = mine ? lo : -1;
  }
}

__global__ void __launch_bounds__(kBlock)
systematic_kernel(const double* __restrict__ cdf, int64_t n, double u0, int64_t m, int64_t* __restrict__ idx,
                  int* __restrict__ overflow) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double pos = __ddiv_rn(__dadd_rn(u0, (double)k), (double)m);   // tools.py:217
    int64_t lo = 0, hi = n;  // first j with NOT(pos > cdf_j)
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (pos > __ldg(cdf + mid)) lo = mid + 1; else hi = mid;
    }
    if (lo >= n) { *overflow = 1; lo = n - 1; }
    idx[k] = lo;
  }
}

__global__ void __launch_bounds__(kBlock)
gather_rows_kernel(const double* __restrict__ hu, const double* __restrict__ hl, int d,
                   const int64_t* __restrict__ idx, int64_t m, double* __restrict__ au, double* __restrict__ al) {
  const int64_t total = m * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t k = e / d;
    const int j = (int)(e - k * d);
    const int64_t src = __ldg(idx + k);
    au[e] = __ldg(hu + src * d + j);
    if (j == 0 && al != nullptr) al[k] = __ldg(hl + src);
  }
}

}  // namespace

extern "C" {

int tb_cdf_sequential(const double* p, int64_t n, double* cdf, tb_stream_t stream) {
  if (n <= 0 || !p || !cdf) return TB_ERR_ARG;
  cdf_sequential_kernel<<<1, 32, 0, as_stream(stream)>>>(p, n, cdf);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_search_right(const double* cdf, int64_t n, const double* draws, int64_t m, int64_t* idx,
                    tb_stream_t stream) {
  if (n <= 0 || m < 0 || !cdf || (m > 0 && (!draws || !idx))) return TB_ERR_ARG;
  if (m == 0) return TB_OK;
  search_right_kernel<<<stream_grid(m, kBlock, 16), kBlock, 0, as_stream(stream)>>>(cdf, n, draws, m, idx);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_search_guide_bytes(int32_t bits) { return sizeof(int64_t) * (((size_t)1 << bits) + 1); }

int tb_search_right_guided(const double* cdf, int64_t n, const double* draws, int64_t m, void* guide, int32_t bits,
                           int64_t* idx, tb_stream_t stream) {
  if (n <= 0 || m < 0 || !cdf || !guide || bits < 4 || bits > 24 || (m > 0 && (!draws || !idx))) return TB_ERR_ARG;
  if (m == 0) return TB_OK;
  const int64_t G = (int64_t)1 << bits;
  cudaStream_t st = as_stream(stream);
  search_guide_kernel<<<stream_grid(G + 1, kBlock, 16), kBlock, 0, st>>>(cdf, n, G, (int64_t*)guide);
  search_right_guided_kernel<<<stream_grid(m, kBlock, 16), kBlock, 0, st>>>(cdf, n, draws, m, (const int64_t*)guide, G, idx);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_search_right_sharded(const double* cdf, int64_t n, const int64_t* seg_begin, const double* seg_shift,
                            const double* seg_start, int32_t n_seg, double total, const double* draws, int64_t m,
                            int64_t* idx, tb_stream_t stream) {
  if (n <= 0 || m < 0 || n_seg <= 0 || !cdf || !seg_begin || !seg_shift || !seg_start || (m > 0 && (!draws || !idx)))
    return TB_ERR_ARG;
  if (m == 0) return TB_OK;
  search_right_sharded_kernel<<<stream_grid(m, kBlock, 16), kBlock, 0, as_stream(stream)>>>(
      cdf, n, seg_begin, seg_shift, seg_start, n_seg, total, draws, m, idx);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_search_right_sharded_guided(const double* cdf, int64_t n, const int64_t* seg_begin, const double* seg_shift,
                                   const double* seg_start, int32_t n_seg, double total, const double* draws,
                                   int64_t m, void* guide, int32_t bits, int64_t* idx, tb_stream_t stream) {
  if (n <= 0 || m < 0 || n_seg <= 0 || !cdf || !seg_begin || !seg_shift || !seg_start || !guide || bits < 4 || bits > 24 ||
      (m > 0 && (!draws || !idx)))
    return TB_ERR_ARG;
  if (m == 0) return TB_OK;
  const int64_t G = (int64_t)1 << bits;
  cudaStream_t st = as_stream(stream);
  search_sharded_guide_kernel<<<stream_grid(G + 1, kBlock, 16), kBlock, 0, st>>>(cdf, n, seg_begin, seg_shift, n_seg, total,
                                                                                G, (int64_t*)guide);
  search_right_sharded_guided_kernel<<<stream_grid(m, kBlock, 16), kBlock, 0, st>>>(
      cdf, n, seg_begin, seg_shift, seg_start, n_seg, total, draws, m, (const int64_t*)guide, G, idx);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_systematic(const double* cdf, int64_t n, double u0, int64_t m, int64_t* idx, int32_t* overflow,
                  tb_stream_t stream) {
  if (n <= 0 || m <= 0 || !cdf || !idx || !overflow) return TB_ERR_ARG;
  cudaError_t e = cudaMemsetAsync(overflow, 0, sizeof(int32_t), as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  systematic_kernel<<<stream_grid(m, kBlock, 16), kBlock, 0, as_stream(stream)>>>(cdf, n, u0, m, idx, overflow);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_gather_rows(const double* hu, const double* hl, int32_t d, const int64_t* idx, int64_t m, double* au,
                   double* al, tb_stream_t stream) {
  if (d <= 0 || m < 0 || !hu || !idx || !au) return TB_ERR_ARG;
  if (m == 0) return TB_OK;
  gather_rows_kernel<<<stream_grid(m * d, kBlock, 16), kBlock, 0, as_stream(stream)>>>(hu, hl, d, idx, m, au, al);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // extern "C"
