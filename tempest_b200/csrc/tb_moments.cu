// Weighted moments, volume variation, trim_weights building blocks, order statistics.
//   ref: tempest/tools.py:10-55 (trim_weights), :58-117 (volume_variation),
//        tempest/student.py:62-63 (median / covariance of the 4n-row resampled set),
//        tempest/modes.py:221-288 (from_global)
//
// HBM layout: u[N_total][d] row-major fp64, w[N_total] fp64.  All reductions publish one
// partial per CTA and the last CTA to finish merges them in a fixed order (deterministic).
#include "tb_common.cuh"
#include <string.h>

namespace {
using namespace tb;

struct ReduceWs {
  unsigned int ticket;
  unsigned int pad[3];
  double partial[kMaxPartials][4];
};

// ---- sum / normalise ------------------------------------------------------------------
// mode 0: out = {sum w, sum w^2, max w}   (no write)
// mode 1: masked: only w >= thr contribute; out = {count, sum w, sum w^2}
template <int MODE>
__global__ void __launch_bounds__(kBlock)
reduce3_kernel(const double* __restrict__ w, int64_t n, double thr, ReduceWs* ws, double* __restrict__ out) {
  __shared__ double smem[40];
  double a = 0.0, b = 0.0, c = (MODE == 0) ? -INFINITY : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = __ldg(w + i);
    if (MODE == 0) { a += v; b += v * v; c = fmax(c, v); }
    else if (v >= thr) { c += 1.0; a += v; b += v * v; }
  }
  a = block_sum(a, smem);
  b = block_sum(b, smem);
  c = (MODE == 0) ? block_max(c, smem) : block_sum(c, smem);
  if (threadIdx.x == 0) { double* p = ws->partial[blockIdx.x]; p[0] = a; p[1] = b; p[2] = c; }
  if (last_block_arrives(&ws->ticket)) {
    a = 0.0; b = 0.0; c = (MODE == 0) ? -INFINITY : 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      a += __ldcg(&ws->partial[i][0]); b += __ldcg(&ws->partial[i][1]);
      double cc = __ldcg(&ws->partial[i][2]);
      c = (MODE == 0) ? fmax(c, cc) : c + cc;
    }
    a = block_sum(a, smem);
    b = block_sum(b, smem);
    c = (MODE == 0) ? block_max(c, smem) : block_sum(c, smem);
    if (threadIdx.x == 0) {
      if (MODE == 0) { out[0] = a; out[1] = b; out[2] = c; }
      else { out[0] = c; out[1] = a; out[2] = b; }
    }
  }
}

// w /= stats[0]; then stats <- {old sum, sum (w/sum)^2, max (w/sum)}  (second reduction fused)
__global__ void __launch_bounds__(kBlock)
scale_reduce_kernel(double* __restrict__ w, int64_t n, const double* __restrict__ sum_in, ReduceWs* ws,
                    double* __restrict__ out) {
  __shared__ double smem[40];
  const double s = sum_in[0];
  double b = 0.0, c = -INFINITY;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = w[i] / s;   // tools.py:36
    w[i] = v;
    b += v * v;
    c = fmax(c, v);
  }
  b = block_sum(b, smem);
  c = block_max(c, smem);
  if (threadIdx.x == 0) { double* p = ws->partial[blockIdx.x]; p[1] = b; p[2] = c; }
  if (last_block_arrives(&ws->ticket)) {
    b = 0.0; c = -INFINITY;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      b += __ldcg(&ws->partial[i][1]);
      c = fmax(c, __ldcg(&ws->partial[i][2]));
    }
    b = block_sum(b, smem);
    c = block_max(c, smem);
    if (threadIdx.x == 0) { out[0] = s; out[1] = b; out[2] = c; }
  }
}

__global__ void __launch_bounds__(kBlock) scale_dev_kernel(double* __restrict__ w, int64_t n, const double* __restrict__ denom_p) {
  const double denom = denom_p[0];       // the (all-reduced) sum lives on the device: no host round trip for it
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) w[i] = w[i] / denom;
}
__global__ void __launch_bounds__(kBlock) scale_kernel(double* __restrict__ w, int64_t n, double denom) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) w[i] = w[i] / denom;
}

// ---- binade histogram -------------------------------------------------------------------
// bin = biased exponent (0..2047) of each non-negative weight; per-bin count, sum w, sum w^2.
// CTA-private shared histograms (uniform warps take a shuffle-reduced fast path), one global
// atomic flush per non-empty bin per CTA.
__global__ void __launch_bounds__(kBlock)
binade_hist_kernel(const double* __restrict__ w, int64_t n, unsigned long long* __restrict__ gcount,
                   double* __restrict__ gs1, double* __restrict__ gs2) {
  __shared__ unsigned int cnt[2048];
  __shared__ double s1[2048];
  __shared__ double s2[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) { cnt[i] = 0u; s1[i] = 0.0; s2[i] = 0.0; }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    const bool on = i < n;
    const double v = on ? __ldg(w + i) : 0.0;
    const int bin = on ? (int)((__double_as_longlong(v) >> 52) & 0x7ff) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (peers == 0xffffffffu && bin >= 0) {       // whole warp in one binade: shuffle-reduce first
      const double a = warp_sum(v), b = warp_sum(v * v);
      if (lane == 0) { atomicAdd(&cnt[bin], 32u); atomicAdd(&s1[bin], a); atomicAdd(&s2[bin], b); }
    } else if (bin >= 0) {
      atomicAdd(&cnt[bin], 1u); atomicAdd(&s1[bin], v); atomicAdd(&s2[bin], v * v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    if (cnt[i]) {
      atomicAdd(&gcount[i], (unsigned long long)cnt[i]);
      atomicAdd(&gs1[i], s1[i]);
      atomicAdd(&gs2[i], s2[i]);
    }
  }
}

// second level: the 2048 linear sub-bins (top 11 mantissa bits) of ONE binade, same three sums per bin.
// Narrows the trim bracket from a binade to 1/2048 of it, so that only the one or two percentile grid
// points inside the flipping sub-bin need an exact evaluation.
__global__ void __launch_bounds__(kBlock)
subbin_hist_kernel(const double* __restrict__ w, int64_t n, int binade, unsigned long long* __restrict__ gcount,
                   double* __restrict__ gs1, double* __restrict__ gs2) {
  __shared__ unsigned int cnt[2048];
  __shared__ double s1[2048];
  __shared__ double s2[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) { cnt[i] = 0u; s1[i] = 0.0; s2[i] = 0.0; }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = __ldg(w + i);
    const long long bits = __double_as_longlong(v);
    if ((int)((bits >> 52) & 0x7ff) != binade) continue;
    const int sub = (int)((bits >> 41) & 0x7ff);
    atomicAdd(&cnt[sub], 1u); atomicAdd(&s1[sub], v); atomicAdd(&s2[sub], v * v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    if (cnt[i]) {
      atomicAdd(&gcount[i], (unsigned long long)cnt[i]);
      atomicAdd(&gs1[i], s1[i]);
      atomicAdd(&gs2[i], s2[i]);
    }
  }
}

// ---- ordered compaction of {w >= thr} ---------------------------------------------------
constexpr int kCompactItems = 2048;  // elements per CTA
__global__ void __launch_bounds__(kBlock)
compact_count_kernel(const double* __restrict__ w, int64_t n, double thr, int64_t* __restrict__ block_count) {
  __shared__ double smem[40];
  const int64_t base = (int64_t)blockIdx.x * kCompactItems;
  double c = 0.0;
  for (int k = threadIdx.x; k < kCompactItems; k += blockDim.x) {
    int64_t i = base + k;
    if (i < n && __ldg(w + i) >= thr) c += 1.0;
  }
  c = block_sum(c, smem);
  if (threadIdx.x == 0) block_count[blockIdx.x] = (int64_t)c;
}

__global__ void __launch_bounds__(1024)
compact_scan_kernel(int64_t* __restrict__ block_count, int64_t nb, int64_t* __restrict__ n_out) {
  __shared__ long long part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (nb + blockDim.x - 1) / blockDim.x;
  const int64_t lo = (int64_t)tid * per, hi = (lo + per < nb) ? lo + per : nb;
  long long acc = 0;
  for (int64_t t = lo; t < hi; ++t) acc += block_count[t];
  part[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    long long run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { long long v = part[i]; part[i] = run; run += v; }
    *n_out = run;
  }
  __syncthreads();
  long long pre = part[tid];
  for (int64_t t = lo; t < hi; ++t) { long long v = block_count[t]; block_count[t] = pre; pre += v; }
}

__global__ void __launch_bounds__(kBlock)
compact_emit_kernel(const double* __restrict__ w, int64_t n, double thr, double denom,
                    const int64_t* __restrict__ block_off, int64_t* __restrict__ idx_out, double* __restrict__ w_out) {
  __shared__ int warp_cnt[kBlock / 32];
  __shared__ int base_sh;
  const int64_t base = (int64_t)blockIdx.x * kCompactItems;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_sh = 0;
  __syncthreads();
  int64_t out0 = block_off[blockIdx.x];
  for (int k0 = 0; k0 < kCompactItems; k0 += blockDim.x) {
    const int64_t i = base + k0 + threadIdx.x;
    const double v = (i < n) ? __ldg(w + i) : -1.0;
    const bool keep = (i < n) && (v >= thr);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int off = base_sh;
    for (int q = 0; q < wid; ++q) off += warp_cnt[q];
    if (keep) {
      const int64_t o = out0 + off + __popc(bal & ((1u << lane) - 1u));
      if (idx_out) idx_out[o] = i;
      if (w_out) w_out[o] = v / denom;   // tools.py:49
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < kBlock / 32; ++q) t += warp_cnt[q]; base_sh += t; }
    __syncthreads();
  }
}

// ---- exact order statistics: 6 MSD radix levels of 11 bits --------------------------------
constexpr int kSelBins = 2048;
constexpr int kSelLevels = 6;
struct SelState {             // per (column, rank slot)
  unsigned long long prefix;  // resolved high bits (low bits zero)
  long long rank;             // remaining rank inside the prefix class
};
__host__ __device__ inline int sel_shift(int level) { int s = 64 - 11 * (level + 1); return s < 0 ? 0 : s; }
__host__ __device__ inline int sel_bits(int level) { return (level == kSelLevels - 1) ? 9 : 11; }

__global__ void sel_init_kernel(SelState* st, const int64_t* __restrict__ ranks, int ncols, int nranks) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncols * nranks) { st[i].prefix = 0ull; st[i].rank = ranks[i % nranks]; }
}

__global__ void __launch_bounds__(kBlock)
sel_hist_kernel(const double* __restrict__ base, const int64_t* __restrict__ rows, int64_t stride_elems,
                int64_t n, int ncols, const int* __restrict__ mult, const SelState* __restrict__ st, int nranks,
                int level, unsigned int* __restrict__ hist /* [ncols][nranks][kSelBins] */) {
  const int shift = sel_shift(level);
  const int bits = sel_bits(level);
  const int hi_shift = shift + bits;  // bits above this are the resolved prefix
  const int64_t total = n * ncols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x; e0 < total; e0 += stride) {
    const int64_t e = e0 + threadIdx.x;
    int key_slot[4] = {-1, -1, -1, -1};
    unsigned int m = 0;
    int c = 0;
    if (e < total) {
      const int64_t j = e / ncols;
      c = (int)(e - j * ncols);
      m = mult ? (unsigned int)__ldg(mult + j) : 1u;
      if (m) {
        const int64_t r = rows ? __ldg(rows + j) : j;
        const unsigned long long key = (unsigned long long)__double_as_longlong(__ldg(base + r * stride_elems + c));
        const int digit = (int)((key >> shift) & ((1u << bits) - 1u));
        for (int q = 0; q < nranks; ++q) {
          const unsigned long long pre = st[c * nranks + q].prefix;
          const bool match = (hi_shift >= 64) ? true : ((key >> hi_shift) == (pre >> hi_shift));
          if (match) key_slot[q] = ((c * nranks + q) << 11) | digit;
        }
      }
    }
    for (int q = 0; q < nranks; ++q) {
      // skip a slot whose class is identical to the previous slot's (shared histogram not needed:
      // each slot owns its histogram; duplicates simply count twice, once per slot)
      const int k = key_slot[q];
      const unsigned peers = __match_any_sync(0xffffffffu, k);
      if (k >= 0) {
        // warp-aggregated add: the lowest lane of each equal-key group adds the group's total
        const int leader = __ffs(peers) - 1;
        const unsigned int sum = __reduce_add_sync(peers, m);
        if ((threadIdx.x & 31) == leader) atomicAdd(&hist[k], sum);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
sel_pick_kernel(SelState* st, const unsigned int* __restrict__ hist, int level) {
  // one CTA per (column, rank slot): find the bin where the cumulative count passes the rank
  __shared__ unsigned long long part[256];
  const unsigned int* h = hist + (size_t)blockIdx.x * kSelBins;
  const int per = kSelBins / 256;
  unsigned long long acc = 0;
  for (int i = 0; i < per; ++i) acc += h[threadIdx.x * per + i];
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    SelState s = st[blockIdx.x];
    unsigned long long run = 0;
    int t = 0;
    for (; t < 256; ++t) { if (run + part[t] > (unsigned long long)s.rank) break; run += part[t]; }
    if (t == 256) { t = 255; run -= part[255]; }   // rank beyond the population: clamp to the last bin
    int b = t * per;
    for (; b < t * per + per - 1; ++b) { if (run + h[b] > (unsigned long long)s.rank) break; run += h[b]; }
    s.prefix |= ((unsigned long long)b) << sel_shift(level);
    s.rank -= (long long)run;
    st[blockIdx.x] = s;
  }
}

__global__ void sel_finish_kernel(const SelState* __restrict__ st, double* __restrict__ out, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = __longlong_as_double((long long)st[i].prefix);
}

// ---- adjacent order-statistic PAIR (ranks r, r+1): 4 MSD radix levels of 16 bits ------------------
// np.percentile's two neighbours and np.median's middle pair are always adjacent ranks, so one
// histogram per level resolves rank r exactly; rank r+1 is either the same value (multiplicity) or
// the smallest key above it (one extra min pass).
constexpr int kPairBins = 65536;
constexpr int kPairLevels = 4;
struct PairState {
  unsigned long long prefix;   // resolved high bits of the rank-r key
  long long rank;              // remaining rank inside the prefix class
  unsigned long long next;     // min key strictly above (pass 5)
  int need_next;               // 1: rank r+1 is not covered by the multiplicity of the rank-r value
  int pad;
};

__global__ void pair_init_kernel(PairState* st, long long rank, int ncols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncols) { st[i].prefix = 0ull; st[i].rank = rank; st[i].next = ~0ull; st[i].need_next = 0; }
}

__global__ void __launch_bounds__(kBlock)
pair_hist_kernel(const double* __restrict__ base, const int64_t* __restrict__ rows, int64_t stride_elems, int64_t n,
                 int ncols, const int* __restrict__ mult, const PairState* __restrict__ st, int level,
                 unsigned int* __restrict__ hist /* [ncols][65536] */) {
  const int shift = 48 - 16 * level;
  const int64_t total = n * ncols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x; e0 < total; e0 += stride) {
    const int64_t e = e0 + threadIdx.x;
    int slot = -1;
    unsigned int m = 0;
    if (e < total) {
      const int64_t j = e / ncols;
      const int c = (int)(e - j * ncols);
      m = mult ? (unsigned int)__ldg(mult + j) : 1u;
      if (m) {
        const int64_t r = rows ? __ldg(rows + j) : j;
        const unsigned long long key = (unsigned long long)__double_as_longlong(__ldg(base + r * stride_elems + c));
        const bool match = (level == 0) || ((key >> (shift + 16)) == (st[c].prefix >> (shift + 16)));
        if (match) slot = (c << 16) | (int)((key >> shift) & 0xffffu);
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    if (slot >= 0) {
      const unsigned int sum = __reduce_add_sync(peers, m);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[slot], sum);
    }
  }
}

__global__ void __launch_bounds__(1024)
pair_pick_kernel(PairState* st, const unsigned int* __restrict__ hist, int level) {
  __shared__ unsigned long long part[1024];
  const unsigned int* h = hist + (size_t)blockIdx.x * kPairBins;
  const int per = kPairBins / 1024;
  unsigned long long acc = 0;
  for (int i = 0; i < per; ++i) acc += h[threadIdx.x * per + i];
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    PairState s = st[blockIdx.x];
    unsigned long long run = 0;
    int t = 0;
    for (; t < 1024; ++t) { if (run + part[t] > (unsigned long long)s.rank) break; run += part[t]; }
    if (t == 1024) { t = 1023; run -= part[1023]; }
    int b = t * per;
    for (; b < t * per + per - 1; ++b) { if (run + h[b] > (unsigned long long)s.rank) break; run += h[b]; }
    s.prefix |= ((unsigned long long)b) << (48 - 16 * level);
    s.rank -= (long long)run;
    if (level == kPairLevels - 1) s.need_next = ((unsigned long long)s.rank + 1ull < (unsigned long long)h[b]) ? 0 : 1;
    st[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kBlock)
pair_next_kernel(const double* __restrict__ base, const int64_t* __restrict__ rows, int64_t stride_elems, int64_t n,
                 int ncols, const int* __restrict__ mult, PairState* __restrict__ st) {
  const int64_t total = n * ncols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t j = e / ncols;
    const int c = (int)(e - j * ncols);
    if (!st[c].need_next) continue;
    if (mult && __ldg(mult + j) == 0) continue;
    const int64_t r = rows ? __ldg(rows + j) : j;
    const unsigned long long key = (unsigned long long)__double_as_longlong(__ldg(base + r * stride_elems + c));
    if (key > st[c].prefix && key < st[c].next) atomicMin(&st[c].next, key);
  }
}

__global__ void pair_finish_kernel(const PairState* __restrict__ st, double* __restrict__ out, int ncols, int same) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols) {
    const double v1 = __longlong_as_double((long long)st[c].prefix);
    out[2 * c] = v1;
    out[2 * c + 1] = (same || !st[c].need_next) ? v1 : __longlong_as_double((long long)st[c].next);
  }
}

// ---- middle pair of unit-interval columns (np.median of the 4n-row multiset, student.py:62) -------
// Values live in [0,1] (unit-cube coordinates), so a 16-bit fixed-point bucket is a monotone key that
// spreads evenly: one histogram pass, a pick, a compaction of the one or two buckets holding ranks
// r and r+1, and an exact in-block radix select over the few thousand candidates.
constexpr int kMedBuckets = 65536;
constexpr int kMedCap = 1 << 16;     // candidates kept per column
struct MedSel { int b1, b2; long long rank_local; unsigned int count; int overflow; };

// monotone 16-bit bucket of a non-negative double: shift < 0 -> unit-interval fixed point
// floor(v * 65536); otherwise (bits(v) - key_lo) >> shift (piecewise-linear in log space, for weights)
struct BucketMap { unsigned long long key_lo; int shift; };
__device__ __forceinline__ int med_bucket(double v, BucketMap bm) {
  if (bm.shift < 0) {
    int b = (int)(v * 65536.0);
    return b < 0 ? 0 : (b > 65535 ? 65535 : b);
  }
  const unsigned long long key = (unsigned long long)__double_as_longlong(v);
  if (key <= bm.key_lo) return 0;
  const unsigned long long b = (key - bm.key_lo) >> bm.shift;
  return b > 65535ull ? 65535 : (int)b;
}

__global__ void __launch_bounds__(kBlock)
med_hist_kernel(const double* __restrict__ u, const int64_t* __restrict__ rows, const int* __restrict__ mult, int64_t n,
                int d, unsigned int* __restrict__ hist, BucketMap bm) {
  const int64_t total = n * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t j = e / d;
    const int c = (int)(e - j * d);
    const unsigned int m = mult ? (unsigned int)__ldg(mult + j) : 1u;
    if (!m) continue;
    const int64_t r = rows ? __ldg(rows + j) : j;
    atomicAdd(&hist[(size_t)c * kMedBuckets + med_bucket(__ldg(u + r * d + c), bm)], m);
  }
}

__global__ void __launch_bounds__(1024)
med_pick_kernel(const unsigned int* __restrict__ hist, long long rank_lo, MedSel* __restrict__ sel) {
  __shared__ unsigned long long part[1024];
  const unsigned int* h = hist + (size_t)blockIdx.x * kMedBuckets;
  const int per = kMedBuckets / 1024;
  unsigned long long acc = 0;
  for (int i = 0; i < per; ++i) acc += h[threadIdx.x * per + i];
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    int t = 0;
    for (; t < 1024; ++t) { if (run + part[t] > (unsigned long long)rank_lo) break; run += part[t]; }
    if (t == 1024) { t = 1023; run -= part[1023]; }
    int b = t * per;
    for (; b < t * per + per - 1; ++b) { if (run + h[b] > (unsigned long long)rank_lo) break; run += h[b]; }
    MedSel s;
    s.b1 = b; s.rank_local = rank_lo - (long long)run; s.count = 0u; s.overflow = 0;
    s.b2 = b;
    if ((unsigned long long)s.rank_local + 1ull >= (unsigned long long)h[b]) {   // rank r+1 is in a later bucket
      int nb = b + 1;
      while (nb < kMedBuckets && h[nb] == 0u) ++nb;
      s.b2 = nb < kMedBuckets ? nb : b;
    }
    sel[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kBlock)
med_compact_kernel(const double* __restrict__ u, const int64_t* __restrict__ rows, const int* __restrict__ mult,
                   int64_t n, int d, MedSel* __restrict__ sel, double* __restrict__ cval, unsigned int* __restrict__ cmul,
                   BucketMap bm) {
  const int64_t total = n * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t j = e / d;
    const int c = (int)(e - j * d);
    const unsigned int m = mult ? (unsigned int)__ldg(mult + j) : 1u;
    if (!m) continue;
    const int64_t r = rows ? __ldg(rows + j) : j;
    const double v = __ldg(u + r * d + c);
    const int b = med_bucket(v, bm);
    if (b == sel[c].b1 || b == sel[c].b2) {
      const unsigned int pos = atomicAdd(&sel[c].count, 1u);
      if (pos < (unsigned int)kMedCap) { cval[(size_t)c * kMedCap + pos] = v; cmul[(size_t)c * kMedCap + pos] = m; }
      else sel[c].overflow = 1;
    }
  }
}

// exact (rank_local, rank_local + 1) over one column's candidates: 8 MSD radix levels of 8 bits in smem
__global__ void __launch_bounds__(1024)
med_small_kernel(const MedSel* __restrict__ sel, const double* __restrict__ cval, const unsigned int* __restrict__ cmul,
                 double* __restrict__ out, int* __restrict__ overflow, int same) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long prefix, nextkey;
  __shared__ long long rank;
  __shared__ int need_next;
  const MedSel s = sel[blockIdx.x];
  if (s.overflow) { if (threadIdx.x == 0) *overflow = 1; return; }
  const double* v = cval + (size_t)blockIdx.x * kMedCap;
  const unsigned int* mu = cmul + (size_t)blockIdx.x * kMedCap;
  const int cnt = (int)s.count;
  if (threadIdx.x == 0) { prefix = 0ull; rank = s.rank_local; nextkey = ~0ull; need_next = 0; }
  __syncthreads();
  for (int level = 0; level < 8; ++level) {
    const int shift = 56 - 8 * level;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long pre = prefix;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(v[i]);
      if (level == 0 || (key >> (shift + 8)) == (pre >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 0xffu], mu[i]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long run = 0;
      int b = 0;
      for (; b < 255; ++b) { if (run + hist[b] > (unsigned long long)rank) break; run += hist[b]; }
      prefix |= ((unsigned long long)b) << shift;
      rank -= (long long)run;
      if (level == 7) need_next = ((unsigned long long)rank + 1ull < (unsigned long long)hist[b]) ? 0 : 1;
    }
    __syncthreads();
  }
  if (need_next) {
    const unsigned long long v1 = prefix;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(v[i]);
      if (key > v1) atomicMin(&nextkey, key);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double a = __longlong_as_double((long long)prefix);
    out[2 * blockIdx.x] = a;
    out[2 * blockIdx.x + 1] = (need_next && nextkey != ~0ull && !same) ? __longlong_as_double((long long)nextkey) : a;
  }
}

// Sharded candidate merge: every rank compacted the candidates of the selected bucket(s) of ITS rows; the all-gathered
// lists (a fixed `capx` slots per column and rank, so no count has to visit the host first) are concatenated in rank
// order into this rank's candidate arrays and the global count is set -- the exact in-block select then runs on
// identical input on every rank.  One CTA per column.
__global__ void __launch_bounds__(256)
med_merge_kernel(const int* __restrict__ gsel /* [G][d][2]: count, overflow */, const double* __restrict__ gval,
                 const unsigned int* __restrict__ gmul, int G, int d, int capx, MedSel* __restrict__ sel,
                 double* __restrict__ cval, unsigned int* __restrict__ cmul) {
  __shared__ int off[64 + 1];
  const int c = blockIdx.x;
  if (threadIdx.x == 0) {
    int total = 0, ovf = 0;
    for (int r = 0; r < G; ++r) {
      int cnt = gsel[((size_t)r * d + c) * 2];
      if (gsel[((size_t)r * d + c) * 2 + 1] != 0 || cnt > capx || cnt < 0) { ovf = 1; cnt = 0; }
      off[r] = total;
      total += cnt;
    }
    off[G] = total;
    if (total > kMedCap) { ovf = 1; total = 0; }
    sel[c].count = (unsigned int)total;
    sel[c].overflow = ovf;
  }
  __syncthreads();
  if (sel[c].overflow) return;
  for (int r = 0; r < G; ++r) {
    const int base = off[r], cnt = off[r + 1] - off[r];
    const double* sv = gval + ((size_t)r * d + c) * capx;
    const unsigned int* sm = gmul + ((size_t)r * d + c) * capx;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      cval[(size_t)c * kMedCap + base + i] = sv[i];
      cmul[(size_t)c * kMedCap + base + i] = sm[i];
    }
  }
}

// ---- multiplicities ---------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
count_indices_kernel(const int64_t* __restrict__ idx, int64_t m, int* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const int64_t j = __ldg(idx + k);
    if (j >= 0) atomicAdd(&counts[j], 1);     // -1: draw owned by another shard
  }
}

// ---- weighted / counted moments -----------------------------------------------------------
struct MomWs {
  unsigned int ticket;
  unsigned int pad[3];
  double partial[1];  // [grid][P] follows
};
constexpr int kMomMaxD = 128;
constexpr int kMomRows = 32;       // rows staged per tile
constexpr int kMomAccMax = 34;     // ceil(128*129/2 / 256)

template <typename WT>
__device__ __forceinline__ double wt_load(const WT* w, int64_t j) { return (double)__ldg(w + j); }

// pass 1: column sums  sum_j w_j u_j[c]  (and sum w) -> mean = sums * inv_norm
template <typename WT>
__global__ void __launch_bounds__(kBlock)
mom_mean_kernel(const double* __restrict__ u, const int64_t* __restrict__ rows, const WT* __restrict__ w,
                int64_t n, int d, double inv_norm, MomWs* ws, double* __restrict__ mean) {
  __shared__ double red[kBlock];
  // threads [0, act) each own one column (tid % d) and walk rows tid/d, tid/d + rpb, ...
  const int rpb = kBlock / d;             // rows per sweep (d <= 128 -> rpb >= 2)
  const int act = rpb * d;
  const int col = threadIdx.x % d, r0 = threadIdx.x / d;
  double acc = 0.0;
  if (threadIdx.x < act) {
    // four independent accumulators: four rows of loads in flight per thread
    const int64_t step = (int64_t)gridDim.x * rpb;
    int64_t j = (int64_t)blockIdx.x * rpb + r0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (; j + 3 * step < n; j += 4 * step) {
      const int64_t j1 = j + step, j2 = j + 2 * step, j3 = j + 3 * step;
      const int64_t q0 = rows ? __ldg(rows + j) : j, q1 = rows ? __ldg(rows + j1) : j1;
      const int64_t q2 = rows ? __ldg(rows + j2) : j2, q3 = rows ? __ldg(rows + j3) : j3;
      a0 += wt_load(w, j) * __ldg(u + q0 * d + col);
      a1 += wt_load(w, j1) * __ldg(u + q1 * d + col);
      a2 += wt_load(w, j2) * __ldg(u + q2 * d + col);
      a3 += wt_load(w, j3) * __ldg(u + q3 * d + col);
    }
    for (; j < n; j += step) {
      const int64_t r = rows ? __ldg(rows + j) : j;
      a0 += wt_load(w, j) * __ldg(u + r * d + col);
    }
    acc = (a0 + a1) + (a2 + a3);
  }
  red[threadIdx.x] = (threadIdx.x < act) ? acc : 0.0;
  __syncthreads();
  double* part = ws->partial + (size_t)blockIdx.x * d;
  if (threadIdx.x < d) {
    double t = 0.0;
    for (int q = 0; q < rpb; ++q) t += red[q * d + threadIdx.x];
    part[threadIdx.x] = t;
  }
  if (last_block_arrives(&ws->ticket)) {
    if (threadIdx.x < d) {
      const double t = fold_column(ws->partial, (int)gridDim.x, (size_t)d, threadIdx.x);
      mean[threadIdx.x] = t * inv_norm;
    }
  }
}

// pass 2: scatter = sum_j w_j (u_j - mean)(u_j - mean)^T, upper triangle pairs spread over threads
template <typename WT>
__global__ void __launch_bounds__(kBlock)
mom_cov_kernel(const double* __restrict__ u, const int64_t* __restrict__ rows, const WT* __restrict__ w,
               int64_t n, int d, const double* __restrict__ mean, MomWs* ws, double* __restrict__ cov) {
  extern __shared__ double sm[];
  double* tile = sm;                         // [kMomRows][d] centred rows
  double* tw = tile + (kMomRows * d > 256 ? kMomRows * d : 256);   // [kMomRows] weights
  double* mu = tw + kMomRows;                // [d]
  short* pj = reinterpret_cast<short*>(mu + d);
  const int P = d * (d + 1) / 2;
  short* pk = pj + P;
  for (int c = threadIdx.x; c < d; c += blockDim.x) mu[c] = mean[c];
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    // unrank p -> (j,k), j <= k, row-major over the upper triangle
    int j = 0, rem = p;
    while (rem >= d - j) { rem -= d - j; ++j; }
    pj[p] = (short)j; pk[p] = (short)(j + rem);
  }
  // small d: P < 256 pairs -> G thread groups split the rows of a tile between them
  const int Pp = (P + 31) / 32 * 32;
  const int G = (Pp <= kBlock) ? (kBlock / Pp) : 1;
  const int grp = (Pp <= kBlock) ? (threadIdx.x / Pp) : 0;
  const int p0 = (Pp <= kBlock) ? (threadIdx.x % Pp) : threadIdx.x;
  const bool live = (Pp <= kBlock) ? (grp < G && p0 < P) : true;
  double acc[kMomAccMax];
#pragma unroll
  for (int a = 0; a < kMomAccMax; ++a) acc[a] = 0.0;
  __syncthreads();
  for (int64_t j0 = (int64_t)blockIdx.x * kMomRows; j0 < n; j0 += (int64_t)gridDim.x * kMomRows) {
    const int nr = (int)((n - j0 < kMomRows) ? (n - j0) : kMomRows);
    for (int e = threadIdx.x; e < kMomRows * d; e += blockDim.x) {
      const int rr = e / d, c = e - rr * d;
      double v = 0.0;
      if (rr < nr) {
        const int64_t r = rows ? __ldg(rows + j0 + rr) : (j0 + rr);
        v = __ldg(u + r * d + c) - mu[c];
      }
      tile[e] = v;
    }
    for (int rr = threadIdx.x; rr < kMomRows; rr += blockDim.x) tw[rr] = (rr < nr) ? wt_load(w, j0 + rr) : 0.0;
    __syncthreads();
    if (Pp <= kBlock) {
      if (live) {
        const int j = pj[p0], k = pk[p0];
        double s = acc[0];
        for (int rr = grp; rr < kMomRows; rr += G) s += (tile[rr * d + j] * tw[rr]) * tile[rr * d + k];
        acc[0] = s;
      }
    } else {
#pragma unroll
      for (int a = 0; a < kMomAccMax; ++a) {
        const int p = threadIdx.x + a * kBlock;
        if (p < P) {
          const int j = pj[p], k = pk[p];
          double s = acc[a];
          for (int rr = 0; rr < kMomRows; ++rr) s += (tile[rr * d + j] * tw[rr]) * tile[rr * d + k];
          acc[a] = s;
        }
      }
    }
    __syncthreads();
  }
  double* part = ws->partial + (size_t)blockIdx.x * P;
  if (Pp <= kBlock) {
    // fold the G groups in a fixed order through the (now idle) tile buffer
    double* fold = tile;   // needs G*P doubles <= kMomRows*d (checked on the host)
    if (live) fold[grp * P + p0] = acc[0];
    __syncthreads();
    if (threadIdx.x < P) {
      double t = 0.0;
      for (int g = 0; g < G; ++g) t += fold[g * P + threadIdx.x];
      part[threadIdx.x] = t;
    }
  } else {
#pragma unroll
    for (int a = 0; a < kMomAccMax; ++a) {
      const int p = threadIdx.x + a * kBlock;
      if (p < P) part[p] = acc[a];
    }
  }
  if (last_block_arrives(&ws->ticket)) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const double t = fold_column(ws->partial, (int)gridDim.x, (size_t)P, p);
      const int j = pj[p], k = pk[p];
      cov[j * d + k] = t;
      cov[k * d + j] = t;
    }
  }
}

// pass 3: cv = 0.5 sqrt( sum w^2 clip(d2 - d, +-1e6)^2 )
__global__ void __launch_bounds__(kBlock)
mahal_cv_kernel(const double* __restrict__ u, const double* __restrict__ w, int64_t n, int d, int rt,
                const double* __restrict__ mean, const double* __restrict__ inv, ReduceWs* ws,
                double* __restrict__ out) {
  extern __shared__ double sm[];
  double* sinv = sm;               // [d][d]
  double* mu = sinv + d * d;       // [d]
  double* tile = mu + d;           // [rt][ld], rt rows per sweep
  const int ld = d | 1;            // odd stride: conflict-free 64-bit row reads
  __shared__ double red[40];
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) sinv[e] = inv[e];
  for (int c = threadIdx.x; c < d; c += blockDim.x) mu[c] = mean[c];
  __syncthreads();
  double acc = 0.0;
  for (int64_t j0 = (int64_t)blockIdx.x * rt; j0 < n; j0 += (int64_t)gridDim.x * rt) {
    const int nr = (int)((n - j0 < rt) ? (n - j0) : rt);
    for (int64_t e = threadIdx.x; e < (int64_t)nr * d; e += blockDim.x) {
      const int rr = (int)(e / d), c = (int)(e - (int64_t)rr * d);
      tile[rr * ld + c] = __ldg(u + j0 * d + e) - mu[c];
    }
    __syncthreads();
    if (threadIdx.x < nr) {
      const double* x = tile + threadIdx.x * ld;
      double d2 = 0.0;
      for (int k = 0; k < d; ++k) {          // (xc @ inv)[k] * xc[k], summed over k (tools.py:111)
        double y = 0.0;
        for (int j = 0; j < d; ++j) y += x[j] * sinv[j * d + k];
        d2 += y * x[k];
      }
      double dev = d2 - (double)d;
      dev = fmin(fmax(dev, -1e6), 1e6);
      const double ww = __ldg(w + j0 + threadIdx.x);
      acc += (ww * ww) * (dev * dev);
    }
    __syncthreads();
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) ws->partial[blockIdx.x][0] = acc;
  if (last_block_arrives(&ws->ticket)) {
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&ws->partial[i][0]);
    t = block_sum(t, red);
    if (threadIdx.x == 0) { out[0] = 0.5 * sqrt(t); out[1] = t; }
  }
}

// ---- small-d fast paths: one row per thread, everything in registers -----------------------------
// scatter (upper triangle) of rows centred at `mean`; HBM-bound: each row is read once (8 d bytes).
constexpr int kSmallBlock = 128;
template <int D, typename WT>
__global__ void __launch_bounds__(kSmallBlock)
mom_cov_small_kernel(const double* __restrict__ u, const int64_t* __restrict__ rows, const WT* __restrict__ w,
                     int64_t n, const double* __restrict__ mean, MomWs* ws, double* __restrict__ cov) {
  constexpr int P = D * (D + 1) / 2;
  __shared__ double red[kSmallBlock / 32][P];
  double mu[D];
#pragma unroll
  for (int c = 0; c < D; ++c) mu[c] = mean[c];
  double acc[P];
#pragma unroll
  for (int p = 0; p < P; ++p) acc[p] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const int64_t r = rows ? __ldg(rows + j) : j;
    const double wj = wt_load(w, j);
    const double* row = u + r * D;
    double x[D];
#pragma unroll
    for (int c = 0; c < D; ++c) x[c] = __ldg(row + c) - mu[c];
    int p = 0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double xa = x[a] * wj;
#pragma unroll
      for (int b = a; b < D; ++b) { acc[p] += xa * x[b]; ++p; }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const double v = warp_sum(acc[p]);
    if (lane == 0) red[wid][p] = v;
  }
  __syncthreads();
  double* part = ws->partial + (size_t)blockIdx.x * P;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    double t = 0.0;
    for (int q = 0; q < kSmallBlock / 32; ++q) t += red[q][p];
    part[p] = t;
  }
  if (last_block_arrives(&ws->ticket)) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const double t = fold_column(ws->partial, (int)gridDim.x, (size_t)P, p);
      int a = 0, rem = p;
      while (rem >= D - a) { rem -= D - a; ++a; }
      const int bcol = a + rem;
      cov[a * D + bcol] = t;
      cov[bcol * D + a] = t;
    }
  }
}

// column sums, same one-row-per-thread streaming pattern (the column-per-thread kernel above peaks at
// ~2.7 TB/s; this one reads rows like mom_cov_small and runs at its ~4.6 TB/s)
template <int D, typename WT>
__global__ void __launch_bounds__(kSmallBlock)
mom_mean_small_kernel(const double* __restrict__ u, const int64_t* __restrict__ rows, const WT* __restrict__ w,
                      int64_t n, double inv_norm, MomWs* ws, double* __restrict__ mean) {
  __shared__ double red[kSmallBlock / 32][D];
  double acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const int64_t r = rows ? __ldg(rows + j) : j;
    const double wj = wt_load(w, j);
    const double* row = u + r * D;
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] += wj * __ldg(row + c);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const double v = warp_sum(acc[c]);
    if (lane == 0) red[wid][c] = v;
  }
  __syncthreads();
  double* part = ws->partial + (size_t)blockIdx.x * D;
  if (threadIdx.x < D) {
    double t = 0.0;
    for (int q = 0; q < kSmallBlock / 32; ++q) t += red[q][threadIdx.x];
    part[threadIdx.x] = t;
  }
  if (last_block_arrives(&ws->ticket)) {
    // D columns x 8 strided shares, then a fixed-order sum of the shares
    __shared__ double share[8][D];
    const int c = threadIdx.x % D, q = threadIdx.x / D;
    if (q < 8) {
      double t = 0.0;
      for (int b = q; b < (int)gridDim.x; b += 8) t += __ldcg(ws->partial + (size_t)b * D + c);
      share[q][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < D) {
      double t = 0.0;
      for (int q2 = 0; q2 < 8; ++q2) t += share[q2][threadIdx.x];
      mean[threadIdx.x] = t * inv_norm;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kSmallBlock)
mahal_small_kernel(const double* __restrict__ u, const double* __restrict__ w, int64_t n,
                   const double* __restrict__ mean, const double* __restrict__ inv, ReduceWs* ws,
                   double* __restrict__ out) {
  __shared__ double sinv[D * D];
  __shared__ double red[40];
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) sinv[e] = inv[e];
  double mu[D];
#pragma unroll
  for (int c = 0; c < D; ++c) mu[c] = mean[c];
  __syncthreads();
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const double* row = u + j * D;
    double x[D];
#pragma unroll
    for (int c = 0; c < D; ++c) x[c] = __ldg(row + c) - mu[c];
    double d2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {          // (xc @ inv)[k] * xc[k], summed over k (tools.py:111)
      double y = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) y += x[c] * sinv[c * D + k];
      d2 += y * x[k];
    }
    double dev = d2 - (double)D;
    dev = fmin(fmax(dev, -1e6), 1e6);
    const double ww = __ldg(w + j);
    acc += (ww * ww) * (dev * dev);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) ws->partial[blockIdx.x][0] = acc;
  if (last_block_arrives(&ws->ticket)) {
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&ws->partial[i][0]);
    t = block_sum(t, red);
    if (threadIdx.x == 0) { out[0] = 0.5 * sqrt(t); out[1] = t; }
  }
}

inline int small_grid(int64_t n) {
  int64_t need = (n + kSmallBlock - 1) / kSmallBlock;
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

template <typename WT>
bool launch_cov_small(int d, int grid, const double* u, const int64_t* rows, const WT* w, int64_t n,
                      const double* mean, MomWs* ws, double* cov, cudaStream_t st) {
  switch (d) {
#define TB_CASE(DD) case DD: mom_cov_small_kernel<DD, WT><<<grid, kSmallBlock, 0, st>>>(u, rows, w, n, mean, ws, cov); return true;
    TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8) TB_CASE(9) TB_CASE(10)
#undef TB_CASE
    default: return false;
  }
}

template <typename WT>
bool launch_mean_small(int d, int grid, const double* u, const int64_t* rows, const WT* w, int64_t n, double inv_norm,
                       MomWs* ws, double* mean, cudaStream_t st) {
  switch (d) {
#define TB_CASE(DD) case DD: mom_mean_small_kernel<DD, WT><<<grid, kSmallBlock, 0, st>>>(u, rows, w, n, inv_norm, ws, mean); return true;
    TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8) TB_CASE(9) TB_CASE(10)
    TB_CASE(12) TB_CASE(16)
#undef TB_CASE
    default: return false;
  }
}

inline bool launch_mahal_small(int d, int grid, const double* u, const double* w, int64_t n, const double* mean,
                               const double* inv, ReduceWs* ws, double* out, cudaStream_t st) {
  switch (d) {
#define TB_CASE(DD) case DD: mahal_small_kernel<DD><<<grid, kSmallBlock, 0, st>>>(u, w, n, mean, inv, ws, out); return true;
    TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8) TB_CASE(9) TB_CASE(10)
    TB_CASE(12) TB_CASE(16)
#undef TB_CASE
    default: return false;
  }
}

inline int mom_grid(int64_t n, int d) {
  // keep the partial buffer small for large d: P doubles per CTA
  int per_sm = d <= 16 ? 8 : (d <= 50 ? 2 : 1);
  return stream_grid(n, kMomRows, per_sm);
}

template <typename WT>
int launch_moments(const double* u, const int64_t* rows, const WT* w, int64_t n, int d, double inv_norm,
                   void* workspace, double* mean, double* cov, cudaStream_t st, bool do_mean = true) {
  if (n <= 0 || d <= 0 || d > kMomMaxD || !u || !w || !workspace || !mean) return TB_ERR_ARG;
  MomWs* ws = (MomWs*)workspace;
  const int grid = mom_grid(n, d);
  if (do_mean && !launch_mean_small<WT>(d, small_grid(n), u, rows, w, n, inv_norm, ws, mean, st))
    mom_mean_kernel<WT><<<grid, kBlock, 0, st>>>(u, rows, w, n, d, inv_norm, ws, mean);
  if (cov && launch_cov_small<WT>(d, small_grid(n), u, rows, w, n, mean, ws, cov, st)) {
    // register-resident fast path (d <= 10)
  } else if (cov) {
    const int P = d * (d + 1) / 2;
    const int tile_elems = kMomRows * d > 256 ? kMomRows * d : 256;
    size_t smem = sizeof(double) * (tile_elems + kMomRows + d) + sizeof(short) * 2 * P + 16;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(mom_cov_kernel<WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    mom_cov_kernel<WT><<<grid, kBlock, smem, st>>>(u, rows, w, n, d, mean, ws, cov);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TB_OK : (int)e;
}

}  // namespace

extern "C" {

size_t tb_reduce_workspace_bytes(void) { return sizeof(ReduceWs); }

int tb_normalize_inplace(double* w, int64_t n, void* workspace, double* stats3, tb_stream_t stream) {
  if (n <= 0 || !w || !workspace || !stats3) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const int grid = stream_grid(n, kBlock * 4, 8);
  reduce3_kernel<0><<<grid, kBlock, 0, st>>>(w, n, 0.0, (ReduceWs*)workspace, stats3);
  scale_reduce_kernel<<<grid, kBlock, 0, st>>>(w, n, stats3, (ReduceWs*)workspace, stats3);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_masked_sums(const double* w, int64_t n, double thr, void* workspace, double* out3, tb_stream_t stream) {
  if (n <= 0 || !w || !workspace || !out3) return TB_ERR_ARG;
  const int grid = stream_grid(n, kBlock * 4, 8);
  reduce3_kernel<1><<<grid, kBlock, 0, as_stream(stream)>>>(w, n, thr, (ReduceWs*)workspace, out3);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_binade_hist(const double* w, int64_t n, uint64_t* count2048, double* s1_2048, double* s2_2048,
                   tb_stream_t stream) {
  if (n <= 0 || !w || !count2048 || !s1_2048 || !s2_2048) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(count2048, 0, 2048 * sizeof(uint64_t), st);
  cudaMemsetAsync(s1_2048, 0, 2048 * sizeof(double), st);
  cudaMemsetAsync(s2_2048, 0, 2048 * sizeof(double), st);
  const int grid = stream_grid(n, kBlock * 8, 4);
  binade_hist_kernel<<<grid, kBlock, 0, st>>>(w, n, (unsigned long long*)count2048, s1_2048, s2_2048);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_subbin_hist(const double* w, int64_t n, int32_t binade, uint64_t* count2048, double* s1_2048, double* s2_2048,
                   tb_stream_t stream) {
  if (n <= 0 || !w || !count2048 || !s1_2048 || !s2_2048 || binade < 0 || binade > 2047) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(count2048, 0, 2048 * sizeof(uint64_t), st);
  cudaMemsetAsync(s1_2048, 0, 2048 * sizeof(double), st);
  cudaMemsetAsync(s2_2048, 0, 2048 * sizeof(double), st);
  const int grid = stream_grid(n, kBlock * 8, 4);
  subbin_hist_kernel<<<grid, kBlock, 0, st>>>(w, n, binade, (unsigned long long*)count2048, s1_2048, s2_2048);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_compact_workspace_bytes(int64_t n) {
  int64_t nb = (n + kCompactItems - 1) / kCompactItems;
  return (size_t)(nb + 1) * sizeof(int64_t) + 256;
}

int tb_compact_ge(const double* w, int64_t n, double thr, double denom, void* workspace, int64_t* idx_out,
                  double* w_out, int64_t* n_out, tb_stream_t stream) {
  if (n <= 0 || !w || !workspace || !n_out) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const int64_t nb = (n + kCompactItems - 1) / kCompactItems;
  int64_t* counts = (int64_t*)workspace;
  compact_count_kernel<<<(int)nb, kBlock, 0, st>>>(w, n, thr, counts);
  compact_scan_kernel<<<1, 1024, 0, st>>>(counts, nb, n_out);
  if (idx_out || w_out)
    compact_emit_kernel<<<(int)nb, kBlock, 0, st>>>(w, n, thr, denom, counts, idx_out, w_out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_select_workspace_bytes(int32_t ncols, int32_t nranks) {
  return 256 + sizeof(SelState) * (size_t)ncols * nranks + sizeof(unsigned int) * (size_t)ncols * nranks * kSelBins;
}

int tb_select_ranks(const double* base, const int64_t* rows, int64_t stride, int64_t n, int32_t ncols,
                    const int32_t* mult, const int64_t* ranks, int32_t nranks, void* workspace, double* out,
                    tb_stream_t stream) {
  if (n <= 0 || ncols <= 0 || nranks <= 0 || nranks > 4 || ncols * nranks > (1 << 20) || !base || !ranks ||
      !workspace || !out)
    return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  SelState* state = (SelState*)workspace;
  size_t off = (sizeof(SelState) * (size_t)ncols * nranks + 255) / 256 * 256;
  unsigned int* hist = (unsigned int*)((char*)workspace + off);
  const int slots = ncols * nranks;
  sel_init_kernel<<<(slots + 255) / 256, 256, 0, st>>>(state, ranks, ncols, nranks);
  const int grid = stream_grid(n * ncols, kBlock * 4, 8);
  for (int level = 0; level < kSelLevels; ++level) {
    cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)slots * kSelBins, st);
    sel_hist_kernel<<<grid, kBlock, 0, st>>>(base, rows, stride, n, ncols, mult, state, nranks, level, hist);
    sel_pick_kernel<<<slots, 256, 0, st>>>(state, hist, level);
  }
  sel_finish_kernel<<<(slots + 255) / 256, 256, 0, st>>>(state, out, slots);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_unit_median_workspace_bytes(int32_t d) {
  return 512 + sizeof(unsigned int) * (size_t)d * kMedBuckets + sizeof(MedSel) * (size_t)d +
         (sizeof(double) + sizeof(unsigned int)) * (size_t)d * kMedCap + 256 * 4;
}

static int bucket_pair_launch(const double* u, const int64_t* rows, const int32_t* mult, int64_t n, int32_t d,
                              int64_t rank_lo, int same, BucketMap bm, int build_hist, void* workspace, double* out,
                              int32_t* overflow, cudaStream_t st) {
  char* p = (char*)workspace;
  unsigned int* hist = (unsigned int*)p;             p += (sizeof(unsigned int) * (size_t)d * kMedBuckets + 255) / 256 * 256;
  MedSel* sel = (MedSel*)p;                          p += (sizeof(MedSel) * (size_t)d + 255) / 256 * 256;
  double* cval = (double*)p;                         p += sizeof(double) * (size_t)d * kMedCap;
  unsigned int* cmul = (unsigned int*)p;
  cudaMemsetAsync(overflow, 0, sizeof(int32_t), st);
  const int grid = stream_grid(n * d, kBlock * 4, 8);
  if (build_hist) {
    cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)d * kMedBuckets, st);
    med_hist_kernel<<<grid, kBlock, 0, st>>>(u, rows, mult, n, d, hist, bm);
  }
  med_pick_kernel<<<d, 1024, 0, st>>>(hist, rank_lo, sel);
  med_compact_kernel<<<grid, kBlock, 0, st>>>(u, rows, mult, n, d, sel, cval, cmul, bm);
  med_small_kernel<<<d, 1024, 0, st>>>(sel, cval, cmul, out, overflow, same);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TB_OK : (int)e;
}

// The same pipeline stage by stage for sharded ensembles: the callers all-reduce the histogram after stage 0
// and merge the ranks' candidate lists after stage 2 (tempest_b200/sharded.py).
int tb_bucket_offsets(int32_t d, int64_t* out4) {
  if (d <= 0 || !out4) return TB_ERR_ARG;
  size_t off = 0;
  out4[0] = (int64_t)off;  off += (sizeof(unsigned int) * (size_t)d * kMedBuckets + 255) / 256 * 256;   // hist u32[d][65536]
  out4[1] = (int64_t)off;  off += (sizeof(MedSel) * (size_t)d + 255) / 256 * 256;                       // MedSel[d] (24 B each)
  out4[2] = (int64_t)off;  off += sizeof(double) * (size_t)d * kMedCap;                                 // candidates f64[d][65536]
  out4[3] = (int64_t)off;                                                                               // multiplicities u32[d][65536]
  return TB_OK;
}

int tb_bucket_stage(const double* u, const int64_t* rows, const int32_t* mult, int64_t n, int32_t d, int64_t rank_lo,
                    int32_t same, double lo_value, double hi_value, int32_t unit_map, int32_t stage, void* workspace,
                    double* out, int32_t* overflow, tb_stream_t stream) {
  if (d <= 0 || d > 4096 || n < 0 || rank_lo < 0 || !workspace || stage < 0 || stage > 3) return TB_ERR_ARG;
  BucketMap bm;
  if (unit_map) { bm.key_lo = 0ull; bm.shift = -1; }
  else {
    if (!(lo_value >= 0.0) || !(hi_value >= lo_value)) return TB_ERR_ARG;
    unsigned long long klo, khi;
    memcpy(&klo, &lo_value, 8);
    memcpy(&khi, &hi_value, 8);
    bm.key_lo = klo; bm.shift = 0;
    while (((khi - klo) >> bm.shift) > 65535ull) ++bm.shift;
  }
  int64_t off[4];
  tb_bucket_offsets(d, off);
  char* p = (char*)workspace;
  unsigned int* hist = (unsigned int*)(p + off[0]);
  MedSel* sel = (MedSel*)(p + off[1]);
  double* cval = (double*)(p + off[2]);
  unsigned int* cmul = (unsigned int*)(p + off[3]);
  cudaStream_t st = as_stream(stream);
  const int grid = stream_grid((n > 0 ? n : 1) * d, kBlock * 4, 8);
  if (stage == 0) {
    cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)d * kMedBuckets, st);
    if (n > 0) { if (!u) return TB_ERR_ARG; med_hist_kernel<<<grid, kBlock, 0, st>>>(u, rows, mult, n, d, hist, bm); }
  } else if (stage == 1) {
    med_pick_kernel<<<d, 1024, 0, st>>>(hist, rank_lo, sel);
  } else if (stage == 2) {
    if (n > 0) { if (!u) return TB_ERR_ARG; med_compact_kernel<<<grid, kBlock, 0, st>>>(u, rows, mult, n, d, sel, cval, cmul, bm); }
  } else {
    if (!out || !overflow) return TB_ERR_ARG;
    cudaMemsetAsync(overflow, 0, sizeof(int32_t), st);
    med_small_kernel<<<d, 1024, 0, st>>>(sel, cval, cmul, out, overflow, same);
  }
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_bucket_merge(const int32_t* gsel, const double* gval, const uint32_t* gmul, int32_t world, int32_t d, int32_t capx,
                    void* workspace, tb_stream_t stream) {
  if (!gsel || !gval || !gmul || world < 1 || world > 64 || d <= 0 || d > 4096 || capx < 1 || capx > kMedCap || !workspace)
    return TB_ERR_ARG;
  int64_t off[4];
  tb_bucket_offsets(d, off);
  char* p = (char*)workspace;
  med_merge_kernel<<<d, 256, 0, as_stream(stream)>>>(gsel, gval, gmul, world, d, capx, (MedSel*)(p + off[1]),
                                                      (double*)(p + off[2]), (unsigned int*)(p + off[3]));
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_unit_median_pair(const double* u, const int64_t* rows, const int32_t* mult, int64_t n, int32_t d,
                        int64_t rank_lo, void* workspace, double* out, int32_t* overflow, tb_stream_t stream) {
  if (n <= 0 || d <= 0 || d > 4096 || rank_lo < 0 || !u || !workspace || !out || !overflow) return TB_ERR_ARG;
  BucketMap bm; bm.key_lo = 0ull; bm.shift = -1;
  return bucket_pair_launch(u, rows, mult, n, d, rank_lo, 0, bm, 1, workspace, out, overflow, as_stream(stream));
}

int tb_bucket_select_pair(const double* v, int64_t n, int64_t rank_lo, int32_t same, double lo_value, double hi_value,
                          int32_t build_hist, void* workspace, double* out2, int32_t* overflow, tb_stream_t stream) {
  if (n <= 0 || rank_lo < 0 || !v || !workspace || !out2 || !overflow || !(lo_value >= 0.0) || !(hi_value >= lo_value))
    return TB_ERR_ARG;
  unsigned long long klo, khi;
  memcpy(&klo, &lo_value, 8);
  memcpy(&khi, &hi_value, 8);
  BucketMap bm; bm.key_lo = klo; bm.shift = 0;
  while (((khi - klo) >> bm.shift) > 65535ull) ++bm.shift;
  return bucket_pair_launch(v, nullptr, nullptr, n, 1, rank_lo, same, bm, build_hist, workspace, out2, overflow,
                            as_stream(stream));
}

size_t tb_select_pair_workspace_bytes(int32_t ncols) {
  return 256 + (sizeof(PairState) * (size_t)ncols + 255) / 256 * 256 + sizeof(unsigned int) * (size_t)ncols * kPairBins;
}

int tb_select_pair(const double* base, const int64_t* rows, int64_t stride, int64_t n, int32_t ncols,
                   const int32_t* mult, int64_t rank_lo, int32_t same, void* workspace, double* out,
                   tb_stream_t stream) {
  if (n <= 0 || ncols <= 0 || ncols > 4096 || rank_lo < 0 || !base || !workspace || !out) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  PairState* state = (PairState*)workspace;
  unsigned int* hist = (unsigned int*)((char*)workspace + (sizeof(PairState) * (size_t)ncols + 255) / 256 * 256);
  pair_init_kernel<<<(ncols + 255) / 256, 256, 0, st>>>(state, rank_lo, ncols);
  const int grid = stream_grid(n * ncols, kBlock * 4, 8);
  for (int level = 0; level < kPairLevels; ++level) {
    cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)ncols * kPairBins, st);
    pair_hist_kernel<<<grid, kBlock, 0, st>>>(base, rows, stride, n, ncols, mult, state, level, hist);
    pair_pick_kernel<<<ncols, 1024, 0, st>>>(state, hist, level);
  }
  if (!same) pair_next_kernel<<<grid, kBlock, 0, st>>>(base, rows, stride, n, ncols, mult, state);
  pair_finish_kernel<<<(ncols + 255) / 256, 256, 0, st>>>(state, out, ncols, same);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_select_stage(const double* base, const int64_t* rows, int64_t stride, int64_t n, int32_t ncols,
                    const int32_t* mult, const int64_t* ranks, int32_t nranks, void* workspace, double* out,
                    int32_t stage, int32_t level, tb_stream_t stream) {
  if (ncols <= 0 || nranks <= 0 || nranks > 4 || !workspace || level < 0 || level >= kSelLevels) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  SelState* state = (SelState*)workspace;
  size_t off = (sizeof(SelState) * (size_t)ncols * nranks + 255) / 256 * 256;
  unsigned int* hist = (unsigned int*)((char*)workspace + off);
  const int slots = ncols * nranks;
  if (stage == 0) {
    if (!ranks) return TB_ERR_ARG;
    sel_init_kernel<<<(slots + 255) / 256, 256, 0, st>>>(state, ranks, ncols, nranks);
  } else if (stage == 1) {      // local histogram of this level (callers all-reduce it across ranks before stage 2)
    cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)slots * kSelBins, st);
    if (n > 0) {
      if (!base) return TB_ERR_ARG;
      const int grid = stream_grid(n * ncols, kBlock * 4, 8);
      sel_hist_kernel<<<grid, kBlock, 0, st>>>(base, rows, stride, n, ncols, mult, state, nranks, level, hist);
    }
  } else if (stage == 2) {
    sel_pick_kernel<<<slots, 256, 0, st>>>(state, hist, level);
  } else if (stage == 3) {
    if (!out) return TB_ERR_ARG;
    sel_finish_kernel<<<(slots + 255) / 256, 256, 0, st>>>(state, out, slots);
  } else return TB_ERR_ARG;
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_select_hist_offset(int32_t ncols, int32_t nranks) {
  return (sizeof(SelState) * (size_t)ncols * nranks + 255) / 256 * 256;
}

int tb_scale_inplace_dev(double* w, int64_t n, const double* denom, tb_stream_t stream) {
  if (n < 0 || (n > 0 && !w) || !denom) return TB_ERR_ARG;
  if (n == 0) return TB_OK;
  scale_dev_kernel<<<stream_grid(n, kBlock, 16), kBlock, 0, as_stream(stream)>>>(w, n, denom);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_scale_inplace(double* w, int64_t n, double denom, tb_stream_t stream) {
  if (n <= 0 || !w) return TB_ERR_ARG;
  scale_kernel<<<stream_grid(n, kBlock, 16), kBlock, 0, as_stream(stream)>>>(w, n, denom);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_moments_partial(const double* u, const int64_t* rows, const double* w, const int32_t* mult, int64_t n,
                       int32_t d, double inv_norm, int32_t do_mean, int32_t do_cov, void* workspace, double* mean,
                       double* cov, tb_stream_t stream) {
  if ((w == nullptr) == (mult == nullptr)) return TB_ERR_ARG;
  if (w) return launch_moments<double>(u, rows, w, n, d, inv_norm, workspace, mean, do_cov ? cov : nullptr,
                                       as_stream(stream), do_mean != 0);
  return launch_moments<int32_t>(u, rows, mult, n, d, inv_norm, workspace, mean, do_cov ? cov : nullptr,
                                 as_stream(stream), do_mean != 0);
}

int tb_count_indices(const int64_t* idx, int64_t m, int32_t* counts, int64_t n, tb_stream_t stream) {
  if (m < 0 || n <= 0 || !counts || (m > 0 && !idx)) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(counts, 0, sizeof(int32_t) * n, st);
  if (m > 0) count_indices_kernel<<<stream_grid(m, kBlock, 16), kBlock, 0, st>>>(idx, m, counts);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_moments_workspace_bytes(int32_t d) {
  const size_t P = (size_t)d * (d + 1) / 2;
  const size_t per_sm = d <= 16 ? 8 : (d <= 50 ? 2 : 1);
  size_t cap = (size_t)tb::sm_count() * per_sm;
  if (cap > (size_t)kMaxPartials) cap = kMaxPartials;
  return 256 + sizeof(double) * cap * (P > (size_t)d ? P : (size_t)d);
}

int tb_weighted_moments(const double* u, const double* w, int64_t n, int32_t d, void* workspace, double* mean,
                        double* cov, tb_stream_t stream) {
  return launch_moments<double>(u, nullptr, w, n, d, 1.0, workspace, mean, cov, as_stream(stream));
}

int tb_counted_moments(const double* u, const int64_t* rows, const int32_t* mult, int64_t n, int32_t d,
                       double inv_total, void* workspace, double* mean, double* scatter, tb_stream_t stream) {
  return launch_moments<int32_t>(u, rows, mult, n, d, inv_total, workspace, mean, scatter, as_stream(stream));
}

int tb_mahalanobis_cv(const double* u, const double* w, int64_t n, int32_t d, const double* mean,
                      const double* cov_inv, void* workspace, double* cv_out, tb_stream_t stream) {
  if (n <= 0 || d <= 0 || d > kMomMaxD || !u || !w || !mean || !cov_inv || !workspace || !cv_out) return TB_ERR_ARG;
  if (launch_mahal_small(d, small_grid(n), u, w, n, mean, cov_inv, (ReduceWs*)workspace, cv_out, as_stream(stream))) {
    TB_CHECK_LAUNCH();
    return TB_OK;
  }
  const int ld = d | 1;
  const int rt = d <= 32 ? kBlock : 64;
  size_t smem = sizeof(double) * ((size_t)d * d + d + (size_t)rt * ld);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mahal_cv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = stream_grid(n, rt, d <= 16 ? 4 : 1);
  mahal_cv_kernel<<<grid, kBlock, smem, as_stream(stream)>>>(u, w, n, d, rt, mean, cov_inv, (ReduceWs*)workspace, cv_out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // extern "C"
