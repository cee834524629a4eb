// Resampling kernels (a8, a9).
//   ref: tempest/steps/resample.py:52-99, tempest/tools.py:178-228,
//        numpy legacy RandomState.choice (cumsum -> /= last -> searchsorted 'right').
//
// Bit-exact sequential cumsum in parallel ("binade-segmented integer scan")
// -------------------------------------------------------------------------------------
// numpy's cumsum is s_j = fl(s_{j-1} + p_j), strictly left to right; a parallel scan
// associates differently and flips resampling indices.  While the running
// sum stays inside one binade [2^E, 2^{E+1}) its ulp q = 2^{E-52} is constant, s = S*q with S
// an integer in [2^52, 2^53), and for 0 <= p < 2^{E+1}
//     fl(s + p) = (S + inc(p)) * q,  inc(p) = floor(p/q) + [frac(p/q) > 1/2]   (ties: see below)
// as long as the result stays below 2^{E+1}.  Integer addition IS associative, so inside a
// binade the sequential cumsum is an int64 prefix sum.  The kernels below
//   K1 sum tiles in fp64 (approximate prefix, only used to GUESS each tile's binade),
//   K2 scan the tile sums, classify tiles easy (one binade, far from its edges) / hard,
//   K3 integer tile totals of inc() for easy tiles (tiles containing a round-half tie or an
//      element >= 2^{E+1} are demoted to hard),
//   K4 one warp walks the runs of easy tiles (int64 segmented prefix) and the few hard
//      tiles in order, carrying the EXACT fp64 running sum; hard tiles are re-tried in
//      32-element sub-tiles and fall back to literal sequential adds only around a binade
//      crossing / tie; every hypothesis is validated against the exact running sum,
//   K5 easy tiles: int64 in-tile scan -> cdf_j = (S_tile + incl_j) * q.
// Exactness never depends on the guess: a wrong guess only costs a sequential fallback.
#include "tb_common.cuh"

namespace {
using namespace tb;

constexpr int kTile = 1024;       // elements per tile
constexpr int kTileThreads = 256; // 4 elements per thread
constexpr int kHard = INT32_MIN;  // tile_E marker
constexpr int kMinE = -960;       // below this the power-of-two scale factors leave the normal range

struct CdfWs {
  double* tile_sum;    // [nt]
  double* tile_start;  // [nt] exact running sum BEFORE the tile's first element (easy tiles)
  long long* tile_F;   // [nt] integer total of the tile / later: within-run exclusive prefix
  int* tile_E;         // [nt] binade hypothesis or kHard
  int* run_head;       // [nt] first tile of the run an easy tile belongs to
  int* run_next;       // [nt] at run heads: first tile after the run
  // after K4a: tile_sum[head] holds the run's integer total (bit pattern), tile_start[head] the
  // exact running sum at the run's first element (written by the walker)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline CdfWs carve(void* base, int64_t nt) {
  char* p = reinterpret_cast<char*>(base);
  CdfWs w;
  size_t o = 0;
  w.tile_sum = reinterpret_cast<double*>(p + o);     o += align_up(sizeof(double) * nt, 256);
  w.tile_start = reinterpret_cast<double*>(p + o);   o += align_up(sizeof(double) * nt, 256);
  w.tile_F = reinterpret_cast<long long*>(p + o);    o += align_up(sizeof(long long) * nt, 256);
  w.tile_E = reinterpret_cast<int*>(p + o);          o += align_up(sizeof(int) * nt, 256);
  w.run_head = reinterpret_cast<int*>(p + o);        o += align_up(sizeof(int) * nt, 256);
  w.run_next = reinterpret_cast<int*>(p + o);
  return w;
}

__device__ __forceinline__ double pow2(int e) {  // 2^e for -1022 <= e <= 1023
  return __longlong_as_double((long long)(e + 1023) << 52);
}
__device__ __forceinline__ int exponent_of(double x) {  // floor(log2 x) for normal x > 0
  return (int)((__double_as_longlong(x) >> 52) & 0x7ff) - 1023;
}

// inc(p) under binade E.  Returns -1 when the element cannot be handled by the integer rule
// (negative / NaN / >= 2^{E+1} / exact round-half tie).
__device__ __forceinline__ long long inc_of(double p, double up /* 2^(52-E) */, double top /* 2^(E+1) */) {
  if (!(p >= 0.0) || !(p < top)) return -1;
  double sc = p * up;            // exact power-of-two scaling (a denormal p may round: then sc << 1/2)
  double fl = floor(sc);
  double fr = sc - fl;           // exact
  if (fr == 0.5) return -1;
  return (long long)fl + (fr > 0.5 ? 1 : 0);
}

// ---- K1 -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
cdf_tile_sum_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ tile_sum) {
  __shared__ double smem[40];
  const int64_t base = (int64_t)blockIdx.x * kTile;
  double v = 0.0;
#pragma unroll
  for (int k = 0; k < kTile / kTileThreads; ++k) {
    int64_t i = base + k * kTileThreads + threadIdx.x;
    if (i < n) v += __ldg(p + i);
  }
  v = block_sum(v, smem);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = v;
}

// ---- K2: exclusive scan of tile sums + classification (single CTA) ----------------------
__global__ void __launch_bounds__(1024)
cdf_classify_kernel(CdfWs w, int64_t nt) {
  __shared__ double part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (nt + blockDim.x - 1) / blockDim.x;
  const int64_t lo = (int64_t)tid * per, hi = (lo + per < nt) ? lo + per : nt;
  double acc = 0.0;
  for (int64_t t = lo; t < hi; ++t) acc += w.tile_sum[t];
  part[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) { double v = part[i]; part[i] = run; run += v; }
  }
  __syncthreads();
  double pre = part[tid];
  for (int64_t t = lo; t < hi; ++t) {
    const double ps = pre, pe = pre + w.tile_sum[t];
    pre = pe;
    int E = kHard;
    // relative guard 1e-7 >> worst-case fp64 summation error for n < 2^29 non-negative terms
    const double a = ps * (1.0 - 1e-7), b = pe * (1.0 + 1e-7);
    if (ps > 0.0 && isfinite(b) && a >= 2.2250738585072014e-308) {
      int ea = exponent_of(a), eb = exponent_of(b);
      if (ea == eb && ea >= kMinE && ea <= 1000) E = ea;
    }
    w.tile_E[t] = E;
  }
}

// ---- K3: integer tile totals ---------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
cdf_tile_inc_kernel(const double* __restrict__ p, int64_t n, CdfWs w) {
  __shared__ long long sm[kTileThreads / 32];
  __shared__ int bad;
  const int E = w.tile_E[blockIdx.x];
  if (E == kHard) return;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  const double up = pow2(52 - E), top = pow2(E + 1);
  const int64_t base = (int64_t)blockIdx.x * kTile;
  long long tot = 0;
  int mybad = 0;
#pragma unroll
  for (int k = 0; k < kTile / kTileThreads; ++k) {
    int64_t i = base + k * kTileThreads + threadIdx.x;
    if (i < n) {
      long long inc = inc_of(__ldg(p + i), up, top);
      if (inc < 0) mybad = 1; else tot += inc;
    }
  }
  if (mybad) bad = 1;
  tot = (long long)warp_sum_u64((unsigned long long)tot);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < kTileThreads / 32; ++i) t += sm[i];
    if (bad) w.tile_E[blockIdx.x] = kHard; else w.tile_F[blockIdx.x] = t;
  }
}

// ---- K4a: segmented exclusive prefix of tile_F over runs of equal E (single CTA) ---------
// Also records, per run, its head tile, the tile after it and its integer total, so that the
// walker (K4b) visits one entry per RUN instead of one per tile.
__global__ void __launch_bounds__(1024)
cdf_run_prefix_kernel(CdfWs w, int64_t nt) {
  __shared__ long long tail_sum[1024];   // sum of the (open) last segment of each thread's chunk
  __shared__ int tail_E[1024];           // E of that segment (kHard if the chunk ends with a hard tile / is empty)
  __shared__ int tail_head[1024];        // its first tile
  __shared__ int whole[1024];            // 1 if the chunk is a single easy segment (no boundary inside)
  __shared__ long long carry_in[1024];
  __shared__ int carry_E[1024];
  __shared__ int carry_head[1024];
  const int tid = threadIdx.x;
  const int64_t per = (nt + blockDim.x - 1) / blockDim.x;
  const int64_t lo = (int64_t)tid * per, hi = (lo + per < nt) ? lo + per : nt;
  {
    long long s = 0; int curE = kHard; int single = 1; bool first = true; int head = -1;
    for (int64_t t = lo; t < hi; ++t) {
      int E = w.tile_E[t];
      if (first) { curE = E; first = false; s = 0; head = (int)t; if (E == kHard) { single = 0; head = -1; } }
      else if (E != curE || E == kHard) { single = 0; curE = E; s = 0; head = (E == kHard) ? -1 : (int)t; }
      if (E != kHard) s += w.tile_F[t];
    }
    tail_sum[tid] = (lo < hi) ? s : 0; tail_E[tid] = (lo < hi) ? curE : kHard;
    tail_head[tid] = (lo < hi) ? head : -1; whole[tid] = (lo < hi) ? single : 0;
  }
  __syncthreads();
  if (tid == 0) {
    long long c = 0; int cE = kHard, cH = -1;   // open segment entering chunk i
    for (int i = 0; i < (int)blockDim.x; ++i) {
      carry_in[i] = c; carry_E[i] = cE; carry_head[i] = cH;
      int64_t l = (int64_t)i * per;
      if (l >= nt) break;
      if (whole[i] && tail_E[i] == cE && cE != kHard) c += tail_sum[i];
      else { c = tail_sum[i]; cE = tail_E[i]; cH = tail_head[i]; }
    }
  }
  __syncthreads();
  if (lo < hi) {
    long long s = carry_in[tid]; int curE = carry_E[tid]; int head = carry_head[tid];
    for (int64_t t = lo; t < hi; ++t) {
      int E = w.tile_E[t];
      if (E == kHard || E != curE) {            // the open run (if any) ends right before t
        if (curE != kHard && head >= 0) { w.run_next[head] = (int)t; w.tile_sum[head] = __longlong_as_double(s); }
        if (E == kHard) { curE = kHard; head = -1; s = 0; continue; }
        curE = E; head = (int)t; s = 0;
      }
      long long f = w.tile_F[t];
      w.tile_F[t] = s;            // exclusive within-run prefix
      w.run_head[t] = head;
      s += f;
    }
    if (hi == nt && curE != kHard && head >= 0) { w.run_next[head] = (int)nt; w.tile_sum[head] = __longlong_as_double(s); }
  }
}

// ---- K4b: one warp carries the exact running sum through runs and hard tiles -------------
__device__ __forceinline__ long long warp_incl_scan_ll(long long v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// process elements [beg,end) literally or by validated 32-wide integer sub-tiles; writes cdf
// The span is staged through shared memory in chunks (independent, fully pipelined loads) so the
// serial sub-tile loop below never waits on HBM latency.
constexpr int kStage = 2048;
__device__ void hard_span(const double* __restrict__ p, double* __restrict__ cdf, int64_t beg, int64_t end,
                          double& s, bool& have, int lane, double* stage) {
 for (int64_t c0 = beg; c0 < end; c0 += kStage) {
  const int64_t c1 = (c0 + kStage < end) ? c0 + kStage : end;
  __syncwarp();
  for (int64_t i = c0 + lane; i < c1; i += 32) stage[i - c0] = __ldg(p + i);
  __syncwarp();
  for (int64_t b = c0; b < c1; b += 32) {
    const int64_t i = b + lane;
    const double v = (i < c1) ? stage[i - c0] : 0.0;
    bool ok = false;
    if (have && s >= 2.2250738585072014e-308) {
      const int E = exponent_of(s);
      if (E >= kMinE && E <= 1000) {
        const double up = pow2(52 - E), top = pow2(E + 1), q = pow2(E - 52);
        long long inc = (i < c1) ? inc_of(v, up, top) : 0;
        const bool bad = __any_sync(0xffffffffu, inc < 0);
        if (!bad) {
          const long long incl = warp_incl_scan_ll(inc, lane);
          const long long S0 = __double2ll_rn(s * up);
          const long long tot = __shfl_sync(0xffffffffu, incl, 31);
          if (S0 + tot < (1LL << 53)) {
            if (i < c1) cdf[i] = (double)(S0 + incl) * q;
            s = (double)(S0 + tot) * q;
            ok = true;
          }
        }
      }
    }
    if (!ok) {  // literal sequential adds (binade crossing, tie, leading zeros, tiny sums)
      double mine = 0.0;
      for (int k = 0; k < 32; ++k) {
        const double vk = __shfl_sync(0xffffffffu, v, k);
        if (b + k < c1) {
          if (!have) { s = vk; have = true; } else s = __dadd_rn(s, vk);
          if (lane == k) mine = s;
        }
      }
      if (i < c1) cdf[i] = mine;
    }
  }
 }
}

__global__ void __launch_bounds__(32)
cdf_walk_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ cdf, CdfWs w, int64_t nt) {
  __shared__ double stage[kStage];
  const int lane = threadIdx.x;
  double s = 0.0;
  bool have = false;   // numpy: cdf_0 = p_0 (no 0 + p_0)
  int64_t t = 0;
  while (t < nt) {
    const int E = w.tile_E[t];
    if (E == kHard) {
      const int64_t beg = t * kTile, end = (beg + kTile < n) ? beg + kTile : n;
      hard_span(p, cdf, beg, end, s, have, lane, stage);
      ++t;
      continue;
    }
    // t heads a run of easy tiles with hypothesis E: validate it against the exact running sum
    const int64_t t1 = w.run_next[t];
    const long long total = __double_as_longlong(w.tile_sum[t]);
    bool valid = have && s >= 2.2250738585072014e-308 && exponent_of(s) == E;
    if (valid) {
      const long long S0 = __double2ll_rn(s * pow2(52 - E));
      valid = (S0 + total) < (1LL << 53);
      if (valid) {
        if (lane == 0) w.tile_start[t] = s;       // exact running sum entering the run; K5 finishes the tiles
        s = (double)(S0 + total) * pow2(E - 52);
        t = t1;
        continue;
      }
    }
    // hypothesis refuted (never observed; guarded by the 1e-7 margin): demote and go literal
    for (int64_t c = t + lane; c < t1; c += 32) w.tile_E[c] = kHard;
    __syncwarp();
    const int64_t beg = t * kTile, end = (t1 * kTile < n) ? t1 * kTile : n;
    hard_span(p, cdf, beg, end, s, have, lane, stage);
    t = t1;
  }
}

// ---- K5: finish easy tiles -----------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
cdf_emit_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ cdf, CdfWs w) {
  __shared__ long long wsum[kTileThreads / 32];
  const int E = w.tile_E[blockIdx.x];
  if (E == kHard) return;
  const double up = pow2(52 - E), top = pow2(E + 1), q = pow2(E - 52);
  const long long S0 = __double2ll_rn(w.tile_start[w.run_head[blockIdx.x]] * up) + w.tile_F[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * (kTile / kTileThreads);
  long long inc[kTile / kTileThreads];
  long long mine = 0;
#pragma unroll
  for (int k = 0; k < kTile / kTileThreads; ++k) {
    int64_t i = base + k;
    inc[k] = (i < n) ? inc_of(__ldg(p + i), up, top) : 0;
    mine += inc[k];
    inc[k] = mine;  // inclusive within thread
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long incl = warp_incl_scan_ll(mine, lane);
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  long long off = incl - mine;
  for (int i = 0; i < wid; ++i) off += wsum[i];
#pragma unroll
  for (int k = 0; k < kTile / kTileThreads; ++k) {
    int64_t i = base + k;
    if (i < n) cdf[i] = (double)(S0 + off + inc[k]) * q;
  }
}

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

size_t tb_cdf_workspace_bytes(int64_t n) {
  int64_t nt = (n + kTile - 1) / kTile;
  if (nt < 1) nt = 1;
  return align_up(sizeof(double) * nt, 256) * 2 + align_up(sizeof(long long) * nt, 256) +
         align_up(sizeof(int) * nt, 256) * 3 + 256;
}

int tb_cdf_exact(const double* p, int64_t n, double* cdf, void* workspace, tb_stream_t stream) {
  if (n <= 0 || !p || !cdf || !workspace) return TB_ERR_ARG;
  const int64_t nt = (n + kTile - 1) / kTile;
  if (nt > 0x7fffffff) return TB_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  CdfWs w = carve(workspace, nt);
  cdf_tile_sum_kernel<<<(int)nt, kTileThreads, 0, st>>>(p, n, w.tile_sum);
  cdf_classify_kernel<<<1, 1024, 0, st>>>(w, nt);
  cdf_tile_inc_kernel<<<(int)nt, kTileThreads, 0, st>>>(p, n, w);
  cdf_run_prefix_kernel<<<1, 1024, 0, st>>>(w, nt);
  cdf_walk_kernel<<<1, 32, 0, st>>>(p, n, cdf, w, nt);
  cdf_emit_kernel<<<(int)nt, kTileThreads, 0, st>>>(p, n, cdf, w);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

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
