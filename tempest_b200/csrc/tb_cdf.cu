// Bit-exact sequential cumulative sum in parallel, on one GPU or over a sharded weight vector (a8, a9;
// App. C.1).   ref: numpy legacy RandomState.choice (cdf = cumsum(p); cdf /= cdf[-1]; searchsorted 'right'),
//                   tempest/steps/resample.py:79-84, tempest/tools.py:178-228, tempest/modes.py:199-201
//
// numpy's cumsum is s_j = fl(s_{j-1} + p_j), strictly left to right; a parallel scan associates differently and
// flips resampling indices.  While the running sum stays inside one binade [2^E, 2^{E+1}) its ulp q = 2^{E-52} is
// constant, s = S q with S an integer in [2^52, 2^53), and for 0 <= p < 2^{E+1}
//     fl(s + p) = (S + inc(p)) q,   inc(p) = floor(p/q) + [frac(p/q) > 1/2]      (exact ties: handled literally)
// as long as the result stays below 2^{E+1}.  Integer addition IS associative, so inside a binade the sequential
// cumsum is an int64 prefix sum; only the ~40 binade crossings of a real weight vector need the literal fp64 add.
//
// Pipeline (tiles of 1024 elements in GLOBAL order; G ranks, rank r stores one segment per generation, global order
// = generation-major, rank-minor = the order of the single-GPU ensemble):
//   plan     global tile table from the ranks' segment lengths                              (single CTA)
//   K1       fp64 tile sums (approximate prefix, only used to GUESS each tile's binade)     (parallel, reads p)
//   K2       scan of the tile sums, classification easy(E) / hard / zero                    (single CTA)
//   K3       int64 tile totals of inc() under E; ties / oversized elements demote the tile  (parallel, reads p)
//   K4       runs of equal E (int64 prefix inside a run); ONE CTA then carries the EXACT running sum through the
//            runs (O(1) each, validated) and the hard tiles: 1024 threads resolve a hard tile in rounds of
//            (inc under the current binade -> block scan -> first crossing / tie -> literal add of that element)
//   K5       easy tiles: int64 in-tile scan -> cdf_j = (S_tile + incl_j) q                  (parallel, reads p, writes cdf)
// Exactness never depends on a guess: every hypothesis is validated against the exact running sum.
// Sharded: the tile tables (8 + 12 bytes per tile) and the elements of the hard tiles are pushed to every rank's
// table memory over NVLink peer stores between the stages (one release flag per stage and rank, no host
// involvement); every rank then runs the single-CTA stages redundantly on identical tables, so all ranks hold the
// same exact tile-end values and the two-level search below returns numpy's global indices on any number of GPUs.
#include <stdlib.h>
#include "tb_common.cuh"
#include "tb_xgpu.cuh"

namespace {
using namespace tb;

constexpr int kTile = 1024;
constexpr int kHard = INT32_MIN;       // tile class: resolved element by element by the walker
constexpr int kZero = INT32_MIN + 1;   // tile class: leading all-zero tile (running sum still 0)
constexpr int kMinE = -960;            // below this the power-of-two scale factors leave the normal range
constexpr int kMaxSeg = 4096;          // layout segments (ranks x generations)
constexpr int kMaxGen = kMaxSeg / kXMaxRanks;
constexpr int kMaxHardX = 1024;        // hard tiles whose elements are exchanged in a sharded call
constexpr int kPhases = 4;
constexpr double kTinyNormal = 2.2250738585072014e-308;

// ---- table memory (exchanged between ranks; one copy per call parity) -----------------------------------------
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
struct XLayout {
  size_t flags, seglen, tsum, E, F, hard, total;
};
__host__ __device__ inline XLayout x_layout(int64_t ntg_cap) {
  XLayout l;
  size_t o = 0;
  l.flags = o;  o += align_up(sizeof(unsigned long long) * kPhases * kXMaxRanks, 256);
  l.seglen = o; o += align_up(sizeof(long long) * kXMaxRanks * kMaxGen, 256);
  l.tsum = o;   o += align_up(sizeof(double) * ntg_cap, 256);
  l.E = o;      o += align_up(sizeof(int) * ntg_cap, 256);
  l.F = o;      o += align_up(sizeof(long long) * ntg_cap, 256);
  l.hard = o;   o += align_up(sizeof(double) * (size_t)kMaxHardX * kTile, 256);
  l.total = o;
  return l;
}

struct Tables {            // resolved pointers into one rank's table memory (selected parity)
  unsigned long long* flags;
  long long* seglen;
  double* tsum;
  int* E;
  long long* F;
  double* hard;
};
__device__ __forceinline__ Tables tables_at(char* base, int64_t ntg_cap) {
  const XLayout l = x_layout(ntg_cap);
  Tables t;
  t.flags = reinterpret_cast<unsigned long long*>(base + l.flags);
  t.seglen = reinterpret_cast<long long*>(base + l.seglen);
  t.tsum = reinterpret_cast<double*>(base + l.tsum);
  t.E = reinterpret_cast<int*>(base + l.E);
  t.F = reinterpret_cast<long long*>(base + l.F);
  t.hard = reinterpret_cast<double*>(base + l.hard);
  return t;
}

// ---- local scratch ---------------------------------------------------------------------------------------------
struct Local {
  int* status;            // [16]: 0 ntg, 1 nseg, 2 nruns, 3 nhard, 4 error, 5.. spare
  unsigned int* tickets;  // [8]
  int* seg_tile0;         // [kMaxSeg + 1]
  int* seg_owner;         // [kMaxSeg]
  long long* seg_len;     // [kMaxSeg]
  long long* seg_local0;  // [kMaxSeg]
  long long* G;           // [ntg_cap + 1] exclusive prefix of F over the easy tiles (global order)
  int* run_of;            // [ntg_cap]
  int* run_head;          // [ntg_cap + 1]
  int* run_E;             // [ntg_cap]
  double* run_s;          // [ntg_cap] exact running sum entering the run
  long long* run_G0;      // [ntg_cap + 1] G at the run's first tile
  int* hard_list;         // [ntg_cap] global tile ids of the hard tiles, in order
  double* tile_end;       // [ntg_cap] exact cdf value of the tile's last element (level 1 of the sharded search)
  long long* tile_off;    // [ntg_cap] local element offset of the tile, -1 when another rank stores it
  int* tile_len;          // [ntg_cap] elements in the tile
  double* total;          // [2]: cdf[-1], spare
};
__host__ __device__ inline size_t local_bytes(int64_t c) {
  return 256 + 256 + align_up(4 * (kMaxSeg + 1), 256) + align_up(4 * kMaxSeg, 256) + 2 * align_up(8 * kMaxSeg, 256) +
         align_up(8 * (c + 1), 256) + align_up(4 * c, 256) + align_up(4 * (c + 1), 256) + align_up(4 * c, 256) +
         align_up(8 * c, 256) + align_up(4 * c, 256) + align_up(8 * c, 256) + align_up(8 * c, 256) + align_up(4 * c, 256) +
         align_up(8 * (c + 1), 256) + 256;
}
__host__ __device__ inline Local local_at(char* p, int64_t c) {
  Local w;
  size_t o = 0;
  w.status = reinterpret_cast<int*>(p + o);             o += 256;
  w.tickets = reinterpret_cast<unsigned int*>(p + o);   o += 256;
  w.seg_tile0 = reinterpret_cast<int*>(p + o);          o += align_up(4 * (kMaxSeg + 1), 256);
  w.seg_owner = reinterpret_cast<int*>(p + o);          o += align_up(4 * kMaxSeg, 256);
  w.seg_len = reinterpret_cast<long long*>(p + o);      o += align_up(8 * kMaxSeg, 256);
  w.seg_local0 = reinterpret_cast<long long*>(p + o);   o += align_up(8 * kMaxSeg, 256);
  w.G = reinterpret_cast<long long*>(p + o);            o += align_up(8 * (c + 1), 256);
  w.run_of = reinterpret_cast<int*>(p + o);             o += align_up(4 * c, 256);
  w.run_head = reinterpret_cast<int*>(p + o);           o += align_up(4 * (c + 1), 256);
  w.run_E = reinterpret_cast<int*>(p + o);              o += align_up(4 * c, 256);
  w.run_s = reinterpret_cast<double*>(p + o);           o += align_up(8 * c, 256);
  w.hard_list = reinterpret_cast<int*>(p + o);          o += align_up(4 * c, 256);
  w.tile_end = reinterpret_cast<double*>(p + o);        o += align_up(8 * c, 256);
  w.tile_off = reinterpret_cast<long long*>(p + o);     o += align_up(8 * c, 256);
  w.tile_len = reinterpret_cast<int*>(p + o);           o += align_up(4 * c, 256);
  w.run_G0 = reinterpret_cast<long long*>(p + o);       o += align_up(8 * (c + 1), 256);
  w.total = reinterpret_cast<double*>(p + o);
  return w;
}

struct CdfArgs {
  const double* p;        // this rank's weights (its segments, generation-major)
  double* cdf;            // this rank's cdf values
  int64_t n_local;
  int64_t ntg_cap;
  char* local;            // Local scratch
  char* xbase[kXMaxRanks];// table memory of every rank as addressed from here (world == 1: [0] = own workspace)
  int rank, world;
  unsigned long long seq; // call number (>= 1): flags carry it
};

__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }
__device__ __forceinline__ int exponent_of(double x) { return (int)((__double_as_longlong(x) >> 52) & 0x7ff) - 1023; }
__device__ __forceinline__ bool in_integer_regime(double s) {
  if (!(s >= kTinyNormal) || !isfinite(s)) return false;
  const int E = exponent_of(s);
  return E >= kMinE && E <= 1000;
}
// inc(p) under binade E; -1 when the element cannot be handled by the integer rule (negative / NaN / at least a
// quarter of 2^(E+1) / exact round-half tie).  round-to-nearest-even of p/q equals floor + [frac > 1/2] except at
// ties, which are excluded, so the magic-number rounding (two DADDs, no fp64 -> int64 conversion) gives inc.
__device__ __forceinline__ long long inc_of(double p, double up /* 2^(52-E) */, double top /* 2^(E+1) */) {
  if (!(p >= 0.0) || !(p < 0.25 * top)) return -1;
  const double sc = p * up;          // exact power-of-two scaling, < 2^51 (a denormal p may round: then sc << 1/2)
  const double magic = 6755399441055744.0;                  // 2^52 + 2^51
  const double t = sc + magic;       // round to nearest even integer
  const double r = t - magic;
  if (fabs(sc - r) == 0.5) return -1;
  return __double_as_longlong(t) - __double_as_longlong(magic);
}

// ---- cross-rank stage flags ------------------------------------------------------------------------------------
// called by ONE thread after this rank's stores of `phase` are complete and fenced (system scope)
__device__ inline void signal_phase(const CdfArgs& a, int phase) {
  for (int r = 0; r < a.world; ++r) {
    Tables t = tables_at(a.xbase[r], a.ntg_cap);
    st_release_sys(t.flags + phase * kXMaxRanks + a.rank, a.seq);
  }
}
// called by all threads of a CTA: returns false on timeout
__device__ inline bool wait_phase(const CdfArgs& a, int phase) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    Tables t = tables_at(a.xbase[a.rank], a.ntg_cap);
    const unsigned long long* f = t.flags + phase * kXMaxRanks + threadIdx.x;
    if (ld_acquire_sys(f) != a.seq) {
      const unsigned long long t0 = global_ns();
      while (ld_acquire_sys(f) != a.seq)
        if (global_ns() - t0 > kXSpinBudgetNs) { s_ok = 0; break; }
    }
  }
  __syncthreads();
  return s_ok != 0;
}
// "all CTAs of this launch have finished their peer stores": last CTA signals the phase
__device__ inline void arrive_and_signal(const CdfArgs& a, unsigned int* ticket, int phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      signal_phase(a, phase);
    }
  }
}

// segment of global tile g (last s with seg_tile0[s] <= g)
__device__ __forceinline__ int seg_of_tile(const int* __restrict__ seg_tile0, int nseg, int g) {
  int lo = 0, hi = nseg;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(seg_tile0 + mid) <= g) lo = mid; else hi = mid; }
  return lo;
}

// ---- plan ------------------------------------------------------------------------------------------------------
// seg_begin[n_gen + 1]: local start positions of this rank's per-generation segments.
__global__ void __launch_bounds__(256)
cdf_plan_kernel(CdfArgs a, const int64_t* __restrict__ seg_begin, int n_gen) {
  Local w = local_at(a.local, a.ntg_cap);
  if (a.world > 1) {
    for (int t = threadIdx.x; t < n_gen; t += blockDim.x) {
      const long long len = seg_begin[t + 1] - seg_begin[t];
      for (int r = 0; r < a.world; ++r) {
        Tables tr = tables_at(a.xbase[r], a.ntg_cap);
        tr.seglen[a.rank * kMaxGen + t] = len;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); signal_phase(a, 0); }
    if (!wait_phase(a, 0)) { if (threadIdx.x == 0) w.status[4] = 3; return; }
  }
  Tables mine = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
  const int G = a.world, nseg = n_gen * G;
  if (threadIdx.x == 0) {
    long long tiles = 0;
    for (int t = 0; t < n_gen; ++t)
      for (int r = 0; r < G; ++r) {
        const int s = t * G + r;
        const long long len = (G > 1) ? mine.seglen[r * kMaxGen + t] : (seg_begin[t + 1] - seg_begin[t]);
        w.seg_len[s] = len;
        w.seg_owner[s] = r;
        w.seg_tile0[s] = (int)tiles;
        tiles += (len + kTile - 1) / kTile;
      }
    w.seg_tile0[nseg] = (int)tiles;
    for (int r = 0; r < G; ++r) {          // local offsets inside the owner's array
      long long off = 0;
      for (int t = 0; t < n_gen; ++t) { w.seg_local0[t * G + r] = off; off += w.seg_len[t * G + r]; }
    }
    w.status[0] = (int)tiles;
    w.status[1] = nseg;
    w.status[2] = 0; w.status[3] = 0;
    w.status[4] = (tiles > a.ntg_cap) ? 4 : 0;
  }
}

// ---- tile map: local offset / length of every global tile (parallel binary searches, once per call) -----------
__global__ void __launch_bounds__(256)
cdf_tile_map_kernel(CdfArgs a) {
  Local w = local_at(a.local, a.ntg_cap);
  const int ntg = w.status[0], nseg = w.status[1];
  if (w.status[4] != 0) return;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ntg; g += gridDim.x * blockDim.x) {
    const int s = seg_of_tile(w.seg_tile0, nseg, g);
    const long long off = (long long)(g - w.seg_tile0[s]) * kTile;
    w.tile_len[g] = (int)min((long long)kTile, w.seg_len[s] - off);
    w.tile_off[g] = (w.seg_owner[s] == a.rank) ? w.seg_local0[s] + off : -1;
  }
}

// The three streaming kernels give one WARP a tile: lane l reads elements 4 (32 k + l) .. + 3 for k = 0..7 (two
// 16-byte loads per step, 1 KB contiguous per warp and step), eight warps per CTA, grid-stride over the tiles.
constexpr int kTileWarps = 8;
__device__ __forceinline__ void load4(const double* __restrict__ src, int len, int base, double (&v)[4]) {
  if (base + 3 < len && ((reinterpret_cast<uintptr_t>(src + base) & 15) == 0)) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(src + base));
    const double2 b = __ldg(reinterpret_cast<const double2*>(src + base) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (base + e < len) ? __ldg(src + base + e) : 0.0;
  }
}

// ---- K1: approximate tile sums ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kTileWarps)
cdf_tile_sum_kernel(CdfArgs a) {
  Local w = local_at(a.local, a.ntg_cap);
  const int ntg = w.status[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (w.status[4] == 0) {
    for (int g = blockIdx.x * kTileWarps + wid; g < ntg; g += gridDim.x * kTileWarps) {
      const long long off = w.tile_off[g];
      if (off < 0) continue;
      const int len = w.tile_len[g];
      const double* src = a.p + off;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double v[4];
        load4(src, len, 4 * (32 * k + lane), v);
        acc += (v[0] + v[1]) + (v[2] + v[3]);
      }
      acc = warp_sum(acc);
      if (lane < a.world) tables_at(a.xbase[lane], a.ntg_cap).tsum[g] = acc;
    }
  }
  if (a.world > 1) arrive_and_signal(a, w.tickets + 0, 1);
}

// ---- warp scans -----------------------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_incl_scan_ll(long long v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_incl_scan_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// ---- K2: scan of the tile sums + classification (single CTA; warp w owns a contiguous range of tiles and walks it
//      32 tiles at a time: coalesced loads, warp-level scans, one cross-warp combine) --------------------------------
__global__ void __launch_bounds__(1024)
cdf_classify_kernel(CdfArgs a) {
  __shared__ double sh[40];
  Local w = local_at(a.local, a.ntg_cap);
  if (a.world > 1 && !wait_phase(a, 1)) { if (threadIdx.x == 0) w.status[4] = 3; return; }
  if (w.status[4] != 0) return;
  const int ntg = w.status[0];
  Tables t = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int per = ((ntg + nw - 1) / nw + 31) & ~31;         // tiles per warp, a multiple of 32
  const int g0 = min(ntg, wid * per), g1 = min(ntg, g0 + per);
  double mine = 0.0;
  for (int g = g0 + lane; g < g1; g += 32) mine += t.tsum[g];
  mine = warp_sum(mine);
  if (lane == 0) sh[wid] = mine;
  __syncthreads();
  double run = 0.0;                                          // approximate running sum before tile g0
  for (int i = 0; i < wid; ++i) run += sh[i];
  for (int base = g0; base < g1; base += 32) {
    const int g = base + lane;
    const double v = (g < g1) ? t.tsum[g] : 0.0;
    const double incl = warp_incl_scan_d(v, lane);
    const double pe = run + incl, ps = pe - v;
    if (g < g1) {
      int E = kHard;
      if (v == 0.0 && ps == 0.0) E = kZero;                 // leading zeros: the running sum is still exactly 0
      else {
        // relative guard 1e-7 >> worst-case deviation of any fp64 summation order for n < 2^29 non-negative terms
        const double lo = ps * (1.0 - 1e-7), hi = pe * (1.0 + 1e-7);
        if (ps > 0.0 && isfinite(hi) && lo >= kTinyNormal) {
          const int ea = exponent_of(lo), eb = exponent_of(hi);
          if (ea == eb && ea >= kMinE && ea <= 1000) E = ea;
        }
      }
      // only the owner's entry: the other ranks' K3 kernels publish the FINAL class of their tiles into this table,
      // possibly before this kernel gets here, and must not be overwritten with the preliminary one
      if (w.tile_off[g] >= 0) t.E[g] = E;
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// ---- K3: integer tile totals -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kTileWarps)
cdf_tile_inc_kernel(CdfArgs a) {
  Local w = local_at(a.local, a.ntg_cap);
  const int ntg = w.status[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (w.status[4] == 0) {
    Tables me = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
    for (int g = blockIdx.x * kTileWarps + wid; g < ntg; g += gridDim.x * kTileWarps) {
      const long long off = w.tile_off[g];
      if (off < 0) continue;
      const int E = me.E[g];
      long long tot = 0;
      int bad = 0;
      if (E != kHard && E != kZero) {
        const int len = w.tile_len[g];
        const double* src = a.p + off;
        const double up = pow2(52 - E), top = pow2(E + 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int base = 4 * (32 * k + lane);
          double v[4];
          load4(src, len, base, v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const long long inc = (base + e < len) ? inc_of(v[e], up, top) : 0;
            if (inc < 0) bad = 1; else tot += inc;
          }
        }
        bad = __any_sync(0xffffffffu, bad);
        tot = (long long)warp_sum_u64((unsigned long long)tot);
      }
      if (lane < a.world) {                // owners publish the final class of their tiles to every rank
        Tables tr = tables_at(a.xbase[lane], a.ntg_cap);
        tr.E[g] = bad ? kHard : E;
        tr.F[g] = (bad || E == kHard || E == kZero) ? 0 : tot;
      }
    }
  }
  if (a.world > 1) arrive_and_signal(a, w.tickets + 1, 2);
}

// ---- K4a: runs of equal class, exclusive int64 prefix of F, list of hard tiles (single CTA of 1024 threads).
//      Warp w owns a contiguous range of tiles and walks it 32 tiles at a time (coalesced loads, warp-level scans);
//      one cross-warp combine ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
cdf_runs_kernel(CdfArgs a) {
  __shared__ long long sh_f[33];
  __shared__ long long sh_c[33];
  Local w = local_at(a.local, a.ntg_cap);
  if (a.world > 1 && !wait_phase(a, 2)) { if (threadIdx.x == 0) w.status[4] = 3; return; }
  if (w.status[4] != 0) return;
  const int ntg = w.status[0];
  Tables t = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int per = ((ntg + nw - 1) / nw + 31) & ~31;
  const int g0 = min(ntg, wid * per), g1 = min(ntg, g0 + per);
  auto classify = [&](int g, int& cls, bool& hard, bool& newrun, long long& f) {
    cls = t.E[g];
    const int prev = (g > 0) ? t.E[g - 1] : kHard;
    hard = cls == kHard;
    newrun = g == 0 || hard || prev == kHard || cls != prev;
    f = (!hard && cls != kZero) ? t.F[g] : 0;
  };
  long long fsum = 0, cnt = 0;           // cnt packs (new runs << 32) | hard tiles
  for (int g = g0 + lane; g < g1; g += 32) {
    int cls; bool hard, newrun; long long f;
    classify(g, cls, hard, newrun, f);
    fsum += f;
    cnt += ((long long)(newrun ? 1 : 0) << 32) | (long long)(hard ? 1 : 0);
  }
  fsum = (long long)warp_sum_u64((unsigned long long)fsum);
  cnt = (long long)warp_sum_u64((unsigned long long)cnt);
  if (lane == 0) { sh_f[wid] = fsum; sh_c[wid] = cnt; }
  __syncthreads();
  if (wid == 0) {
    const long long xf = (lane < nw) ? sh_f[lane] : 0, xc = (lane < nw) ? sh_c[lane] : 0;
    const long long fi = warp_incl_scan_ll(xf, lane), ci = warp_incl_scan_ll(xc, lane);
    sh_f[lane] = fi - xf; sh_c[lane] = ci - xc;
    if (lane == 31) { sh_f[32] = fi; sh_c[32] = ci; }
  }
  __syncthreads();
  long long G = sh_f[wid];
  int run = (int)(sh_c[wid] >> 32) - 1, hidx = (int)(sh_c[wid] & 0xffffffffLL);
  for (int base = g0; base < g1; base += 32) {
    const int g = base + lane;
    int cls = kHard; bool hard = false, newrun = false; long long f = 0;
    if (g < g1) classify(g, cls, hard, newrun, f);
    const long long fi = warp_incl_scan_ll(f, lane);
    const long long ci = warp_incl_scan_ll(((long long)(newrun ? 1 : 0) << 32) | (long long)(hard ? 1 : 0), lane);
    if (g < g1) {
      const int myrun = run + (int)(ci >> 32);
      w.G[g] = G + fi - f;
      w.run_of[g] = myrun;
      if (newrun) { w.run_head[myrun] = g; w.run_E[myrun] = cls; w.run_G0[myrun] = G + fi - f; }
      if (hard) w.hard_list[hidx + (int)(ci & 0xffffffffLL) - 1] = g;
    }
    G += __shfl_sync(0xffffffffu, fi, 31);
    const long long ct = __shfl_sync(0xffffffffu, ci, 31);
    run += (int)(ct >> 32);
    hidx += (int)(ct & 0xffffffffLL);
  }
  if (threadIdx.x == 0) {
    const int nr = (int)(sh_c[32] >> 32), nh = (int)(sh_c[32] & 0xffffffffLL);
    w.G[ntg] = sh_f[32]; w.run_head[nr] = ntg; w.run_G0[nr] = sh_f[32]; w.status[2] = nr; w.status[3] = nh;
    if (a.world > 1 && nh > kMaxHardX) w.status[4] = 5;
  }
}

// ---- K4p: the elements of the hard tiles, gathered into the contiguous staging area of the table memory (of every
//      rank when sharded) in hard-list order: the walker then reads them from a few L2-resident pages instead of
//      paying DRAM + TLB latency per tile on its serial path ---------------------------------------------------------
__global__ void __launch_bounds__(256)
cdf_push_hard_kernel(CdfArgs a) {
  Local w = local_at(a.local, a.ntg_cap);
  if (w.status[4] == 0) {
    const int nhard = min(w.status[3], kMaxHardX);
    for (int h = blockIdx.x; h < nhard; h += gridDim.x) {
      const int g = w.hard_list[h];
      const long long off = w.tile_off[g];
      if (off < 0) continue;
      const int len = w.tile_len[g];
      for (int e = threadIdx.x; e < kTile; e += blockDim.x) {
        const double v = (e < len) ? __ldg(a.p + off + e) : 0.0;
        for (int r = 0; r < a.world; ++r) tables_at(a.xbase[r], a.ntg_cap).hard[(size_t)h * kTile + e] = v;
      }
    }
  }
  if (a.world > 1) arrive_and_signal(a, w.tickets + 2, 3);
}

// ---- K4b: the exact walk (single CTA of 256 threads: eight warps keep the rounds short, four elements per thread)
constexpr int kWalkThreads = 256;
constexpr int kPer = kTile / kWalkThreads;      // elements per thread
constexpr int kHardStage = 1024;      // hard tiles whose descriptors are staged in shared memory at a time
struct WalkShared {
  double v[kTile];            // elements of the hard tile being resolved
  union {
    long long incl[2][kTile]; // inclusive inc prefix of the current round (buffers alternate between rounds)
    double out[kTile];        // cdf values of the serial regime
  };
  long long woff[2][kWalkThreads / 32];       // warp totals of the current round
  double s_bcast;
  int i_bcast;
  int c_bcast[2];
  int rounds, serial;         // diagnostics of the last call
  long long h_off[kHardStage];
  int h_len[kHardStage];
};

// Resolve one tile literally, starting from the exact running sum s (replicated in all threads).
// v[k]: element 4 j + k of thread j (0 beyond len); cdf_out: where this rank stores the tile's cdf values (null:
// another rank's tile, only the running sum is carried).  A round: inc under the current binade (an element the
// integer rule cannot take counts 2^53, so the prefix overflows exactly where it must stop) -> thread-local and
// warp scans -> cross-warp offsets + the warp in which S + incl first reaches 2^53 -> that warp's ballot gives the
// element c -> elements before c are final, element c is added literally (exact fp64 add), the next round starts
// behind it.
__device__ double walk_hard_tile(const double (&v)[kPer], int len, double s, double* cdf_out, WalkShared& S) {
  constexpr int NW = kWalkThreads / 32;
  const int j = threadIdx.x, lane = j & 31, wid = j >> 5, e0 = kPer * j;
#pragma unroll
  for (int k = 0; k < kPer; ++k) S.v[e0 + k] = v[k];
  __syncthreads();
  int start = 0, rnd = 0;
  while (start < len) {
    if (!in_integer_regime(s)) {
      // running sum 0, subnormal, tiny, or non-finite: one thread adds literally until the integer rule applies
      if (j == 0) {
        double t = s;
        int i = start;
        while (i < len && !in_integer_regime(t)) { t = __dadd_rn(t, S.v[i]); S.out[i] = t; ++i; }
        S.s_bcast = t; S.i_bcast = i; S.serial += i - start;
      }
      __syncthreads();
      const int stop = S.i_bcast;
      if (cdf_out) {
#pragma unroll
        for (int k = 0; k < kPer; ++k) if (e0 + k >= start && e0 + k < stop) cdf_out[e0 + k] = S.out[e0 + k];
      }
      s = S.s_bcast;
      start = stop;
      __syncthreads();
      continue;
    }
    if (j == 0) ++S.rounds;
    const int E = exponent_of(s);
    const double up = pow2(52 - E), top = pow2(E + 1), q = pow2(E - 52);
    const long long S0 = __double2ll_rn(s * up);
    const long long room = (1LL << 53) - S0;                    // the prefix is valid while incl < room
    const int buf = rnd & 1;                                    // shared buffers alternate: no barrier at the round's end
    long long* s_incl = S.incl[buf];
    long long* s_woff = S.woff[buf];
    long long l[kPer];
    long long run = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int e = e0 + k;
      long long inc = (e >= start && e < len) ? inc_of(v[k], up, top) : 0;
      if (inc < 0) inc = 1LL << 53;
      run += inc;
      l[k] = run;                                               // inclusive inside the thread
    }
    const long long wi = warp_incl_scan_ll(run, lane);
    if (lane == 31) s_woff[wid] = wi;
    __syncthreads();                                            // (1) warp totals visible
    long long before = wi - run;                                // prefix before this thread's first element
    int cw = NW;                                                // warp in which the prefix stops (every thread finds it)
    {
      long long acc = 0;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const long long tw = s_woff[i];
        if (i < wid) before += tw;
        acc += tw;
        if (cw == NW && acc >= room) cw = i;
      }
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) { l[k] += before; s_incl[e0 + k] = l[k]; }
    if (wid == cw) {
      const unsigned st = __ballot_sync(0xffffffffu, l[kPer - 1] >= room);
      if (lane == __ffs(st) - 1) {
        int k = 0;
#pragma unroll
        for (int kk = kPer - 1; kk >= 0; --kk) if (l[kk] >= room) k = kk;
        S.c_bcast[buf] = e0 + k;
      }
    }
    __syncthreads();                                            // (2) stop element and prefixes visible
    const int c = (cw < NW) ? S.c_bcast[buf] : len;             // first crossing / tie / oversized element
    if (cdf_out) {
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int e = e0 + k;
        if (e >= start && e < len && e < c) cdf_out[e] = (double)(S0 + l[k]) * q;
      }
    }
    if (c < len) {                                              // literal add of element c
      const double s_prev = (c > start) ? (double)(S0 + s_incl[c - 1]) * q : s;
      s = __dadd_rn(s_prev, S.v[c]);
      if (cdf_out && j == c / kPer) cdf_out[c] = s;
      start = c + 1;
    } else {
      if (len > start) s = (double)(S0 + s_incl[len - 1]) * q;
      start = len;
    }
    ++rnd;
  }
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(kWalkThreads)
cdf_walk_kernel(CdfArgs a) {
  extern __shared__ __align__(16) unsigned char walk_smem[];
  WalkShared& S = *reinterpret_cast<WalkShared*>(walk_smem);
  Local w = local_at(a.local, a.ntg_cap);
  if (w.status[4] != 0) return;
  Tables t = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
  if (threadIdx.x == 0) { S.rounds = 0; S.serial = 0; }
  const int nruns = w.status[2], nhard = w.status[3];
  const int e0 = kPer * threadIdx.x;
  const unsigned long long tb0 = global_ns();
  if (a.world > 1) {
    if (!wait_phase(a, 3)) { if (threadIdx.x == 0) w.status[4] = 3; return; }
    if (nhard > kMaxHardX) return;       // status 5 was set by cdf_runs_kernel
  }
  // ---- the walk: run descriptors and hard-tile descriptors are staged through shared memory (no dependent global
  //      loads per run); the elements of the next hard tile are fetched while the current one is resolved ---------
  double s = 0.0;          // exact running sum (numpy: cdf_0 = p_0 = 0 + p_0)
  constexpr int kChunk = 512;
  __shared__ int c_head[kChunk + 1];
  __shared__ int c_E[kChunk];
  __shared__ long long c_G[kChunk + 1];
  int hptr = 0;            // next entry of the hard list
  int hbase = 0;           // first hard tile whose descriptor is staged
  auto stage_hard = [&](int from) {      // all threads
    __syncthreads();
    for (int i = threadIdx.x; i < kHardStage && from + i < nhard; i += blockDim.x) {
      const int g = w.hard_list[from + i];
      S.h_off[i] = w.tile_off[g];
      S.h_len[i] = w.tile_len[g];
    }
    hbase = from;
    __syncthreads();
  };
  auto fetch = [&](int h, double (&v)[kPer], int& len, long long& off) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) v[k] = 0.0;
    len = 0; off = -1;
    if (h >= nhard) return;
    len = S.h_len[h - hbase];
    off = S.h_off[h - hbase];
    const double* src = (h < kMaxHardX) ? t.hard + (size_t)h * kTile : a.p + off;    // staged by cdf_push_hard_kernel
#pragma unroll
    for (int k = 0; k < kPer; ++k) if (e0 + k < len) v[k] = src[e0 + k];
  };
  double v_next[kPer]; int len_next; long long off_next;
  stage_hard(0);
  fetch(0, v_next, len_next, off_next);
  for (int r0 = 0; r0 < nruns; r0 += kChunk) {
    const int cnt = min(kChunk, nruns - r0);
    __syncthreads();
    for (int i = threadIdx.x; i <= cnt; i += blockDim.x) {
      c_head[i] = w.run_head[r0 + i];
      c_G[i] = w.run_G0[r0 + i];
      if (i < cnt) c_E[i] = w.run_E[r0 + i];
    }
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      const int r = r0 + i;
      const int head = c_head[i], next = c_head[i + 1], cls = c_E[i];
      if (cls == kHard) {                // a hard tile is a run of its own, the next entry of the hard list
        double v[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) v[k] = v_next[k];
        const int len = len_next; const long long off = off_next;
        ++hptr;
        if (hptr < nhard && hptr - hbase >= kHardStage) stage_hard(hptr);
        fetch(hptr, v_next, len_next, off_next);
        s = walk_hard_tile(v, len, s, off >= 0 ? a.cdf + off : nullptr, S);
        if (threadIdx.x == 0) w.tile_end[head] = s;
        continue;
      }
      if (cls == kZero) { if (threadIdx.x == 0) w.run_s[r] = s; continue; }
      const long long total = c_G[i + 1] - c_G[i];
      bool valid = in_integer_regime(s) && exponent_of(s) == cls;
      if (valid) {
        const long long S0 = __double2ll_rn(s * pow2(52 - cls));
        valid = (S0 + total) < (1LL << 53);
        if (valid) {
          if (threadIdx.x == 0) w.run_s[r] = s;
          s = (double)(S0 + total) * pow2(cls - 52);
        }
      }
      if (!valid) {      // hypothesis refuted (guarded by the 1e-7 margin; never observed): resolve the run literally
        if (a.world > 1) { if (threadIdx.x == 0) w.status[4] = 6; return; }
        for (int g = head + threadIdx.x; g < next; g += blockDim.x) t.E[g] = kHard;
        __syncthreads();
        for (int g = head; g < next; ++g) {
          const int len = w.tile_len[g];
          const long long off = w.tile_off[g];
          double v[kPer];
#pragma unroll
          for (int k = 0; k < kPer; ++k) v[k] = (e0 + k < len) ? a.p[off + e0 + k] : 0.0;
          s = walk_hard_tile(v, len, s, a.cdf + off, S);
          if (threadIdx.x == 0) w.tile_end[g] = s;
          __syncthreads();
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    w.total[0] = s;
    w.status[9] = (int)(global_ns() - tb0); w.status[11] = S.rounds; w.status[12] = S.serial;   // diagnostics

  }
}

// ---- K4c: exact tile-end values of the easy tiles (level 1 of the sharded search; not needed on one GPU) --------
__global__ void __launch_bounds__(256)
cdf_tile_end_kernel(CdfArgs a) {
  Local w = local_at(a.local, a.ntg_cap);
  if (w.status[4] != 0) return;
  const int ntg = w.status[0];
  Tables t = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ntg; g += gridDim.x * blockDim.x) {
    const int cls = t.E[g];
    if (cls == kHard) continue;
    const int r = w.run_of[g];
    const double rs = w.run_s[r];
    if (cls == kZero) { w.tile_end[g] = rs; continue; }
    const long long S0 = __double2ll_rn(rs * pow2(52 - cls));
    w.tile_end[g] = (double)(S0 + (w.G[g] - w.run_G0[r]) + t.F[g]) * pow2(cls - 52);
  }
}

// ---- K5: finish easy tiles ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kTileWarps)
cdf_emit_kernel(CdfArgs a) {
  Local w = local_at(a.local, a.ntg_cap);
  const int ntg = w.status[0];
  if (w.status[4] != 0) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  Tables t = tables_at(a.xbase[a.world > 1 ? a.rank : 0], a.ntg_cap);
  for (int g = blockIdx.x * kTileWarps + wid; g < ntg; g += gridDim.x * kTileWarps) {
    const long long off = w.tile_off[g];
    if (off < 0) continue;
    const int E = t.E[g];
    if (E == kHard) continue;
    const int len = w.tile_len[g];
    const double* src = a.p + off;
    double* dst = a.cdf + off;
    const int r = w.run_of[g];
    const double rs = w.run_s[r];
    if (E == kZero) {
      for (int i = lane; i < len; i += 32) dst[i] = rs;
      continue;
    }
    const double up = pow2(52 - E), top = pow2(E + 1), q = pow2(E - 52);
    long long carry = __double2ll_rn(rs * up) + (w.G[g] - w.run_G0[r]);     // S before the tile's first element
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
      const int base = 4 * (32 * k + lane);
      double v[4];
      load4(src, len, base, v);
      long long inc[4];
      long long run = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        run += (base + e < len) ? inc_of(v[e], up, top) : 0;
        inc[e] = run;                                      // inclusive within the lane's four elements
      }
      const long long incl = warp_incl_scan_ll(run, lane);
      const long long before = carry + incl - run;
      double o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (double)(before + inc[e]) * q;
      if (base + 3 < len && aligned) {
        reinterpret_cast<double2*>(dst + base)[0] = make_double2(o[0], o[1]);
        reinterpret_cast<double2*>(dst + base)[1] = make_double2(o[2], o[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (base + e < len) dst[base + e] = o[e];
      }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

// ---- two-level search in the global cdf (sharded runs) ----------------------------------------------------------
// Level 1: first tile whose LAST cdf value exceeds the draw (tile_end[], identical on every rank); level 2: inside
// that tile, in the owner's local cdf.  MODE 0: multinomial, count of cdf_j / cdf[-1] <= u (numpy searchsorted
// 'right' after `cdf /= cdf[-1]`); MODE 1: systematic, first j with not (pos > cdf_j) (tools.py:219-226).
// idx = LOCAL index of the ancestor, or -1 when it lives on another rank.
template <int MODE>
__global__ void __launch_bounds__(kBlock)
cdf_search_x_kernel(CdfArgs a, const double* __restrict__ draws, int64_t m, double u0, int64_t* __restrict__ idx,
                    int* __restrict__ overflow) {
  Local w = local_at(a.local, a.ntg_cap);
  const int ntg = w.status[0];
  const double total = w.total[0];
  const double* __restrict__ tend = w.tile_end;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double u = MODE == 0 ? __ldg(draws + k) : __ddiv_rn(__dadd_rn(u0, (double)k), (double)m);
    auto passes = [&](double c) -> bool { return MODE == 0 ? (__ddiv_rn(c, total) <= u) : (u > c); };
    int lo = 0, hi = ntg;             // first tile whose last element does not pass
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (passes(__ldg(tend + mid))) lo = mid + 1; else hi = mid; }
    if (lo >= ntg) { if (MODE == 1) *overflow = 1; idx[k] = -1; continue; }
    const int g = lo;
    const long long off = w.tile_off[g];
    if (off < 0) { idx[k] = -1; continue; }
    const int len = w.tile_len[g];
    const double* c = a.cdf + off;
    int l = 0, h = len - 1;           // the last element is known not to pass
    while (l < h) { const int mid = (l + h) >> 1; if (passes(__ldg(c + mid))) l = mid + 1; else h = mid; }
    idx[k] = off + l;
  }
}


// =================================================================================================================
// Single-GPU path: a guess pass (read p) and ONE streaming kernel (read p, write cdf) -- a chained scan with decoupled
// look-back in which the carried quantity is the EXACT running sum.
//
// A warp owns a 1024-element tile (tiles are handed out by a ticket counter, so a warp only ever waits for tiles that
// are already running).  Per tile:
//   1. load the tile into registers; publish its AGGREGATE: the int64 totals of inc() under two candidate binades
//      (E, E + 1), where E is a guess of the binade the tile starts in (cdf_guess_* kernels: a 1e-7-accurate prefix of
//      1024-element sums -- one extra read of p; a tile whose guess has no safe value publishes "no aggregate").  Each
//      64-bit descriptor word is self-describing (binade | total), so readers never see a torn aggregate;
//   2. look back over the predecessors' descriptors (32 per round, one per lane) to the nearest tile that has published
//      its exact inclusive PREFIX s; the tiles in between must all offer a total under the binade of s and the integer
//      sum must stay below 2^53 -- then s_in = (S + sum of totals) q is exactly numpy's running sum at the tile start.
//      Anything else (a binade crossing or a tie in between, stale candidates, a tile not yet there) is retried; a tile
//      whose candidates went stale because the hint moved on re-publishes its aggregate while it waits;
//   3. with s_in known, the tile's own end value is published FIRST when the tile is plain (one binade, no tie), then
//      the cdf values are produced by validated rounds: inc under the current binade -> warp scan -> first element that
//      would reach 2^(E+1) / tie / oversized element -> literal fp64 add of that element -> next round behind it.
// Exactness never depends on a guess: guesses only decide how early the successors can proceed.
constexpr int kChainWarps = 8;
constexpr int kCK = 4;                       // 128-element chunks per tile (a lane holds 4 kCK elements in registers)
constexpr int kWarpTile = 128 * kCK;          // elements per warp
constexpr int kChainTile = kWarpTile * kChainWarps;   // elements per CTA tile (one descriptor)
constexpr unsigned long long kWordF = (1ull << 53) - 1ull;
constexpr unsigned long long kCodeZero = 2046ull, kCodeInvalid = 2047ull;   // binade field of an aggregate word
constexpr int kENone = INT32_MIN;

struct ChainHead {               // 256 bytes, zeroed by cdf_chain_init_kernel
  unsigned int ticket, pad0;
  unsigned long long hint;       // bits of the largest exact prefix published so far
  unsigned long long diag[24];   // 0 tiles with more than one round, 1 rounds, 2 look-back retries, 3 re-publications,
                                 // 4 serial-regime elements, 5 late prefix publications, 6 look-back windows read,
                                 // 7 / 8 / 9 ns summed over tiles: ticket -> aggregate, aggregate -> start value, start -> done
};
__host__ __device__ inline size_t chain_bytes(int64_t ntiles) { return 256 + 3 * align_up(8 * (size_t)(ntiles + 1), 256); }
// after the descriptors: sub-tile sums (double) and binade guesses (int), kChainWarps / 2 sub-tiles of 1024 per CTA tile
__host__ __device__ inline size_t chain_guess_bytes(int64_t ntiles) { return align_up(12 * (size_t)(ntiles + 1) * 4, 256) + 256; }
struct ChainWs {
  ChainHead* head;
  unsigned long long *P, *A0, *A1;
};
__host__ __device__ inline ChainWs chain_at(char* base, int64_t ntiles) {
  ChainWs w;
  const size_t stride = align_up(8 * (size_t)(ntiles + 1), 256);
  w.head = reinterpret_cast<ChainHead*>(base);
  w.P = reinterpret_cast<unsigned long long*>(base + 256);
  w.A0 = reinterpret_cast<unsigned long long*>(base + 256 + stride);
  w.A1 = reinterpret_cast<unsigned long long*>(base + 256 + 2 * stride);
  return w;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long word_of(int E, long long F) {
  if (F < 0) return kCodeInvalid << 53;
  return ((unsigned long long)(E + 1023) << 53) | (unsigned long long)F;
}
// total offered by an aggregate word under binade E, or -1
__device__ __forceinline__ long long offer_of(unsigned long long w, int E) {
  const unsigned long long code = w >> 53;
  if (code == kCodeZero) return 0;
  if (code == (unsigned long long)(E + 1023)) return (long long)(w & kWordF);
  return -1;
}

// ---- binade guesses for the chained kernel: approximate prefix of 1024-element sub-tile sums ---------------------------
// (one extra read of p.  The running sum of a PS weight vector is a staircase: it jumps by many binades at the ~50 record
// weights and is flat in between, so "the binade of the latest published prefix" is useless for the ~300 tiles in flight
// behind a jump -- measured: every jump drained the pipeline for ~7 us.  A 1e-7-accurate prefix per sub-tile gives every
// tile its own binade up front; exactness still only rests on the validated look-back.)
__global__ void __launch_bounds__(32 * kTileWarps)
cdf_guess_sum_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ tsum, int64_t nsub) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t g = (int64_t)blockIdx.x * kTileWarps + wid; g < nsub; g += (int64_t)gridDim.x * kTileWarps) {
    const int64_t off = g * kTile;
    const int len = (int)min((int64_t)kTile, n - off);
    const double* src = p + off;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double v[4];
      load4(src, len, 4 * (32 * k + lane), v);
      acc += (v[0] + v[1]) + (v[2] + v[3]);
    }
    acc = warp_sum(acc);
    if (lane == 0) tsum[g] = acc;
  }
}
// single CTA: scan of the sub-tile sums; guess[g] = binade of the running sum over the whole of sub-tile g when the
// 1e-7 guard band says it cannot change inside it, kHard otherwise, kZero while the running sum is still exactly 0
__global__ void __launch_bounds__(1024)
cdf_guess_scan_kernel(const double* __restrict__ tsum, int* __restrict__ guess, int64_t nsub) {
  __shared__ double sh[40];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t per = (((nsub + nw - 1) / nw) + 31) & ~(int64_t)31;
  const int64_t g0 = min(nsub, wid * per), g1 = min(nsub, g0 + per);
  double mine = 0.0;
  for (int64_t g = g0 + lane; g < g1; g += 32) mine += tsum[g];
  mine = warp_sum(mine);
  if (lane == 0) sh[wid] = mine;
  __syncthreads();
  double run = 0.0;
  for (int i = 0; i < wid; ++i) run += sh[i];
  for (int64_t base = g0; base < g1; base += 32) {
    const int64_t g = base + lane;
    const double v = (g < g1) ? tsum[g] : 0.0;
    const double incl = warp_incl_scan_d(v, lane);
    const double pe = run + incl, ps = pe - v;
    if (g < g1) {
      int E = kHard;
      if (v == 0.0 && ps == 0.0) E = kZero;
      else {
        const double lo = ps * (1.0 - 1e-7), hi = pe * (1.0 + 1e-7);
        if (ps > 0.0 && isfinite(hi) && lo >= kTinyNormal) {
          const int ea = exponent_of(lo), eb = exponent_of(hi);
          if (ea == eb && ea >= kMinE && ea <= 1000) E = ea;
        }
      }
      guess[g] = E;
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
}

__global__ void __launch_bounds__(256)
cdf_chain_init_kernel(char* base, int64_t ntiles, int* status) {
  const size_t words = chain_bytes(ntiles) / 8;
  unsigned long long* w = reinterpret_cast<unsigned long long*>(base);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) w[i] = 0ull;
  if (blockIdx.x == 0 && threadIdx.x < 16) status[threadIdx.x] = (threadIdx.x == 0) ? (int)ntiles : 0;
}

// int64 total of inc() over the tile under binade E, or -1 (an element the integer rule cannot take, or a total that
// cannot stay inside the binade).  Warp-uniform result.
__device__ __forceinline__ long long chain_total(const double (&v)[kCK][4], int len, int lane, int E) {
  if (E < kMinE || E > 1000) return -1;
  const double up = pow2(52 - E), top = pow2(E + 1);
  long long tot = 0;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < kCK; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = 128 * k + 4 * lane + j;
      const long long inc = (e < len) ? inc_of(v[k][j], up, top) : 0;
      if (inc < 0) bad = 1; else tot += inc;
    }
  bad = __any_sync(0xffffffffu, bad);
  tot = (long long)warp_sum_u64((unsigned long long)tot);
  return (bad || tot >= (1LL << 53)) ? -1 : tot;
}

// element e of the tile (warp-uniform e): the owner lane selects it from its registers, everyone receives it
__device__ __forceinline__ double chain_element(const double (&v)[kCK][4], int lane, int e) {
  double mine = 0.0;
#pragma unroll
  for (int k = 0; k < kCK; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (128 * k + 4 * lane + j == e) mine = v[k][j];
  return __shfl_sync(0xffffffffu, mine, (e >> 2) & 31);
}

// cdf values of one tile from the exact running sum s entering it; returns the exact running sum leaving it.
__device__ __forceinline__ double chain_emit(const double (&v)[kCK][4], int len, double s, double* __restrict__ dst, int lane,
                                             unsigned long long* diag) {
  const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  int start = 0, rounds = 0;
  while (start < len) {
    if (!in_integer_regime(s)) {
      // running sum 0 / subnormal / tiny / non-finite: literal adds, skipping stretches of zeros (s + 0 = s)
      int cand = INT32_MAX;
#pragma unroll
      for (int k = 0; k < kCK; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = 128 * k + 4 * lane + j;
          if (e >= start && e < len && !(v[k][j] == 0.0)) cand = min(cand, e);
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      const int stop = min(cand, len);
#pragma unroll
      for (int k = 0; k < kCK; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = 128 * k + 4 * lane + j;
          if (e >= start && e < stop) dst[e] = s;
        }
      if (stop < len) {
        const double pc = chain_element(v, lane, stop);
        s = __dadd_rn(s, pc);
        if (lane == 0) { dst[stop] = s; atomicAdd(diag + 4, 1ull); }
        start = stop + 1;
      } else start = len;
      continue;
    }
    ++rounds;
    const int E = exponent_of(s);
    const double up = pow2(52 - E), top = pow2(E + 1), q = pow2(E - 52);
    long long carry = __double2ll_rn(s * up);                 // S: s = S q exactly
    bool stopped = false;
#pragma unroll
    for (int k = 0; k < kCK; ++k) {
      if (stopped || 128 * (k + 1) <= start || 128 * k >= len) continue;      // warp-uniform
      const int base = 128 * k + 4 * lane;
      long long inc[4];
      long long run = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = base + j;
        long long x = (e >= start && e < len) ? inc_of(v[k][j], up, top) : 0;
        if (x < 0) x = 1LL << 53;                                // must be added literally: the prefix stops here
        run += x;
        inc[j] = run;
      }
      const long long incl = warp_incl_scan_ll(run, lane);
      const long long before = carry + incl - run;
      const unsigned st = __ballot_sync(0xffffffffu, before + run >= (1LL << 53));
      if (st == 0u) {
        double o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (double)(before + inc[j]) * q;
        if (aligned && base >= start && base + 3 < len) {
          reinterpret_cast<double2*>(dst + base)[0] = make_double2(o[0], o[1]);
          reinterpret_cast<double2*>(dst + base)[1] = make_double2(o[2], o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (base + j >= start && base + j < len) dst[base + j] = o[j];
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
      } else {
        const int lc = __ffs(st) - 1;
        int jc = 0;
        double s_new = 0.0;
        if (lane < lc) {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (base + j >= start && base + j < len) dst[base + j] = (double)(before + inc[j]) * q;
        } else if (lane == lc) {
          jc = 3;
#pragma unroll
          for (int j = 3; j >= 0; --j) if (before + inc[j] >= (1LL << 53)) jc = j;
          long long prev = before;
          double pc = v[k][0];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < jc) { if (base + j >= start) dst[base + j] = (double)(before + inc[j]) * q; prev = before + inc[j]; }
            if (j == jc) pc = v[k][j];
          }
          // the stop element is >= start by construction (elements before start contribute 0)
          s_new = __dadd_rn((double)prev * q, pc);
          dst[base + jc] = s_new;
        }
        s = __shfl_sync(0xffffffffu, s_new, lc);
        start = 128 * k + 4 * lc + __shfl_sync(0xffffffffu, jc, lc) + 1;
        stopped = true;
      }
    }
    if (!stopped) { s = (double)carry * q; start = len; }
  }
  if (lane == 0 && rounds > 1) { atomicAdd(diag + 0, 1ull); atomicAdd(diag + 1, (unsigned long long)rounds); }
  return s;
}

// One look-back attempt by a single warp over the CTA-tile descriptors before tile g.  Returns 1 with the exact start
// value in s_in; 0 when it has to be retried, with found / sP = the nearest exact prefix it saw (a lower bound of the
// tile's own start value, used to refresh stale candidates).
__device__ __forceinline__ int chain_look_back(const ChainWs& w, long long g, int lane, double& s_in, bool& found, double& sP,
                                               unsigned& nwin) {
  long long sumA = 0, sumB = 0;
  bool okA = true, okB = true, fail = false;
  int why = 14;
  int Ea = kENone;
  found = false;
  sP = 0.0;
  // two windows of 32 predecessors per L2 round trip (the chain of exact prefixes advances one round trip per window)
  for (long long j0 = g - 1; !found && !fail; j0 -= 64) {
    unsigned long long wPs[2] = {0ull, 0ull}, a0s[2] = {0ull, 0ull}, a1s[2] = {0ull, 0ull};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long idx = j0 - 32 * h - lane;
      if (idx >= 0) { wPs[h] = ld_relaxed_u64(w.P + idx); a0s[h] = ld_relaxed_u64(w.A0 + idx); a1s[h] = ld_relaxed_u64(w.A1 + idx); }
    }
    ++nwin;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (found || fail) break;
      const long long j = j0 - 32 * h;
      if (j < 0) { fail = true; why = 14; break; }
      const long long idx = j - lane;
      const unsigned long long wP = wPs[h], a0 = a0s[h], a1 = a1s[h];
      const unsigned haveP = __ballot_sync(0xffffffffu, wP != 0ull);
      const int lp = haveP ? __ffs(haveP) - 1 : 32;
      if (lp < 32) {
        const unsigned long long wsel = __shfl_sync(0xffffffffu, wP, lp);
        sP = __longlong_as_double((long long)(wsel & ~(1ull << 63)));
        found = true;
      }
      const bool between = lane < lp && idx >= 0;                     // tiles strictly after the prefix tile
      if (__any_sync(0xffffffffu, between && (a0 == 0ull || a1 == 0ull))) { fail = true; why = 10; break; }
      if (lp == 32 && j - 31 <= 0) { fail = true; why = 11; break; }             // reached tile 0 and it has no prefix yet
      if (Ea == kENone) {          // binade candidates of this look-back: from the nearest tile that offers any
        int mine = INT32_MAX;
        if (between) {
          const unsigned long long c0 = a0 >> 53, c1 = a1 >> 53;
          if (c0 != kCodeZero && c0 != kCodeInvalid) mine = (int)c0 - 1023;
          else if (c1 != kCodeZero && c1 != kCodeInvalid) mine = (int)c1 - 1023;
        }
        const unsigned has = __ballot_sync(0xffffffffu, mine != INT32_MAX);
        if (has) Ea = __shfl_sync(0xffffffffu, mine, __ffs(has) - 1);
        else if (__any_sync(0xffffffffu, between && ((a0 >> 53) == kCodeInvalid))) { fail = true; why = 12; break; }
      }
      if (Ea != kENone) {
        long long fa = 0, fb = 0;
        bool ba = false, bb = false;
        if (between) {
          long long x = offer_of(a0, Ea); if (x < 0) x = offer_of(a1, Ea);
          long long y = offer_of(a0, Ea + 1); if (y < 0) y = offer_of(a1, Ea + 1);
          ba = x < 0; bb = y < 0;
          fa = ba ? 0 : x; fb = bb ? 0 : y;
        }
        okA = okA && !__any_sync(0xffffffffu, ba);
        okB = okB && !__any_sync(0xffffffffu, bb);
        sumA += (long long)warp_sum_u64((unsigned long long)fa);
        sumB += (long long)warp_sum_u64((unsigned long long)fb);
        if (!okA && !okB) { fail = true; why = 13; break; }
      }
    }
  }
  if (!found || fail) { if (lane == 0) atomicAdd(w.head->diag + why, 1ull); return 0; }
  if (Ea == kENone) { s_in = sP; return 1; }                          // only all-zero tiles in between
  if (!in_integer_regime(sP)) { if (lane == 0) atomicAdd(w.head->diag + 15, 1ull); return 0; }
  const int E = exponent_of(sP);
  const bool useA = (E == Ea) && okA, useB = (E == Ea + 1) && okB;
  if (!useA && !useB) { if (lane == 0) atomicAdd(w.head->diag + 16, 1ull); return 0; }
  const long long tot = __double2ll_rn(sP * pow2(52 - E)) + (useA ? sumA : sumB);
  if (tot >= (1LL << 53)) { if (lane == 0) atomicAdd(w.head->diag + 17, 1ull); return 0; }
  s_in = (double)tot * pow2(E - 52);
  return 1;
}

struct ChainShared {
  long long g;
  unsigned long long t_start;                   // diagnostics: time at which this tile's start value became known
  long long F0[kChainWarps], F1[kChainWarps];   // per-warp totals under the candidate binades (Ec, Ec + 1); -1: none
  long long FE[kChainWarps];                    // per-warp totals under the binade of the tile's exact start value
  int zero[kChainWarps];                        // warp's part of the tile is all zeros (or empty)
  double s_in;
  int cmd, cmd_E;                               // outcome of warp 0's look-back: 1 done, 2 re-publish under cmd_E, 3 timeout
  volatile double hand[kChainWarps + 1];        // exact start value handed from warp to warp when it cannot be computed
  volatile int hand_ready[kChainWarps + 1];     //   in parallel (a crossing / tie inside the CTA tile)
};

__global__ void __launch_bounds__(32 * kChainWarps, 2)
cdf_chain_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ cdf, char* chain_base, int64_t ntiles,
                 int* __restrict__ status, const int* __restrict__ guess) {
  __shared__ ChainShared sm;
  const ChainWs w = chain_at(chain_base, ntiles);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (;;) {
    __syncthreads();                                     // the previous tile's shared state is no longer in use
    if (threadIdx.x == 0) {
      sm.g = (long long)atomicAdd(&w.head->ticket, 1u);
    }
    __syncthreads();
    const long long g = sm.g;
    if (g >= ntiles) return;
    const unsigned long long t_ticket = global_ns();
    const int64_t off = g * (int64_t)kChainTile + (int64_t)wid * kWarpTile;      // this warp's part of the CTA tile
    const int len = (int)max((int64_t)0, min((int64_t)kWarpTile, n - off));
    const double* src = p + off;
    double v[kCK][4];
#pragma unroll
    for (int k = 0; k < kCK; ++k) load4(src, len, 4 * (32 * k + lane), v[k]);
    bool nz = false;
#pragma unroll
    for (int k = 0; k < kCK; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) nz = nz || !(v[k][j] == 0.0);
    const bool wzero = !__any_sync(0xffffffffu, nz);

    int Ec = kENone;                 // lower candidate binade the tile's aggregate is published under (CTA-uniform)
    long long F0 = -1, F1 = -1;      // this warp's totals under (Ec, Ec + 1)
    auto totals = [&](int E) {       // all warps: candidate totals -> shared
      Ec = E;
      F0 = wzero ? 0 : chain_total(v, len, lane, E);
      F1 = wzero ? 0 : chain_total(v, len, lane, E + 1);
      if (lane == 0) { sm.F0[wid] = F0; sm.F1[wid] = F1; }
    };
    auto publish_aggregate = [&]() { // warp 0, after a barrier
      bool allz = true;
      long long t0 = 0, t1 = 0;
      for (int i = 0; i < kChainWarps; ++i) {
        allz = allz && sm.zero[i] != 0;
        if (Ec != kENone) {
          const long long a = sm.F0[i], b = sm.F1[i];
          t0 = (t0 < 0 || a < 0) ? -1 : t0 + a;
          t1 = (t1 < 0 || b < 0) ? -1 : t1 + b;
        }
      }
      if (lane == 0) {
        unsigned long long w0 = kCodeInvalid << 53, w1 = kCodeInvalid << 53;
        if (allz) { w0 = kCodeZero << 53; w1 = kCodeZero << 53; }
        else if (Ec != kENone) {
          w0 = word_of(Ec, (t0 >= 0 && t0 < (1LL << 53)) ? t0 : -1);
          w1 = word_of(Ec + 1, (t1 >= 0 && t1 < (1LL << 53)) ? t1 : -1);
        }
        st_relaxed_u64(w.A0 + g, w0);
        st_relaxed_u64(w.A1 + g, w1);
      }
    };
    double s_in = 0.0;
    if (lane == 0) sm.zero[wid] = wzero ? 1 : 0;
    if (g > 0) {
      // ---- 1. aggregate under the candidate binades ----------------------------------------------------------------
      {
        // candidates (E, E + 1): E = the guessed binade of the running sum over the tile's first 1024 elements
        const int Eg = __ldg(guess + g * (kChainTile / kTile));
        if (Eg != kHard && Eg != kZero) totals(Eg);
      }
      __syncthreads();
      const unsigned long long t_agg = global_ns();
      // ---- 2. look back (warp 0); the other warps wait at the barrier ----------------------------------------------
      for (;;) {
        if (wid == 0) {
          publish_aggregate();
          unsigned tries = 0, nwin = 0;
          int cmd = 0, cmd_E = 0;
          double s_found = 0.0;
          while (cmd == 0) {
            bool found; double sP;
            if (chain_look_back(w, g, lane, s_found, found, sP, nwin)) { cmd = 1; break; }
            // stale or missing candidates?  An INVALID tile cannot have been passed by a successor, so the global hint is
            // still a lower bound of its start value; otherwise only a predecessor's exact prefix is.
            int En = kENone;
            if (Ec == kENone) {
              unsigned long long hb = 0;
              if (lane == 0) hb = ld_relaxed_u64(&w.head->hint);
              hb = __shfl_sync(0xffffffffu, hb, 0);
              const double h = __longlong_as_double((long long)hb);
              if (in_integer_regime(h)) En = exponent_of(h);
            }
            if (found && in_integer_regime(sP)) {
              const int Es = exponent_of(sP);
              if (Ec == kENone || Es > Ec + 1) En = (En == kENone || Es > En) ? Es : En;
            }
            bool allz = true;
            for (int i = 0; i < kChainWarps; ++i) allz = allz && sm.zero[i] != 0;
            if (En != kENone && !allz) { cmd = 2; cmd_E = En; break; }
            ++tries;
            if (tries > 8000000u) { cmd = 3; break; }      // a predecessor never published (cannot happen; never hang)
            __nanosleep(tries < 64 ? 20 : 100);
          }
          if (lane == 0) {
            sm.cmd = cmd; sm.cmd_E = cmd_E; sm.s_in = s_found;
            if (tries) atomicAdd(w.head->diag + 2, (unsigned long long)tries);
            atomicAdd(w.head->diag + 6, (unsigned long long)nwin);
            if (cmd == 2) atomicAdd(w.head->diag + 3, 1ull);
            if (cmd == 3) { status[4] = 7; sm.s_in = __longlong_as_double(0x7ff8000000000000LL); }
          }
        }
        __syncthreads();
        if (sm.cmd != 2) break;
        totals(sm.cmd_E);
        __syncthreads();
      }
      s_in = sm.s_in;
      if (threadIdx.x == 0) {
        const unsigned long long t_in = global_ns();
        atomicAdd(w.head->diag + 7, t_agg - t_ticket);
        atomicAdd(w.head->diag + 8, t_in - t_agg);
        sm.t_start = t_in;
      }
    }
    // ---- 3. per-warp exact start values, the tile's end value first when it is plain, then the cdf values ------------
    const bool reg_in = in_integer_regime(s_in);
    const int E = reg_in ? exponent_of(s_in) : kENone;
    {
      long long FE = -1;
      if (wzero) FE = 0;
      else if (reg_in) FE = (E == Ec) ? F0 : ((Ec != kENone && E == Ec + 1) ? F1 : chain_total(v, len, lane, E));
      if (lane == 0) { sm.FE[wid] = FE; sm.hand_ready[wid + 1] = 0; }
    }
    __syncthreads();
    // start value of this warp: s_in carried through the preceding warps' totals while everything stays plain
    bool known = true;
    long long S = reg_in ? __double2ll_rn(s_in * pow2(52 - E)) : 0;
    for (int i = 0; i < wid && known; ++i) {
      if (sm.zero[i]) continue;
      const long long f = sm.FE[i];
      if (!reg_in || f < 0 || S + f >= (1LL << 53)) known = false; else S += f;
    }
    double s_start;
    if (known) s_start = reg_in ? (double)S * pow2(E - 52) : s_in;
    else {
      if (lane == 0) { while (sm.hand_ready[wid] == 0) __nanosleep(20); }
      __syncwarp();
      s_start = sm.hand[wid];
    }
    // end value without walking the elements?
    bool early = false;
    double s_end = s_start;
    if (wzero) early = true;
    else if (in_integer_regime(s_start)) {
      const int Es = exponent_of(s_start);
      const long long f = (Es == E) ? sm.FE[wid] : chain_total(v, len, lane, Es);
      if (f >= 0) {
        const long long tot = __double2ll_rn(s_start * pow2(52 - Es)) + f;
        if (tot < (1LL << 53)) { s_end = (double)tot * pow2(Es - 52); early = true; }
      }
    }
    auto hand_on = [&](double s) {
      if (lane == 0) {
        if (wid + 1 < kChainWarps) {
          sm.hand[wid + 1] = s;
          __threadfence_block();
          sm.hand_ready[wid + 1] = 1;
        } else {               // last warp: the CTA tile's exact inclusive prefix
          const unsigned long long bits = (unsigned long long)__double_as_longlong(s) & ~(1ull << 63);
          st_relaxed_u64(w.P + g, bits | (1ull << 63));
          atomicMax(&w.head->hint, bits);
        }
      }
    };
    if (early) hand_on(s_end);
    if (len > 0) {
      const double s_walk = chain_emit(v, len, s_start, cdf + off, lane, w.head->diag);
      if (!early) s_end = s_walk;
    }
    if (!early) {
      hand_on(s_end);
      if (lane == 0) atomicAdd(w.head->diag + 5, 1ull);
    }
    if (g > 0 && wid == kChainWarps - 1 && lane == 0) atomicAdd(w.head->diag + 9, global_ns() - sm.t_start);
  }
}

__global__ void cdf_set_bounds_kernel(int64_t* seg, int64_t n) { seg[0] = 0; seg[1] = n; }

CdfArgs make_args(const double* p, int64_t n_local, double* cdf, void* workspace, int64_t ntg_cap, const tb_cdf_x* x) {
  CdfArgs a;
  a.p = p; a.cdf = cdf; a.n_local = n_local; a.ntg_cap = ntg_cap;
  a.local = reinterpret_cast<char*>(workspace);
  for (int r = 0; r < kXMaxRanks; ++r) a.xbase[r] = nullptr;
  if (x && x->world > 1) {
    a.rank = x->rank; a.world = x->world; a.seq = x->seq;
    const size_t par = (x->seq & 1ull) ? x_layout(ntg_cap).total : 0;
    for (int r = 0; r < x->world; ++r) a.xbase[r] = reinterpret_cast<char*>(x->peer[r]) + par;
  } else {
    a.rank = 0; a.world = 1; a.seq = 1;
    a.xbase[0] = a.local + local_bytes(ntg_cap);
  }
  return a;
}

}  // namespace

extern "C" {

int64_t tb_cdf_tile_cap(int64_t n_global, int32_t n_segments) {
  return (n_global + kTile - 1) / kTile + (n_segments > 0 ? n_segments : 1);
}
size_t tb_cdf_x_workspace_bytes(int64_t ntg_cap) { return local_bytes(ntg_cap) + x_layout(ntg_cap).total + 256; }
size_t tb_cdf_x_table_bytes(int64_t ntg_cap) { return 2 * x_layout(ntg_cap).total; }
// single-GPU workspace: the multi-kernel pipeline's scratch followed by the descriptors of the chained kernel
size_t tb_cdf_workspace_bytes(int64_t n) {
  const int64_t cap = tb_cdf_tile_cap(n, 1);
  const int64_t nt = (n + kChainTile - 1) / kChainTile + 1;
  return align_up(tb_cdf_x_workspace_bytes(cap), 256) + chain_bytes(nt) + chain_guess_bytes(nt);
}
static int g_cdf_chain = -1;     // 0: multi-kernel pipeline (default: 0.60 ms at 3.8e7 weights), 1: chained kernel
                                 // (TB_CDF_CHAIN=1; 0.75 ms on the same vector, profiles/r02_cdf_paths.txt)
void tb_cdf_set_chain(int32_t on) { g_cdf_chain = on ? 1 : 0; }

// status of the last call on this workspace: {tiles, segments, runs, hard tiles, error, ...} (device ints)
int32_t* tb_cdf_status_ptr(void* workspace) { return reinterpret_cast<int32_t*>(workspace); }
// diagnostics of the chained kernel's last call on this workspace (device words: multi-round tiles, rounds, look-back
// retries, re-publications, serial-regime elements, late prefix publications)
uint64_t* tb_cdf_chain_diag_ptr(void* workspace, int64_t n) {
  const int64_t cap = tb_cdf_tile_cap(n, 1);
  char* chain = reinterpret_cast<char*>(workspace) + align_up(tb_cdf_x_workspace_bytes(cap), 256);
  return reinterpret_cast<uint64_t*>(reinterpret_cast<ChainHead*>(chain)->diag);
}
double* tb_cdf_total_ptr(void* workspace, int64_t ntg_cap) { return local_at(reinterpret_cast<char*>(workspace), ntg_cap).total; }

int tb_cdf_exact_x(const double* p, int64_t n_local, const int64_t* seg_begin, int32_t n_gen, int64_t n_global,
                   int64_t ntg_cap, double* cdf, void* workspace, const tb_cdf_x* x, tb_stream_t stream) {
  if (n_local < 0 || n_gen <= 0 || !seg_begin || !workspace || ntg_cap <= 0 || (n_local > 0 && (!p || !cdf))) return TB_ERR_ARG;
  const int world = (x && x->world > 1) ? x->world : 1;
  if (world > kXMaxRanks || n_gen * world > kMaxSeg || ntg_cap > 0x7ffffff0) return TB_ERR_UNSUPPORTED;
  if (world > 1 && x->seq < 1) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const CdfArgs a = make_args(p, n_local, cdf, workspace, ntg_cap, x);
  cudaError_t e = cudaMemsetAsync(workspace, 0, 512, st);          // status + tickets
  if (e != cudaSuccess) return (int)e;
  // the tile kernels are launched over an upper bound of the global tile count (n_global: the same number on every
  // rank); the exact count is only known on the device, surplus CTAs exit at once
  int64_t grid = tb_cdf_tile_cap(n_global > 0 ? n_global : 1, n_gen * world);
  if (grid > ntg_cap) grid = ntg_cap;
  cdf_plan_kernel<<<1, 256, 0, st>>>(a, seg_begin, n_gen);
  const int cap_ctas = sm_count() * 8;
  int tgrid = (int)((grid + kTileWarps - 1) / kTileWarps);
  if (tgrid > cap_ctas) tgrid = cap_ctas;
  cdf_tile_map_kernel<<<(int)((grid + 255) / 256) < cap_ctas ? (int)((grid + 255) / 256) : cap_ctas, 256, 0, st>>>(a);
  cdf_tile_sum_kernel<<<tgrid, 32 * kTileWarps, 0, st>>>(a);
  cdf_classify_kernel<<<1, 1024, 0, st>>>(a);
  cdf_tile_inc_kernel<<<tgrid, 32 * kTileWarps, 0, st>>>(a);
  cdf_runs_kernel<<<1, 1024, 0, st>>>(a);
  {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e2 = cudaFuncSetAttribute(cdf_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WalkShared));
      if (e2 != cudaSuccess) return (int)e2;
      attr_set = true;
    }
  }
  cdf_push_hard_kernel<<<128, 256, 0, st>>>(a);
  cdf_walk_kernel<<<1, kWalkThreads, sizeof(WalkShared), st>>>(a);
  if (world > 1) cdf_tile_end_kernel<<<(int)((grid + 255) / 256) < cap_ctas ? (int)((grid + 255) / 256) : cap_ctas, 256, 0, st>>>(a);
  cdf_emit_kernel<<<tgrid, 32 * kTileWarps, 0, st>>>(a);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_cdf_exact(const double* p, int64_t n, double* cdf, void* workspace, tb_stream_t stream) {
  if (n <= 0 || !p || !cdf || !workspace) return TB_ERR_ARG;
  const int64_t cap = tb_cdf_tile_cap(n, 1);
  if (g_cdf_chain < 0) {
    const char* e = getenv("TB_CDF_CHAIN");
    g_cdf_chain = (e && e[0] == '1') ? 1 : 0;
  }
  if (g_cdf_chain) {
    cudaStream_t st = as_stream(stream);
    const int64_t ntiles = (n + kChainTile - 1) / kChainTile;
    char* chain = reinterpret_cast<char*>(workspace) + align_up(tb_cdf_x_workspace_bytes(cap), 256);
    static int per_sm = 0;
    if (per_sm == 0) {
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cdf_chain_kernel, 32 * kChainWarps, 0);
      if (e != cudaSuccess) return (int)e;
      if (per_sm < 1) per_sm = 1;
    }
    const size_t words = chain_bytes(ntiles) / 8;
    int igrid = (int)((words + 256 * 8 - 1) / (256 * 8));
    if (igrid > sm_count() * 4) igrid = sm_count() * 4;
    cdf_chain_init_kernel<<<igrid < 1 ? 1 : igrid, 256, 0, st>>>(chain, ntiles, reinterpret_cast<int*>(workspace));
    int64_t grid = ntiles;                                  // one CTA per tile in flight, tiles by ticket
    if (grid > (int64_t)per_sm * sm_count()) grid = (int64_t)per_sm * sm_count();
    // binade guesses: sub-tile sums -> scan + classification (the arrays follow the descriptors, sized for ntiles + 1)
    const int64_t nsub = (n + kTile - 1) / kTile;
    double* tsum = reinterpret_cast<double*>(chain + chain_bytes(ntiles + 1));
    int* guess = reinterpret_cast<int*>(chain + chain_bytes(ntiles + 1) + align_up(8 * (size_t)(ntiles + 1) * 4, 256));
    int sgrid = (int)((nsub + kTileWarps - 1) / kTileWarps);
    if (sgrid > sm_count() * 8) sgrid = sm_count() * 8;
    cdf_guess_sum_kernel<<<sgrid, 32 * kTileWarps, 0, st>>>(p, n, tsum, nsub);
    cdf_guess_scan_kernel<<<1, 1024, 0, st>>>(tsum, guess, nsub);
    cdf_chain_kernel<<<(unsigned)grid, 32 * kChainWarps, 0, st>>>(p, n, cdf, chain, ntiles, reinterpret_cast<int*>(workspace),
                                                                  guess);
    TB_CHECK_LAUNCH();
    return TB_OK;
  }
  // single segment [0, n): the bounds live at the end of the workspace (written on the stream)
  int64_t* seg = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(workspace) + local_bytes(cap) + x_layout(cap).total);
  cdf_set_bounds_kernel<<<1, 1, 0, as_stream(stream)>>>(seg, n);
  return tb_cdf_exact_x(p, n, seg, 1, n, cap, cdf, workspace, nullptr, stream);
}

int tb_cdf_search_x(const double* p, int64_t n_local, double* cdf, void* workspace, int64_t ntg_cap, const tb_cdf_x* x,
                    const double* draws, int64_t m, int32_t systematic, double u0, int64_t* idx, int32_t* overflow,
                    tb_stream_t stream) {
  if (!workspace || m < 0 || (m > 0 && !idx) || (!systematic && m > 0 && !draws) || (systematic && !overflow)) return TB_ERR_ARG;
  if (m == 0) return TB_OK;
  const CdfArgs a = make_args(p, n_local, cdf, workspace, ntg_cap, x);
  cudaStream_t st = as_stream(stream);
  if (systematic) {
    cudaError_t e = cudaMemsetAsync(overflow, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    cdf_search_x_kernel<1><<<stream_grid(m, kBlock, 16), kBlock, 0, st>>>(a, nullptr, m, u0, idx, overflow);
  } else {
    cdf_search_x_kernel<0><<<stream_grid(m, kBlock, 16), kBlock, 0, st>>>(a, draws, m, 0.0, idx, nullptr);
  }
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // extern "C"
