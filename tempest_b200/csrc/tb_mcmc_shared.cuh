// Shared pieces of the mutation kernels: control block, argument block, boundary maps, production
// variate transforms, the per-step update (sigma adaptation + stop rule) and the persistent step loop.
//   ref: tempest/mcmc.py:104-135, 142-208, 326-411
#pragma once
#include "tb_like.cuh"
#include "tb_xgpu.cuh"

namespace tb {

constexpr int kMcmcBlock = 128;      // walkers per CTA of the per-launch runtime-dimension kernel (split step)
constexpr int kRunWarpsMax = 16;     // upper bound on BODY::kWarps (warps per CTA of the persistent kernels)
constexpr int kMaxModes = 64;
constexpr int kMaxAttempts = 100000;
constexpr int kDeferCap = 64;        // deferred-redraw list entries per warp (flushed whenever 32 are waiting)

// control block indices (doubles)
enum { C_STEPS = 0, C_DONE = 1, C_NACC = 2, C_MEAN_ALPHA = 3, C_ERR = 4, C_NPROP = 5, C_SIGMA0 = 6, C_ATT_RATIO = 7, C_BASE = 8 };
__host__ __device__ inline int ctrl_doubles(int K) { return C_BASE + 4 * K + 3; }

struct McmcWs {          // workspace of the per-launch kernel (tb_mcmc_propose / tb_mcmc_accept)
  unsigned int ticket;
  unsigned int pad[3];
  double partial[1];     // [grid][K+3]: sum alpha per mode, n accepted, n proposals, error flag
};

struct StepArgs {
  int64_t n;
  tb_mcmc_params p;
  tb_tape tape;
  const int32_t* assign;
  double* u;
  double* logl;
  double* qcur;
  void* ws;           // McmcWs (per-launch kernel) or GridSync (persistent kernels)
  double* ctrl;
  tb_xgpu x;          // world > 1: the per-step totals are exchanged over peer memory inside the kernel
  int max_steps;      // persistent kernels: Metropolis steps this launch may run
  // split step for caller-evaluated likelihoods (tb_mcmc_propose / tb_mcmc_accept)
  double* ext_prop;          // [n][d] proposals in the unit cube
  const double* ext_logl;    // [n] log-likelihood of the proposals, filled by the caller
  int32_t* ext_meta;         // [n] proposals drawn (> 0) or -error
};

__device__ __forceinline__ double bc_apply(double v, int kind) {
  if (kind == 1) {                       // periodic: numpy float `% 1.0`
    double m = fmod(v, 1.0);
    if (m != 0.0) { if (m < 0.0) m += 1.0; } else m = 0.0;
    return m;
  }
  if (kind == 2) {                       // reflective: floor-parity fold (mcmc.py:356-364)
    const double fl = floor(v);
    const double r = v - fl;
    const long long k = (long long)fl;
    return ((k & 1LL) == 0) ? r : 1.0 - r;
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// Production variates of the fused step kernels (Philox mode).  Tape mode replaces exactly these three
// draws by the recorded ones; everything downstream is the same code.  tb_debug_variates exposes them
// to the tests (tests/test_gpu_kernels.py: numpy restatement in oracle/philox.py, KS / moment checks).
//
// Normals: Box-Muller on two 32-bit Philox words, radius and angle in fp32 (MUFU log / sin / cos, as
// curand_normal does), promoted to fp64: 24-bit resolution, |z| <= 6.76.  TB_NORMALS_F64 builds the
// fp64 Box-Muller instead (A/B: tools/ab_normals.py; profiles/r02_normals_ab.txt).
__device__ __forceinline__ void bm_pair32(uint32_t a, uint32_t b, double& z0, double& z1) {
#ifdef TB_NORMALS_F64
  const double u1 = ((double)a + 0.5) * 2.3283064365386963e-10;
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(((double)(int32_t)b) * 4.656612873077393e-10, &s, &c);    // angle pi * b / 2^31 in [-pi, pi)
  z0 = r * c;
  z1 = r * s;
#else
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]
  const float r = __fsqrt_rn(-2.0f * __logf(u1));
  float s, c;
  __sincosf(((float)(int32_t)b) * 1.4629180792671596e-9f, &s, &c);  // angle in [-pi, pi): 2*pi*b/2^32 (signed)
  z0 = (double)(r * c);
  z1 = (double)(r * s);
#endif
}

// Standard gamma variate of shape (d + nu)/2 >= 1 by Marsaglia-Tsang (2000): one Philox block per trial
// gives the normal (words 0,1), the uniform of the test (word 2) and, on acceptance, the 32-bit uniform
// of the walker's Metropolis test (word 3; returned through acc_word / have).  The exact test
//   log U < x^2/2 + dd (1 - v + log v),  v = (1 + t)^3,  t = cc x
// is evaluated in fp64.  For |t| < 2^-6 (every shape above ~2e4; the sampler's nu = 1e6 gives t ~ 1e-3)
// the x^2/2 term cancels the t^2 term of the bracket exactly and the rest is the alternating series
//   R = -3 dd t^4 (1/4 - t/5 + t^2/6 - t^3/7 + t^4/8 - t^5/9),   |R| < 1e-3,
// so "log U < R" is "U < exp(R)" with exp by its cubic Taylor polynomial: no transcendental at all
// (an fp32 log here loses ~0.03 absolute at dd = 5e5 through dd * log v).
__device__ __forceinline__ double gamma_mt(const Philox& rng, uint64_t slot, uint32_t step, double shape,
                                           uint32_t& acc_word, bool& have) {
  const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
  double g = dd;
  have = false;
  for (uint32_t trial = 0; trial < 64; ++trial) {
    const uint4 r = rng.block((uint32_t)slot, (uint32_t)(slot >> 32), step, (RNG_GAMMA << 24) | trial);
    double n0, n1;
    bm_pair32(r.x, r.y, n0, n1);
    const double t = cc * n0;
    const double v1 = 1.0 + t;
    if (v1 <= 0.0) continue;
    const double v = v1 * v1 * v1;
    const double uu = ((double)r.z + 0.5) * 2.3283064365386963e-10;
    const double x2 = n0 * n0;
    bool ok = uu < 1.0 - 0.0331 * x2 * x2;                  // squeeze: decides ~90 % of the trials
    if (!ok) {
      if (fabs(t) < 0.015625) {
        const double t2 = t * t;
        double s = 1.0 / 8.0 - t * (1.0 / 9.0);
        s = 1.0 / 7.0 - t * s; s = 1.0 / 6.0 - t * s; s = 1.0 / 5.0 - t * s; s = 1.0 / 4.0 - t * s;   // Horner, alternating
        const double R = -3.0 * dd * (t2 * t2) * s;
        ok = uu < 1.0 + R * (1.0 + R * (0.5 + R * (1.0 / 6.0)));
      } else {
        ok = log(uu) < 0.5 * x2 + dd * (1.0 - v + log(v));
      }
    }
    if (ok) { g = dd * v; acc_word = r.w; have = true; break; }
  }
  return g;
}

// D normals of (walker slot, step, attempt): ceil(D/4) Philox blocks, four normals each.
template <int D>
__device__ __forceinline__ void normals_fixed(const Philox& rng, uint64_t slot, uint32_t step, int attempt, double (&z)[D]) {
  constexpr int NCALL = (D + 3) / 4;
#pragma unroll
  for (int cidx = 0; cidx < NCALL; ++cidx) {
    const uint4 r = rng.block((uint32_t)slot, (uint32_t)(slot >> 32), step,
                              (RNG_NORMAL << 24) | (uint32_t)((attempt * NCALL + cidx) & 0xffffff));
    double n0, n1, n2, n3;
    bm_pair32(r.x, r.y, n0, n1);
    bm_pair32(r.z, r.w, n2, n3);
    if (4 * cidx + 0 < D) z[4 * cidx + 0] = n0;
    if (4 * cidx + 1 < D) z[4 * cidx + 1] = n1;
    if (4 * cidx + 2 < D) z[4 * cidx + 2] = n2;
    if (4 * cidx + 3 < D) z[4 * cidx + 3] = n3;
  }
}

// accept uniform: the spare word of the gamma block when there is one, else a dedicated block (RWM)
__device__ __forceinline__ double accept_uniform(const Philox& rng, uint64_t slot, uint32_t step, uint32_t acc_word, bool have) {
  if (have) return ((double)acc_word + 0.5) * 2.3283064365386963e-10;   // 32-bit uniform in (0,1)
  const uint4 r = rng.block((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_ACCEPT << 24);
  return u53(r.x, r.y);
}

// ------------------------------------------------------------------------------------------
// adapt sigma and evaluate the stop rule from the (global) per-mode totals (mcmc.py:180-194, 104-135)
//   tot[0..K) sum alpha per mode, tot[K] accepted, tot[K+1] proposals drawn, tot[K+2] error code (max)
static __device__ __noinline__ void apply_step_update(const tb_mcmc_params& p, double* ctrl, const double* tot, int K) {
  const int d = p.n_dim;
  const int it = (int)ctrl[C_STEPS] + 1;
  const double sigma0 = 2.38 / sqrt((double)d);
  double* sigma = ctrl + C_BASE;
  const double* count = ctrl + C_BASE + K;
  double* salpha = ctrl + C_BASE + 2 * K;
  double all_alpha = 0.0;
  const double rate = 1.0 / (double)(it + 1);
  for (int c = 0; c < K; ++c) {
    salpha[c] = tot[c];
    all_alpha += tot[c];
    if (count[c] > 0.0) {
      const double mean_alpha = tot[c] / count[c];
      double s = sigma[c] + rate * (mean_alpha - 0.234);
      if (p.sampler == TB_SAMPLE_TPCN) s = fmin(fmax(s, 0.0), fmin(sigma0, 0.99));
      sigma[c] = s;
    }
  }
  const double n_all = (double)p.n_global;
  const double acc = tot[K] / n_all;
  // weighted sigma over the first n_nonempty sigmas (reference quirk: sigmas[:len(sizes)])
  double sw = 0.0, ws = 0.0;
  int j = 0;
  for (int c = 0; c < K; ++c) if (count[c] > 0.0) { ws += sigma[j] * count[c]; sw += count[c]; ++j; }
  const double wsig = ws / sw;
  const double n_min = (double)(p.n_steps * d);
  const double ratio = sigma0 / fmax(1e-6, wsig);
  const double n_adapt = (double)(p.n_steps * d) * (0.234 / fmax(0.01, acc)) * (ratio * ratio);
  const double n_cap = (double)(p.n_max * d);
  const double n_final = fmin(fmax(n_min, n_adapt), n_cap);
  const int stop_at = (int)n_final;   // Python int() truncation
  ctrl[C_STEPS] = (double)it;
  ctrl[C_NACC] = tot[K];
  ctrl[C_MEAN_ALPHA] = all_alpha / n_all;
  ctrl[C_NPROP] += tot[K + 1];
  ctrl[C_ATT_RATIO] = tot[K + 1] / n_all;          // proposals drawn per walker in this step (1 = no redraws)
  if (tot[K + 2] != 0.0) ctrl[C_ERR] = tot[K + 2];
  if (it >= stop_at || tot[K + 2] != 0.0) ctrl[C_DONE] = 1.0;
}

// ------------------------------------------------------------------------------------------
// Persistent step loop shared by the compile-time-dimension and the wide kernels.
//
// One cooperative launch runs Metropolis steps until the stop rule fires (or max_steps).  Warps own
// 32-walker tiles (tile = global warp index + i * total warps: a fixed assignment, so every sum below has
// a fixed order).  Per step: every warp runs its tiles and accumulates (sum alpha per mode, accepted,
// proposals, error) -> CTA row -> grid_xreduce (one ticket + one flag round trip for the whole job, peers
// included) -> every CTA applies the same sigma / stop update to its own shared copy of the control block.
// Block 0 writes the control block back when the loop ends; the host reads it once per mutation.
//
// BODY: struct with
//   static constexpr int kWarps (warps per CTA), bool kSingleMode (K == 1 known at compile time)
//   static __device__ void tile(const StepArgs&, double* cta_smem, double* warp_smem, const double* ctrl_s,
//                               int64_t tile, int step, Acc& acc, double* warp_alpha /* [K] or null */)
struct TileAcc {
  double alpha;      // K == 1 fast path: this lane's running sum of alpha
  int accepted, nprop, err;
};

template <class BODY>
__device__ __forceinline__ void run_steps(const StepArgs& a, double* dyn_smem) {
  const int K = a.p.n_modes, W = K + 3;
  const int nctrl = ctrl_doubles(K);
  // shared layout: ctrl copy | part[W] | tot[W] | warp rows [kRunWarps][W] | body CTA area | per-warp areas
  double* s_ctrl = dyn_smem;
  double* s_part = s_ctrl + nctrl;
  double* s_tot = s_part + W;
  constexpr int NWARP = BODY::kWarps;
  double* s_wrow = s_tot + W;
  double* s_body = s_wrow + NWARP * W;
  const size_t body_cta = BODY::cta_doubles(a.p);
  const size_t body_warp = BODY::warp_doubles(a.p);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* s_warp = s_body + body_cta + (size_t)wid * body_warp;
  for (int e = threadIdx.x; e < nctrl; e += blockDim.x) s_ctrl[e] = a.ctrl[e];
  BODY::stage(a, s_body);
  __syncthreads();
  if (s_ctrl[C_DONE] != 0.0) return;
  GridSync* gs = reinterpret_cast<GridSync*>(a.ws);
  const int64_t n_tiles = (a.n + 31) / 32;
  const int64_t total_warps = (int64_t)gridDim.x * NWARP;
  const int64_t gw = (int64_t)blockIdx.x * NWARP + wid;
  ColumnFold fold;
  fold.max_cols = 1ull << (K + 2 < 64 ? K + 2 : 63);
  __shared__ int s_left[kRunWarpsMax];       // deferred walkers each warp has left at the end of a step
  for (int s = 0; s < a.max_steps; ++s) {
    const int step = (int)s_ctrl[C_STEPS];
    TileAcc acc;
    acc.alpha = 0.0; acc.accepted = 0; acc.nprop = 0; acc.err = 0;
    double* warp_alpha = s_wrow + wid * W;               // [K] per-mode sums of this warp (K > 1)
    if (lane < W) warp_alpha[lane] = 0.0;
    for (int c = 32 + lane; c < W; c += 32) warp_alpha[c] = 0.0;
    __syncwarp();
    if constexpr (BODY::kDeferred) {
      // Redraws are pooled over the warp's tiles.  A tile takes ONE round of proposals; walkers that are still outside
      // the unit cube go to a per-warp list (walker, attempts consumed) and are redrawn 32 at a time -- a full warp of
      // cooperative attempts -- as soon as 32 are waiting, the rest at the end of the step.  Per-tile redraw loops
      // otherwise cost a whole extra round whenever ONE of 32 walkers misses (half of all tiles at a 2 % miss rate)
      // and the spread of those rounds between warps is paid at the grid barrier of every step.  When most proposals
      // miss (first iterations) the tiles run their redraw loops to completion as before.  Which lane evaluates an
      // attempt never matters: variates are keyed by (walker slot, step, attempt).
      const double ratio = s_ctrl[C_ATT_RATIO];
      const bool defer = !(ratio >= 1.3);
      int* list_k = BODY::defer_list(s_warp);
      int* list_att = list_k + kDeferCap;
      int cnt = 0;
      int64_t tile = gw;
      int64_t pooled_k = -2;
      int pooled_att = 0;
      bool last = false;
      for (;;) {                       // ONE call site of the (large, fully unrolled) pass body
        int64_t k;
        int att0;
        bool single;
        if (cnt >= 32) {                 // a full warp of deferred walkers
          cnt -= 32;
          k = (int64_t)list_k[cnt + lane];
          att0 = list_att[cnt + lane];
          single = false;
          __syncwarp();
        } else if (pooled_k != -2) {     // this warp's share of the CTA's pooled leftovers (set below, once per step)
          k = pooled_k;
          att0 = pooled_att;
          single = false;
          pooled_k = -2;
          last = true;
        } else if (tile < n_tiles) {
          k = tile * 32 + lane;
          if (k >= a.n) k = -1;
          att0 = 0;
          single = defer;
          tile += total_warps;
        } else {
          if (last) break;
          // leftovers (< 32 per warp) are pooled over the CTA's warps: at a 0.3 % miss rate every warp would otherwise
          // spend a whole pass on one or two walkers at the end of every step
          last = true;
          if (lane == 0) s_left[wid] = cnt;
          __syncthreads();
          int total = 0, mine_w = -1, mine_i = 0;
          const int want = 32 * wid + lane;                  // warp c takes entries [32 c, 32 c + 32) of the concatenation
          for (int wv = 0; wv < NWARP; ++wv) {
            const int c = s_left[wv];
            if (want >= total && want < total + c) { mine_w = wv; mine_i = want - total; }
            total += c;
          }
          cnt = 0;
          if (32 * wid >= total) break;
          if (mine_w >= 0) {
            const int* lk = BODY::defer_list(s_body + body_cta + (size_t)mine_w * body_warp);
            pooled_k = lk[mine_i];
            pooled_att = lk[kDeferCap + mine_i];
          } else pooled_k = -1;
          continue;
        }
        unsigned def = 0u;
        int att = 0;
        BODY::pass(a, s_body, s_warp, s_ctrl, k, att0, single, step, acc, warp_alpha, def, att);
        if (def) {
          if ((def >> lane) & 1u) {
            const int pos = cnt + __popc(def & ((1u << lane) - 1u));
            list_k[pos] = (int)k;
            list_att[pos] = att;
          }
          cnt += __popc(def);
          __syncwarp();
        }
      }
    } else {
      for (int64_t tile = gw; tile < n_tiles; tile += total_warps)
        BODY::tile(a, s_body, s_warp, s_ctrl, tile, step, acc, warp_alpha);
    }
    // warp row (fixed lane order), CTA row (fixed warp order)
    {
      const double na = warp_sum((double)acc.accepted), npr = warp_sum((double)acc.nprop), ne = warp_max((double)acc.err);
      const double al = warp_sum(acc.alpha);
      if (lane == 0) {
        if (BODY::kSingleMode) warp_alpha[0] = al;
        warp_alpha[K] = na; warp_alpha[K + 1] = npr; warp_alpha[K + 2] = ne;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
      double v = s_wrow[c];
      if (c == K + 2) { for (int w = 1; w < NWARP; ++w) v = fmax(v, s_wrow[w * W + c]); }
      else { for (int w = 1; w < NWARP; ++w) v += s_wrow[w * W + c]; }
      s_part[c] = v;
    }
    __syncthreads();
    int rc;
    {
      const tb_xgpu xg = a.x;          // local copy: the kernel parameters themselves stay in the constant bank
      rc = grid_xreduce(gs, xg, s, W, s_part, s_tot, fold);
    }
    if (threadIdx.x == 0 && rc) s_tot[K + 2] = (double)rc;
    __syncthreads();
    if (a.p.defer_update) {
      // host-driven collective (no peer memory): leave this rank's totals for an all-reduce + tb_mcmc_update
      if (blockIdx.x == 0) for (int c = threadIdx.x; c < W; c += blockDim.x) a.ctrl[C_BASE + 3 * K + c] = s_tot[c];
      return;
    }
    if (threadIdx.x == 0) {
      const tb_mcmc_params pp = a.p;
      apply_step_update(pp, s_ctrl, s_tot, K);
    }
    __syncthreads();
    if (s_ctrl[C_DONE] != 0.0) break;
  }
  if (blockIdx.x == 0) for (int e = threadIdx.x; e < nctrl; e += blockDim.x) a.ctrl[e] = s_ctrl[e];
}

template <class BODY>
__host__ inline size_t run_smem_bytes(const tb_mcmc_params& p) {
  const int K = p.n_modes, W = K + 3;
  return sizeof(double) * ((size_t)ctrl_doubles(K) + 2 * W + (size_t)BODY::kWarps * W + BODY::cta_doubles(p) +
                           (size_t)BODY::kWarps * BODY::warp_doubles(p));
}

// runtime-dimension body with the warp-cooperative redraw (tb_mcmc_wide.cu)
int launch_wide(const StepArgs& a, cudaStream_t st);

// compile-time-dimension fast path, instantiated per dimension in tb_mcmc_fast_*.cu
template <int D>
int launch_fast(const StepArgs& a, cudaStream_t st);

// test hook: 0 = route single-mode runs through the multi-mode (shared-memory operand) instantiation
extern int g_allow_kone;

// cooperative launch of a persistent step kernel: as many CTAs as are co-resident, never more than there are tiles
template <class KERNEL>
__host__ inline int launch_persistent(KERNEL kernel, int warps, const StepArgs& a, size_t smem, cudaStream_t st) {
  const int kRunWarps = warps, kRunBlock = 32 * warps;
  int dev = 0, per_sm = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRunBlock, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) return TB_ERR_UNSUPPORTED;
  const int64_t n_tiles = (a.n + 31) / 32;
  int64_t grid = (n_tiles + kRunWarps - 1) / kRunWarps;
  const int64_t cap = (int64_t)per_sm * sm_count();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const int W = a.p.n_modes + 3;
  e = cudaMemsetAsync(a.ws, 0, sizeof(GridSync), st);
  if (e != cudaSuccess) return (int)e;
  (void)W;
  StepArgs copy = a;
  void* args[] = {(void*)&copy};
  e = cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)grid), dim3(kRunBlock), args, smem, st);
  return e == cudaSuccess ? TB_OK : (int)e;
}

}  // namespace tb
