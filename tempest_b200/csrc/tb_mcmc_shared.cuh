// Shared pieces of the mutation kernels: control block, per-step argument block, boundary maps,
// the per-step fold (sigma adaptation + stop rule) and the fused cross-GPU exchange.
//   ref: tempest/mcmc.py:104-135, 180-194, 326-411
#pragma once
#include "tb_like.cuh"
#include "tb_xgpu.cuh"

namespace tb {

constexpr int kMcmcBlock = 128;      // walkers per CTA of the runtime-dimension kernels
#ifndef TB_FAST_BLOCK
#define TB_FAST_BLOCK 32
#endif
constexpr int kFastBlock = TB_FAST_BLOCK;   // walkers per CTA of the compile-time-dimension kernel
constexpr int kFoldGroup = 32;              // CTAs whose partials one group leader folds
constexpr int kMaxModes = 64;
constexpr int kMaxAttempts = 100000;

// control block indices (doubles)
enum { C_STEPS = 0, C_DONE = 1, C_NACC = 2, C_MEAN_ALPHA = 3, C_ERR = 4, C_NPROP = 5, C_SIGMA0 = 6, C_BASE = 8 };

struct McmcWs {
  unsigned int ticket;
  unsigned int pad[3];
  double partial[1];  // [grid][K+3]: sum alpha per mode, n accepted, n proposals, error flag
};

struct StepArgs {
  int64_t n;
  tb_mcmc_params p;
  tb_tape tape;
  const int32_t* assign;
  double* u;
  double* logl;
  double* qcur;
  McmcWs* ws;
  double* ctrl;
  tb_xgpu x;          // world > 1: the per-step totals are exchanged over peer memory inside the kernel
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

__device__ inline void apply_step_update(const tb_mcmc_params& p, double* ctrl, const double* tot, int K);

// Workspace of the fast kernel (after the 16-byte header with the global ticket): one ticket per group
// of kFoldGroup CTAs, the group partials [ngroups][W], then the CTA partials [grid][W].
__host__ __device__ inline int fold_groups(int grid) { return (grid + kFoldGroup - 1) / kFoldGroup; }
__host__ __device__ inline size_t fold_ticket_bytes(int grid) { return ((size_t)fold_groups(grid) * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t fold_workspace_bytes(int grid, int W) {
  return 16 + fold_ticket_bytes(grid) + sizeof(double) * (size_t)W * ((size_t)fold_groups(grid) + (size_t)grid);
}
__device__ inline unsigned int* fold_tickets(McmcWs* ws) { return reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + 16); }
__device__ inline double* fold_group_partials(McmcWs* ws, int grid) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 16 + fold_ticket_bytes(grid));
}
__device__ inline double* fold_cta_partials(McmcWs* ws, int grid, int W) {
  return fold_group_partials(ws, grid) + (size_t)fold_groups(grid) * W;
}

__device__ inline void finish_tail(const StepArgs& a, int K, double* tot, int nth);

// fold the per-CTA partials, adapt sigma and evaluate the stop rule (mcmc.py:180-194, 104-135)
__device__ inline void finish_step(const StepArgs& a, int K, int nparts) {
  __shared__ double tot[kMaxModes + 3];
  __shared__ double red[40];
  const int W = K + 3;
  // every thread sums a strided share of the rows, then a fixed-order CTA sum per column (a single
  // thread walking all rows costs ~30 ns per dependent L2 load: 0.25 ms at 8192 CTAs)
  for (int c = 0; c < W; ++c) {
    double t = 0.0;
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) t += __ldcg(a.ws->partial + (size_t)b * W + c);
    t = block_sum(t, red);
    if (threadIdx.x == 0) tot[c] = t;
  }
  __syncthreads();
  finish_tail(a, K, tot, blockDim.x);
}

// Hierarchical, order-fixed fold for the fast kernel, executed by warp 0 of every CTA after the CTA's
// partial row is in memory: the last CTA of each group of kFoldGroup folds the group, the last group
// leader folds the groups and applies the update.  No CTA ever waits for another.
__device__ inline void arrive_and_fold(const StepArgs& a, int K, double* tot /* shared, >= K+3 */) {
  const int W = K + 3, grid = gridDim.x, lane = threadIdx.x & 31;
  const int ngroups = fold_groups(grid);
  unsigned int* gticket = fold_tickets(a.ws);
  double* gpart = fold_group_partials(a.ws, grid);
  const double* cpart = fold_cta_partials(a.ws, grid, W);
  const int g = blockIdx.x / kFoldGroup;
  const int gsize = min(kFoldGroup, grid - g * kFoldGroup);
  __threadfence();
  __syncwarp();
  unsigned int tk = 0;
  if (lane == 0) tk = atomicAdd(&gticket[g], 1u);
  tk = __shfl_sync(0xffffffffu, tk, 0);
  if (tk != (unsigned)(gsize - 1)) return;
  if (lane == 0) gticket[g] = 0u;
  __threadfence();
  for (int c = 0; c < W; ++c) {
    double v = (lane < gsize) ? __ldcg(cpart + ((size_t)g * kFoldGroup + lane) * W + c) : 0.0;
    v = warp_sum(v);
    if (lane == 0) gpart[(size_t)g * W + c] = v;
  }
  __threadfence();
  __syncwarp();
  unsigned int t2 = 0;
  if (lane == 0) t2 = atomicAdd(&a.ws->ticket, 1u);
  t2 = __shfl_sync(0xffffffffu, t2, 0);
  if (t2 != (unsigned)(ngroups - 1)) return;
  if (lane == 0) a.ws->ticket = 0u;
  __threadfence();
  for (int c = 0; c < W; ++c) {
    double v = 0.0;
    for (int q = lane; q < ngroups; q += 32) v += __ldcg(gpart + (size_t)q * W + c);
    v = warp_sum(v);
    if (lane == 0) tot[c] = v;
  }
  __syncwarp();
  finish_tail(a, K, tot, 32);
}

// `nth` = number of threads (threadIdx.x < nth) executing this call
__device__ inline void finish_tail(const StepArgs& a, int K, double* tot, int nth) {
  const int W = K + 3;
  if (a.x.world > 1) {
    // fused collective: exchange this rank's (sum alpha per mode, accepted, proposals, error) with every
    // peer over NVLink and fold them in rank order, then adapt sigma / evaluate the stop rule right here
    if (threadIdx.x == 0) {
      double all[kXMaxRanks * 15];
      const unsigned long long seq = a.x.seq + (unsigned long long)a.ctrl[C_STEPS];
      xgpu_allgather(a.x, seq, tot, W, all);
      for (int c = 0; c < W; ++c) {
        double t = 0.0;
        for (int r = 0; r < a.x.world; ++r) t += all[r * W + c];
        tot[c] = t;
      }
      apply_step_update(a.p, a.ctrl, tot, K);
    }
    return;
  }
  if (a.p.defer_update) {
    // sharded run: leave this rank's totals for the host to all-reduce; tb_mcmc_update applies them
    for (int c = threadIdx.x; c < W; c += nth) a.ctrl[C_BASE + 3 * K + c] = tot[c];
    return;
  }
  if (threadIdx.x == 0) apply_step_update(a.p, a.ctrl, tot, K);
}

// adapt sigma and evaluate the stop rule from the (global) per-mode totals
__device__ inline void apply_step_update(const tb_mcmc_params& p, double* ctrl, const double* tot, int K) {
  {
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
    if (tot[K + 2] != 0.0) ctrl[C_ERR] = tot[K + 2];
    if (it >= stop_at || tot[K + 2] != 0.0) ctrl[C_DONE] = 1.0;
  }
}


// runtime-dimension step with the warp-cooperative redraw (tb_mcmc_wide.cu)
int launch_wide(const StepArgs& a, int count, cudaStream_t st);

// compile-time-dimension fast path, instantiated per dimension in tb_mcmc_fast_*.cu
template <int D>
int launch_fast(const StepArgs& a, int count, cudaStream_t st);

}  // namespace tb
