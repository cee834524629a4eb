// Shared pieces of the mutation kernels: control block, per-step argument block, boundary maps,
// the per-step fold (sigma adaptation + stop rule) and the fused cross-GPU exchange.
//   ref: tempest/mcmc.py:104-135, 180-194, 326-411
#pragma once
#include "tb_like.cuh"
#include "tb_xgpu.cuh"

namespace tb {

constexpr int kMcmcBlock = 128;
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

// fold the per-CTA partials, adapt sigma and evaluate the stop rule (mcmc.py:180-194, 104-135)
__device__ inline void finish_step(const StepArgs& a, int K, int nparts) {
  __shared__ double tot[kMaxModes + 3];
  const int W = K + 3;
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    double t = 0.0;
    for (int b = 0; b < nparts; ++b) t += __ldcg(a.ws->partial + (size_t)b * W + c);
    tot[c] = t;
  }
  __syncthreads();
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
    for (int c = threadIdx.x; c < W; c += blockDim.x) a.ctrl[C_BASE + 3 * K + c] = tot[c];
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


// compile-time-dimension fast path, instantiated per dimension in tb_mcmc_fast_*.cu
template <int D>
int launch_fast(const StepArgs& a, int count, cudaStream_t st);

}  // namespace tb
