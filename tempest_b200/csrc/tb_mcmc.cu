// Fused mutation kernels (a14, a15).
//   ref: tempest/steps/mutate.py:76-200 (warm-up draw, bookkeeping)
//        tempest/mcmc.py:142-208 (step loop), :225-288 (tpCN), :301-323 (RWM),
//        :104-135 (adaptive step count), :326-411 (boundaries)
//
// Registry likelihoods: ONE persistent cooperative launch per mutation (tb_mcmc_fast.cuh for compile-time
// n_dim <= 16, tb_mcmc_wide.cu for the rest; step loop, grid-wide + cross-GPU fold, sigma adaptation and
// stop rule in tb_mcmc_shared.cuh).  A step proposes (Student-t scale, Cholesky-preconditioned pCN or
// random-walk move, boundary map, redraw until inside the unit cube), evaluates prior transform +
// likelihood in-kernel and accepts / rejects; proposals, normals and likelihood terms never leave
// registers / shared memory, the active set (u row, logl, q) is read once per step and written on accept.
// Caller-evaluated likelihoods use the per-launch runtime-dimension kernel below (split propose / accept step),
// which is also the independent implementation the fused kernels are cross-checked against in the tests.
#include "tb_mcmc_shared.cuh"

static int tb_force_generic_mcmc = 0;
static int tb_wide_mcmc = 0;     // n_dim without a compile-time instantiation: warp-cooperative kernel instead of the per-thread one
namespace tb { int g_allow_kone = 1; }

namespace {
using namespace tb;

// fold the per-CTA partials of the per-launch kernel, adapt sigma and evaluate the stop rule (mcmc.py:180-194, 104-135)
__device__ inline void finish_step(const StepArgs& a, int K, int nparts) {
  __shared__ double tot[kMaxModes + 3];
  __shared__ double red[40];
  const int W = K + 3;
  McmcWs* ws = reinterpret_cast<McmcWs*>(a.ws);
  // every thread sums a strided share of the rows, then a fixed-order CTA sum per column
  for (int c = 0; c < W; ++c) {
    double t = 0.0;
    const bool mx = c == K + 2;       // error codes fold with max
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) {
      const double r = __ldcg(ws->partial + (size_t)b * W + c);
      t = mx ? fmax(t, r) : t + r;
    }
    t = mx ? block_max(t, red) : block_sum(t, red);
    if (threadIdx.x == 0) tot[c] = t;
  }
  __syncthreads();
  if (a.p.defer_update) {
    for (int c = threadIdx.x; c < W; c += blockDim.x) a.ctrl[C_BASE + 3 * K + c] = tot[c];
    return;
  }
  if (threadIdx.x == 0) apply_step_update(a.p, a.ctrl, tot, K);
}

// Runtime-dimension step kernel (any n_dim <= 128).
// PHASE 0: whole step with the in-kernel registry likelihood.  PHASE 1: proposal only (written to
// ext_prop / ext_meta).  PHASE 2: accept / reject against caller-evaluated ext_logl, sigma adaptation, stop rule.
template <int DP, int PHASE>
__global__ void __launch_bounds__(kMcmcBlock)
mcmc_step_kernel(StepArgs a) {
  if (a.ctrl[C_DONE] != 0.0) return;
  extern __shared__ double sm[];
  const int d = a.p.n_dim, K = a.p.n_modes;
  const bool tpcn = a.p.sampler == TB_SAMPLE_TPCN;
  const int step = (int)a.ctrl[C_STEPS];          // steps completed so far
  // stage mode statistics: mean[K][d], chol[K][d][d], inv[K][d][d], dof[K], sigma[K]
  double* s_mean = sm;
  double* s_chol = s_mean + K * d;
  double* s_inv = s_chol + K * d * d;
  double* s_dof = s_inv + K * d * d;
  double* s_sig = s_dof + K;
  double* s_part = s_sig + K;                      // [K+3] CTA partial sums
  for (int e = threadIdx.x; e < K * d; e += blockDim.x) s_mean[e] = __ldg(a.p.mode_mean + e);
  for (int e = threadIdx.x; e < K * d * d; e += blockDim.x) {
    s_chol[e] = __ldg(a.p.mode_chol + e);
    s_inv[e] = __ldg(a.p.mode_inv + e);
  }
  for (int e = threadIdx.x; e < K; e += blockDim.x) { s_dof[e] = __ldg(a.p.mode_dof + e); s_sig[e] = a.ctrl[C_BASE + e]; }
  for (int e = threadIdx.x; e < K + 3; e += blockDim.x) s_part[e] = 0.0;
  __syncthreads();

  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double alpha = 0.0;
  int accepted = 0, nprop = 0, err = 0, c = 0;
  if (k < a.n) {
    c = a.assign ? a.assign[k] : 0;
    const double* mu = s_mean + c * d;
    const double* L = s_chol + c * d * d;
    const double* IV = s_inv + c * d * d;
    const double sig = s_sig[c], dof = s_dof[c];
    double u[DP], diff[DP], z[DP], prop[DP];
    double* urow = a.u + k * d;
#pragma unroll
    for (int i = 0; i < DP; ++i) if (i < d) { u[i] = urow[i]; diff[i] = u[i] - mu[i]; }
    const double logl = a.logl[k];
    const uint64_t slot = (uint64_t)(a.p.slot_offset + k);
    const Philox rng(a.p.seed, a.p.iteration);
    const bool tape = a.p.rng_mode == TB_RNG_TAPE;
    const int64_t tix = (int64_t)step * a.n + k;
    if (tape && step >= a.tape.steps) err = 1;

    double scale_s = 0.0, q = 0.0, keep = 0.0;
    if (tpcn) q = a.qcur[k];
    if (PHASE == 2) {
      const int meta = a.ext_meta[k];
      if (meta > 0) nprop = meta; else err = -meta;
#pragma unroll
      for (int i = 0; i < DP; ++i) if (i < d) prop[i] = a.ext_prop[k * d + i];
    }
    if (tpcn && PHASE != 2) {
      // s = 1 / Gamma(shape=(d+nu)/2, scale=2/(nu+q))   (mcmc.py:233-236)
      double g;
      if (tape) g = err ? 1.0 : a.tape.gamma[tix];
      else {
        // Marsaglia-Tsang (shape >= 1 always: d + nu >= 2)
        const double shape = 0.5 * ((double)d + dof);
        const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        g = dd;
        for (uint32_t trial = 0; trial < 64; ++trial) {
          double n0, n1, u0, u1;
          philox_n2(rng, slot, (uint32_t)step, 0x800000u | trial, n0, n1);
          philox_u2(rng, slot, (uint32_t)step, RNG_GAMMA, trial, u0, u1, true);
          const double v1 = 1.0 + cc * n0;
          if (v1 <= 0.0) continue;
          const double v = v1 * v1 * v1;
          if (log(u0) < 0.5 * n0 * n0 + dd - dd * v + dd * log(v)) { g = dd * v; break; }
        }
      }
      const double gscale = 2.0 / (dof + q);
      const double s = 1.0 / (gscale * g);
      scale_s = sig * sqrt(s);
      keep = sqrt(__dsub_rn(1.0, __dmul_rn(sig, sig)));
    }
    // propose until inside the cube (redraw z only; mcmc.py:239-249 / :306-312)
    const int n_att_tape = tape && !err ? a.tape.z_cnt[tix] : 0;
    const double* ztape = tape && !err ? a.tape.z + a.tape.z_off[tix] : nullptr;
    bool inside = false;
    int attempt = 0;
    while (PHASE != 2 && !inside && !err) {
      if (tape) {
        if (attempt >= n_att_tape) { err = 1; break; }
#pragma unroll
        for (int i = 0; i < DP; ++i) if (i < d) z[i] = ztape[attempt * d + i];
      } else {
#pragma unroll
        for (int i = 0; i < DP; i += 2) if (i < d) {
          double z0, z1;
          philox_n2(rng, slot, (uint32_t)step, (uint32_t)(attempt * ((DP + 1) / 2) + i / 2), z0, z1);
          z[i] = z0;
          if (i + 1 < DP) z[i + 1] = z1;
        }
      }
      ++nprop;
      inside = true;
#pragma unroll
      for (int i = 0; i < DP; ++i) if (i < d) {
        double lz = 0.0;
        const double cmul = tpcn ? scale_s : sig;
        for (int j = 0; j <= i; ++j) lz += (cmul * L[i * d + j]) * z[j];
        double v = tpcn ? ((mu[i] + keep * diff[i]) + lz) : (u[i] + lz);
        const int kind = a.p.bc_kind ? a.p.bc_kind[i] : 0;
        v = bc_apply(v, kind);
        if (kind == 0 && !(v >= 0.0 && v <= 1.0)) inside = false;
        prop[i] = v;
      }
      ++attempt;
      if (attempt >= kMaxAttempts) { err = 2; break; }
    }
    if (PHASE == 1) {
      a.ext_meta[k] = err ? -err : nprop;
#pragma unroll
      for (int i = 0; i < DP; ++i) if (i < d) a.ext_prop[k * d + i] = err ? u[i] : prop[i];
    }
    if (PHASE != 1 && !err) {
      double logl_new;
      if (PHASE == 2) logl_new = a.ext_logl[k];
      else {   // prior transform + likelihood in registers
        double x[DP];
#pragma unroll
        for (int i = 0; i < DP; ++i) if (i < d) x[i] = prior_affine(a.p.prior_params, d, i, prop[i]);
        logl_new = eval_like(a.p.like_id, a.p.like_params, d, x);
      }
      double factor = 0.0, q_new = 0.0;
      if (tpcn) {   // Student-t density ratio (mcmc.py:251-279)
        double dn[DP];
#pragma unroll
        for (int i = 0; i < DP; ++i) if (i < d) dn[i] = prop[i] - mu[i];
        for (int j = 0; j < d; ++j) {
          double y = 0.0;
          for (int i = 0; i < d; ++i) y += dn[i] * IV[i * d + j];
          q_new += y * dn[j];
        }
        const double hd = -0.5 * ((double)d + dof);
        const double A = hd * log(1.0 + q_new / dof);
        const double B = hd * log(1.0 + q / dof);
        factor = __dadd_rn(-A, B);
      }
      double al = exp(__dadd_rn(__dmul_rn(a.p.beta, __dsub_rn(logl_new, logl)), factor));
      al = fmin(1.0, al);
      if (isnan(al)) al = 0.0;
      alpha = al;
      double ur;
      if (tape) ur = a.tape.acc_u[tix];
      else { double dummy; philox_u2(rng, slot, (uint32_t)step, RNG_ACCEPT, 0, ur, dummy, false); }
      if (ur < al) {
        accepted = 1;
#pragma unroll
        for (int i = 0; i < DP; ++i) if (i < d) urow[i] = prop[i];
        a.logl[k] = logl_new;
        if (tpcn) a.qcur[k] = q_new;
      }
    }
  }
  if (PHASE == 1) return;
  // CTA partials: shared atomics would make the order run-dependent; use a fixed-order fold
  __shared__ double fold[kMcmcBlock];
  __shared__ int foldc[kMcmcBlock];
  fold[threadIdx.x] = alpha;
  foldc[threadIdx.x] = (k < a.n) ? c : -1;
  __syncthreads();
  for (int m = threadIdx.x; m < K; m += blockDim.x) {
    double t = 0.0;
    for (int i = 0; i < kMcmcBlock; ++i) if (foldc[i] == m) t += fold[i];
    s_part[m] = t;
  }
  __shared__ double red[40];
  const double nacc = block_sum((double)accepted, red);
  const double npr = block_sum((double)nprop, red);
  const double ner = block_max((double)err, red);
  const int W = K + 3;
  if (threadIdx.x == 0) { s_part[K] = nacc; s_part[K + 1] = npr; s_part[K + 2] = ner; }
  __syncthreads();
  McmcWs* ws = reinterpret_cast<McmcWs*>(a.ws);
  for (int e = threadIdx.x; e < W; e += blockDim.x) ws->partial[(size_t)blockIdx.x * W + e] = s_part[e];
  if (last_block_arrives(&ws->ticket)) finish_step(a, K, gridDim.x);
}

// q_k = (u_k - mu_c)^T Sigma_c^{-1} (u_k - mu_c) for the initial state; per-mode walker counts; sigma init
__global__ void __launch_bounds__(kMcmcBlock)
mcmc_begin_kernel(int64_t n, tb_mcmc_params p, const int32_t* __restrict__ assign, const double* __restrict__ u,
                  double* __restrict__ qcur, double* __restrict__ ctrl) {
  const int d = p.n_dim, K = p.n_modes;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    const int c = assign ? assign[k] : 0;
    if (qcur) {
      const double* mu = p.mode_mean + c * d;
      const double* IV = p.mode_inv + (size_t)c * d * d;
      double q = 0.0;
      for (int j = 0; j < d; ++j) {
        double y = 0.0;
        for (int i = 0; i < d; ++i) y += (u[k * d + i] - __ldg(mu + i)) * __ldg(IV + i * d + j);
        q += y * (u[k * d + j] - __ldg(mu + j));
      }
      qcur[k] = q;
    }
    if (assign) atomicAdd(ctrl + C_BASE + K + c, 1.0);   // exact: integer counts < 2^53
  }
  if (!assign && blockIdx.x == 0 && threadIdx.x == 0) ctrl[C_BASE + K] = (double)n;   // single mode: every walker
}

__global__ void mcmc_update_kernel(tb_mcmc_params p, double* __restrict__ ctrl) {
  if (threadIdx.x == 0 && ctrl[C_DONE] == 0.0) apply_step_update(p, ctrl, ctrl + C_BASE + 3 * p.n_modes, p.n_modes);
}

__global__ void mcmc_init_ctrl_kernel(tb_mcmc_params p, double* __restrict__ ctrl) {
  const int K = p.n_modes;
  const double sigma0 = 2.38 / sqrt((double)p.n_dim);
  for (int e = threadIdx.x; e < C_BASE + 4 * K + 3; e += blockDim.x) {
    double v = 0.0;
    if (e == C_SIGMA0) v = sigma0;
    if (e >= C_BASE && e < C_BASE + K) v = (p.sampler == TB_SAMPLE_TPCN) ? fmin(sigma0, 0.99) : sigma0;  // mcmc.py:222-223, 298-299
    ctrl[e] = v;
  }
}

// warm-up: u ~ U(0,1)^d, x = prior(u), logl = L(x)   (mutate.py:100-105)
__global__ void __launch_bounds__(kMcmcBlock)
prior_draw_kernel(int64_t n, tb_mcmc_params p, const double* __restrict__ tape_u, const double* __restrict__ u_in,
                  double* __restrict__ u, double* __restrict__ x, double* __restrict__ logl) {
  const int d = p.n_dim;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const Philox rng(p.seed, p.iteration);
  const uint64_t slot = (uint64_t)(p.slot_offset + k);
  double xs[128];
  for (int i = 0; i < d; i += 2) {
    double a = 0.0, b = 0.0;
    if (u_in) { a = u_in[k * d + i]; if (i + 1 < d) b = u_in[k * d + i + 1]; }
    else if (tape_u) { a = tape_u[k * d + i]; if (i + 1 < d) b = tape_u[k * d + i + 1]; }
    else philox_u2(rng, slot, 0u, RNG_PRIOR, (uint32_t)(i / 2), a, b, false);
    if (u) u[k * d + i] = a;
    if (i + 1 < d && u) u[k * d + i + 1] = b;
    if (p.prior_params) {                      // NULL: caller-evaluated prior, only the uniforms are wanted
      xs[i] = prior_affine(p.prior_params, d, i, a);
      if (i + 1 < d) xs[i + 1] = prior_affine(p.prior_params, d, i + 1, b);
    }
  }
  if (x && p.prior_params) for (int i = 0; i < d; ++i) x[k * d + i] = xs[i];
  if (logl && p.prior_params && p.like_params && p.like_id >= 0) logl[k] = eval_like(p.like_id, p.like_params, d, xs);
}

// uniforms for the host-driven resampling / training draws (same Philox stream family)
__global__ void __launch_bounds__(kBlock)
philox_uniform_kernel(uint64_t seed, uint64_t iteration, uint32_t purpose, int64_t offset, int64_t n,
                      double* __restrict__ out) {
  const Philox rng(seed, iteration);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double a, b;
    philox_u2(rng, (uint64_t)(offset + i), 0u, purpose, 0u, a, b, false);
    out[i] = a;
  }
}

// the three variates a production step consumes, as the step kernels generate them (family 0: fused kernels,
// fp32 Box-Muller / Marsaglia-Tsang / spare-word accept uniform; family 1: per-launch kernel, all fp64)
__global__ void __launch_bounds__(kBlock)
debug_variates_kernel(uint64_t seed, uint64_t iteration, int64_t slot_offset, int64_t n, int step, int d, double shape,
                      int attempt, int family, double* __restrict__ gamma, double* __restrict__ z,
                      double* __restrict__ acc_u) {
  const Philox rng(seed, iteration);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint64_t slot = (uint64_t)(slot_offset + k);
    if (family == 0) {
      uint32_t w = 0u; bool have = false;
      const double g = gamma_mt(rng, slot, (uint32_t)step, shape, w, have);
      if (gamma) gamma[k] = g;
      if (acc_u) acc_u[k] = accept_uniform(rng, slot, (uint32_t)step, w, have);
      if (z) {
        const int ncall = (d + 3) / 4;
        for (int cidx = 0; cidx < ncall; ++cidx) {
          const uint4 r = rng.block((uint32_t)slot, (uint32_t)(slot >> 32), (uint32_t)step,
                                    (RNG_NORMAL << 24) | (uint32_t)((attempt * ncall + cidx) & 0xffffff));
          double nn[4];
          bm_pair32(r.x, r.y, nn[0], nn[1]);
          bm_pair32(r.z, r.w, nn[2], nn[3]);
          for (int e = 0; e < 4; ++e) if (4 * cidx + e < d) z[k * d + 4 * cidx + e] = nn[e];
        }
      }
    } else {
      const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
      double g = dd;
      for (uint32_t trial = 0; trial < 64; ++trial) {
        double n0, n1, u0, u1;
        philox_n2(rng, slot, (uint32_t)step, 0x800000u | trial, n0, n1);
        philox_u2(rng, slot, (uint32_t)step, RNG_GAMMA, trial, u0, u1, true);
        const double v1 = 1.0 + cc * n0;
        if (v1 <= 0.0) continue;
        const double v = v1 * v1 * v1;
        if (log(u0) < 0.5 * n0 * n0 + dd - dd * v + dd * log(v)) { g = dd * v; break; }
      }
      if (gamma) gamma[k] = g;
      if (acc_u) { double a, b; philox_u2(rng, slot, (uint32_t)step, RNG_ACCEPT, 0, a, b, false); acc_u[k] = a; }
      if (z) {
        const int dp = d <= 2 ? 2 : d <= 4 ? 4 : d <= 6 ? 6 : d <= 8 ? 8 : d <= 10 ? 10 : d <= 16 ? 16 : d <= 32 ? 32 : d <= 64 ? 64 : 128;
        for (int i = 0; i < d; i += 2) {
          double z0, z1;
          philox_n2(rng, slot, (uint32_t)step, (uint32_t)(attempt * ((dp + 1) / 2) + i / 2), z0, z1);
          z[k * d + i] = z0;
          if (i + 1 < d) z[k * d + i + 1] = z1;
        }
      }
    }
  }
}

template <int DP, int PHASE>
int launch_steps(const StepArgs& a, int count, cudaStream_t st) {
  const int d = a.p.n_dim, K = a.p.n_modes;
  const size_t smem = sizeof(double) * ((size_t)K * d + 2 * (size_t)K * d * d + 2 * K + K + 3);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_step_kernel<DP, PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = (int)((a.n + kMcmcBlock - 1) / kMcmcBlock);
  for (int s = 0; s < count; ++s) mcmc_step_kernel<DP, PHASE><<<grid, kMcmcBlock, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TB_OK : (int)e;
}

template <int PHASE>
int launch_generic(const StepArgs& a, int count, cudaStream_t st) {
  const int d = a.p.n_dim;
  if (d <= 2) return launch_steps<2, PHASE>(a, count, st);
  if (d <= 4) return launch_steps<4, PHASE>(a, count, st);
  if (d <= 6) return launch_steps<6, PHASE>(a, count, st);
  if (d <= 8) return launch_steps<8, PHASE>(a, count, st);
  if (d <= 10) return launch_steps<10, PHASE>(a, count, st);
  if (d <= 16) return launch_steps<16, PHASE>(a, count, st);
  if (d <= 32) return launch_steps<32, PHASE>(a, count, st);
  if (d <= 64) return launch_steps<64, PHASE>(a, count, st);
  return launch_steps<128, PHASE>(a, count, st);
}


inline bool has_fast_path(int d) {
  return d == 2 || d == 3 || d == 4 || d == 5 || d == 6 || d == 8 || d == 10 || d == 12 || d == 16;
}

bool params_ok(int64_t n, const tb_mcmc_params* p) {
  return p && n > 0 && p->n_dim > 0 && p->n_dim <= 128 && p->n_modes > 0 && p->n_modes <= kMaxModes &&
         p->prior_params && p->like_params;
}

}  // namespace

// ---- benchmark of the in-kernel synchronisation point (tools/xgpu_bench.py) ----------------------------------------------
// `count` back-to-back grid_xreduce calls of a 4-double row in one cooperative launch: with one CTA this is the pure
// cross-GPU exchange, with a full grid it is what every Metropolis step / ESS pass pays (row publish + ticket + last-CTA
// fold + peer stores / flags + broadcast to the waiting CTAs).
namespace tb {
__global__ void __launch_bounds__(128)
xgpu_bench_kernel(GridSync* gs, tb_xgpu x, int count, double* __restrict__ out) {
  __shared__ double s_part[4], s_tot[4];
  ColumnFold fold;
  fold.max_cols = 1ull << 3;
  double check = 0.0;
  const unsigned long long t0 = global_ns();
  for (int k = 0; k < count; ++k) {
    if (threadIdx.x < 4) s_part[threadIdx.x] = (threadIdx.x == 3) ? 0.0 : (double)(k & 7);
    __syncthreads();
    const int rc = grid_xreduce(gs, x, k, 4, s_part, s_tot, fold);
    if (rc) break;
    check += s_tot[0];
    __syncthreads();
  }
  const unsigned long long t1 = global_ns();
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = (double)(t1 - t0) / (double)count; out[1] = check; }
}
}  // namespace tb

extern "C" {

int tb_set_mcmc_generic(int32_t on) { tb_force_generic_mcmc = on ? 1 : 0; return TB_OK; }
int tb_set_mcmc_wide(int32_t on) { tb_wide_mcmc = on ? 1 : 0; return TB_OK; }
int tb_set_mcmc_kone(int32_t on) { tb::g_allow_kone = on ? 1 : 0; return TB_OK; }

size_t tb_mcmc_workspace_bytes(int64_t n, int32_t n_modes) {
  const int64_t grid = (n + kMcmcBlock - 1) / kMcmcBlock;
  const size_t flat = 256 + sizeof(double) * (size_t)grid * (n_modes + 3);           // per-launch kernel
  int64_t ctas = (n + 31) / 32;                                                         // persistent kernels (>= 1 warp per CTA)
  const int64_t cap = (int64_t)tb::sm_count() * 32;
  if (ctas > cap) ctas = cap;
  const size_t pers = 256 + grid_sync_bytes((int)ctas, n_modes + 3);
  return flat > pers ? flat : pers;
}
size_t tb_mcmc_ctrl_doubles(int32_t n_modes) { return (size_t)ctrl_doubles(n_modes); }

int tb_mcmc_update(const tb_mcmc_params* p, double* ctrl, tb_stream_t stream) {
  if (!p || !ctrl || p->n_modes <= 0 || p->n_modes > kMaxModes) return TB_ERR_ARG;
  mcmc_update_kernel<<<1, 32, 0, as_stream(stream)>>>(*p, ctrl);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_prior_draw(int64_t n, const tb_mcmc_params* p, const double* prior_u_tape, double* u, double* x,
                  double* logl, tb_stream_t stream) {
  if (!p || n <= 0 || p->n_dim <= 0 || p->n_dim > 128 || !u) return TB_ERR_ARG;
  if ((x && !p->prior_params) || (logl && (!p->prior_params || !p->like_params || p->like_id < 0))) return TB_ERR_ARG;
  const int grid = (int)((n + kMcmcBlock - 1) / kMcmcBlock);
  prior_draw_kernel<<<grid, kMcmcBlock, 0, as_stream(stream)>>>(n, *p, prior_u_tape, nullptr, u, x, logl);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_transform(const double* u, int64_t n, const tb_mcmc_params* p, double* x, double* logl, tb_stream_t stream) {
  if (!p || n <= 0 || p->n_dim <= 0 || p->n_dim > 128 || !u || !p->prior_params) return TB_ERR_ARG;
  if (logl && (!p->like_params || p->like_id < 0)) return TB_ERR_ARG;
  const int grid = (int)((n + kMcmcBlock - 1) / kMcmcBlock);
  prior_draw_kernel<<<grid, kMcmcBlock, 0, as_stream(stream)>>>(n, *p, nullptr, u, nullptr, x, logl);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_philox_uniform(uint64_t seed, uint64_t iteration, uint32_t purpose, int64_t offset, int64_t n, double* out,
                      tb_stream_t stream) {
  if (n <= 0 || !out) return TB_ERR_ARG;
  philox_uniform_kernel<<<stream_grid(n, kBlock, 16), kBlock, 0, as_stream(stream)>>>(seed, iteration, purpose,
                                                                                     offset, n, out);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_debug_variates(uint64_t seed, uint64_t iteration, int64_t slot_offset, int64_t n, int32_t step, int32_t d,
                      double shape, int32_t attempt, int32_t family, double* gamma, double* z, double* acc_u,
                      tb_stream_t stream) {
  if (n <= 0 || d <= 0 || d > 128 || !(shape >= 1.0) || attempt < 0 || (family != 0 && family != 1)) return TB_ERR_ARG;
  debug_variates_kernel<<<stream_grid(n, kBlock, 16), kBlock, 0, as_stream(stream)>>>(
      seed, iteration, slot_offset, n, step, d, shape, attempt, family, gamma, z, acc_u);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_mcmc_begin(int64_t n, const tb_mcmc_params* p, const int32_t* assign, const double* u, double* qcur,
                  void* workspace, double* ctrl, tb_stream_t stream) {
  if (!p || n <= 0 || p->n_dim <= 0 || p->n_dim > 128 || p->n_modes <= 0 || p->n_modes > kMaxModes || !u || !workspace ||
      !ctrl)
    return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, 16, st);       // ticket of the per-launch kernel
  if (e != cudaSuccess) return (int)e;
  mcmc_init_ctrl_kernel<<<1, 256, 0, st>>>(*p, ctrl);
  const int grid = (int)((n + kMcmcBlock - 1) / kMcmcBlock);
  // the compile-time-dimension kernel computes q itself on its first step (the others read it from qcur;
  // like_id < 0 marks caller-evaluated likelihoods)
  const bool need_q = p->sampler == TB_SAMPLE_TPCN &&
                      (tb_force_generic_mcmc || !has_fast_path(p->n_dim) || p->like_id < 0);
  mcmc_begin_kernel<<<grid, kMcmcBlock, 0, st>>>(n, *p, assign, u, need_q ? qcur : nullptr, ctrl);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_mcmc_steps(int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign, double* u,
                  double* logl, double* qcur, void* workspace, double* ctrl, int32_t count, tb_stream_t stream) {
  if (!params_ok(n, p) || !u || !logl || !workspace || !ctrl || count < 0) return TB_ERR_ARG;
  if (p->sampler == TB_SAMPLE_TPCN && !qcur) return TB_ERR_ARG;
  if (p->rng_mode == TB_RNG_TAPE && !tape) return TB_ERR_ARG;
  if (count == 0) return TB_OK;
  StepArgs a;
  a.n = n; a.p = *p;
  if (tape) a.tape = *tape; else { a.tape.gamma = nullptr; a.tape.acc_u = nullptr; a.tape.z = nullptr;
                                    a.tape.z_off = nullptr; a.tape.z_cnt = nullptr; a.tape.steps = 0; }
  a.assign = assign; a.u = u; a.logl = logl; a.qcur = qcur; a.ws = workspace; a.ctrl = ctrl;
  a.ext_prop = nullptr; a.ext_logl = nullptr; a.ext_meta = nullptr;
  a.max_steps = count;
  if (p->xgpu) {
    a.x = *p->xgpu;
    if (a.x.world < 1 || a.x.world > kXMaxRanks || a.x.seq < 1 || p->n_modes + 3 > kXMaxPayload) return TB_ERR_ARG;
    a.p.defer_update = 0;
  } else { a.x.rank = 0; a.x.world = 1; a.x.seq = 1; for (int i = 0; i < 8; ++i) a.x.peer[i] = nullptr; }
  a.p.xgpu = nullptr;
  cudaStream_t st = as_stream(stream);
  const int d = p->n_dim;
  if (!tb_force_generic_mcmc) {
    switch (d) {   // compile-time dimension: fully unrolled, register-resident fast path
      case 2: return launch_fast<2>(a, st);
      case 3: return launch_fast<3>(a, st);
      case 4: return launch_fast<4>(a, st);
      case 5: return launch_fast<5>(a, st);
      case 6: return launch_fast<6>(a, st);
      case 8: return launch_fast<8>(a, st);
      case 10: return launch_fast<10>(a, st);
      case 12: return launch_fast<12>(a, st);
      case 16: return launch_fast<16>(a, st);
      default: break;
    }
  }
  if (tb_wide_mcmc && !tb_force_generic_mcmc && p->like_id >= 0) return launch_wide(a, st);
  return launch_generic<0>(a, count, st);
}

// one Metropolis step split around a caller-evaluated likelihood (arbitrary user callables):
//   tb_mcmc_propose -> caller: x = prior(u_prop), logl_prop = L(x) -> tb_mcmc_accept
static int split_args(StepArgs& a, int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign,
                      double* u, double* logl, double* qcur, void* workspace, double* ctrl) {
  if (!p || n <= 0 || p->n_dim <= 0 || p->n_dim > 128 || p->n_modes <= 0 || p->n_modes > kMaxModes || !u || !workspace ||
      !ctrl)
    return TB_ERR_ARG;
  if (p->sampler == TB_SAMPLE_TPCN && !qcur) return TB_ERR_ARG;
  if (p->rng_mode == TB_RNG_TAPE && !tape) return TB_ERR_ARG;
  if (p->xgpu) return TB_ERR_ARG;             // the split step is single-GPU
  a.n = n; a.p = *p;
  if (tape) a.tape = *tape; else { a.tape.gamma = nullptr; a.tape.acc_u = nullptr; a.tape.z = nullptr;
                                    a.tape.z_off = nullptr; a.tape.z_cnt = nullptr; a.tape.steps = 0; }
  a.assign = assign; a.u = u; a.logl = logl; a.qcur = qcur; a.ws = workspace; a.ctrl = ctrl; a.max_steps = 1;
  a.x.rank = 0; a.x.world = 1; a.x.seq = 1;
  for (int i = 0; i < 8; ++i) a.x.peer[i] = nullptr;
  return TB_OK;
}

int tb_mcmc_propose(int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign, const double* u,
                    const double* logl, const double* qcur, void* workspace, double* ctrl, double* u_prop,
                    int32_t* meta, tb_stream_t stream) {
  StepArgs a;
  if (int rc = split_args(a, n, p, tape, assign, const_cast<double*>(u), const_cast<double*>(logl),
                          const_cast<double*>(qcur), workspace, ctrl))
    return rc;
  if (!u_prop || !meta || !logl) return TB_ERR_ARG;
  a.ext_prop = u_prop; a.ext_logl = nullptr; a.ext_meta = meta;
  return launch_generic<1>(a, 1, as_stream(stream));
}

int tb_mcmc_accept(int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign, double* u,
                   double* logl, double* qcur, void* workspace, double* ctrl, const double* u_prop,
                   const double* logl_prop, const int32_t* meta, tb_stream_t stream) {
  StepArgs a;
  if (int rc = split_args(a, n, p, tape, assign, u, logl, qcur, workspace, ctrl)) return rc;
  if (!u_prop || !logl_prop || !meta || !logl) return TB_ERR_ARG;
  a.ext_prop = const_cast<double*>(u_prop); a.ext_logl = logl_prop; a.ext_meta = const_cast<int32_t*>(meta);
  return launch_generic<2>(a, 1, as_stream(stream));
}

size_t tb_xgpu_bench_workspace_bytes(int32_t grid) { return tb::grid_sync_bytes(grid, 4); }

// out2 (device): {ns per synchronisation point, checksum}.  x may be NULL (one GPU).  Consumes `count` exchanges.
int tb_xgpu_bench(const tb_xgpu* xgpu, int32_t grid, int32_t count, void* workspace, double* out2, tb_stream_t stream) {
  if (grid < 1 || count < 1 || !workspace || !out2) return TB_ERR_ARG;
  tb_xgpu xg;
  if (xgpu) xg = *xgpu;
  else { xg.rank = 0; xg.world = 1; xg.seq = 1; for (int i = 0; i < 8; ++i) xg.peer[i] = nullptr; }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tb::xgpu_bench_kernel, 128, 0);
  if (e != cudaSuccess) return (int)e;
  if (grid > per_sm * tb::sm_count()) return TB_ERR_UNSUPPORTED;
  cudaStream_t st = tb::as_stream(stream);
  tb::GridSync* gs = reinterpret_cast<tb::GridSync*>(workspace);
  e = cudaMemsetAsync(gs, 0, sizeof(tb::GridSync), st);
  if (e != cudaSuccess) return (int)e;
  void* args[] = {(void*)&gs, (void*)&xg, (void*)&count, (void*)&out2};
  e = cudaLaunchCooperativeKernel((void*)tb::xgpu_bench_kernel, dim3(grid), dim3(128), args, 0, st);
  return e == cudaSuccess ? TB_OK : (int)e;
}

}  // extern "C"
