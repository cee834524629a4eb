// Fused mutation kernels (a14, a15).
//   ref: tempest/steps/mutate.py:76-200 (warm-up draw, bookkeeping)
//        tempest/mcmc.py:142-208 (step loop), :225-288 (tpCN), :301-323 (RWM),
//        :104-135 (adaptive step count), :326-411 (boundaries)
//
// One thread per walker; one launch per Metropolis step.  A step proposes (Student-t scale,
// Cholesky-preconditioned pCN or random-walk move, boundary map, redraw until inside the unit
// cube), evaluates prior transform + likelihood in-kernel, accepts/rejects and leaves per-mode
// partial sums; the last CTA folds them in a fixed order, adapts sigma_c and evaluates the stop
// rule into a device-resident control block, so the next launch needs no host round trip (a
// launch made after the stop rule fired returns immediately).
// HBM traffic per step: the active set (u row, logl, q) is read once and written on accept;
// proposals, normals and likelihood terms never leave registers.
#include "tb_like.cuh"
#include "tb_xgpu.cuh"

static int tb_force_generic_mcmc = 0;

namespace {
using namespace tb;

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

__device__ void apply_step_update(const tb_mcmc_params& p, double* ctrl, const double* tot, int K);

// fold the per-CTA partials, adapt sigma and evaluate the stop rule (mcmc.py:180-194, 104-135)
__device__ void finish_step(const StepArgs& a, int K, int nparts) {
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
__device__ void apply_step_update(const tb_mcmc_params& p, double* ctrl, const double* tot, int K) {
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

// ------------------------------------------------------------------------------------------
// Fast path (compile-time D <= 16): three phases per CTA of 128 walkers.
//   A  one walker per thread: load state, Student-t scale (gamma draw), proposal centre -> smem
//   B  warp-cooperative rejection sampling: the lanes of a warp are dealt (walker, attempt) pairs
//      over the warp's UNFINISHED walkers, so lanes whose own walker is already inside the cube
//      evaluate further attempts of the others instead of idling in a divergent redraw loop; the
//      lowest attempt index that lands inside wins, exactly as the sequential redraw would
//   C  one walker per thread: prior transform, likelihood, Student-t ratio, accept/reject
// Philox counters are (walker slot, step, attempt), so which lane evaluates an attempt is irrelevant.
// Production normals: Box-Muller on 32-bit uniforms with fp32 log/sincos (like curand_normal),
// promoted to fp64; all state, likelihood and acceptance arithmetic is fp64.  Tape mode reads the
// recorded fp64 variates instead and is bit-compatible with the generic kernel.
__device__ __forceinline__ void bm_pair32(uint32_t a, uint32_t b, double& z0, double& z1) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]
  // MUFU-based log / sin / cos (abs. error ~2^-21 on [-pi, pi]): 5-7 % faster steps than logf / sincospif,
  // and far below the fp32 resolution the proposal noise already has
  const float r = __fsqrt_rn(-2.0f * __logf(u1));
  float s, c;
  __sincosf(((float)(int32_t)b) * 1.4629180792671596e-9f, &s, &c);  // angle in [-pi, pi): 2*pi*b/2^32 (signed)
  z0 = (double)(r * c);
  z1 = (double)(r * s);
}

// resident CTAs per SM the register allocation is held to (measured at D = 10: 6 CTAs / 80 registers
// beat 5 CTAs / 96 registers by ~3 %)
template <int D>
constexpr int mcmc_min_ctas() { return D <= 10 ? 6 : (D <= 12 ? 4 : 3); }

template <int D, bool TPCN, bool TAPE>
__global__ void __launch_bounds__(kMcmcBlock, mcmc_min_ctas<D>())
mcmc_step_fast(StepArgs a) {
  if (a.ctrl[C_DONE] != 0.0) return;
  constexpr int B = kMcmcBlock;
  constexpr int NCALL = (D + 3) / 4;                 // Philox blocks per attempt (4 normals each)
  extern __shared__ double sm[];
  __shared__ double s_x[D][B];                       // proposal centre, later the winning proposal
  __shared__ double s_cm[B];
  __shared__ int s_used[B];
  __shared__ int s_mode[B];
  const int K = a.p.n_modes;
  constexpr bool tpcn = TPCN;
  constexpr bool tape = TAPE;
  const int step = (int)a.ctrl[C_STEPS];
  double* s_mean = sm;
  double* s_chol = s_mean + K * D;
  double* s_inv = s_chol + K * D * D;                // symmetric form: diagonal as is, off-diagonals doubled (upper)
  double* s_dof = s_inv + K * D * D;
  double* s_sig = s_dof + K;
  double* s_part = s_sig + K;                        // [4][K] per-warp alpha sums
  double* s_prior = s_part + 4 * K;                  // [2*D] lo, scale
  for (int e = threadIdx.x; e < K * D; e += B) s_mean[e] = __ldg(a.p.mode_mean + e);
  for (int e = threadIdx.x; e < K * D * D; e += B) {
    s_chol[e] = __ldg(a.p.mode_chol + e);
    const int i = (e % (D * D)) / D, j = e % D;
    const double v = __ldg(a.p.mode_inv + e);
    s_inv[e] = (j > i) ? 2.0 * v : v;               // q = sum_i d_i * sum_{j>=i} S_ij d_j
  }
  for (int e = threadIdx.x; e < K; e += B) { s_dof[e] = __ldg(a.p.mode_dof + e); s_sig[e] = a.ctrl[C_BASE + e]; }
  for (int e = threadIdx.x; e < 2 * D; e += B) s_prior[e] = __ldg(a.p.prior_params + e);
  __syncthreads();

  const int t = threadIdx.x, lane = t & 31, wbase = t & ~31;
  const int64_t k = (int64_t)blockIdx.x * B + t;
  const bool valid = k < a.n;
  const Philox rng(a.p.seed, a.p.iteration);
  const uint64_t slot0 = (uint64_t)(a.p.slot_offset + (int64_t)blockIdx.x * B);   // slot of CTA-local walker 0
  const bool tape_over = tape && step >= a.tape.steps;
  int err = tape_over ? 1 : 0;
  int c = 0;
  double logl = 0.0, q = 0.0;
  uint32_t acc_word = 0u;
  bool have_acc_word = false;
  // ---- phase A -------------------------------------------------------------------------------
  if (valid) {
    c = a.assign ? a.assign[k] : 0;
    const double* mu = s_mean + c * D;
    const double sig = s_sig[c], dof = s_dof[c];
    const double* urow = a.u + k * D;
    logl = a.logl[k];
    double cm = sig, keep = 0.0;
    if (tpcn) {
      if (step == 0) {     // first step of this mutation: q of the freshly resampled state
        const double* IV = s_inv + c * D * D;
        double dq[D];
#pragma unroll
        for (int i = 0; i < D; ++i) dq[i] = urow[i] - mu[i];
        q = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double y = 0.0;
#pragma unroll
          for (int j = i; j < D; ++j) y += IV[i * D + j] * dq[j];
          q += dq[i] * y;
        }
        a.qcur[k] = q;
      } else q = a.qcur[k];
      double g;
      if (tape) g = tape_over ? 1.0 : a.tape.gamma[(int64_t)step * a.n + k];
      else {   // Marsaglia-Tsang with the squeeze test; shape = (D + nu)/2 >= 1
        const double shape = 0.5 * ((double)D + dof);
        const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        g = dd;
        for (uint32_t trial = 0; trial < 64; ++trial) {
          const uint4 r = rng.block((uint32_t)(slot0 + t), (uint32_t)((slot0 + t) >> 32), (uint32_t)step,
                                    (RNG_GAMMA << 24) | trial);
          double n0, n1;
          bm_pair32(r.x, r.y, n0, n1);
          const double v1 = 1.0 + cc * n0;
          if (v1 <= 0.0) continue;
          const double v = v1 * v1 * v1;
          const double uu = ((double)r.z + 0.5) * 2.3283064365386963e-10;
          const double x2 = n0 * n0;
          if (uu < 1.0 - 0.0331 * x2 * x2 || log(uu) < 0.5 * x2 + dd * (1.0 - v + log(v))) {
            g = dd * v; acc_word = r.w; have_acc_word = true;     // the 4th word of this block feeds the accept test
            break;
          }
        }
      }
      const double gscale = 2.0 / (dof + q);
      cm = sig * sqrt(1.0 / (gscale * g));
      keep = sqrt(__dsub_rn(1.0, __dmul_rn(sig, sig)));
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double ui = urow[i];
      s_x[i][t] = tpcn ? (mu[i] + keep * (ui - mu[i])) : ui;
    }
    s_cm[t] = cm;
    s_mode[t] = c;
  }
  s_used[t] = 0;
  __syncwarp();
  // ---- phase B: warp-cooperative redraw --------------------------------------------------------
  {
    int att = 0;                                          // next attempt index of MY walker
    const int n_att_tape = (tape && valid && !tape_over) ? a.tape.z_cnt[(int64_t)step * a.n + k] : 0;
    unsigned pending = __ballot_sync(0xffffffffu, valid && !tape_over);
    while (pending) {
      const int np = __popc(pending);
      const int sl = lane % np, off = lane / np;
      const int wsel = __fns(pending, 0, sl + 1);         // lane of the walker I work for
      const int aidx = __shfl_sync(0xffffffffu, att, wsel) + off;
      const int natt_sel = __shfl_sync(0xffffffffu, n_att_tape, wsel);
      const int wl = wbase + wsel;                        // CTA-local walker index
      const int cs = s_mode[wl];
      const double* L = s_chol + cs * D * D;
      const double cmul = s_cm[wl];
      double z[D];
      bool have = true;
      if (tape) {
        have = aidx < natt_sel;
        const int64_t gk = (int64_t)blockIdx.x * B + wl;
        const double* zt = a.tape.z + a.tape.z_off[(int64_t)step * a.n + gk] + (int64_t)(have ? aidx : 0) * D;
#pragma unroll
        for (int i = 0; i < D; ++i) z[i] = zt[i];
      } else {
        const uint64_t slot = slot0 + (uint64_t)wl;
#pragma unroll
        for (int cidx = 0; cidx < NCALL; ++cidx) {
          const uint4 r = rng.block((uint32_t)slot, (uint32_t)(slot >> 32), (uint32_t)step,
                                    (RNG_NORMAL << 24) | (uint32_t)((aidx * NCALL + cidx) & 0xffffff));
          double n0, n1, n2, n3;
          bm_pair32(r.x, r.y, n0, n1);
          bm_pair32(r.z, r.w, n2, n3);
          if (4 * cidx + 0 < D) z[4 * cidx + 0] = n0;
          if (4 * cidx + 1 < D) z[4 * cidx + 1] = n1;
          if (4 * cidx + 2 < D) z[4 * cidx + 2] = n2;
          if (4 * cidx + 3 < D) z[4 * cidx + 3] = n3;
        }
      }
      double prop[D];
      bool inside = have;
      if (!TAPE) {
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] *= cmul;        // (c L) z == L (c z) up to rounding; tape mode keeps the reference order
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double lz = 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) lz += TAPE ? (cmul * L[i * D + j]) * z[j] : L[i * D + j] * z[j];
        double v = s_x[i][wl] + lz;
        const int kind = a.p.bc_kind ? a.p.bc_kind[i] : 0;
        v = bc_apply(v, kind);
        if (kind == 0 && !(v >= 0.0 && v <= 1.0)) inside = false;
        prop[i] = v;
      }
      const unsigned okm = __ballot_sync(0xffffffffu, inside);
      const unsigned grp = __match_any_sync(0xffffffffu, wsel);
      const unsigned win = okm & grp;
      __syncwarp();
      if (win && lane == __ffs(win) - 1) {                // lowest attempt inside the cube wins
#pragma unroll
        for (int i = 0; i < D; ++i) s_x[i][wl] = prop[i];
        s_used[wl] = aidx + 1;
      }
      // bookkeeping for MY walker
      bool still = (pending >> lane) & 1u;
      const int my_sl = still ? __popc(pending & ((1u << lane) - 1u)) : 0;
      const unsigned my_grp = __shfl_sync(0xffffffffu, grp, my_sl);     // lane my_sl works for me at offset 0
      if (still) {
        if (okm & my_grp) still = false;
        else {
          att += __popc(my_grp);
          if (tape && att >= n_att_tape) { err = 1; still = false; s_used[t] = -1; }
          else if (att >= kMaxAttempts) { err = 2; still = false; s_used[t] = -1; }
        }
      }
      pending = __ballot_sync(0xffffffffu, still);
    }
  }
  __syncwarp();
  // ---- phase C -------------------------------------------------------------------------------
  double alpha = 0.0;
  int accepted = 0, nprop = 0;
  if (valid && s_used[t] > 0) {
    nprop = s_used[t];
    const double* mu = s_mean + c * D;
    const double* IV = s_inv + c * D * D;
    const double dof = s_dof[c];
    double prop[D], x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { prop[i] = s_x[i][t]; x[i] = __dadd_rn(s_prior[i], __dmul_rn(s_prior[D + i], prop[i])); }
    const double logl_new = eval_like(a.p.like_id, a.p.like_params, D, x);
    double factor = 0.0, q_new = 0.0;
    if (tpcn) {
      double dn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) dn[i] = prop[i] - mu[i];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double y = 0.0;
#pragma unroll
        for (int j = i; j < D; ++j) y += IV[i * D + j] * dn[j];
        q_new += dn[i] * y;
      }
      const double hd = -0.5 * ((double)D + dof);
      const double A = hd * log(1.0 + q_new / dof);
      const double Bq = hd * log(1.0 + q / dof);
      factor = __dadd_rn(-A, Bq);
    }
    double al = exp(__dadd_rn(__dmul_rn(a.p.beta, __dsub_rn(logl_new, logl)), factor));
    al = fmin(1.0, al);
    if (isnan(al)) al = 0.0;
    alpha = al;
    double ur;
    if (tape) ur = a.tape.acc_u[(int64_t)step * a.n + k];
    else if (have_acc_word) ur = ((double)acc_word + 0.5) * 2.3283064365386963e-10;   // 32-bit uniform in (0,1)
    else {
      const uint4 r = rng.block((uint32_t)(slot0 + t), (uint32_t)((slot0 + t) >> 32), (uint32_t)step, RNG_ACCEPT << 24);
      ur = u53(r.x, r.y);
    }
    if (ur < al) {
      accepted = 1;
      double* urow = a.u + k * D;
#pragma unroll
      for (int i = 0; i < D; ++i) urow[i] = prop[i];
      a.logl[k] = logl_new;
      if (tpcn) a.qcur[k] = q_new;
    }
  }
  // ---- CTA partials (fixed order) ---------------------------------------------------------------
  const int wid = t >> 5;
  for (int m = 0; m < K; ++m) {
    const double v = warp_sum((valid && c == m) ? alpha : 0.0);
    if (lane == 0) s_part[wid * K + m] = v;
  }
  __shared__ double s_tot[4][3];
  {
    const double na = warp_sum((double)accepted), npr = warp_sum((double)nprop), ne = warp_max((double)err);
    if (lane == 0) { s_tot[wid][0] = na; s_tot[wid][1] = npr; s_tot[wid][2] = ne; }
  }
  __syncthreads();
  const int W = K + 3;
  double* part = a.ws->partial + (size_t)blockIdx.x * W;
  for (int m = t; m < K; m += B) part[m] = ((s_part[m] + s_part[K + m]) + s_part[2 * K + m]) + s_part[3 * K + m];
  if (t == 0) {
    part[K] = ((s_tot[0][0] + s_tot[1][0]) + s_tot[2][0]) + s_tot[3][0];
    part[K + 1] = ((s_tot[0][1] + s_tot[1][1]) + s_tot[2][1]) + s_tot[3][1];
    part[K + 2] = fmax(fmax(s_tot[0][2], s_tot[1][2]), fmax(s_tot[2][2], s_tot[3][2]));
  }
  if (last_block_arrives(&a.ws->ticket)) finish_step(a, K, gridDim.x);
}

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
  for (int e = threadIdx.x; e < W; e += blockDim.x) a.ws->partial[(size_t)blockIdx.x * W + e] = s_part[e];
  if (last_block_arrives(&a.ws->ticket)) finish_step(a, K, gridDim.x);
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

template <int D, bool TPCN, bool TAPE>
int launch_fast_variant(const StepArgs& a, int count, cudaStream_t st) {
  const int K = a.p.n_modes;
  const size_t smem = sizeof(double) * ((size_t)K * D + 2 * (size_t)K * D * D + 2 * K + 4 * K + 2 * D);
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_step_fast<D, TPCN, TAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = (int)((a.n + kMcmcBlock - 1) / kMcmcBlock);
  for (int s = 0; s < count; ++s) mcmc_step_fast<D, TPCN, TAPE><<<grid, kMcmcBlock, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TB_OK : (int)e;
}

template <int D>
int launch_fast(const StepArgs& a, int count, cudaStream_t st) {
  const bool tpcn = a.p.sampler == TB_SAMPLE_TPCN, tape = a.p.rng_mode == TB_RNG_TAPE;
  if (tpcn) return tape ? launch_fast_variant<D, true, true>(a, count, st) : launch_fast_variant<D, true, false>(a, count, st);
  return tape ? launch_fast_variant<D, false, true>(a, count, st) : launch_fast_variant<D, false, false>(a, count, st);
}

inline bool has_fast_path(int d) {
  return d == 2 || d == 3 || d == 4 || d == 5 || d == 6 || d == 8 || d == 10 || d == 12 || d == 16;
}

bool params_ok(int64_t n, const tb_mcmc_params* p) {
  return p && n > 0 && p->n_dim > 0 && p->n_dim <= 128 && p->n_modes > 0 && p->n_modes <= kMaxModes &&
         p->prior_params && p->like_params;
}

}  // namespace

extern "C" {

int tb_set_mcmc_generic(int32_t on) { tb_force_generic_mcmc = on ? 1 : 0; return TB_OK; }

size_t tb_mcmc_workspace_bytes(int64_t n, int32_t n_modes) {
  const int64_t grid = (n + kMcmcBlock - 1) / kMcmcBlock;
  return 256 + sizeof(double) * (size_t)grid * (n_modes + 3);
}
size_t tb_mcmc_ctrl_doubles(int32_t n_modes) { return C_BASE + 4 * (size_t)n_modes + 3; }

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

int tb_mcmc_begin(int64_t n, const tb_mcmc_params* p, const int32_t* assign, const double* u, double* qcur,
                  void* workspace, double* ctrl, tb_stream_t stream) {
  if (!p || n <= 0 || p->n_dim <= 0 || p->n_dim > 128 || p->n_modes <= 0 || p->n_modes > kMaxModes || !u || !workspace ||
      !ctrl)
    return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, 16, st);
  if (e != cudaSuccess) return (int)e;
  mcmc_init_ctrl_kernel<<<1, 256, 0, st>>>(*p, ctrl);
  const int grid = (int)((n + kMcmcBlock - 1) / kMcmcBlock);
  // the fast step kernel computes q itself on its first step
  // (the generic and the split step kernels read it from qcur; like_id < 0 marks caller-evaluated likelihoods)
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
  StepArgs a;
  a.n = n; a.p = *p;
  if (tape) a.tape = *tape; else { a.tape.gamma = nullptr; a.tape.acc_u = nullptr; a.tape.z = nullptr;
                                    a.tape.z_off = nullptr; a.tape.z_cnt = nullptr; a.tape.steps = 0; }
  a.assign = assign; a.u = u; a.logl = logl; a.qcur = qcur; a.ws = (McmcWs*)workspace; a.ctrl = ctrl;
  a.ext_prop = nullptr; a.ext_logl = nullptr; a.ext_meta = nullptr;
  if (p->xgpu) {
    a.x = *p->xgpu;
    if (a.x.world < 1 || a.x.world > kXMaxRanks || a.x.seq < 1 || p->n_modes + 3 > 15) return TB_ERR_ARG;
    a.p.defer_update = 0;
  } else { a.x.rank = 0; a.x.world = 1; a.x.seq = 1; for (int i = 0; i < 8; ++i) a.x.peer[i] = nullptr; }
  a.p.xgpu = nullptr;
  cudaStream_t st = as_stream(stream);
  const int d = p->n_dim;
  if (!tb_force_generic_mcmc) {
    switch (d) {   // compile-time dimension: fully unrolled, register-resident fast path
      case 2: return launch_fast<2>(a, count, st);
      case 3: return launch_fast<3>(a, count, st);
      case 4: return launch_fast<4>(a, count, st);
      case 5: return launch_fast<5>(a, count, st);
      case 6: return launch_fast<6>(a, count, st);
      case 8: return launch_fast<8>(a, count, st);
      case 10: return launch_fast<10>(a, count, st);
      case 12: return launch_fast<12>(a, count, st);
      case 16: return launch_fast<16>(a, count, st);
      default: break;
    }
  }
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
  a.assign = assign; a.u = u; a.logl = logl; a.qcur = qcur; a.ws = (McmcWs*)workspace; a.ctrl = ctrl;
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

}  // extern "C"
