// Compile-time-dimension Metropolis step (the hot kernel of the mutation stage).  Included by the
// tb_mcmc_fast_*.cu translation units, one or two dimensions each, so they compile in parallel.
#pragma once
#include "tb_mcmc_shared.cuh"

namespace tb {

// ------------------------------------------------------------------------------------------
// Fast path (compile-time D <= 16): one warp works on one tile of 32 walkers in three phases.
//   A  one walker per lane: load state, Student-t scale (gamma draw), proposal centre -> smem
//   B  warp-cooperative rejection sampling: the lanes are dealt (walker, attempt) pairs over the tile's
//      UNFINISHED walkers, so lanes whose own walker is already inside the cube evaluate further attempts
//      of the others instead of idling in a divergent redraw loop; the lowest attempt index that lands
//      inside wins, exactly as the sequential redraw (mcmc.py:239-249) would
//   C  one walker per lane: prior transform, likelihood, Student-t ratio, accept/reject
// Philox counters are (walker slot, step, attempt), so which lane evaluates an attempt is irrelevant.
// TAPE only changes where the three variates of a step come from (recorded gamma / normals / accept
// uniform instead of gamma_mt / normals_fixed / accept_uniform): the parity tests therefore run the
// production arithmetic, including the constant-memory single-mode variant and the compile-time likelihood.

// Two CTAs per SM; warps per CTA such that the register allocation is held to 72 / 128 / 168 registers
// (measured at D = 10 in round 1: 28 warps / 72 registers beat 24 / 80).  Few, large CTAs keep the per-step grid
// synchronisation short: 296 rows to fold instead of 1036 (tools/xgpu_bench.py: 4.8 vs 7.2 us per step).
template <int D>
constexpr int mcmc_warps() { return (D <= 10 ? 14 : (D <= 12 ? 8 : 6)); }

// Single-mode runs (K = 1, the clustering=False headline path) read the mode statistics and the prior box
// from constant memory: the operands fold into the DFMAs, so the ~3 D^2/2 shared-memory loads per
// walker-step and the staging loop disappear.  Layout: mean[D], chol[D*D], inv[D*D] (row-major, as stored),
// prior lo[D], scale[D], dof.  One copy per translation unit (static).
constexpr int kConstDoubles = 16 + 2 * 256 + 32 + 1;
static __constant__ double c_mode[kConstDoubles];

// LIKE >= 0: the registry likelihood is fixed at compile time (single-mode variants), so the other
// likelihood bodies are not in the instruction stream; LIKE = -1 switches on a.p.like_id.
template <int D, bool TPCN, bool TAPE, bool KONE, int LIKE>
struct FastBody {
  static constexpr bool kSingleMode = KONE;
  static constexpr bool kDeferred = true;       // run_steps drives pass() with the deferred-redraw list
  static constexpr int kWarps = mcmc_warps<D>();
  static constexpr int CM_CHOL = D, CM_INV = D + D * D, CM_PRIOR = D + 2 * D * D, CM_DOF = D + 2 * D * D + 2 * D;

  // CTA area (multi-mode variant only): mean[K][D], chol[K][D][D], inv[K][D][D] (symmetric form: off-diagonals
  // of the upper triangle doubled), dof[K], prior lo[D], scale[D]
  __host__ __device__ static size_t cta_doubles(const tb_mcmc_params& p) {
    return KONE ? 0 : (size_t)p.n_modes * D + 2 * (size_t)p.n_modes * D * D + p.n_modes + 2 * D;
  }
  // per-warp area: proposal centre / winning proposal [D][32], proposal scale [32], attempts used [32], mode [32],
  // deferred-redraw list (walker, attempts consumed) [kDeferCap] each
  __host__ __device__ static size_t warp_doubles(const tb_mcmc_params&) { return (size_t)D * 32 + 32 + 32 + kDeferCap; }
  __device__ static int* defer_list(double* s_warp) { return reinterpret_cast<int*>(s_warp + D * 32 + 32 + 32); }

  __device__ static void stage(const StepArgs& a, double* s_body) {
    if (KONE) return;
    const int K = a.p.n_modes, B = blockDim.x;
    double* s_mean = s_body;
    double* s_chol = s_mean + K * D;
    double* s_inv = s_chol + K * D * D;
    double* s_dof = s_inv + K * D * D;
    double* s_prior = s_dof + K;
    for (int e = threadIdx.x; e < K * D; e += B) s_mean[e] = __ldg(a.p.mode_mean + e);
    for (int e = threadIdx.x; e < K * D * D; e += B) {
      s_chol[e] = __ldg(a.p.mode_chol + e);
      const int i = (e % (D * D)) / D, j = e % D;
      const double v = __ldg(a.p.mode_inv + e);
      s_inv[e] = (j > i) ? 2.0 * v : v;             // q = sum_i d_i * sum_{j>=i} S_ij d_j
    }
    for (int e = threadIdx.x; e < K; e += B) s_dof[e] = __ldg(a.p.mode_dof + e);
    for (int e = threadIdx.x; e < 2 * D; e += B) s_prior[e] = __ldg(a.p.prior_params + e);
  }

  // One pass over up to 32 walkers: lane's walker k (< 0: none) with att0 attempts already consumed.  single_round:
  // evaluate ONE round of the cooperative redraw only and report the walkers that are still outside the cube in
  // `deferred` (their consumed attempts in att_out) instead of redrawing them here; they skip phase C.
  __device__ static __forceinline__ void pass(const StepArgs& a, double* s_body, double* s_warp, const double* s_ctrl,
                                              const int64_t k, const int att0, const bool single_round, int step,
                                              TileAcc& acc, double* warp_alpha, unsigned& deferred, int& att_out) {
    const int K = a.p.n_modes;
    const double* s_mean = s_body;
    const double* s_chol = s_mean + K * D;
    const double* s_inv = s_chol + K * D * D;
    const double* s_dof = s_inv + K * D * D;
    const double* s_prior = s_dof + K;
    double* s_x = s_warp;                              // [D][32]
    double* s_cm = s_x + D * 32;                       // [32]
    int* s_used = reinterpret_cast<int*>(s_cm + 32);   // [32]
    int* s_mode = s_used + 32;                         // [32]
    // quadratic form (u - mu)^T Sigma^-1 (u - mu) of mode `cm` from centred coordinates
    auto quad = [&](const double (&dv)[D], int cm) -> double {
      double r = 0.0;
      if (KONE) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double y = 0.0;
#pragma unroll
          for (int j = i + 1; j < D; ++j) y += c_mode[CM_INV + i * D + j] * dv[j];
          r += dv[i] * (c_mode[CM_INV + i * D + i] * dv[i] + 2.0 * y);
        }
      } else {
        const double* IV = s_inv + cm * D * D;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double y = 0.0;
#pragma unroll
          for (int j = i; j < D; ++j) y += IV[i * D + j] * dv[j];
          r += dv[i] * y;
        }
      }
      return r;
    };

    const int lane = threadIdx.x & 31;
    const bool valid = k >= 0 && k < a.n;
    const Philox rng(a.p.seed, a.p.iteration);
    const uint64_t slot = (uint64_t)(a.p.slot_offset + (valid ? k : 0));   // global walker slot: the Philox counter
    const bool tape_over = TAPE && step >= a.tape.steps;
    int err = tape_over ? 1 : 0;
    int c = 0;
    double logl = 0.0, q = 0.0;
    uint32_t acc_word = 0u;
    bool have_acc_word = false;
    // ---- phase A -------------------------------------------------------------------------------
    if (valid) {
      c = (!KONE && a.assign) ? a.assign[k] : 0;
      const double* mu = KONE ? nullptr : s_mean + c * D;
      auto mean_of = [&](int i) -> double { if (KONE) return c_mode[i]; else return mu[i]; };
      const double sig = s_ctrl[C_BASE + c], dof = KONE ? c_mode[CM_DOF] : s_dof[c];
      const double* urow = a.u + k * D;
      logl = a.logl[k];
      double cm = sig, keep = 0.0;
      if (TPCN) {
        if (step == 0) {     // first step of this mutation: q (and the Student-t term B) of the freshly resampled state
          double dq[D];
#pragma unroll
          for (int i = 0; i < D; ++i) dq[i] = urow[i] - mean_of(i);
          q = quad(dq, c);
          a.qcur[k] = q;
          a.qcur[a.n + k] = (-0.5 * ((double)D + dof)) * log(1.0 + q / dof);
        } else q = a.qcur[k];
        double g;
        if (TAPE) g = tape_over ? 1.0 : a.tape.gamma[(int64_t)step * a.n + k];
        else g = gamma_mt(rng, slot, (uint32_t)step, 0.5 * ((double)D + dof), acc_word, have_acc_word);
        const double gscale = 2.0 / (dof + q);
        cm = sig * sqrt(1.0 / (gscale * g));
        keep = sqrt(__dsub_rn(1.0, __dmul_rn(sig, sig)));
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double ui = urow[i];
        s_x[i * 32 + lane] = TPCN ? (mean_of(i) + keep * (ui - mean_of(i))) : ui;
      }
      s_cm[lane] = cm;
      s_mode[lane] = c;
    }
    s_used[lane] = 0;
    __syncwarp();
    // ---- phase B: warp-cooperative redraw --------------------------------------------------------
    {
      int att = att0;                                       // next attempt index of MY walker
      const int n_att_tape = (TAPE && valid && !tape_over) ? a.tape.z_cnt[(int64_t)step * a.n + k] : 0;
      unsigned pending = __ballot_sync(0xffffffffu, valid && !tape_over);
      while (pending) {
        const int np = __popc(pending);
        const int sl = lane % np, off = lane / np;
        const int wsel = __fns(pending, 0, sl + 1);         // lane of the walker I work for
        const int aidx = __shfl_sync(0xffffffffu, att, wsel) + off;
        const int natt_sel = __shfl_sync(0xffffffffu, n_att_tape, wsel);
        const int cs = KONE ? 0 : s_mode[wsel];
        const double* L = s_chol + cs * D * D;
        const double cmul = s_cm[wsel];
        double z[D];
        bool have = true;
        if (TAPE) {
          have = aidx < natt_sel;
          const int64_t gk = __shfl_sync(0xffffffffu, k, wsel);
          const double* zt = a.tape.z + a.tape.z_off[(int64_t)step * a.n + gk] + (int64_t)(have ? aidx : 0) * D;
#pragma unroll
          for (int i = 0; i < D; ++i) z[i] = zt[i];
        } else {
          normals_fixed<D>(rng, __shfl_sync(0xffffffffu, slot, wsel), (uint32_t)step, aidx, z);
        }
        double prop[D];
        bool inside = have;
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] *= cmul;          // L (c z): D multiplications instead of D (D + 1) / 2
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double lz = 0.0;
#pragma unroll
          for (int j = 0; j <= i; ++j) lz += (KONE ? c_mode[CM_CHOL + i * D + j] : L[i * D + j]) * z[j];
          double v = s_x[i * 32 + wsel] + lz;
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
          for (int i = 0; i < D; ++i) s_x[i * 32 + wsel] = prop[i];
          s_used[wsel] = aidx + 1;
        }
        // bookkeeping for MY walker
        bool still = (pending >> lane) & 1u;
        const int my_sl = still ? __popc(pending & ((1u << lane) - 1u)) : 0;
        const unsigned my_grp = __shfl_sync(0xffffffffu, grp, my_sl);     // lane my_sl works for me at offset 0
        if (still) {
          if (okm & my_grp) still = false;
          else {
            att += __popc(my_grp);
            if (TAPE && att >= n_att_tape) { err = 1; still = false; s_used[lane] = -1; }
            else if (att >= kMaxAttempts) { err = 2; still = false; s_used[lane] = -1; }
          }
        }
        pending = __ballot_sync(0xffffffffu, still);
        if (single_round) break;
      }
      deferred = pending;            // non-zero only after a single round
      att_out = att;
    }
    __syncwarp();
    // ---- phase C -------------------------------------------------------------------------------
    double alpha = 0.0;
    if (valid && s_used[lane] > 0) {
      acc.nprop += s_used[lane];
      const double* mu = KONE ? nullptr : s_mean + c * D;
      const double dof = KONE ? c_mode[CM_DOF] : s_dof[c];
      double prop[D], x[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        prop[i] = s_x[i * 32 + lane];
        x[i] = KONE ? __dadd_rn(c_mode[CM_PRIOR + i], __dmul_rn(c_mode[CM_PRIOR + D + i], prop[i]))
                    : __dadd_rn(s_prior[i], __dmul_rn(s_prior[D + i], prop[i]));
      }
      const double logl_new = eval_like(LIKE >= 0 ? LIKE : a.p.like_id, a.p.like_params, D, x);
      double factor = 0.0, q_new = 0.0, A_new = 0.0;
      if (TPCN) {
        double dn[D];
#pragma unroll
        for (int i = 0; i < D; ++i) dn[i] = prop[i] - (KONE ? c_mode[i] : mu[i]);
        q_new = quad(dn, c);
        const double hd = -0.5 * ((double)D + dof);
        A_new = hd * log(1.0 + q_new / dof);
        const double Bq = a.qcur[a.n + k];               // hd * log(1 + q / dof) of the current state, cached
        factor = __dadd_rn(-A_new, Bq);
      }
      double al = exp(__dadd_rn(__dmul_rn(a.p.beta, __dsub_rn(logl_new, logl)), factor));
      al = fmin(1.0, al);
      if (isnan(al)) al = 0.0;
      alpha = al;
      double ur;
      if (TAPE) ur = a.tape.acc_u[(int64_t)step * a.n + k];
      else ur = accept_uniform(rng, slot, (uint32_t)step, acc_word, have_acc_word);
      if (ur < al) {
        acc.accepted += 1;
        double* urow = a.u + k * D;
#pragma unroll
        for (int i = 0; i < D; ++i) urow[i] = prop[i];
        a.logl[k] = logl_new;
        if (TPCN) { a.qcur[k] = q_new; a.qcur[a.n + k] = A_new; }
      }
    }
    acc.err = max(acc.err, err);
    if (KONE) acc.alpha += alpha;
    else {
      for (int m = 0; m < K; ++m) {
        const double v = warp_sum((valid && c == m) ? alpha : 0.0);
        if (lane == 0) warp_alpha[m] += v;
      }
    }
    __syncwarp();      // s_x / s_used are reused by this warp's next tile
  }
};

template <int D, bool TPCN, bool TAPE, bool KONE, int LIKE>
__global__ void __launch_bounds__(32 * mcmc_warps<D>(), 2)
mcmc_run_fast(const StepArgs a) {
  extern __shared__ double dyn_smem[];
  run_steps<FastBody<D, TPCN, TAPE, KONE, LIKE>>(a, dyn_smem);
}

template <int D, bool TPCN, bool TAPE, bool KONE, int LIKE = -1>
int launch_fast_variant(const StepArgs& a, cudaStream_t st) {
  using Body = FastBody<D, TPCN, TAPE, KONE, LIKE>;
  const size_t smem = run_smem_bytes<Body>(a.p);
  if (smem > 200 * 1024) return TB_ERR_UNSUPPORTED;
  if (KONE && !a.p.reserved) {   // device -> constant copies, ordered on the stream before the launch
                                 // (reserved != 0: the caller states they are unchanged since its last call)
    const cudaMemcpyKind kd = cudaMemcpyDeviceToDevice;
    cudaMemcpyToSymbolAsync(c_mode, a.p.mode_mean, sizeof(double) * D, 0, kd, st);
    cudaMemcpyToSymbolAsync(c_mode, a.p.mode_chol, sizeof(double) * D * D, sizeof(double) * D, kd, st);
    cudaMemcpyToSymbolAsync(c_mode, a.p.mode_inv, sizeof(double) * D * D, sizeof(double) * (D + D * D), kd, st);
    cudaMemcpyToSymbolAsync(c_mode, a.p.prior_params, sizeof(double) * 2 * D, sizeof(double) * (D + 2 * D * D), kd, st);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_mode, a.p.mode_dof, sizeof(double), sizeof(double) * (D + 2 * D * D + 2 * D),
                                            kd, st);
    if (e != cudaSuccess) return (int)e;
  }
  return launch_persistent(mcmc_run_fast<D, TPCN, TAPE, KONE, LIKE>, Body::kWarps, a, smem, st);
}

template <int D, bool TPCN, bool TAPE>
int launch_fast_like(const StepArgs& a, cudaStream_t st) {
  const bool kone = g_allow_kone && a.p.n_modes == 1 && a.assign == nullptr;
  if (!kone) return launch_fast_variant<D, TPCN, TAPE, false>(a, st);
  switch (a.p.like_id) {   // single-mode path: likelihood known at compile time
    case TB_LIKE_ROSENBROCK: return launch_fast_variant<D, TPCN, TAPE, true, TB_LIKE_ROSENBROCK>(a, st);
    case TB_LIKE_GAUSSIAN: return launch_fast_variant<D, TPCN, TAPE, true, TB_LIKE_GAUSSIAN>(a, st);
    case TB_LIKE_ISO_MIXTURE: return launch_fast_variant<D, TPCN, TAPE, true, TB_LIKE_ISO_MIXTURE>(a, st);
    case TB_LIKE_TWIN_SHELLS: return launch_fast_variant<D, TPCN, TAPE, true, TB_LIKE_TWIN_SHELLS>(a, st);
    default: return launch_fast_variant<D, TPCN, TAPE, true>(a, st);
  }
}

template <int D>
int launch_fast(const StepArgs& a, cudaStream_t st) {
  const bool tpcn = a.p.sampler == TB_SAMPLE_TPCN, tape = a.p.rng_mode == TB_RNG_TAPE;
  if (tpcn) return tape ? launch_fast_like<D, true, true>(a, st) : launch_fast_like<D, true, false>(a, st);
  // random-walk Metropolis: single-mode constant-memory variant with the runtime likelihood switch
  const bool kone = g_allow_kone && a.p.n_modes == 1 && a.assign == nullptr;
  if (tape) return kone ? launch_fast_variant<D, false, true, true>(a, st) : launch_fast_variant<D, false, true, false>(a, st);
  return kone ? launch_fast_variant<D, false, false, true>(a, st) : launch_fast_variant<D, false, false, false>(a, st);
}

}  // namespace tb
