// Runtime-dimension Metropolis step for n_dim without a compile-time instantiation (17 .. 128), with the
// warp-cooperative redraw of the fast path: the lanes of a one-warp CTA are dealt (walker, attempt) pairs over
// the warp's unfinished walkers, so the O(100) redraws per walker that high-dimensional early iterations need
// (mcmc.py:239-249) do not serialise behind the slowest lane.  The per-lane vectors (normals, proposal, centred
// proposal) live in shared-memory columns [d][32] instead of local memory; mode statistics are read through L1
// as warp-uniform broadcasts.  Arithmetic order in tape mode follows mcmc_step_kernel (tb_mcmc.cu).
//   ref: tempest/mcmc.py:142-323
#include "tb_mcmc_fast.cuh"

namespace tb {

namespace {

struct SmemCol {                     // column `lane` of a [d][32] shared array, indexable like a vector
  const double* p;
  __device__ __forceinline__ double operator[](int i) const { return p[i * 32]; }
};

template <bool TPCN, bool TAPE>
struct WideBody {
  static constexpr bool kSingleMode = false;
  static constexpr bool kDeferred = false;
  static constexpr int kWarps = 1;
  __host__ __device__ static size_t cta_doubles(const tb_mcmc_params&) { return 0; }
  // per-warp: proposal centre / winner [d][32], lane-private normals [d][32], lane-private proposal [d][32],
  // proposal scale [32], attempts used + mode [32 + 32 ints]
  __host__ __device__ static size_t warp_doubles(const tb_mcmc_params& p) { return 3 * 32 * (size_t)p.n_dim + 32 + 32; }
  __device__ static void stage(const StepArgs&, double*) {}

  __device__ static __forceinline__ void tile(const StepArgs& a, double*, double* s_warp, const double* s_ctrl,
                                              int64_t tile, int step, TileAcc& acc, double* warp_alpha) {
  const int d = a.p.n_dim, K = a.p.n_modes;
  double* s_x = s_warp;              // [d][32] proposal centre, later the winning proposal
  double* s_z = s_x + d * 32;        // [d][32] lane-private normals; prior-transformed point in phase C
  double* s_p = s_z + d * 32;        // [d][32] lane-private proposal; centred proposal in phase C
  double* s_cm = s_p + d * 32;
  int* s_used = reinterpret_cast<int*>(s_cm + 32);
  int* s_mode = s_used + 32;
  const int lane = threadIdx.x & 31;
  const int64_t k = tile * 32 + lane;
  const bool valid = k < a.n;
  const Philox rng(a.p.seed, a.p.iteration);
  const uint64_t slot0 = (uint64_t)(a.p.slot_offset + tile * 32);
  const bool tape_over = TAPE && step >= a.tape.steps;
  const int ncall = (d + 3) / 4;
  int err = tape_over ? 1 : 0;
  int c = 0;
  double logl = 0.0, q = 0.0;
  uint32_t acc_word = 0u;
  bool have_acc_word = false;
  // ---- phase A: one walker per lane ----------------------------------------------------------
  if (valid) {
    c = a.assign ? a.assign[k] : 0;
    const double* mu = a.p.mode_mean + (size_t)c * d;
    const double sig = s_ctrl[C_BASE + c], dof = __ldg(a.p.mode_dof + c);
    const double* urow = a.u + k * d;
    logl = a.logl[k];
    double cm = sig, keep = 0.0;
    if (TPCN) {
      q = a.qcur[k];                                   // tb_mcmc_begin computed it for the starting state
      double g;
      if (TAPE) g = tape_over ? 1.0 : a.tape.gamma[(int64_t)step * a.n + k];
      else g = gamma_mt(rng, slot0 + lane, (uint32_t)step, 0.5 * ((double)d + dof), acc_word, have_acc_word);
      const double gscale = 2.0 / (dof + q);
      cm = sig * sqrt(1.0 / (gscale * g));
      keep = sqrt(__dsub_rn(1.0, __dmul_rn(sig, sig)));
    }
    for (int i = 0; i < d; ++i) {
      const double ui = urow[i];
      const double m = __ldg(mu + i);
      s_x[i * 32 + lane] = TPCN ? (m + keep * (ui - m)) : ui;
    }
    s_cm[lane] = cm;
    s_mode[lane] = c;
  }
  s_used[lane] = 0;
  __syncwarp();
  // ---- phase B: warp-cooperative redraw (see mcmc_step_fast) --------------------------------------
  {
    int att = 0;
    const int n_att_tape = (TAPE && valid && !tape_over) ? a.tape.z_cnt[(int64_t)step * a.n + k] : 0;
    unsigned pending = __ballot_sync(0xffffffffu, valid && !tape_over);
    while (pending) {
      const int np = __popc(pending);
      const int sl = lane % np, off = lane / np;
      const int wsel = __fns(pending, 0, sl + 1);
      const int aidx = __shfl_sync(0xffffffffu, att, wsel) + off;
      const int natt_sel = __shfl_sync(0xffffffffu, n_att_tape, wsel);
      const int cs = s_mode[wsel];
      const double* L = a.p.mode_chol + (size_t)cs * d * d;
      const double cmul = s_cm[wsel];
      bool have = true;
      if (TAPE) {
        have = aidx < natt_sel;
        const int64_t gk = tile * 32 + wsel;
        const double* zt = a.tape.z + a.tape.z_off[(int64_t)step * a.n + gk] + (int64_t)(have ? aidx : 0) * d;
        for (int i = 0; i < d; ++i) s_z[i * 32 + lane] = zt[i];
      } else {
        const uint64_t slot = slot0 + (uint64_t)wsel;
        for (int cidx = 0; cidx < ncall; ++cidx) {
          const uint4 r = rng.block((uint32_t)slot, (uint32_t)(slot >> 32), (uint32_t)step,
                                    (RNG_NORMAL << 24) | (uint32_t)((aidx * ncall + cidx) & 0xffffff));
          double n0, n1, n2, n3;
          bm_pair32(r.x, r.y, n0, n1);
          bm_pair32(r.z, r.w, n2, n3);
          const int b = 4 * cidx;
          if (b + 0 < d) s_z[(b + 0) * 32 + lane] = n0;
          if (b + 1 < d) s_z[(b + 1) * 32 + lane] = n1;
          if (b + 2 < d) s_z[(b + 2) * 32 + lane] = n2;
          if (b + 3 < d) s_z[(b + 3) * 32 + lane] = n3;
        }
      }
      bool inside = have;
      for (int j = 0; j < d; ++j) s_z[j * 32 + lane] *= cmul;     // L (c z), as the compile-time-dimension kernel
      for (int i = 0; i < d; ++i) {
        double lz = 0.0;
        const double* Li = L + (size_t)i * d;
        for (int j = 0; j <= i; ++j) lz += __ldg(Li + j) * s_z[j * 32 + lane];
        double v = s_x[i * 32 + wsel] + lz;
        const int kind = a.p.bc_kind ? a.p.bc_kind[i] : 0;
        v = bc_apply(v, kind);
        if (kind == 0 && !(v >= 0.0 && v <= 1.0)) inside = false;
        s_p[i * 32 + lane] = v;
      }
      const unsigned okm = __ballot_sync(0xffffffffu, inside);
      const unsigned grp = __match_any_sync(0xffffffffu, wsel);
      const unsigned win = okm & grp;
      __syncwarp();
      if (win && lane == __ffs(win) - 1) {                // lowest attempt inside the cube wins
        for (int i = 0; i < d; ++i) s_x[i * 32 + wsel] = s_p[i * 32 + lane];
        s_used[wsel] = aidx + 1;
      }
      bool still = (pending >> lane) & 1u;
      const int my_sl = still ? __popc(pending & ((1u << lane) - 1u)) : 0;
      const unsigned my_grp = __shfl_sync(0xffffffffu, grp, my_sl);
      if (still) {
        if (okm & my_grp) still = false;
        else {
          att += __popc(my_grp);
          if (TAPE && att >= n_att_tape) { err = 1; still = false; s_used[lane] = -1; }
          else if (att >= kMaxAttempts) { err = 2; still = false; s_used[lane] = -1; }
        }
      }
      pending = __ballot_sync(0xffffffffu, still);
    }
  }
  __syncwarp();
  // ---- phase C: likelihood, Student-t ratio, accept -----------------------------------------------
  double alpha = 0.0;
  if (valid && s_used[lane] > 0) {
    acc.nprop += s_used[lane];
    const double* mu = a.p.mode_mean + (size_t)c * d;
    const double* IV = a.p.mode_inv + (size_t)c * d * d;
    const double dof = __ldg(a.p.mode_dof + c);
    for (int i = 0; i < d; ++i) {
      const double pi = s_x[i * 32 + lane];
      s_z[i * 32 + lane] = prior_affine(a.p.prior_params, d, i, pi);
      s_p[i * 32 + lane] = pi - __ldg(mu + i);
    }
    SmemCol xcol; xcol.p = s_z + lane;
    const double logl_new = eval_like(a.p.like_id, a.p.like_params, d, xcol);
    double factor = 0.0, q_new = 0.0;
    if (TPCN) {   // Student-t density ratio (mcmc.py:251-279), same order as mcmc_step_kernel
      for (int j = 0; j < d; ++j) {
        double y = 0.0;
        for (int i = 0; i < d; ++i) y += s_p[i * 32 + lane] * __ldg(IV + (size_t)i * d + j);
        q_new += y * s_p[j * 32 + lane];
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
    if (TAPE) ur = a.tape.acc_u[(int64_t)step * a.n + k];
    else ur = accept_uniform(rng, slot0 + lane, (uint32_t)step, acc_word, have_acc_word);
    if (ur < al) {
      acc.accepted += 1;
      double* urow = a.u + k * d;
      for (int i = 0; i < d; ++i) urow[i] = s_x[i * 32 + lane];
      a.logl[k] = logl_new;
      if (TPCN) a.qcur[k] = q_new;
    }
  }
  acc.err = max(acc.err, err);
  for (int m = 0; m < K; ++m) {
    const double v = warp_sum((valid && c == m) ? alpha : 0.0);
    if (lane == 0) warp_alpha[m] += v;
  }
  __syncwarp();
  }
};

template <bool TPCN, bool TAPE>
__global__ void __launch_bounds__(32)
mcmc_run_wide(const StepArgs a) {
  extern __shared__ double dyn_smem[];
  run_steps<WideBody<TPCN, TAPE>>(a, dyn_smem);
}

template <bool TPCN, bool TAPE>
int launch_wide_variant(const StepArgs& a, cudaStream_t st) {
  using Body = WideBody<TPCN, TAPE>;
  const size_t smem = run_smem_bytes<Body>(a.p);
  if (smem > 220 * 1024) return TB_ERR_UNSUPPORTED;
  return launch_persistent(mcmc_run_wide<TPCN, TAPE>, Body::kWarps, a, smem, st);
}

}  // namespace

int launch_wide(const StepArgs& a, cudaStream_t st) {
  const bool tpcn = a.p.sampler == TB_SAMPLE_TPCN, tape = a.p.rng_mode == TB_RNG_TAPE;
  if (tpcn) return tape ? launch_wide_variant<true, true>(a, st) : launch_wide_variant<true, false>(a, st);
  return tape ? launch_wide_variant<false, true>(a, st) : launch_wide_variant<false, false>(a, st);
}

}  // namespace tb
