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
__global__ void __launch_bounds__(32)
mcmc_step_wide(StepArgs a) {
  if (a.ctrl[C_DONE] != 0.0) return;
  extern __shared__ double sm[];
  const int d = a.p.n_dim, K = a.p.n_modes;
  double* s_x = sm;                  // [d][32] proposal centre, later the winning proposal
  double* s_z = s_x + d * 32;        // [d][32] lane-private normals; prior-transformed point in phase C
  double* s_p = s_z + d * 32;        // [d][32] lane-private proposal; centred proposal in phase C
  __shared__ double s_cm[32];
  __shared__ int s_used[32];
  __shared__ int s_mode[32];
  __shared__ double s_fold[kMaxModes + 3];
  const int lane = threadIdx.x;
  const int64_t k = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = k < a.n;
  const int step = (int)a.ctrl[C_STEPS];
  const Philox rng(a.p.seed, a.p.iteration);
  const uint64_t slot0 = (uint64_t)(a.p.slot_offset + (int64_t)blockIdx.x * 32);
  const bool tape_over = TAPE && step >= a.tape.steps;
  const int ncall = (d + 3) / 4;
  int err = tape_over ? 1 : 0;
  int c = 0;
  double logl = 0.0, q = 0.0;
  // ---- phase A: one walker per lane ----------------------------------------------------------
  if (valid) {
    c = a.assign ? a.assign[k] : 0;
    const double* mu = a.p.mode_mean + (size_t)c * d;
    const double sig = a.ctrl[C_BASE + c], dof = __ldg(a.p.mode_dof + c);
    const double* urow = a.u + k * d;
    logl = a.logl[k];
    double cm = sig, keep = 0.0;
    if (TPCN) {
      q = a.qcur[k];                                   // tb_mcmc_begin computed it for the starting state
      double g;
      if (TAPE) g = tape_over ? 1.0 : a.tape.gamma[(int64_t)step * a.n + k];
      else {   // Marsaglia-Tsang, shape = (d + nu)/2 >= 1
        const double shape = 0.5 * ((double)d + dof);
        const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        g = dd;
        for (uint32_t trial = 0; trial < 64; ++trial) {
          const uint4 r = rng.block((uint32_t)(slot0 + lane), (uint32_t)((slot0 + lane) >> 32), (uint32_t)step,
                                    (RNG_GAMMA << 24) | trial);
          double n0, n1;
          bm_pair32(r.x, r.y, n0, n1);
          const double v1 = 1.0 + cc * n0;
          if (v1 <= 0.0) continue;
          const double v = v1 * v1 * v1;
          const double uu = ((double)r.z + 0.5) * 2.3283064365386963e-10;
          const double x2 = n0 * n0;
          if (uu < 1.0 - 0.0331 * x2 * x2 || log(uu) < 0.5 * x2 + dd * (1.0 - v + log(v))) { g = dd * v; break; }
        }
      }
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
        const int64_t gk = (int64_t)blockIdx.x * 32 + wsel;
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
      for (int i = 0; i < d; ++i) {
        double lz = 0.0;
        const double* Li = L + (size_t)i * d;
        for (int j = 0; j <= i; ++j) lz += (cmul * __ldg(Li + j)) * s_z[j * 32 + lane];
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
  int accepted = 0, nprop = 0;
  if (valid && s_used[lane] > 0) {
    nprop = s_used[lane];
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
    else {
      const uint4 r = rng.block((uint32_t)(slot0 + lane), (uint32_t)((slot0 + lane) >> 32), (uint32_t)step, RNG_ACCEPT << 24);
      ur = u53(r.x, r.y);
    }
    if (ur < al) {
      accepted = 1;
      double* urow = a.u + k * d;
      for (int i = 0; i < d; ++i) urow[i] = s_x[i * 32 + lane];
      a.logl[k] = logl_new;
      if (TPCN) a.qcur[k] = q_new;
    }
  }
  // ---- CTA partial row and the hierarchical fold ------------------------------------------------
  const int W = K + 3;
  double* part = fold_cta_partials(a.ws, gridDim.x, W) + (size_t)blockIdx.x * W;
  for (int m = 0; m < K; ++m) {
    const double v = warp_sum((valid && c == m) ? alpha : 0.0);
    if (lane == 0) part[m] = v;
  }
  {
    const double na = warp_sum((double)accepted), npr = warp_sum((double)nprop), ne = warp_max((double)err);
    if (lane == 0) { part[K] = na; part[K + 1] = npr; part[K + 2] = ne; }
  }
  arrive_and_fold(a, K, s_fold);
}

template <bool TPCN, bool TAPE>
int launch_wide_variant(const StepArgs& a, int count, cudaStream_t st) {
  const size_t smem = sizeof(double) * 3 * 32 * (size_t)a.p.n_dim;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_step_wide<TPCN, TAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = (int)((a.n + 31) / 32);
  for (int s = 0; s < count; ++s) mcmc_step_wide<TPCN, TAPE><<<grid, 32, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TB_OK : (int)e;
}

}  // namespace

int launch_wide(const StepArgs& a, int count, cudaStream_t st) {
  const bool tpcn = a.p.sampler == TB_SAMPLE_TPCN, tape = a.p.rng_mode == TB_RNG_TAPE;
  if (tpcn) return tape ? launch_wide_variant<true, true>(a, count, st) : launch_wide_variant<true, false>(a, count, st);
  return tape ? launch_wide_variant<false, true>(a, count, st) : launch_wide_variant<false, false>(a, count, st);
}

}  // namespace tb
