// Compile-time-dimension Metropolis step (the hot kernel of the mutation stage).  Included by the
// tb_mcmc_fast_*.cu translation units, one or two dimensions each, so they compile in parallel.
#pragma once
#include "tb_mcmc_shared.cuh"

namespace tb {

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
// (measured at D = 10: 28 one-warp CTAs / 72 registers + a compile-time likelihood: 0.120 -> 0.110 ms per step)
template <int D>
constexpr int mcmc_min_ctas() { return (D <= 10 ? 7 : (D <= 12 ? 4 : 3)) * (128 / kFastBlock); }

// Single-mode production runs (K = 1, the clustering=False headline path) read the mode statistics and the
// prior box from constant memory: the operands fold into the DFMAs, so the ~3 D^2/2 shared-memory loads per
// walker-step, the per-CTA staging loop and its barrier disappear.  Layout: mean[D], chol[D*D], inv[D*D]
// (row-major, as stored), prior lo[D], scale[D], dof.  One copy per translation unit (static).
constexpr int kConstDoubles = 16 + 2 * 256 + 32 + 1;
static __constant__ double c_mode[kConstDoubles];

// LIKE >= 0: the registry likelihood is fixed at compile time (single-mode production variants), so the
// other three likelihood bodies are not in the instruction stream; LIKE = -1 switches on a.p.like_id.
template <int D, bool TPCN, bool TAPE, bool KONE, int LIKE>
__global__ void __launch_bounds__(kFastBlock, mcmc_min_ctas<D>())
mcmc_step_fast(StepArgs a) {
  if (a.ctrl[C_DONE] != 0.0) return;
  constexpr int CM_CHOL = D, CM_INV = D + D * D, CM_PRIOR = D + 2 * D * D, CM_DOF = D + 2 * D * D + 2 * D;
  constexpr int B = kFastBlock, NW = B / 32;         // one-warp CTAs by default: no CTA-wide barrier anywhere
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
  double* s_part = s_sig + K;                        // [NW][K] per-warp alpha sums
  double* s_prior = s_part + NW * K;                 // [2*D] lo, scale
  if (!KONE) {
    for (int e = threadIdx.x; e < K * D; e += B) s_mean[e] = __ldg(a.p.mode_mean + e);
    for (int e = threadIdx.x; e < K * D * D; e += B) {
      s_chol[e] = __ldg(a.p.mode_chol + e);
      const int i = (e % (D * D)) / D, j = e % D;
      const double v = __ldg(a.p.mode_inv + e);
      s_inv[e] = (j > i) ? 2.0 * v : v;             // q = sum_i d_i * sum_{j>=i} S_ij d_j
    }
    for (int e = threadIdx.x; e < K; e += B) { s_dof[e] = __ldg(a.p.mode_dof + e); s_sig[e] = a.ctrl[C_BASE + e]; }
    for (int e = threadIdx.x; e < 2 * D; e += B) s_prior[e] = __ldg(a.p.prior_params + e);
    if (NW > 1) __syncthreads(); else __syncwarp();
  }
  // quadratic form (u - mu)^T Sigma^-1 (u - mu) of mode `cm` from centred coordinates
  auto quad = [&](const double (&dv)[D], int cm) -> double {
    double acc = 0.0;
    if (KONE) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double y = 0.0;
#pragma unroll
        for (int j = i + 1; j < D; ++j) y += c_mode[CM_INV + i * D + j] * dv[j];
        acc += dv[i] * (c_mode[CM_INV + i * D + i] * dv[i] + 2.0 * y);
      }
    } else {
      const double* IV = s_inv + cm * D * D;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double y = 0.0;
#pragma unroll
        for (int j = i; j < D; ++j) y += IV[i * D + j] * dv[j];
        acc += dv[i] * y;
      }
    }
    return acc;
  };

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
    c = (!KONE && a.assign) ? a.assign[k] : 0;
    const double* mu = KONE ? nullptr : s_mean + c * D;
    auto mean_of = [&](int i) -> double { if (KONE) return c_mode[i]; else return mu[i]; };
    const double sig = KONE ? a.ctrl[C_BASE] : s_sig[c], dof = KONE ? c_mode[CM_DOF] : s_dof[c];
    const double* urow = a.u + k * D;
    logl = a.logl[k];
    double cm = sig, keep = 0.0;
    if (tpcn) {
      if (step == 0) {     // first step of this mutation: q (and the Student-t term B) of the freshly resampled state
        double dq[D];
#pragma unroll
        for (int i = 0; i < D; ++i) dq[i] = urow[i] - mean_of(i);
        q = quad(dq, c);
        a.qcur[k] = q;
        a.qcur[a.n + k] = (-0.5 * ((double)D + dof)) * log(1.0 + q / dof);
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
          // squeeze first; the exact test only decides ~8 % of the trials, in fp32 (the variate itself is fp64)
          if (uu < 1.0 - 0.0331 * x2 * x2 ||
              __logf((float)uu) < (float)(0.5 * x2 + dd * (1.0 - v)) + (float)dd * __logf((float)v)) {
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
      s_x[i][t] = tpcn ? (mean_of(i) + keep * (ui - mean_of(i))) : ui;
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
      const int cs = KONE ? 0 : s_mode[wl];
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
        for (int j = 0; j <= i; ++j)
          lz += TAPE ? (cmul * L[i * D + j]) * z[j] : (KONE ? c_mode[CM_CHOL + i * D + j] : L[i * D + j]) * z[j];
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
    const double* mu = KONE ? nullptr : s_mean + c * D;
    const double dof = KONE ? c_mode[CM_DOF] : s_dof[c];
    double prop[D], x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      prop[i] = s_x[i][t];
      x[i] = KONE ? __dadd_rn(c_mode[CM_PRIOR + i], __dmul_rn(c_mode[CM_PRIOR + D + i], prop[i]))
                  : __dadd_rn(s_prior[i], __dmul_rn(s_prior[D + i], prop[i]));
    }
    const double logl_new = eval_like(LIKE >= 0 ? LIKE : a.p.like_id, a.p.like_params, D, x);
    double factor = 0.0, q_new = 0.0, A_new = 0.0;
    if (tpcn) {
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
      if (tpcn) { a.qcur[k] = q_new; a.qcur[a.n + k] = A_new; }
    }
  }
  // ---- CTA partials (fixed order) ---------------------------------------------------------------
  const int wid = t >> 5;
  for (int m = 0; m < K; ++m) {
    const double v = warp_sum((valid && c == m) ? alpha : 0.0);
    if (lane == 0) s_part[wid * K + m] = v;
  }
  __shared__ double s_tot[NW][3];
  __shared__ double s_fold[kMaxModes + 3];
  {
    const double na = warp_sum((double)accepted), npr = warp_sum((double)nprop), ne = warp_max((double)err);
    if (lane == 0) { s_tot[wid][0] = na; s_tot[wid][1] = npr; s_tot[wid][2] = ne; }
  }
  if (NW > 1) __syncthreads(); else __syncwarp();
  const int W = K + 3;
  double* part = fold_cta_partials(a.ws, gridDim.x, W) + (size_t)blockIdx.x * W;
  for (int m = t; m < K; m += B) {
    double v = s_part[m];
#pragma unroll
    for (int w = 1; w < NW; ++w) v += s_part[w * K + m];
    part[m] = v;
  }
  if (t == 0) {
    double na = s_tot[0][0], npr = s_tot[0][1], ne = s_tot[0][2];
#pragma unroll
    for (int w = 1; w < NW; ++w) { na += s_tot[w][0]; npr += s_tot[w][1]; ne = fmax(ne, s_tot[w][2]); }
    part[K] = na; part[K + 1] = npr; part[K + 2] = ne;
  }
  if (NW > 1) __syncthreads();
  if (wid == 0) arrive_and_fold(a, K, s_fold);
}


template <int D, bool TPCN, bool TAPE, bool KONE, int LIKE = -1>
int launch_fast_variant(const StepArgs& a, int count, cudaStream_t st) {
  const int K = a.p.n_modes;
  const size_t smem = sizeof(double) * ((size_t)K * D + 2 * (size_t)K * D * D + 2 * K + (kFastBlock / 32) * K + 2 * D);
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mcmc_step_fast<D, TPCN, TAPE, KONE, LIKE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  if (KONE && !a.p.reserved) {   // device -> constant copies, ordered on the stream before the step launches
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
  const int grid = (int)((a.n + kFastBlock - 1) / kFastBlock);
  for (int s = 0; s < count; ++s) mcmc_step_fast<D, TPCN, TAPE, KONE, LIKE><<<grid, kFastBlock, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TB_OK : (int)e;
}

template <int D>
int launch_fast(const StepArgs& a, int count, cudaStream_t st) {
  const bool tpcn = a.p.sampler == TB_SAMPLE_TPCN, tape = a.p.rng_mode == TB_RNG_TAPE;
  const bool kone = !tape && a.p.n_modes == 1 && a.assign == nullptr;
  if (tpcn) {
    if (tape) return launch_fast_variant<D, true, true, false>(a, count, st);
    if (!kone) return launch_fast_variant<D, true, false, false>(a, count, st);
    switch (a.p.like_id) {   // single-mode tpCN production path: likelihood known at compile time
      case TB_LIKE_ROSENBROCK: return launch_fast_variant<D, true, false, true, TB_LIKE_ROSENBROCK>(a, count, st);
      case TB_LIKE_GAUSSIAN: return launch_fast_variant<D, true, false, true, TB_LIKE_GAUSSIAN>(a, count, st);
      case TB_LIKE_ISO_MIXTURE: return launch_fast_variant<D, true, false, true, TB_LIKE_ISO_MIXTURE>(a, count, st);
      case TB_LIKE_TWIN_SHELLS: return launch_fast_variant<D, true, false, true, TB_LIKE_TWIN_SHELLS>(a, count, st);
      default: return launch_fast_variant<D, true, false, true>(a, count, st);
    }
  }
  if (tape) return launch_fast_variant<D, false, true, false>(a, count, st);
  return kone ? launch_fast_variant<D, false, false, true>(a, count, st)
              : launch_fast_variant<D, false, false, false>(a, count, st);
}


}  // namespace tb
