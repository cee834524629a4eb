// Persistent-ensemble reweighting kernels (a1-a5).
//   ref: tempest/state_manager.py:418-480 (compute_logw_and_logz)
//        tempest/steps/reweight.py:88-297 (probe, ESS bracket, bisection)
//        tempest/tools.py:120-135 (effective_sample_size)
//
// HBM layout: logl[N_total], C[N_total] fp64 SoA; per-generation scalars in tiny device arrays.
// Probe = 16 B / particle streamed once (HBM-bound); the N_total x T log-sum-exp of the
// reference is folded into the cached column C (fp64-pipe bound, done once per generation).
#include "tb_common.cuh"
#include "tb_xgpu.cuh"

namespace {
using namespace tb;

__device__ __forceinline__ double mix_term(double l, double b, double z, double ln) {
  // (l*beta_t - logZ_t) + log n_t, rounded after every operation like numpy's temporaries
  return __dadd_rn(__dsub_rn(__dmul_rn(l, b), z), ln);
}

__global__ void __launch_bounds__(kBlock)
mixture_kernel(const double* __restrict__ logl, double* __restrict__ C, int64_t n_old, int64_t n_all,
               const double* __restrict__ gb, const double* __restrict__ gz,
               const double* __restrict__ gn, int T) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double bT = __ldg(gb + T - 1), zT = __ldg(gz + T - 1), nT = __ldg(gn + T - 1);
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_all; s += stride) {
    const double l = __ldg(logl + s);
    if (s < n_old) {
      C[s] = np_logaddexp(C[s], mix_term(l, bT, zT, nT));
    } else {
      double acc = mix_term(l, __ldg(gb), __ldg(gz), __ldg(gn));
      for (int t = 1; t < T; ++t) acc = np_logaddexp(acc, mix_term(l, __ldg(gb + t), __ldg(gz + t), __ldg(gn + t)));
      C[s] = acc;
    }
  }
}

// ---- probe -------------------------------------------------------------------------
struct ProbeWs {            // workspace layout of probe_kernel (device)
  unsigned int ticket;      // last-block ticket
  unsigned int pad[3];
  double partial[kMaxPartials][4];     // {m, S1, S2, n_nonfinite} per CTA
};
// next_beta_kernel uses a GridSync + rows[2][grid][4] (tb_xgpu.cuh) placed after this struct

// exp(d) for d in (-40, 0]: Cody-Waite reduction by ln2, degree-13 Taylor polynomial on |r| <= ln2/2
// (remainder < 5e-18), scaled by 2^k through the exponent field (k >= -58, no denormals).  <= 1 ulp of
// libdevice's exp on this range at roughly half its instruction count (no special-case handling).
__constant__ double kExpC[18] = {
    6755399441055744.0, 1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    1.6059043836821613e-10, 2.08767569878680989792e-09, 2.50521083854417187751e-08, 2.75573192239858906526e-07,
    2.75573192239858906526e-06, 2.48015873015873015873e-05, 1.98412698412698412698e-04, 1.38888888888888888889e-03,
    8.33333333333333333333e-03, 4.16666666666666666667e-02, 1.66666666666666666667e-01, 0.5, 1.0, -40.0};
// (coefficients live in the constant bank so every DFMA takes its constant as an operand instead of
// materialising a 64-bit immediate with two extra moves)
__device__ __forceinline__ double exp_neg40(double d) {
  const double t = fma(d, kExpC[1], kExpC[0]);                     // 1.5 * 2^52: round-to-nearest-integer trick
  const double kf = t - kExpC[0];
  const int k = __double2loint(t);                                 // low word of t holds the integer
  double r = fma(kf, kExpC[2], d);
  r = fma(kf, kExpC[3], r);
  double p = kExpC[4];                                             // 1/13!
#pragma unroll
  for (int i = 5; i <= 15; ++i) p = fma(p, r, kExpC[i]);
  p = fma(p, r, kExpC[16]);
  p = fma(p, r, kExpC[16]);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// Streams (logl, C) once.  Four particles per iteration share one running-max update (rare), then
// each contributes exp(a - m) unless it is below e^-40 (cannot move sums that contain the 1.0 of the
// maximum).  `chk` turns NaN as soon as any a is non-finite (a * 0).
__device__ __forceinline__ void probe4(Ess3& e, double& chk, double a0, double a1, double a2, double a3) {
  chk = fma(a0, 0.0, chk); chk = fma(a1, 0.0, chk); chk = fma(a2, 0.0, chk); chk = fma(a3, 0.0, chk);
  const double mx = fmax(fmax(a0, a1), fmax(a2, a3));
  if (mx > e.m) {
    const double r = (e.m == -INFINITY) ? 0.0 : exp(e.m - mx);
    e.s1 *= r;
    e.s2 *= r * r;
    e.m = mx;
  }
  const double d0 = a0 - e.m, d1 = a1 - e.m, d2 = a2 - e.m, d3 = a3 - e.m;
  if (d0 > -40.0) { const double w = exp_neg40(d0); e.s1 += w; e.s2 = fma(w, w, e.s2); }
  if (d1 > -40.0) { const double w = exp_neg40(d1); e.s1 += w; e.s2 = fma(w, w, e.s2); }
  if (d2 > -40.0) { const double w = exp_neg40(d2); e.s1 += w; e.s2 = fma(w, w, e.s2); }
  if (d3 > -40.0) { const double w = exp_neg40(d3); e.s1 += w; e.s2 = fma(w, w, e.s2); }
}

__device__ __forceinline__ void probe1(Ess3& e, double& chk, double a) {
  chk = fma(a, 0.0, chk);
  if (a > e.m) {
    const double r = (e.m == -INFINITY) ? 0.0 : exp(e.m - a);
    e.s1 *= r;
    e.s2 *= r * r;
    e.m = a;
  }
  const double d = a - e.m;
  if (d > -40.0) { const double w = exp_neg40(d); e.s1 += w; e.s2 = fma(w, w, e.s2); }
}

// NB probes (betas) share ONE pass over (logl, C): the bytes are read once, every beta keeps its own accumulator
// and sees the particles in the same order as a single-beta pass, so each triple is bitwise what NB = 1 gives.
template <int NB>
__device__ __forceinline__ void probe_slice(const double* __restrict__ logl, const double* __restrict__ C,
                                            int64_t n, const double (&beta)[NB], Ess3 (&e)[NB], double& bad,
                                            bool reverse = false) {
  // vectorised 2 x fp64 loads, 4 independent 16-byte loads in flight per thread
  const int64_t n2 = n >> 1;
  const double2* l2 = reinterpret_cast<const double2*>(logl);
  const double2* c2 = reinterpret_cast<const double2*>(C);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double chk = 0.0;
  auto four = [&](const double2& la, const double2& ca, const double2& lb, const double2& cb) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      probe4(e[j], chk, __dsub_rn(__dmul_rn(la.x, beta[j]), ca.x), __dsub_rn(__dmul_rn(la.y, beta[j]), ca.y),
             __dsub_rn(__dmul_rn(lb.x, beta[j]), cb.x), __dsub_rn(__dmul_rn(lb.y, beta[j]), cb.y));
  };
  auto two = [&](const double2& la, const double2& ca) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      probe1(e[j], chk, __dsub_rn(__dmul_rn(la.x, beta[j]), ca.x));
      probe1(e[j], chk, __dsub_rn(__dmul_rn(la.y, beta[j]), ca.y));
    }
  };
  if (!reverse) {
    for (; i + stride < n2; i += 2 * stride) {
      const double2 la = __ldg(l2 + i), ca = __ldg(c2 + i);
      const double2 lb = __ldg(l2 + i + stride), cb = __ldg(c2 + i + stride);
      four(la, ca, lb, cb);
    }
    for (; i < n2; i += stride) {
      const double2 la = __ldg(l2 + i), ca = __ldg(c2 + i);
      two(la, ca);
    }
  } else if (i < n2) {
    // same elements, last to first: the tail of the previous (forward) pass is still in L2
    int64_t cnt = (n2 - 1 - i) / stride + 1;
    int64_t j = i + (cnt - 1) * stride;
    for (; cnt >= 2; cnt -= 2, j -= 2 * stride) {
      const double2 la = __ldg(l2 + j), ca = __ldg(c2 + j);
      const double2 lb = __ldg(l2 + j - stride), cb = __ldg(c2 + j - stride);
      four(la, ca, lb, cb);
    }
    if (cnt == 1) {
      const double2 la = __ldg(l2 + j), ca = __ldg(c2 + j);
      two(la, ca);
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NB; ++j) probe1(e[j], chk, __dsub_rn(__dmul_rn(logl[n - 1], beta[j]), C[n - 1]));
  }
  if (chk != chk) bad += 1.0;      // some a_s was NaN or +-inf (reported as a non-zero count)
}
__device__ __forceinline__ void probe_slice(const double* __restrict__ logl, const double* __restrict__ C,
                                            int64_t n, double beta, Ess3& e, double& bad, bool reverse = false) {
  const double b1[1] = {beta};
  Ess3 e1[1];
  e1[0] = e;
  probe_slice<1>(logl, C, n, b1, e1, bad, reverse);
  e = e1[0];
}

__device__ __forceinline__ void write_probe_result(double* out, const Ess3& e, double bad) {
  out[0] = e.m; out[1] = e.s1; out[2] = e.s2;
  out[3] = (e.s1 * e.s1) / e.s2;      // ESS = 1 / sum (w/S1)^2
  out[4] = e.m + log(e.s1);           // logZ(beta) = LSE_s(a_s)   (log N_total cancels, :475)
  out[5] = bad;
}

// Row fold of grid_xreduce for (m, S1, S2, n_nonfinite) rows: Ess3 merges in a fixed order.
struct EssFold {
  // rows are [nb][4] = nb x (m, S1, S2, n_nonfinite).  All threads of the CTA; thread t merges rows t, t + B, ... in
  // ascending order, then the fixed-order CTA merge -- per triple exactly what the single-triple fold does.
  int nb = 1;
  __device__ void rows(const double* rows, int nrows, int W, double* tot) const {
    __shared__ double s_m[3 * 32 + 8];
    __shared__ double s_b[40];
    for (int j = 0; j < nb; ++j) {
      Ess3 e; e.init();
      double bad = 0.0;
      for (int b = threadIdx.x; b < nrows; b += blockDim.x) {
        const double* r = rows + (size_t)b * W + 4 * j;
        e.merge(__ldcg(r), __ldcg(r + 1), __ldcg(r + 2));
        bad += __ldcg(r + 3);
      }
      block_merge_ess3(e, s_m);
      bad = block_sum(bad, s_b);
      if (threadIdx.x == 0) { tot[4 * j] = e.m; tot[4 * j + 1] = e.s1; tot[4 * j + 2] = e.s2; tot[4 * j + 3] = bad; }
      __syncthreads();
    }
  }
  __device__ void ranks(const double* slots, int world, int W, double* tot) const {
    const int lane = threadIdx.x & 31;
    if (lane < nb) {
      Ess3 g; g.init();
      double gb = 0.0;
      for (int r = 0; r < world; ++r) {
        const double* slot = slots + (size_t)r * kXSlotDoubles;
        (void)ld_acquire_sys(reinterpret_cast<const unsigned long long*>(slot));
        g.merge(ld_relaxed_sys(slot + 1 + 4 * lane), ld_relaxed_sys(slot + 2 + 4 * lane), ld_relaxed_sys(slot + 3 + 4 * lane));
        gb += ld_relaxed_sys(slot + 4 + 4 * lane);
      }
      tot[4 * lane] = g.m; tot[4 * lane + 1] = g.s1; tot[4 * lane + 2] = g.s2; tot[4 * lane + 3] = gb;
    }
  }
};

__global__ void __launch_bounds__(kBlock, 5)
probe_kernel(const double* __restrict__ logl, const double* __restrict__ C, int64_t n, double beta,
             ProbeWs* ws, double* __restrict__ out) {
  __shared__ double smem[160];
  Ess3 e; e.init();
  double bad = 0.0;
  probe_slice(logl, C, n, beta, e, bad);
  block_merge_ess3(e, smem);
  bad = block_sum(bad, smem + 100);
  if (threadIdx.x == 0) {
    double* p = ws->partial[blockIdx.x];
    p[0] = e.m; p[1] = e.s1; p[2] = e.s2; p[3] = bad;
  }
  if (last_block_arrives(&ws->ticket)) {      // the fold of the search kernel: (m, S1, S2) identical bit for bit
    __shared__ double s_tot[4];
    EssFold().rows(&ws->partial[0][0], gridDim.x, 4, s_tot);
    if (threadIdx.x == 0) { e.m = s_tot[0]; e.s1 = s_tot[1]; e.s2 = s_tot[2]; write_probe_result(out, e, s_tot[3]); }
  }
}

__global__ void __launch_bounds__(kBlock)
weights_kernel(const double* __restrict__ logl, const double* __restrict__ C, int64_t n, double beta,
               const double* __restrict__ stats, double* __restrict__ w, int want_log) {
  const double m = stats[0], s1 = stats[1];
  const double lse = stats[4];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
    double a = __dsub_rn(__dmul_rn(__ldg(logl + s), beta), __ldg(C + s));
    w[s] = want_log ? (a - lse) : exp(a - m) / s1;
  }
}

// ---- device-side next-beta search ---------------------------------------------------
constexpr double kBetaTol = 1e-4, kBetaRtol = 1e-8, kEssTol = 0.01, kMetricAtol = 0.5;  // config.py:233-236
constexpr int kMaxBisect = 200;                                                          // reweight.py:121
constexpr double kTiny = 2.2250738585072014e-308;

// State of the reference's search (reweight.py:255-297 bracket, :165-211 bisection), advanced one probe at a time.
struct BetaSearch {
  int phase;                 // 0: probe beta_prev, 1: probe 1.0, 2: bracket, 3: bisect
  double lo, hi, bmin, bmax;
  int nbis, same;
};
// Consume ESS(beta).  Returns true when the search ends at `beta`; otherwise `next` is the beta to probe next.
__device__ __forceinline__ bool search_advance(BetaSearch& s, double beta, double ess, double target, double beta_prev,
                                               double& next) {
  bool done = false;
  next = beta;
  if (s.phase == 0) {                       // reweight.py:264-266
    if (ess <= target) { done = true; s.same = 1; }
    else { s.phase = 1; next = 1.0; }
  } else if (s.phase == 1) {                // :269-271
    if (ess >= target) { done = true; s.same = 1; }
    else s.phase = 2;
  } else if (s.phase == 2) {                // :288-295
    if (ess >= target) s.lo = beta; else s.hi = beta;
  } else {                                  // :165-211
    const double val = isfinite(ess) ? ess : 1e10;
    const bool metric_ok = fabs(val - target) < fmax(kEssTol * fabs(target), kMetricAtol);
    const double scale = fmax(fmax(fabs(s.bmin), fabs(s.bmax)), kTiny);
    const bool beta_ok = (s.bmax - s.bmin) < fmax(kBetaRtol * scale, kBetaTol * scale);
    ++s.nbis;
    if (metric_ok || beta_ok || beta == 1.0 || s.nbis >= kMaxBisect) done = true;
    else if (val < target) s.bmax = beta; else s.bmin = beta;
  }
  if (!done && s.phase == 2) {              // :277-287
    const double mid = (s.hi + s.lo) * 0.5;
    const double scale = fmax(fmax(fabs(s.lo), fabs(s.hi)), kTiny);
    if (s.hi - s.lo <= fmax(kBetaRtol * scale, kBetaTol * scale)) {
      s.phase = 3; s.bmin = beta_prev; s.bmax = s.hi;   // bisect on [beta_prev, beta_high] (:409-414)
    } else next = mid;
  }
  if (!done && s.phase == 3) next = (s.bmax + s.bmin) * 0.5;
  return done;
}

// The whole search in one cooperative launch.  A probe is a full pass over the ensemble, and a bisection needs ~25 of
// them.  flags & 2 selects SPECULATIVE passes: the beta of the next probe is one of two values that are known before the
// current probe's ESS is (the state machine above, fed "above target" / "below target"), so a pass can evaluate three
// betas on the same bytes -- the current one and both possible successors -- and consume TWO steps of the reference's
// search; probe sequence, comparisons and final beta stay the reference's, bit for bit, with half the passes and half
// the grid-wide / cross-GPU folds.  Measured on one B200 (profiles/r02_next_beta_speculative.txt): a three-beta pass
// costs 2.3x a one-beta pass -- the exps make it fp64-pipe bound (~30 DFMA-class operations per beta and particle;
// the e^-40 cut does not help because live and dead particles share warps) -- so the search got 17 % SLOWER, not
// faster.  It is therefore off by default and kept as an option for runs whose passes are latency-bound.
__global__ void __launch_bounds__(kBlock, 4)
next_beta_kernel(const double* __restrict__ logl, const double* __restrict__ C, int64_t n,
                 double beta_prev, double target, int flags, GridSync* gs, double* __restrict__ result,
                 double* __restrict__ plog, int plog_cap, tb_xgpu xg) {
  __shared__ double smem[160];
  __shared__ double s_part[12], s_tot[12];
  __shared__ double sh_beta[3];
  __shared__ int sh_done, sh_nb;
  // search state, replicated bit-identically in thread 0 of every CTA
  BetaSearch st;
  st.phase = (flags & 1) ? 1 : 0;
  st.lo = beta_prev; st.hi = 1.0; st.bmin = beta_prev; st.bmax = 1.0; st.nbis = 0; st.same = 0;
  int nprobe = 0, npass = 0;
  // candidates of the first pass
  auto plan = [&](const BetaSearch& cur, double beta, double (&b)[3]) -> int {
    // b[0] = beta; b[1] / b[2] = the next beta if ESS(beta) turns out high / low (absent: the search would end)
    b[0] = beta; b[1] = beta; b[2] = beta;
    if (!(flags & 2)) return 1;              // one beta per pass (default, see below)
    BetaSearch hi_s = cur, lo_s = cur;
    double nh = beta, nl = beta;
    const bool dh = search_advance(hi_s, beta, 1e300, target, beta_prev, nh);
    const bool dl = search_advance(lo_s, beta, 0.0, target, beta_prev, nl);
    if (dh && dl) return 1;
    b[1] = dh ? nl : nh;
    b[2] = dl ? nh : nl;
    return 3;
  };
  if (threadIdx.x == 0) {
    double b[3];
    sh_nb = plan(st, (st.phase == 0) ? beta_prev : 1.0, b);
    sh_beta[0] = b[0]; sh_beta[1] = b[1]; sh_beta[2] = b[2];
  }
  __syncthreads();
  for (;;) {
    const int nb = sh_nb;
    double beta[3] = {sh_beta[0], sh_beta[1], sh_beta[2]};
    Ess3 e[3];
    e[0].init(); e[1].init(); e[2].init();
    double bad = 0.0;
    const bool reverse = (npass & 1) != 0;      // alternate direction: reuse what the last pass left in L2
    if (nb == 1) {
      const double b1[1] = {beta[0]};
      Ess3 e1[1];
      e1[0].init();
      probe_slice<1>(logl, C, n, b1, e1, bad, reverse);
      e[0] = e1[0];
    } else {
      probe_slice<3>(logl, C, n, beta, e, bad, reverse);
    }
    bad = block_sum(bad, smem + 100);
#pragma unroll
    for (int j = 0; j < 3; ++j) {              // (static indices keep the accumulators in registers; nb is CTA-uniform)
      if (j < nb) {
        block_merge_ess3(e[j], smem);
        if (threadIdx.x == 0) { s_part[4 * j] = e[j].m; s_part[4 * j + 1] = e[j].s1; s_part[4 * j + 2] = e[j].s2; s_part[4 * j + 3] = bad; }
        __syncthreads();
      }
    }
    // one synchronisation point per pass: grid-wide fold, and on >1 GPU the merge of the ranks' triples over
    // NVLink peer memory, fused (tb_xgpu.cuh); every CTA of every rank ends with the same (m, S1, S2) per beta
    EssFold fold;
    fold.nb = nb;
    const int rc = grid_xreduce(gs, xg, npass, 4 * nb, s_part, s_tot, fold);
    ++npass;
    if (threadIdx.x == 0) {
      int done = 0;
      double cur = beta[0], next = beta[0];
      int j = 0;                               // which evaluated beta is being consumed
      for (int level = 0; level < 2 && !done; ++level) {
        Ess3 r;
        r.m = s_tot[4 * j]; r.s1 = s_tot[4 * j + 1]; r.s2 = s_tot[4 * j + 2];
        double rb = s_tot[4 * j + 3];
        if (rc) rb = NAN;                      // a peer did not answer: stop the search, the host raises
        const double ess = (r.s1 * r.s1) / r.s2;
        if (blockIdx.x == 0 && plog != nullptr && nprobe < plog_cap) { plog[2 * nprobe] = cur; plog[2 * nprobe + 1] = ess; }
        ++nprobe;
        done = search_advance(st, cur, ess, target, beta_prev, next) ? 1 : 0;
        if (rb != rb) done = 1;
        if (done) {
          if (blockIdx.x == 0) {
            result[0] = cur; result[1] = r.m; result[2] = r.s1; result[3] = r.s2; result[4] = ess;
            result[5] = r.m + log(r.s1); result[6] = (double)nprobe; result[7] = (double)st.same;
            result[8] = rb; result[9] = (double)npass;
          }
          break;
        }
        if (level == 0) {
          if (nb == 3 && next == beta[1]) j = 1;
          else if (nb == 3 && next == beta[2]) j = 2;
          else break;                          // successor was not evaluated in this pass (cannot happen with nb == 3)
          cur = next;
        }
      }
      if (!done) {
        double b[3];
        sh_nb = plan(st, next, b);
        sh_beta[0] = b[0]; sh_beta[1] = b[1]; sh_beta[2] = b[2];
      }
      sh_done = done;
    }
    __syncthreads();
    if (sh_done) break;
  }
}

}  // namespace

extern "C" {

int tb_version(void) { return 100; }
int tb_sm_count(void) { return tb::sm_count(); }

int tb_mixture_build(const double* logl, double* C, int64_t n_total, const double* gb, const double* gz,
                     const double* gn, int32_t T, tb_stream_t stream) {
  if (n_total < 0 || T <= 0 || (n_total > 0 && (!logl || !C))) return TB_ERR_ARG;
  if (n_total == 0) return TB_OK;
  int grid = stream_grid(n_total, kBlock, 16);
  mixture_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(logl, C, 0, n_total, gb, gz, gn, T);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_mixture_append(const double* logl, double* C, int64_t n_old, int64_t n_new, const double* gb,
                      const double* gz, const double* gn, int32_t T_new, tb_stream_t stream) {
  if (n_old < 0 || n_new < 0 || T_new <= 0) return TB_ERR_ARG;
  if (n_old + n_new == 0) return TB_OK;
  int grid = stream_grid(n_old + n_new, kBlock, 16);
  mixture_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(logl, C, n_old, n_old + n_new, gb, gz, gn, T_new);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

// one allocation serves both kernels: [ProbeWs of probe_kernel | GridSync + rows of next_beta_kernel]
size_t tb_probe_workspace_bytes(void) { return sizeof(ProbeWs) + grid_sync_bytes(kMaxPartials, 12); }
size_t tb_next_beta_workspace_bytes(void) { return tb_probe_workspace_bytes(); }

int tb_probe(const double* logl, const double* C, int64_t n, double beta, void* workspace, double* out6,
             tb_stream_t stream) {
  if (n <= 0 || !logl || !C || !workspace || !out6) return TB_ERR_ARG;
  // 2 particles per thread per load; 4 CTAs of 256 threads per SM keep >= 32 x 16 B loads in flight per SM
  int grid = stream_grid(n, kBlock * 4, 5);
  probe_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(logl, C, n, beta, (ProbeWs*)workspace, out6);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_weights(const double* logl, const double* C, int64_t n, double beta, const double* stats, double* w,
               tb_stream_t stream) {
  if (n <= 0 || !logl || !C || !stats || !w) return TB_ERR_ARG;
  int grid = stream_grid(n, kBlock, 16);
  weights_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(logl, C, n, beta, stats, w, 0);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_log_weights(const double* logl, const double* C, int64_t n, double beta, const double* stats,
                   double* logw, tb_stream_t stream) {
  if (n <= 0 || !logl || !C || !stats || !logw) return TB_ERR_ARG;
  int grid = stream_grid(n, kBlock, 16);
  weights_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(logl, C, n, beta, stats, logw, 1);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_next_beta(const double* logl, const double* C, int64_t n, double beta_prev, double ess_target,
                 int32_t flags, void* workspace, double* result, double* probe_log, int32_t probe_log_cap,
                 tb_stream_t stream) {
  return tb_next_beta_x(logl, C, n, beta_prev, ess_target, flags, workspace, result, probe_log, probe_log_cap,
                        nullptr, stream);
}

size_t tb_xgpu_buffer_bytes(void) { return sizeof(double) * 2 * kXMaxRanks * kXSlotDoubles; }

int tb_next_beta_x(const double* logl, const double* C, int64_t n, double beta_prev, double ess_target,
                   int32_t flags, void* workspace, double* result, double* probe_log, int32_t probe_log_cap,
                   const tb_xgpu* xgpu, tb_stream_t stream) {
  if (n <= 0 || !logl || !C || !workspace || !result) return TB_ERR_ARG;
  tb_xgpu xg;
  if (xgpu) { xg = *xgpu; if (xg.world < 1 || xg.world > kXMaxRanks || xg.seq < 1) return TB_ERR_ARG; }
  else { xg.rank = 0; xg.world = 1; xg.seq = 1; for (int i = 0; i < 8; ++i) xg.peer[i] = nullptr; }
  static int max_coresident = 0;
  if (max_coresident == 0) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, next_beta_kernel, kBlock, 0);
    if (e != cudaSuccess) return (int)e;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) return TB_ERR_UNSUPPORTED;
    max_coresident = per_sm * tb::sm_count();
    if (max_coresident > kMaxPartials) max_coresident = kMaxPartials;
  }
  int64_t need = (n + kBlock * 4 - 1) / (kBlock * 4);
  int grid = (int)(need < max_coresident ? need : max_coresident);
  if (grid < 1) grid = 1;
  GridSync* ws = reinterpret_cast<GridSync*>(reinterpret_cast<char*>(workspace) + sizeof(ProbeWs));
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(GridSync), as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  void* args[] = {(void*)&logl, (void*)&C, (void*)&n, (void*)&beta_prev, (void*)&ess_target, (void*)&flags,
                  (void*)&ws, (void*)&result, (void*)&probe_log, (void*)&probe_log_cap, (void*)&xg};
  e = cudaLaunchCooperativeKernel((void*)next_beta_kernel, dim3(grid), dim3(kBlock), args, 0, as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  return TB_OK;
}

}  // extern "C"
