// In-kernel priors, benchmark likelihoods and the counter-based RNG.
// Operation order mirrors tempest_b200/registry.py exactly (explicit round-to-nearest
// intrinsics, no FMA contraction) so log-likelihoods are bit-identical to the numpy forms
// wherever only + - * / sqrt are involved.
#pragma once
#include "tb_common.cuh"

namespace tb {

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: results do not depend on the grid,
// the GPU count or the order walkers are processed in.
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed, uint64_t iteration)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32) ^ (uint32_t)iteration) {}
  __device__ __forceinline__ uint4 block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// purposes (counter word 3, high byte)
enum { RNG_PRIOR = 0, RNG_GAMMA = 1, RNG_NORMAL = 2, RNG_ACCEPT = 3, RNG_RESAMPLE = 4, RNG_TRAIN = 5 };

__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {          // [0,1)  (numpy random_sample layout)
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u53_open(uint32_t lo, uint32_t hi) {     // (0,1)
  return ((double)((((uint64_t)hi << 32) | lo) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
// two uniforms of walker `slot` for (step, purpose, sub)
__device__ __forceinline__ void philox_u2(const Philox& g, uint64_t slot, uint32_t step, uint32_t purpose,
                                          uint32_t sub, double& a, double& b, bool open) {
  uint4 r = g.block((uint32_t)slot, (uint32_t)(slot >> 32), step, (purpose << 24) | (sub & 0xffffffu));
  a = open ? u53_open(r.x, r.y) : u53(r.x, r.y);
  b = open ? u53_open(r.z, r.w) : u53(r.z, r.w);
}
// two standard normals (Box-Muller)
__device__ __forceinline__ void philox_n2(const Philox& g, uint64_t slot, uint32_t step, uint32_t sub,
                                          double& z0, double& z1) {
  double a, b;
  philox_u2(g, slot, step, RNG_NORMAL, sub, a, b, true);
  const double r = sqrt(-2.0 * log(a));
  double s, c;
  sincospi(2.0 * b, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// ---- priors --------------------------------------------------------------------------
__device__ __forceinline__ double prior_affine(const double* __restrict__ pp, int d, int j, double u) {
  return __dadd_rn(__ldg(pp + j), __dmul_rn(__ldg(pp + d + j), u));   // lo + scale*u
}

// ---- likelihoods ------------------------------------------------------------------------
template <typename X>
__device__ __forceinline__ double like_rosenbrock(const double* __restrict__ lp, int d, const X& x) {
  const double a = __ldg(lp);
  double acc = 0.0;
  for (int i = 0; i < d / 2; ++i) {
    const double xe = x[2 * i], xo = x[2 * i + 1];
    double t = __dsub_rn(__dmul_rn(xe, xe), xo);
    t = __dmul_rn(a, __dmul_rn(t, t));
    const double s = __dsub_rn(xe, 1.0);
    const double term = __dadd_rn(t, __dmul_rn(s, s));
    acc = (i == 0) ? term : __dadd_rn(acc, term);
  }
  return -acc;
}

template <typename X>
__device__ __forceinline__ double like_gaussian(const double* __restrict__ lp, int d, const X& x) {
  // [const, mean[d], Linv[d*d]]
  const double cst = __ldg(lp);
  const double* mean = lp + 1;
  const double* linv = lp + 1 + d;
  double acc = 0.0;
  for (int i = 0; i < d; ++i) {
    double y = __dmul_rn(__ldg(linv + i * d), __dsub_rn(x[0], __ldg(mean)));
    for (int j = 1; j <= i; ++j)
      y = __dadd_rn(y, __dmul_rn(__ldg(linv + i * d + j), __dsub_rn(x[j], __ldg(mean + j))));
    const double sq = __dmul_rn(y, y);
    acc = (i == 0) ? sq : __dadd_rn(acc, sq);
  }
  return __dadd_rn(__dmul_rn(-0.5, acc), cst);
}

template <typename X>
__device__ __forceinline__ double like_iso_mixture(const double* __restrict__ lp, int d, const X& x) {
  // [K, c[K], h[K], means[K*d]]
  const int K = (int)__ldg(lp);
  const double* c = lp + 1;
  const double* h = lp + 1 + K;
  const double* mu = lp + 1 + 2 * K;
  double m = -INFINITY;
  for (int k = 0; k < K; ++k) {
    double r2 = 0.0;
    for (int j = 0; j < d; ++j) {
      const double dl = __dsub_rn(x[j], __ldg(mu + k * d + j));
      const double sq = __dmul_rn(dl, dl);
      r2 = (j == 0) ? sq : __dadd_rn(r2, sq);
    }
    const double lpk = __dsub_rn(__ldg(c + k), __dmul_rn(r2, __ldg(h + k)));
    m = (k == 0) ? lpk : fmax(m, lpk);
  }
  double s = 0.0;
  for (int k = 0; k < K; ++k) {   // recompute lp_k: K is small and this keeps registers flat
    double r2 = 0.0;
    for (int j = 0; j < d; ++j) {
      const double dl = __dsub_rn(x[j], __ldg(mu + k * d + j));
      const double sq = __dmul_rn(dl, dl);
      r2 = (j == 0) ? sq : __dadd_rn(r2, sq);
    }
    const double lpk = __dsub_rn(__ldg(c + k), __dmul_rn(r2, __ldg(h + k)));
    const double e = exp(__dsub_rn(lpk, m));
    s = (k == 0) ? e : __dadd_rn(s, e);
  }
  return __dadd_rn(m, log(s));
}

template <typename X>
__device__ __forceinline__ double shell_term(const double* __restrict__ c, int d, const X& x, double r, double h,
                                             double cst) {
  double r2 = 0.0;
  for (int j = 0; j < d; ++j) {
    const double dl = __dsub_rn(x[j], __ldg(c + j));
    const double sq = __dmul_rn(dl, dl);
    r2 = (j == 0) ? sq : __dadd_rn(r2, sq);
  }
  const double t = __dsub_rn(sqrt(r2), r);
  return __dsub_rn(cst, __dmul_rn(__dmul_rn(t, t), h));
}

template <typename X>
__device__ __forceinline__ double like_twin_shells(const double* __restrict__ lp, int d, const X& x) {
  // [r, h, const, c1[d], c2[d]]
  const double r = __ldg(lp), h = __ldg(lp + 1), cst = __ldg(lp + 2);
  return np_logaddexp(shell_term(lp + 3, d, x, r, h, cst), shell_term(lp + 3 + d, d, x, r, h, cst));
}

template <typename X>
__device__ __forceinline__ double eval_like(int like_id, const double* __restrict__ lp, int d, const X& x) {
  switch (like_id) {
    case TB_LIKE_ROSENBROCK: return like_rosenbrock(lp, d, x);
    case TB_LIKE_GAUSSIAN: return like_gaussian(lp, d, x);
    case TB_LIKE_ISO_MIXTURE: return like_iso_mixture(lp, d, x);
    case TB_LIKE_TWIN_SHELLS: return like_twin_shells(lp, d, x);
    default: return NAN;
  }
}

}  // namespace tb
