// Shared device helpers for libtempest_b200 (sm_100a).  fp64 throughout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/tempest_b200.h"

#define TB_CHECK_LAUNCH()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace tb {

constexpr int kBlock = 256;          // threads per CTA for the streaming kernels
constexpr int kMaxPartials = 4096;   // upper bound on CTAs that publish partial results

// Number of SMs of the current device (148 on B200); grids are sized as multiples of it.
inline int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

// Persistent-style grid: `per_sm` CTAs per SM, never more CTAs than there is work.
inline int stream_grid(int64_t n, int items_per_block, int per_sm) {
  int64_t need = (n + items_per_block - 1) / items_per_block;
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (cap > kMaxPartials) cap = kMaxPartials;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

inline cudaStream_t as_stream(tb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// np.logaddexp (numpy/_core/src/npymath/npy_math_internal.h.src, npy_logaddexp):
//   x == y -> x + log 2 ; else with t = x - y: t > 0 -> x + log1p(exp(-t)); t <= 0 -> y + log1p(exp(t));
//   NaN otherwise.
__device__ __forceinline__ double np_logaddexp(double x, double y) {
  if (x == y) return x + 0.693147180559945309417232121458176568;
  double t = x - y;
  if (t > 0) return x + log1p(exp(-t));
  if (t <= 0) return y + log1p(exp(t));
  return t;  // NaN
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum: butterfly inside warps, then a fixed-order pass over the warps.
// Result valid in every thread.  `smem` must hold >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += smem[i];
    smem[32] = t;
  }
  __syncthreads();
  return smem[32];
}
__device__ __forceinline__ double block_max(double v, double* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = smem[0];
    for (int i = 1; i < nw; ++i) t = fmax(t, smem[i]);
    smem[32] = t;
  }
  __syncthreads();
  return smem[32];
}

// "last block done" ticket: returns true in every thread of the CTA that arrives last.
// The counter resets itself so the same workspace can be reused by the next launch.
__device__ __forceinline__ bool last_block_arrives(unsigned int* ticket) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Order-fixed sum of column `col` over `rows` partial rows of width `stride`, with eight independent
// accumulators so that eight L2 loads are in flight (a single running sum serialises on the ~0.6 us
// round trip per unrolled group and turns last-CTA epilogues into 30-200 us tails).
__device__ __forceinline__ double fold_column(const double* __restrict__ partial, int rows, size_t stride, int col) {
  double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int b = 0;
  for (; b + 8 <= rows; b += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += __ldcg(partial + (size_t)(b + q) * stride + col);
  }
  double tail = 0.0;
  for (; b < rows; ++b) tail += __ldcg(partial + (size_t)b * stride + col);
  return (((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))) + tail;
}

// Streaming (max, S1, S2) accumulator of w = exp(a - max a) and w^2 with online rescaling.
struct Ess3 {
  double m, s1, s2;
  __device__ __forceinline__ void init() { m = -INFINITY; s1 = 0.0; s2 = 0.0; }
  __device__ __forceinline__ void push(double a) {
    if (a <= m) {
      double d = a - m;
      if (d > -40.0) {   // bitwise neutral: s1, s2 >= 1 (the maximum's own term), and e^-40 < 2^-53 cannot change them
        double w = exp(d);
        s1 += w;
        s2 += w * w;
      }
    } else if (a > m) {
      double r = (m == -INFINITY) ? 0.0 : exp(m - a);
      s1 = s1 * r + 1.0;
      s2 = s2 * (r * r) + 1.0;
      m = a;
    }  // NaN: ignored here, counted separately by the caller
  }
  __device__ __forceinline__ void merge(double m2, double a2, double b2) {
    if (m2 == -INFINITY) return;
    if (m == -INFINITY) { m = m2; s1 = a2; s2 = b2; return; }
    if (m2 <= m) {
      double r = exp(m2 - m);
      s1 += a2 * r;
      s2 += b2 * (r * r);
    } else {
      double r = exp(m - m2);
      s1 = s1 * r + a2;
      s2 = s2 * (r * r) + b2;
      m = m2;
    }
  }
};

// Fixed-order CTA merge of Ess3 triples; result valid in thread 0 (and broadcast via smem[0..2]).
__device__ __forceinline__ void block_merge_ess3(Ess3& e, double* smem /* >= 3*32+3 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double m2 = __shfl_xor_sync(0xffffffffu, e.m, o);
    double a2 = __shfl_xor_sync(0xffffffffu, e.s1, o);
    double b2 = __shfl_xor_sync(0xffffffffu, e.s2, o);
    // symmetric merge so that both partners end with bitwise-identical triples
    Ess3 lo, hi;
    bool self_first = (lane & o) == 0;
    lo.m = self_first ? e.m : m2;  lo.s1 = self_first ? e.s1 : a2;  lo.s2 = self_first ? e.s2 : b2;
    hi.m = self_first ? m2 : e.m;  hi.s1 = self_first ? a2 : e.s1;  hi.s2 = self_first ? b2 : e.s2;
    lo.merge(hi.m, hi.s1, hi.s2);
    e = lo;
  }
  __syncthreads();
  if (lane == 0) { smem[3 * wid] = e.m; smem[3 * wid + 1] = e.s1; smem[3 * wid + 2] = e.s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    Ess3 t; t.m = smem[0]; t.s1 = smem[1]; t.s2 = smem[2];
    for (int i = 1; i < nw; ++i) t.merge(smem[3 * i], smem[3 * i + 1], smem[3 * i + 2]);
    e = t;
  }
}

}  // namespace tb
