// Grid-wide + cross-GPU reduction of a few doubles inside a persistent (cooperative) kernel
// (the per-probe and per-Metropolis-step collectives).
//
// One call = ONE synchronisation point of the whole job:
//   1. every CTA publishes its partial row and takes a ticket (no CTA waits here);
//   2. the CTA that arrives last on this GPU folds the rows in a fixed order and stores the GPU's
//      payload into slot[parity][my_rank] of EVERY rank's exchange buffer (lane r serves rank r: the
//      eight NVLink stores leave in parallel), each followed by a system-scope release store of the
//      sequence number;
//   3. warp 0 of every CTA of every GPU polls its OWN GPU's buffer (lane r watches rank r's flag) until
//      all ranks carry the sequence number, then folds the ranks' payloads in rank order.
// The flag wait IS the grid barrier (a rank's flag is only published after all of its CTAs arrived), so a
// step costs one ticket + one flag round trip instead of barrier -> exchange -> barrier, and every CTA of
// every rank ends with bitwise-identical totals.  With one GPU the "exchange buffer" is a few hundred
// bytes of the kernel's own workspace and the same code runs.
//
// Exchange buffers: `peer[r]` is the address of rank r's buffer as seen from THIS device (symmetric /
// peer-mapped memory).  Sequence numbers are consecutive across launches and slots alternate by parity:
// a rank publishes exchange s+2 only after it has seen every rank's flag of exchange s+1, which a rank
// sets only after all of its CTAs finished reading exchange s -- a slot is never overwritten while read.
// Spins are bounded (kXSpinBudgetNs): a rank that died leaves the others with an error code instead of a
// hung context.
#pragma once
#include "tb_common.cuh"

namespace tb {

constexpr int kXSlotDoubles = 72;   // 1 flag word + up to 71 payload doubles
constexpr int kXMaxRanks = 8;
constexpr int kXMaxPayload = kXSlotDoubles - 1;
constexpr unsigned long long kXSpinBudgetNs = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Workspace of one grid-wide reduction stream (device memory, zeroed by the launcher before each launch).
struct GridSync {
  unsigned int arrive;              // monotone ticket counter of the current launch
  unsigned int pad[3];
  unsigned long long bflag[2];      // local broadcast flags (sequence number of the result in bcast[parity])
  double bcast[2][kXSlotDoubles];   // job-wide result of the current sync point, published to this GPU's CTAs
  // followed by the partial rows: double rows[2][grid][W]
  __device__ __forceinline__ double* rows() { return reinterpret_cast<double*>(this + 1); }
};
__host__ __device__ inline size_t grid_sync_bytes(int grid, int W) {
  return sizeof(GridSync) + sizeof(double) * 2 * (size_t)grid * (size_t)W;
}

// Executed by ALL threads of every CTA, once per synchronisation point `k` = 0, 1, ... of this launch.
// part[W] (shared): this CTA's partial.  On return tot[W] (shared, visible to the whole CTA) holds the job-wide
// result.  `fold` combines rows / rank payloads in a fixed order (ColumnFold below; EssFold in tb_reweight.cu).
//   1. every CTA publishes its row and takes a ticket (nobody waits here);
//   2. the CTA that arrives last on this GPU folds the rows (all its threads: many loads in flight), stores the
//      GPU's payload into slot[parity][my_rank] of EVERY rank's exchange buffer (thread r serves rank r: the NVLink
//      stores leave in parallel) with a system-scope release of the sequence number, waits for the ranks' flags in
//      its OWN buffer (thread r watches rank r), folds the ranks' payloads in rank order and publishes the result to
//      the other CTAs of this GPU through a gpu-scope flag;
//   3. every other CTA spins on that LOCAL flag only (one thread, gpu scope): no system-scope polling by thousands
//      of threads, which delayed the very stores they were waiting for (23 us per step on 8 GPUs).
// The flag wait is the grid barrier.  Returns 0, or 3 when a peer did not answer within the spin budget.
template <class RowFold>
__device__ __noinline__ int grid_xreduce(GridSync* gs, const tb_xgpu& x, int k, int W, const double* part, double* tot,
                                         RowFold fold) {
  __shared__ int s_last, s_bad;
  const int lane = threadIdx.x & 31;
  const int nb = gridDim.x;
  double* rows = gs->rows() + (size_t)(k & 1) * nb * W;
  const unsigned long long seq = x.seq + (unsigned long long)k;
  const int par = (int)(seq & 1ull);
  const int world = x.world;
  if (threadIdx.x < 32) {
    for (int c = lane; c < W; c += 32) __stcg(rows + (size_t)blockIdx.x * W + c, part[c]);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      // (a two-level arrival -- 32 group counters, then one -- was measured SLOWER: +1 us at every grid size; the extra
      //  fence + atomic round trip costs more than a thousand arrivals on one L2 counter do)
      const unsigned int old = atomicAdd(&gs->arrive, 1u);
      s_last = (old == (unsigned)(k + 1) * (unsigned)nb - 1u) ? 1 : 0;     // last CTA of this GPU for sync point k
      s_bad = 0;
    }
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    fold.rows(rows, nb, W, tot);                          // fixed order; result in tot[] (shared); ends with a CTA barrier
    if (world > 1) {
      double* mybuf = x.peer[x.rank];
      if ((int)threadIdx.x < world) {
        double* slot = x.peer[threadIdx.x] + (size_t)(par * kXMaxRanks + x.rank) * kXSlotDoubles;
        for (int i = 0; i < W; ++i) st_relaxed_sys(slot + 1 + i, tot[i]);
        st_release_sys(reinterpret_cast<unsigned long long*>(slot), seq);
        const unsigned long long* flag =
            reinterpret_cast<const unsigned long long*>(mybuf + (size_t)(par * kXMaxRanks + threadIdx.x) * kXSlotDoubles);
        if (ld_acquire_sys(flag) != seq) {
          const unsigned long long t0 = global_ns();
          while (ld_acquire_sys(flag) != seq) {
            if (global_ns() - t0 > kXSpinBudgetNs) { s_bad = 1; break; }
          }
        }
      }
      __syncthreads();
      if (!s_bad && threadIdx.x < 32) fold.ranks(mybuf + (size_t)par * kXMaxRanks * kXSlotDoubles, world, W, tot);
      __syncthreads();
    }
    // publish to the other CTAs of this GPU
    if (threadIdx.x < 32) {
      for (int c = lane; c < W; c += 32) __stcg(&gs->bcast[par][c], tot[c]);
      if (lane == 0 && s_bad) __stcg(&gs->bcast[par][kXSlotDoubles - 1], 3.0);
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_gpu(&gs->bflag[par], seq);
    }
    __syncthreads();
    return s_bad ? 3 : 0;
  }
  if (threadIdx.x == 0) {
    if (ld_acquire_gpu(&gs->bflag[par]) != seq) {
      const unsigned long long t0 = global_ns();
      while (ld_acquire_gpu(&gs->bflag[par]) != seq) {
        if (global_ns() - t0 > 2 * kXSpinBudgetNs) { s_bad = 1; break; }
        __nanosleep(40);      // the pollers of finished CTAs took ~15 % of the issue slots from the warps still working
      }
    }
  }
  __syncthreads();
  if (s_bad) return 3;
  for (int c = threadIdx.x; c < W; c += blockDim.x) tot[c] = __ldcg(&gs->bcast[par][c]);
  const int rc = (__ldcg(&gs->bcast[par][kXSlotDoubles - 1]) == 3.0) ? 3 : 0;
  __syncthreads();
  return rc;
}

// Column-wise fold: sum, except the columns flagged in max_cols (error codes).  Row order is fixed: thread t takes
// rows t, t + B, ... into four independent accumulators (loads in flight), then lanes and warps fold in order.
struct ColumnFold {
  unsigned long long max_cols;
  __device__ void rows(const double* rows, int nb, int W, double* tot) const {
    __shared__ double s_w[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5, B = blockDim.x;
    if (W <= 4 && nw <= 16) {
      // narrow rows (single-mode runs: sum alpha, accepted, proposals, error): every thread folds whole rows, so the
      // columns share one pass over the rows, one warp reduction each and ONE barrier (the per-column form below
      // costs two barriers per column on the critical path of every Metropolis step)
      __shared__ double s_wc[16][4];
      double v[4][2];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool mx = (max_cols >> c) & 1ull;
        v[c][0] = v[c][1] = mx ? -INFINITY : 0.0;
      }
      int b = threadIdx.x;
      for (; b + B < nb; b += 2 * B) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < W) {
            const bool mx = (max_cols >> c) & 1ull;
            const double r0 = __ldcg(rows + (size_t)b * W + c), r1 = __ldcg(rows + (size_t)(b + B) * W + c);
            v[c][0] = mx ? fmax(v[c][0], r0) : v[c][0] + r0;
            v[c][1] = mx ? fmax(v[c][1], r1) : v[c][1] + r1;
          }
        }
      }
      if (b < nb) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < W) {
            const bool mx = (max_cols >> c) & 1ull;
            const double r0 = __ldcg(rows + (size_t)b * W + c);
            v[c][0] = mx ? fmax(v[c][0], r0) : v[c][0] + r0;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool mx = (max_cols >> c) & 1ull;
        double x = mx ? fmax(v[c][0], v[c][1]) : v[c][0] + v[c][1];
        x = mx ? warp_max(x) : warp_sum(x);
        if (lane == 0) s_wc[wid][c] = x;
      }
      __syncthreads();
      if ((int)threadIdx.x < W) {
        const int c = threadIdx.x;
        const bool mx = (max_cols >> c) & 1ull;
        double t = s_wc[0][c];
        for (int i = 1; i < nw; ++i) t = mx ? fmax(t, s_wc[i][c]) : t + s_wc[i][c];
        tot[c] = t;
      }
      __syncthreads();
      return;
    }
    for (int c = 0; c < W; ++c) {
      const bool mx = c < 64 && ((max_cols >> c) & 1ull);
      const double id = mx ? -INFINITY : 0.0;
      double v0 = id, v1 = id, v2 = id, v3 = id;
      int b = threadIdx.x;
      for (; b + 3 * B < nb; b += 4 * B) {
        const double r0 = __ldcg(rows + (size_t)b * W + c), r1 = __ldcg(rows + (size_t)(b + B) * W + c);
        const double r2 = __ldcg(rows + (size_t)(b + 2 * B) * W + c), r3 = __ldcg(rows + (size_t)(b + 3 * B) * W + c);
        v0 = mx ? fmax(v0, r0) : v0 + r0; v1 = mx ? fmax(v1, r1) : v1 + r1;
        v2 = mx ? fmax(v2, r2) : v2 + r2; v3 = mx ? fmax(v3, r3) : v3 + r3;
      }
      for (; b < nb; b += B) { const double r = __ldcg(rows + (size_t)b * W + c); v0 = mx ? fmax(v0, r) : v0 + r; }
      double v = mx ? fmax(fmax(v0, v1), fmax(v2, v3)) : (v0 + v1) + (v2 + v3);
      v = mx ? warp_max(v) : warp_sum(v);
      __syncthreads();
      if (lane == 0) s_w[wid] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = s_w[0];
        for (int i = 1; i < nw; ++i) t = mx ? fmax(t, s_w[i]) : t + s_w[i];
        tot[c] = t;
      }
    }
    __syncthreads();
  }
  __device__ void ranks(const double* slots /* [ranks][kXSlotDoubles] of this parity */, int world, int W, double* tot) const {
    const int lane = threadIdx.x & 31;
    for (int c = lane; c < W; c += 32) {
      const bool mx = c < 64 && ((max_cols >> c) & 1ull);
      double v = mx ? -INFINITY : 0.0;
      for (int r = 0; r < world; ++r) {
        const double* slot = slots + (size_t)r * kXSlotDoubles;
        (void)ld_acquire_sys(reinterpret_cast<const unsigned long long*>(slot));   // orders this lane's payload reads
        const double p = ld_relaxed_sys(slot + 1 + c);
        v = mx ? fmax(v, p) : v + p;
      }
      tot[c] = v;
    }
  }
};

}  // namespace tb
