// Small-message collectives over NVLink peer memory ("all collectives are tiny and latency-bound").
//
// The sharded Trainer issues ~25 all-reduces / all-gathers of a few numbers to a few hundred KB per PS iteration
// (normalisation sums, 2048-bin histograms, order-statistic buckets, moment sums).  Through NCCL each costs a
// launch + protocol latency of 30-50 us; here one exchange is two small kernels on the caller's stream:
//   push    every rank stores its payload into slot[parity][my_rank] of EVERY rank's staging buffer (plain peer
//           stores), the last CTA publishes the call number with a system-scope release store per peer;
//   fold    every CTA waits until all ranks' flags carry the call number, then sums the ranks' slots in rank order
//           (all-reduce: bitwise identical results everywhere) or copies them out (all-gather).
// No host synchronisation, no co-residency requirement (push never waits; fold waits only for kernels that were
// already launched on every rank).  Slots alternate by call parity: a rank can be at most one call ahead of its
// slowest peer (it needs that peer's flag of call n+1, published after the peer's fold of call n was enqueued ahead
// of its push of call n+1 on the same stream), so a slot is never overwritten while it is being read.
#include "tb_common.cuh"
#include "tb_xgpu.cuh"

namespace {
using namespace tb;

constexpr int kCollBlock = 256;

struct CollArgs {
  char* peer[kXMaxRanks];      // staging buffer of every rank as addressed from here
  int rank, world;
  unsigned long long seq;      // call number >= 1
  size_t cap;                  // payload capacity per slot (bytes)
  unsigned int* ticket;        // local device counter (zero between calls)
};
// staging layout: flags[2][8] (u64) | pad to 256 | slot[2][world_max=8][cap]
__host__ __device__ inline size_t coll_header() { return 256; }
__device__ __forceinline__ unsigned long long* coll_flags(char* base, int par) {
  return reinterpret_cast<unsigned long long*>(base) + par * kXMaxRanks;
}
__device__ __forceinline__ char* coll_slot(char* base, size_t cap, int par, int r) {
  return base + coll_header() + ((size_t)par * kXMaxRanks + r) * cap;
}

__global__ void __launch_bounds__(kCollBlock)
xcoll_push_kernel(CollArgs a, const char* __restrict__ src, size_t nbytes) {
  const int par = (int)(a.seq & 1ull);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const bool wide = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (nbytes & 15) == 0;     // slots are 16-byte aligned
  for (int r = 0; r < a.world; ++r) {
    char* dst = coll_slot(a.peer[r], a.cap, par, a.rank);
    if (wide) {
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes / 16; i += stride)
        reinterpret_cast<int4*>(dst)[i] = __ldg(reinterpret_cast<const int4*>(src) + i);
    } else {                                            // payloads are multiples of 4 bytes, 4-byte aligned
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes / 4; i += stride)
        reinterpret_cast<int*>(dst)[i] = __ldg(reinterpret_cast<const int*>(src) + i);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    if (t == gridDim.x - 1) {
      *a.ticket = 0u;
      __threadfence_system();
      for (int r = 0; r < a.world; ++r) st_release_sys(coll_flags(a.peer[r], par) + a.rank, a.seq);
    }
  }
}

__device__ inline bool coll_wait(const CollArgs& a, int* err) {
  __shared__ int s_ok;
  const int par = (int)(a.seq & 1ull);
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const unsigned long long* f = coll_flags(a.peer[a.rank], par) + threadIdx.x;
    if (ld_acquire_sys(f) != a.seq) {
      const unsigned long long t0 = global_ns();
      while (ld_acquire_sys(f) != a.seq)
        if (global_ns() - t0 > kXSpinBudgetNs) { s_ok = 0; break; }
    }
  }
  __syncthreads();
  if (!s_ok && threadIdx.x == 0 && err) *err = 3;
  return s_ok != 0;
}

template <typename T>
__global__ void __launch_bounds__(kCollBlock)
xcoll_sum_kernel(CollArgs a, T* __restrict__ data, size_t n, int* err) {
  if (!coll_wait(a, err)) return;
  const int par = (int)(a.seq & 1ull);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // L2 loads: the slots were written by other devices
    T acc = __ldcg(reinterpret_cast<const T*>(coll_slot(a.peer[a.rank], a.cap, par, 0)) + i);
    for (int r = 1; r < a.world; ++r) acc += __ldcg(reinterpret_cast<const T*>(coll_slot(a.peer[a.rank], a.cap, par, r)) + i);
    data[i] = acc;
  }
}

__global__ void __launch_bounds__(kCollBlock)
xcoll_gather_kernel(CollArgs a, char* __restrict__ out, size_t nbytes, int* err) {
  if (!coll_wait(a, err)) return;
  const int par = (int)(a.seq & 1ull);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < a.world; ++r) {
    const char* src = coll_slot(a.peer[a.rank], a.cap, par, r);
    char* dst = out + (size_t)r * nbytes;
    if ((nbytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes / 16; i += stride)
        reinterpret_cast<int4*>(dst)[i] = __ldcg(reinterpret_cast<const int4*>(src) + i);
    } else {
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes / 4; i += stride)
        reinterpret_cast<int*>(dst)[i] = __ldcg(reinterpret_cast<const int*>(src) + i);
    }
  }
}

// ---- resampled rows: gather from the local history and store straight into the slot owner's active set ----------
// Every rank searched all N global draws; idx[k] >= 0 marks the draws whose ancestor it stores.  Row k = (u[idx[k]],
// logl[idx[k]]) is written into the buffer of the rank that owns walker slot k (dest = k / per) over NVLink -- no
// counts, no host synchronisation, no staging copy.  Buffer layout (doubles): 32 flag words | parity 0: u[per][d],
// logl[per] | parity 1: the same.  The last CTA publishes the call number to every peer; xrows_wait_kernel then
// holds the stream until all ranks' rows have arrived.
struct RowsArgs {
  double* peer[kXMaxRanks];
  int rank, world;
  unsigned long long seq;
  unsigned int* ticket;
  long long per;               // walker slots per rank
};
__device__ __forceinline__ double* rows_u(double* base, long long per, int d, int par) {
  return base + 32 + (size_t)par * ((size_t)per * (d + 1));
}

__global__ void __launch_bounds__(kCollBlock)
xrows_scatter_kernel(RowsArgs a, const double* __restrict__ hu, const double* __restrict__ hl, int d,
                     const int64_t* __restrict__ idx, int64_t n_glob) {
  const int par = (int)(a.seq & 1ull);
  const int64_t total = n_glob * (d + 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t k = e / (d + 1);
    const int c = (int)(e - k * (d + 1));
    const int64_t src = __ldg(idx + k);
    if (src < 0) continue;
    const int dest = (int)(k / a.per);
    const int64_t j = k - (int64_t)dest * a.per;
    double* ub = rows_u(a.peer[dest], a.per, d, par);
    if (c < d) ub[j * d + c] = __ldg(hu + src * d + c);
    else ub[(size_t)a.per * d + j] = __ldg(hl + src);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    if (t == gridDim.x - 1) {
      *a.ticket = 0u;
      __threadfence_system();
      for (int r = 0; r < a.world; ++r)
        st_release_sys(reinterpret_cast<unsigned long long*>(a.peer[r]) + par * kXMaxRanks + a.rank, a.seq);
    }
  }
}

__global__ void xrows_wait_kernel(RowsArgs a, int* err) {
  const int par = (int)(a.seq & 1ull);
  if ((int)threadIdx.x < a.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.peer[a.rank]) + par * kXMaxRanks + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) != a.seq)
      if (global_ns() - t0 > kXSpinBudgetNs) { if (err) *err = 3; break; }
  }
}

int make(CollArgs& a, const tb_xcoll* x) {
  if (!x || x->world < 2 || x->world > kXMaxRanks || x->seq < 1 || !x->ticket || x->cap_bytes == 0) return TB_ERR_ARG;
  for (int r = 0; r < kXMaxRanks; ++r) a.peer[r] = reinterpret_cast<char*>(x->peer[r]);
  a.rank = x->rank; a.world = x->world; a.seq = x->seq; a.cap = x->cap_bytes; a.ticket = reinterpret_cast<unsigned int*>(x->ticket);
  return TB_OK;
}
int grid_for(size_t nbytes) {
  size_t g = (nbytes / 16 + kCollBlock * 4 - 1) / (kCollBlock * 4);
  if (g < 1) g = 1;
  if (g > 64) g = 64;
  return (int)g;
}

}  // namespace

extern "C" {

size_t tb_xcoll_buffer_bytes(size_t cap_bytes) { return coll_header() + 2 * (size_t)kXMaxRanks * cap_bytes; }

int tb_xcoll_allreduce_sum(void* data, int64_t n, int32_t dtype, const tb_xcoll* x, int32_t* err, tb_stream_t stream) {
  CollArgs a;
  if (int rc = make(a, x)) return rc;
  const size_t esz = dtype == 2 ? 4 : 8;
  if (n <= 0 || !data || (size_t)n * esz > a.cap || (dtype != 0 && dtype != 1 && dtype != 2)) return TB_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(data) & (esz - 1)) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const size_t nbytes = (size_t)n * esz;
  const int g = grid_for(nbytes);
  xcoll_push_kernel<<<g, kCollBlock, 0, st>>>(a, reinterpret_cast<const char*>(data), nbytes);
  if (dtype == 0) xcoll_sum_kernel<double><<<g, kCollBlock, 0, st>>>(a, reinterpret_cast<double*>(data), (size_t)n, err);
  else if (dtype == 1) xcoll_sum_kernel<long long><<<g, kCollBlock, 0, st>>>(a, reinterpret_cast<long long*>(data), (size_t)n, err);
  else xcoll_sum_kernel<int><<<g, kCollBlock, 0, st>>>(a, reinterpret_cast<int*>(data), (size_t)n, err);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

size_t tb_xrows_buffer_bytes(int64_t per, int32_t d) { return sizeof(double) * (32 + 2 * (size_t)per * (size_t)(d + 1)); }
int64_t tb_xrows_offset(int64_t per, int32_t d, int32_t parity) { return 32 + (int64_t)parity * per * (d + 1); }

int tb_xrows_scatter(const double* hu, const double* hl, int32_t d, const int64_t* idx, int64_t n_glob, int64_t per,
                     const tb_xcoll* x, int32_t* err, tb_stream_t stream) {
  if (!x || x->world < 2 || x->world > kXMaxRanks || x->seq < 1 || !x->ticket || !hu || !hl || !idx || d <= 0 || n_glob <= 0 ||
      per <= 0 || per * x->world != n_glob)
    return TB_ERR_ARG;
  RowsArgs a;
  for (int r = 0; r < kXMaxRanks; ++r) a.peer[r] = reinterpret_cast<double*>(x->peer[r]);
  a.rank = x->rank; a.world = x->world; a.seq = x->seq; a.ticket = reinterpret_cast<unsigned int*>(x->ticket); a.per = per;
  cudaStream_t st = as_stream(stream);
  xrows_scatter_kernel<<<stream_grid(n_glob * (d + 1), kCollBlock, 8), kCollBlock, 0, st>>>(a, hu, hl, d, idx, n_glob);
  xrows_wait_kernel<<<1, 32, 0, st>>>(a, err);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

int tb_xcoll_allgather(const void* src, int64_t nbytes, void* out, const tb_xcoll* x, int32_t* err, tb_stream_t stream) {
  CollArgs a;
  if (int rc = make(a, x)) return rc;
  if (nbytes <= 0 || (nbytes & 3) || !src || !out || (size_t)nbytes > a.cap || (reinterpret_cast<uintptr_t>(src) & 3)) return TB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for((size_t)nbytes);
  xcoll_push_kernel<<<g, kCollBlock, 0, st>>>(a, reinterpret_cast<const char*>(src), (size_t)nbytes);
  xcoll_gather_kernel<<<g, kCollBlock, 0, st>>>(a, reinterpret_cast<char*>(out), (size_t)nbytes, err);
  TB_CHECK_LAUNCH();
  return TB_OK;
}

}  // extern "C"
