// In-kernel all-gather of a few doubles across the GPUs of one NVSwitch domain.
//
// Every rank owns a small exchange buffer in symmetric (peer-mapped) memory; `peer[r]` is the address
// of rank r's buffer as seen from THIS device (NVLink P2P stores / loads).  One exchange = every rank
// stores its payload into slot[parity][my_rank] of every peer, fences, publishes the sequence number
// with a system-scope release store, then spins (acquire loads of its OWN buffer) until all ranks'
// flags carry that sequence number, and reads the payloads in rank order -> every rank folds the same
// values in the same order, so decisions taken from them are bitwise identical everywhere.
// Sequence numbers are consecutive across launches, slots alternate by parity: a rank can run at most
// one exchange ahead of the slowest rank, so a slot is never overwritten while it is being read.
// Replaces a kernel -> NCCL all-reduce -> kernel round trip (~30-50 us of launch + host latency) by
// ~2-4 us of NVLink latency inside the kernel that produced the partial result.
#pragma once
#include "tb_common.cuh"

namespace tb {

constexpr int kXSlotDoubles = 16;   // 1 flag word + up to 15 payload doubles
constexpr int kXMaxRanks = 8;

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

// Called by ONE thread.  `all` receives world*k doubles in rank order.  seq must be >= 1.
__device__ inline void xgpu_allgather(const tb_xgpu& x, unsigned long long seq, const double* mine, int k, double* all) {
  const int par = (int)(seq & 1ull);
  for (int r = 0; r < x.world; ++r) {
    double* slot = x.peer[r] + (size_t)(par * kXMaxRanks + x.rank) * kXSlotDoubles;
    for (int i = 0; i < k; ++i) st_relaxed_sys(slot + 1 + i, mine[i]);
  }
  __threadfence_system();
  for (int r = 0; r < x.world; ++r) {
    double* slot = x.peer[r] + (size_t)(par * kXMaxRanks + x.rank) * kXSlotDoubles;
    st_release_sys(reinterpret_cast<unsigned long long*>(slot), seq);
  }
  double* my = x.peer[x.rank];
  for (int r = 0; r < x.world; ++r) {
    const double* slot = my + (size_t)(par * kXMaxRanks + r) * kXSlotDoubles;
    while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(slot)) != seq) { __nanosleep(40); }
    for (int i = 0; i < k; ++i) all[r * k + i] = ld_relaxed_sys(slot + 1 + i);
  }
}

}  // namespace tb
