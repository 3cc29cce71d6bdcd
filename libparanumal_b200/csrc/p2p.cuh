// Device-side primitives of the NVLink peer-window protocol (see common.hpp: WinHeader / WinAR).
#pragma once
#include "common.hpp"

namespace libp_b200 {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// data that a peer GPU wrote into my window: always read from L2 (never a stale L1 line)
__device__ __forceinline__ double ld_peer_written(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Bounded wait for *p >= target.  Returns false (and raises the communicator's error word) when it lasted longer than
// `timeout` SM cycles or another wait already failed: a peer that threw between collectives, or a protocol bug, then
// shows up as an error instead of a hung GPU.
__device__ __forceinline__ bool wait_ge_sys(const unsigned long long* p, unsigned long long target, int* err,
                                            long long timeout) {
  if (ld_acquire_sys(p) >= target) return true;
  const long long t0 = clock64();
  unsigned it = 0;
  while (ld_acquire_sys(p) < target) {
    if ((++it & 0xff) == 0) {
      if (err != nullptr && *reinterpret_cast<volatile int*>(err) != 0) return false;
      if (timeout > 0 && clock64() - t0 > timeout) {
        if (err != nullptr) atomicExch(err, 1);
        return false;
      }
    }
  }
  return true;
}

// All-reduce (sum) of n <= kWinMaxVals doubles held in shared memory `vals`, by ONE thread block per rank.
// Every rank stores its values into every peer's mailbox, publishes the sequence number, waits for all peers'
// flags and sums in rank order - every rank gets the bit-identical result.  Two mailbox parities suffice: a rank
// can only contribute to all-reduce s+1 after it finished reading s, and nobody completes s+1 without that
// contribution, so the slot of s is free again when s+2 is written.
__device__ __forceinline__ void win_allreduce_sum(const WinAR& w, double* vals, int n) {
  if (w.size == 1) { __syncthreads(); return; }
  __syncthreads();
  const unsigned long long seq = *w.seq + 1;
  const int par = (int)(seq & 1);
  const int t = threadIdx.x;
  if (t < w.size) {
    WinHeader* h = reinterpret_cast<WinHeader*>(w.peer_win[t]);
    for (int i = 0; i < n; ++i) st_peer(&h->mbox[par][w.rank][i], vals[i]);
    __threadfence_system();
    st_release_sys(&h->mflag[par][w.rank], seq);
  }
  WinHeader* me = reinterpret_cast<WinHeader*>(w.peer_win[w.rank]);
  if (t < w.size) {
    wait_ge_sys(&me->mflag[par][t], seq, w.err, w.timeout);
  }
  __syncthreads();
  if (t < n) {
    double acc = 0.0;
    for (int r = 0; r < w.size; ++r) acc += ld_peer_written(&me->mbox[par][r][t]);
    vals[t] = acc;
  }
  __syncthreads();
  if (t == 0) *w.seq = seq;
}

}  // namespace libp_b200
