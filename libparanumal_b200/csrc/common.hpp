// Shared internals of libparanumal_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/libp_b200.h"

typedef libp_dlong dlong;
typedef libp_hlong hlong;
typedef libp_dfloat dfloat;

namespace libp_b200 {

void set_error(const std::string& msg);

struct error : public std::runtime_error {
  explicit error(const std::string& m) : std::runtime_error(m) {}
};

#define LIBP_CHECK(cond, msg)                                                              \
  do {                                                                                     \
    if (!(cond)) throw ::libp_b200::error(std::string(__func__) + ": " + (msg));           \
  } while (0)

#define CUDA_CHECK(call)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw ::libp_b200::error(std::string(__func__) + ": CUDA error '" +                  \
                               cudaGetErrorString(_e) + "' at " __FILE__ ":" +             \
                               std::to_string(__LINE__));                                  \
  } while (0)

// Wraps a C-ABI body: converts exceptions into LIBP_ERROR + thread-local message.
#define LIBP_API_BEGIN try {
#define LIBP_API_END                                                                       \
    return LIBP_SUCCESS;                                                                   \
  } catch (const std::exception& e) {                                                      \
    ::libp_b200::set_error(e.what());                                                      \
    return LIBP_ERROR;                                                                     \
  } catch (...) {                                                                          \
    ::libp_b200::set_error("unknown exception");                                           \
    return LIBP_ERROR;                                                                     \
  }

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Simple RAII device buffer
template <typename T>
struct dev_buf {
  T* p = nullptr;
  size_t n = 0;
  dev_buf() = default;
  dev_buf(const dev_buf&) = delete;
  dev_buf& operator=(const dev_buf&) = delete;
  ~dev_buf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const T* host, size_t count) {
    alloc(count);
    if (count) CUDA_CHECK(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
  }
  void upload(const std::vector<T>& v) { upload(v.data(), v.size()); }
};

int sm_count();

// Derivative matrix + even-odd factors handed to the Ax kernels by value (see ax_hex3d.cu)
struct AxD {
  double D[81];
  double De[16], Do[16], Dc[4], Dr[4];
  void set(int Nq, const double* D_host);
};

// Zero-ahead arguments of the fused Ax kernel (protocol: ax_hex3d.cu, kZA)
struct ZeroAhead {
  int* ctr = nullptr;           // device: [0] ticket, [1] complete leading groups, [2] error, [3] pad, [4+g] done[g]
  const dlong* zoff = nullptr;  // device, nblocks+1: rows first touched by virtual block b = [zoff[b], zoff[b+1])
  int nblocks = 0, delta = 0, group = 0;
};

// Fixed header at the start of every peer window: mailboxes of the scalar all-reduce.
constexpr int kWinMaxRanks = 64;
constexpr int kWinMaxVals = 8;
struct WinHeader {
  double mbox[2][kWinMaxRanks][kWinMaxVals];   // [parity][source rank][value]
  unsigned long long mflag[2][kWinMaxRanks];   // sequence number of the last all-reduce rank r contributed to
};

// What a kernel needs to all-reduce a few doubles across ranks through the peer windows.
struct WinAR {
  int rank = 0, size = 1;
  char* const* peer_win = nullptr;   // device table
  unsigned long long* seq = nullptr; // device counter (shared by every user of the communicator)
  int* err = nullptr;                // device error word of the communicator (a bounded wait timed out)
  long long timeout = 0;             // SM cycles a wait may last
};

}  // namespace libp_b200

// ------------------------------------------------------------------ communicator
struct libp_comm_s {
  int rank = 0, size = 1;
  libp_host_collectives_t host{};
  bool has_host = false;
  void* nccl = nullptr;           // ncclComm_t
  cudaStream_t comm_stream = nullptr;  // high-priority side stream for exchanges (the reference's dataStream)
  // ---- NVLink peer window (libp_comm_p2p_init): one device allocation per rank, mapped by every peer through
  // CUDA IPC.  Kernels exchange halo values and reduction scalars with plain stores into the peers' windows and
  // release/acquire flags - no NCCL launch, no host involvement.  Layout: [WinHeader | bump-allocated regions].
  bool p2p = false;
  char* win = nullptr;
  size_t win_bytes = 0, win_used = 0;
  std::vector<char*> peer_win;                  // peer_win[r] = rank r's window in my address space (self included)
  libp_b200::dev_buf<char*> d_peer_win;         // the same table on the device
  unsigned long long* d_ar_seq = nullptr;       // device counter of window all-reduces
  // every in-kernel wait on a peer's flag is bounded: on a time-out the kernel sets this word and stops waiting
  // (results of the communicator are then invalid; libp_comm_p2p_status reports it, solves check it)
  int* d_p2p_err = nullptr;
  long long p2p_timeout_cycles = 0;
  int p2p_error() const;                        // current value of the error word (synchronises the device)
  size_t win_alloc(size_t bytes);               // offset of a new 256-byte aligned region of my window
  // host collectives with size==1 shortcuts
  void alltoall(const void* send, void* recv, size_t bytes_per_rank) const;
  void alltoallv(const void* send, const int64_t* sc, const int64_t* so, void* recv, const int64_t* rc,
                 const int64_t* ro) const;
  void allreduce_i64(int64_t* inout, int n, int op) const;
  void allreduce_f64(double* inout, int n, int op) const;
  // device collectives (NCCL); no-ops with size==1
  void allreduce_sum_dev(double* buf, int n, cudaStream_t s) const;
  void allreduce_dev(double* buf, int n, int op, cudaStream_t s) const;
  void group_start() const;
  void group_end() const;
  void send(const void* buf, size_t bytes, int peer, cudaStream_t s) const;
  void recv(void* buf, size_t bytes, int peer, cudaStream_t s) const;
};
