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

}  // namespace libp_b200

// ------------------------------------------------------------------ communicator
struct libp_comm_s {
  int rank = 0, size = 1;
  libp_host_collectives_t host{};
  bool has_host = false;
  void* nccl = nullptr;           // ncclComm_t
  cudaStream_t comm_stream = nullptr;  // side stream for exchanges (the reference's dataStream)
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
