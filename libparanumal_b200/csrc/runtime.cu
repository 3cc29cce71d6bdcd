// Runtime: error state, device selection, communicator (host callbacks + NCCL via dlopen).
#include <dlfcn.h>
#include <mutex>

#include "common.hpp"

namespace libp_b200 {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
}  // namespace libp_b200
using namespace libp_b200;

extern "C" const char* libp_last_error(void) { return g_err.c_str(); }
extern "C" const char* libp_b200_version(void) { return "libparanumal_b200 0.1 sm_100a " __DATE__; }

extern "C" int libp_b200_init(int device_id) {
  LIBP_API_BEGIN
  int n = 0;
  CUDA_CHECK(cudaGetDeviceCount(&n));
  LIBP_CHECK(n > 0, "no CUDA device visible; this library has no CPU fallback");
  LIBP_CHECK(device_id >= 0 && device_id < n, "device id out of range");
  CUDA_CHECK(cudaSetDevice(device_id));
  CUDA_CHECK(cudaFree(0));
  LIBP_API_END
}

extern "C" int libp_b200_finish(void* stream) {
  LIBP_API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  LIBP_API_END
}

// ------------------------------------------------------------------ NCCL (loaded lazily)
namespace {
typedef struct { char internal[128]; } ncclUniqueId_;
typedef void* ncclComm_;
enum { ncclInt8_ = 0, ncclFloat64_ = 8 };
enum { ncclSum_ = 0, ncclProd_ = 1, ncclMax_ = 2, ncclMin_ = 3 };
struct nccl_api {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_*) = nullptr;
  int (*CommInitRank)(ncclComm_*, int, ncclUniqueId_, int) = nullptr;
  int (*CommDestroy)(ncclComm_) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
nccl_api& nccl() {
  static nccl_api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.h) break;
    }
    if (!a.h) return;
#define L(sym) a.sym = reinterpret_cast<decltype(a.sym)>(dlsym(a.h, "nccl" #sym))
    L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(AllReduce); L(Send); L(Recv);
    L(GroupStart); L(GroupEnd); L(GetErrorString);
#undef L
  });
  return a;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    const char* s = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
    throw error(std::string("NCCL ") + what + " failed: " + s);
  }
}
}  // namespace

extern "C" int libp_comm_create(int rank, int size, const libp_host_collectives_t* host, libp_comm_t* comm) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm != nullptr, "null output");
  LIBP_CHECK(size >= 1 && rank >= 0 && rank < size, "bad rank/size");
  LIBP_CHECK(size == 1 || host != nullptr, "size>1 needs host collectives for setup");
  auto* c = new libp_comm_s();
  c->rank = rank;
  c->size = size;
  if (host) { c->host = *host; c->has_host = true; }
  *comm = c;
  LIBP_API_END
}

extern "C" int libp_comm_free(libp_comm_t comm) {
  LIBP_API_BEGIN
  if (comm) {
    if (comm->nccl && nccl().CommDestroy) nccl().CommDestroy(comm->nccl);
    if (comm->comm_stream) cudaStreamDestroy(comm->comm_stream);
    if (comm->p2p) {
      cudaDeviceSynchronize();
      for (int r = 0; r < comm->size; ++r)
        if (r != comm->rank && comm->peer_win[r]) cudaIpcCloseMemHandle(comm->peer_win[r]);
      cudaFree(comm->win);
      cudaFree(comm->d_ar_seq);
      if (comm->d_p2p_err) cudaFree(comm->d_p2p_err);
    }
    delete comm;
  }
  LIBP_API_END
}

extern "C" int libp_comm_rank(libp_comm_t comm, int* rank, int* size) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm, "null comm");
  if (rank) *rank = comm->rank;
  if (size) *size = comm->size;
  LIBP_API_END
}

extern "C" int libp_comm_nccl_unique_id(void* uid128) {
  LIBP_API_BEGIN
  LIBP_CHECK(nccl().h && nccl().GetUniqueId, "libnccl.so.2 could not be loaded");
  ncclUniqueId_ id;
  nccl_check(nccl().GetUniqueId(&id), "GetUniqueId");
  memcpy(uid128, &id, 128);
  LIBP_API_END
}

extern "C" int libp_comm_nccl_init(libp_comm_t comm, const void* uid128) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm, "null comm");
  LIBP_CHECK(nccl().h && nccl().CommInitRank, "libnccl.so.2 could not be loaded");
  ncclUniqueId_ id;
  memcpy(&id, uid128, 128);
  ncclComm_ c = nullptr;
  nccl_check(nccl().CommInitRank(&c, comm->size, id, comm->rank), "CommInitRank");
  comm->nccl = c;
  if (!comm->comm_stream) {
    // highest priority: the small pack / exchange kernels must not queue behind a full-GPU Ax launch
    int least = 0, greatest = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CUDA_CHECK(cudaStreamCreateWithPriority(&comm->comm_stream, cudaStreamNonBlocking, greatest));
  }
  LIBP_API_END
}

// ------------------------------------------------------------------ NVLink peer window
size_t libp_comm_s::win_alloc(size_t bytes) {
  const size_t off = (win_used + 255) & ~size_t(255);
  LIBP_CHECK(off + bytes <= win_bytes, "peer window exhausted (raise window_bytes / LIBP_P2P_WINDOW_MB)");
  win_used = off + bytes;
  return off;
}

extern "C" int libp_comm_p2p_init(libp_comm_t comm, size_t window_bytes) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm, "null comm");
  if (comm->size == 1 || comm->p2p) return LIBP_SUCCESS;
  LIBP_CHECK(comm->size <= kWinMaxRanks, "peer window supports at most 64 ranks");
  if (window_bytes == 0) {
    const char* e = getenv("LIBP_P2P_WINDOW_MB");
    window_bytes = (size_t)(e && atoi(e) > 0 ? atoi(e) : 256) << 20;
  }
  LIBP_CHECK(window_bytes >= sizeof(WinHeader) + 4096, "window too small");
  if (!comm->comm_stream) {
    int least = 0, greatest = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CUDA_CHECK(cudaStreamCreateWithPriority(&comm->comm_stream, cudaStreamNonBlocking, greatest));
  }
  CUDA_CHECK(cudaMalloc(&comm->win, window_bytes));
  CUDA_CHECK(cudaMemset(comm->win, 0, window_bytes));
  CUDA_CHECK(cudaDeviceSynchronize());
  comm->win_bytes = window_bytes;
  comm->win_used = sizeof(WinHeader);
  // exchange IPC handles (host all-to-all of 64-byte handles) and map every peer's window
  cudaIpcMemHandle_t mine;
  CUDA_CHECK(cudaIpcGetMemHandle(&mine, comm->win));
  std::vector<cudaIpcMemHandle_t> sendh((size_t)comm->size, mine), recvh((size_t)comm->size);
  comm->alltoall(sendh.data(), recvh.data(), sizeof(cudaIpcMemHandle_t));
  comm->peer_win.assign((size_t)comm->size, nullptr);
  int ok = 1;
  std::string why;
  for (int r = 0; r < comm->size; ++r) {
    if (r == comm->rank) { comm->peer_win[r] = comm->win; continue; }
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, recvh[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { ok = 0; why = cudaGetErrorString(e); cudaGetLastError(); break; }
    comm->peer_win[r] = static_cast<char*>(p);
  }
  // every rank must agree, otherwise some would wait on flags nobody writes
  int64_t agree = ok;
  comm->allreduce_i64(&agree, 1, LIBP_MIN);
  if (!agree) {
    for (int r = 0; r < comm->size; ++r)
      if (r != comm->rank && comm->peer_win[r]) cudaIpcCloseMemHandle(comm->peer_win[r]);
    comm->peer_win.clear();
    cudaFree(comm->win);
    comm->win = nullptr;
    throw error("CUDA IPC peer mapping failed on some rank" + (why.empty() ? std::string() : (": " + why)));
  }
  comm->d_peer_win.upload(comm->peer_win);
  CUDA_CHECK(cudaMalloc(&comm->d_ar_seq, sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemset(comm->d_ar_seq, 0, sizeof(unsigned long long)));
  CUDA_CHECK(cudaMalloc(&comm->d_p2p_err, sizeof(int)));
  CUDA_CHECK(cudaMemset(comm->d_p2p_err, 0, sizeof(int)));
  {
    // bound of every in-kernel wait on a peer (LIBP_P2P_TIMEOUT_MS, default 30 s), in SM cycles
    const char* e = getenv("LIBP_P2P_TIMEOUT_MS");
    const long long ms = (e && atoll(e) > 0) ? atoll(e) : 30000;
    int khz = 0, dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    comm->p2p_timeout_cycles = ms * (long long)(khz > 0 ? khz : 1500000);
  }
  CUDA_CHECK(cudaDeviceSynchronize());
  comm->p2p = true;
  LIBP_API_END
}

int libp_comm_s::p2p_error() const {
  if (!d_p2p_err) return 0;
  int e = 0;
  CUDA_CHECK(cudaMemcpy(&e, d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
  return e;
}

extern "C" int libp_comm_p2p_status(libp_comm_t comm, int* timed_out) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm && timed_out, "null argument");
  *timed_out = comm->p2p_error();
  LIBP_API_END
}

extern "C" int libp_comm_p2p_reset(libp_comm_t comm) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm, "null comm");
  if (comm->d_p2p_err) CUDA_CHECK(cudaMemset(comm->d_p2p_err, 0, sizeof(int)));
  LIBP_API_END
}

extern "C" int libp_comm_p2p_enabled(libp_comm_t comm, int* enabled) {
  LIBP_API_BEGIN
  LIBP_CHECK(comm && enabled, "null argument");
  *enabled = comm->p2p ? 1 : 0;
  LIBP_API_END
}

// ------------------------------------------------------------------ comm methods
void libp_comm_s::alltoall(const void* send, void* recv, size_t bytes) const {
  if (size == 1) { memcpy(recv, send, bytes); return; }
  LIBP_CHECK(has_host && host.alltoall, "no host alltoall");
  LIBP_CHECK(host.alltoall(host.ctx, send, recv, bytes) == 0, "host alltoall failed");
}
void libp_comm_s::alltoallv(const void* send, const int64_t* sc, const int64_t* so, void* recv, const int64_t* rc,
                            const int64_t* ro) const {
  if (size == 1) {
    if (sc[0]) memcpy(static_cast<char*>(recv) + ro[0], static_cast<const char*>(send) + so[0], (size_t)sc[0]);
    return;
  }
  LIBP_CHECK(has_host && host.alltoallv, "no host alltoallv");
  LIBP_CHECK(host.alltoallv(host.ctx, send, sc, so, recv, rc, ro) == 0, "host alltoallv failed");
}
void libp_comm_s::allreduce_i64(int64_t* inout, int n, int op) const {
  if (size == 1) return;
  LIBP_CHECK(has_host && host.allreduce_i64, "no host allreduce");
  LIBP_CHECK(host.allreduce_i64(host.ctx, inout, n, op) == 0, "host allreduce failed");
}
void libp_comm_s::allreduce_f64(double* inout, int n, int op) const {
  if (size == 1) return;
  LIBP_CHECK(has_host && host.allreduce_f64, "no host allreduce");
  LIBP_CHECK(host.allreduce_f64(host.ctx, inout, n, op) == 0, "host allreduce failed");
}
void libp_comm_s::allreduce_dev(double* buf, int n, int op, cudaStream_t s) const {
  if (size == 1) return;
  if (!nccl && has_host && host.allreduce_f64 && op != LIBP_MUL) {
    // no NCCL (e.g. several ranks on one GPU): stage the few doubles through the host collectives
    std::vector<double> h((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), buf, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    allreduce_f64(h.data(), n, op);
    CUDA_CHECK(cudaMemcpyAsync(buf, h.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return;
  }
  LIBP_CHECK(nccl, "communicator has size>1 but NCCL was not initialised (libp_comm_nccl_init)");
  int nop = op == LIBP_ADD ? ncclSum_ : op == LIBP_MUL ? ncclProd_ : op == LIBP_MAX ? ncclMax_ : ncclMin_;
  nccl_check(::nccl().AllReduce(buf, buf, (size_t)n, ncclFloat64_, nop, nccl, s), "AllReduce");
}
void libp_comm_s::allreduce_sum_dev(double* buf, int n, cudaStream_t s) const { allreduce_dev(buf, n, LIBP_ADD, s); }
void libp_comm_s::group_start() const { if (size > 1) nccl_check(::nccl().GroupStart(), "GroupStart"); }
void libp_comm_s::group_end() const { if (size > 1) nccl_check(::nccl().GroupEnd(), "GroupEnd"); }
void libp_comm_s::send(const void* buf, size_t bytes, int peer, cudaStream_t s) const {
  LIBP_CHECK(nccl, "NCCL not initialised");
  nccl_check(::nccl().Send(buf, bytes, ncclInt8_, peer, nccl, s), "Send");
}
void libp_comm_s::recv(void* buf, size_t bytes, int peer, cudaStream_t s) const {
  LIBP_CHECK(nccl, "NCCL not initialised");
  nccl_check(::nccl().Recv(buf, bytes, ncclInt8_, peer, nccl, s), "Recv");
}
