// Device gather / scatter / gather-scatter and the NCCL pairwise halo exchange.
//
// Semantics follow ogs_t (libs/ogs/ogs.cpp:39-488), halo_t (libs/ogs/ogsHalo.cpp:46-143),
// ogsOperator_t (libs/ogs/ogsOperator.cpp:64-479) and ogsPairwise_t::Start/Finish
// (libs/ogs/ogsPairwise.cpp:116-183, the GPU-aware flavour): per gathered row the values are
// combined left to right in colIds order starting from the op's identity, halo rows combine
// the rank's own partial first and then the received partials in ascending source rank.
//
// B200 mapping: rows are short (1..8 entries for hex meshes), so one thread owns one (row, k)
// pair and walks its few columns; row starts and column ids stream coalesced, values are
// sector-sized random reads that mostly hit L2.  The exchange posts one ncclGroup of
// send/recv pairs straight from/into device buffers on a side stream (no host staging).
#include <algorithm>
#include <limits>

#include "ogs.hpp"
#include "p2p.cuh"

using namespace libp_b200;

namespace {

template <typename T> struct OpAdd { __device__ static T init() { return T(0); } __device__ static T apply(T a, T b) { return a + b; } };
template <typename T> struct OpMul { __device__ static T init() { return T(1); } __device__ static T apply(T a, T b) { return a * b; } };
template <typename T> struct OpMax {
  __device__ static T init() { return -lim(); }
  __device__ static T lim();
  __device__ static T apply(T a, T b) { return (b > a) ? b : a; }
};
template <typename T> struct OpMin {
  __device__ static T init() { return OpMax<T>::lim(); }
  __device__ static T apply(T a, T b) { return (b < a) ? b : a; }
};
template <> __device__ float OpMax<float>::lim() { return 3.402823466e+38f; }
template <> __device__ double OpMax<double>::lim() { return 1.7976931348623157e+308; }
template <> __device__ int OpMax<int>::lim() { return 2147483647; }
template <> __device__ long long OpMax<long long>::lim() { return 9223372036854775807LL; }

constexpr int kBlock = 256;

// gv[row*K + k] = OP_{g in row} v[colIds[g]*K + k]
template <typename T, class Op>
__global__ void __launch_bounds__(kBlock) gather_kernel(dlong Nrows, int K, const dlong* __restrict__ rowStarts,
                                                        const dlong* __restrict__ colIds, const T* __restrict__ v,
                                                        T* __restrict__ gv) {
  const size_t total = (size_t)Nrows * K;
  for (size_t t = (size_t)blockIdx.x * kBlock + threadIdx.x; t < total; t += (size_t)gridDim.x * kBlock) {
    const dlong row = (dlong)(t / K);
    const int k = (int)(t - (size_t)row * K);
    const dlong s = rowStarts[row], e = rowStarts[row + 1];
    T val = Op::init();
    for (dlong g = s; g < e; ++g) val = Op::apply(val, v[(size_t)colIds[g] * K + k]);
    gv[t] = val;
  }
}

// v[colIds[g]*K + k] = gv[row*K + k]
template <typename T>
__global__ void __launch_bounds__(kBlock) scatter_kernel(dlong Nrows, int K, const dlong* __restrict__ rowStarts,
                                                         const dlong* __restrict__ colIds, const T* __restrict__ gv,
                                                         T* __restrict__ v) {
  const size_t total = (size_t)Nrows * K;
  for (size_t t = (size_t)blockIdx.x * kBlock + threadIdx.x; t < total; t += (size_t)gridDim.x * kBlock) {
    const dlong row = (dlong)(t / K);
    const int k = (int)(t - (size_t)row * K);
    const T val = gv[t];
    for (dlong g = rowStarts[row]; g < rowStarts[row + 1]; ++g) v[(size_t)colIds[g] * K + k] = val;
  }
}

// in-place gather then scatter with (possibly) different maps
template <typename T, class Op>
__global__ void __launch_bounds__(kBlock) gather_scatter_kernel(dlong Nrows, int K, const dlong* __restrict__ gRowStarts,
                                                                const dlong* __restrict__ gColIds,
                                                                const dlong* __restrict__ sRowStarts,
                                                                const dlong* __restrict__ sColIds, T* v) {
  const size_t total = (size_t)Nrows * K;
  for (size_t t = (size_t)blockIdx.x * kBlock + threadIdx.x; t < total; t += (size_t)gridDim.x * kBlock) {
    const dlong row = (dlong)(t / K);
    const int k = (int)(t - (size_t)row * K);
    T val = Op::init();
    for (dlong g = gRowStarts[row]; g < gRowStarts[row + 1]; ++g) val = Op::apply(val, v[(size_t)gColIds[g] * K + k]);
    for (dlong g = sRowStarts[row]; g < sRowStarts[row + 1]; ++g) v[(size_t)sColIds[g] * K + k] = val;
  }
}

// out[n*K + k] = in[ids[n]*K + k]   (ogsKernels.okl:167-177 "extract")
template <typename T>
__global__ void __launch_bounds__(kBlock) extract_kernel(dlong N, int K, const dlong* __restrict__ ids,
                                                         const T* __restrict__ in, T* __restrict__ out) {
  const size_t total = (size_t)N * K;
  for (size_t t = (size_t)blockIdx.x * kBlock + threadIdx.x; t < total; t += (size_t)gridDim.x * kBlock) {
    const dlong n = (dlong)(t / K);
    const int k = (int)(t - (size_t)n * K);
    out[t] = in[(size_t)ids[n] * K + k];
  }
}

inline int grid_for(size_t total) {
  size_t b = (total + kBlock - 1) / kBlock;
  const size_t cap = (size_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
void op_gather(const OgsOperator& op, T* gv, const T* v, int K, int opk, int trans, cudaStream_t s) {
  const bool useN = (trans == LIBP_NOTRANS);
  const dlong Nrows = useN ? op.NrowsN : op.NrowsT;
  if (Nrows == 0) return;
  const dlong* rs = useN ? op.d_rowStartsN.p : op.d_rowStartsT.p;
  const dlong* ci = useN ? op.d_colIdsN.p : op.d_colIdsT.p;
  const int g = grid_for((size_t)Nrows * K);
  switch (opk) {
    case LIBP_ADD: gather_kernel<T, OpAdd<T>><<<g, kBlock, 0, s>>>(Nrows, K, rs, ci, v, gv); break;
    case LIBP_MUL: gather_kernel<T, OpMul<T>><<<g, kBlock, 0, s>>>(Nrows, K, rs, ci, v, gv); break;
    case LIBP_MAX: gather_kernel<T, OpMax<T>><<<g, kBlock, 0, s>>>(Nrows, K, rs, ci, v, gv); break;
    case LIBP_MIN: gather_kernel<T, OpMin<T>><<<g, kBlock, 0, s>>>(Nrows, K, rs, ci, v, gv); break;
    default: throw error("bad ogs op");
  }
  CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void op_scatter(const OgsOperator& op, T* v, const T* gv, int K, int trans, cudaStream_t s) {
  const bool useN = (trans == LIBP_TRANS);
  const dlong Nrows = useN ? op.NrowsN : op.NrowsT;
  if (Nrows == 0) return;
  const dlong* rs = useN ? op.d_rowStartsN.p : op.d_rowStartsT.p;
  const dlong* ci = useN ? op.d_colIdsN.p : op.d_colIdsT.p;
  scatter_kernel<T><<<grid_for((size_t)Nrows * K), kBlock, 0, s>>>(Nrows, K, rs, ci, gv, v);
  CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void op_gather_scatter(const OgsOperator& op, T* v, int K, int opk, int trans, cudaStream_t s) {
  dlong Nrows;
  const dlong *grs, *gci, *srs, *sci;
  if (trans == LIBP_TRANS) {
    Nrows = op.NrowsN; grs = op.d_rowStartsT.p; gci = op.d_colIdsT.p; srs = op.d_rowStartsN.p; sci = op.d_colIdsN.p;
  } else if (trans == LIBP_SYM) {
    Nrows = op.NrowsT; grs = op.d_rowStartsT.p; gci = op.d_colIdsT.p; srs = op.d_rowStartsT.p; sci = op.d_colIdsT.p;
  } else {
    Nrows = op.NrowsT; grs = op.d_rowStartsN.p; gci = op.d_colIdsN.p; srs = op.d_rowStartsT.p; sci = op.d_colIdsT.p;
  }
  if (Nrows == 0) return;
  const int g = grid_for((size_t)Nrows * K);
  switch (opk) {
    case LIBP_ADD: gather_scatter_kernel<T, OpAdd<T>><<<g, kBlock, 0, s>>>(Nrows, K, grs, gci, srs, sci, v); break;
    case LIBP_MUL: gather_scatter_kernel<T, OpMul<T>><<<g, kBlock, 0, s>>>(Nrows, K, grs, gci, srs, sci, v); break;
    case LIBP_MAX: gather_scatter_kernel<T, OpMax<T>><<<g, kBlock, 0, s>>>(Nrows, K, grs, gci, srs, sci, v); break;
    case LIBP_MIN: gather_scatter_kernel<T, OpMin<T>><<<g, kBlock, 0, s>>>(Nrows, K, grs, gci, srs, sci, v); break;
    default: throw error("bad ogs op");
  }
  CUDA_CHECK(cudaGetLastError());
}

// ---- pairwise exchange over NCCL (buffer layout: [Nhalo own rows | received values]) ----
// Start: on the comm stream (after the producer event) pack the send buffer and post the group.
template <typename T>
void exchange_start(libp_ogs_s& o, T* buf, int K, int trans, cudaStream_t s) {
  const libp_comm_s& c = *o.comm;
  if (c.size == 1) return;
  const ExchangeLists& ex = (trans == LIBP_NOTRANS) ? o.exN : o.exT;
  const dlong Nhalo = o.gatherHalo.NrowsT;
  cudaStream_t cs = c.comm_stream;
  CUDA_CHECK(cudaEventRecord(o.ev_ready, s));
  CUDA_CHECK(cudaStreamWaitEvent(cs, o.ev_ready, 0));
  T* sendBuf = reinterpret_cast<T*>(o.sendBuf.p);
  if (ex.Nsend()) {
    extract_kernel<T><<<grid_for((size_t)ex.Nsend() * K), kBlock, 0, cs>>>(ex.Nsend(), K, ex.d_sendIds.p, buf, sendBuf);
    CUDA_CHECK(cudaGetLastError());
  }
  if (c.nccl == nullptr) {
    // no device transport for generic types / ops (a communicator without NCCL, e.g. several ranks sharing one
    // GPU): stage the exchange through the host collectives.  Setup-time traffic only; the hot path (FP64, k = 1)
    // uses the peer window.
    LIBP_CHECK(c.has_host, "exchange needs NCCL or host collectives");
    const size_t eb = (size_t)K * sizeof(T);
    std::vector<char> hs((size_t)ex.Nsend() * eb), hr((size_t)ex.Nrecv() * eb);
    if (!hs.empty()) CUDA_CHECK(cudaMemcpyAsync(hs.data(), sendBuf, hs.size(), cudaMemcpyDeviceToHost, cs));
    CUDA_CHECK(cudaStreamSynchronize(cs));
    std::vector<int64_t> sc((size_t)c.size, 0), so((size_t)c.size, 0), rc((size_t)c.size, 0), ro((size_t)c.size, 0);
    for (size_t r = 0; r < ex.sendRanks.size(); ++r) {
      sc[(size_t)ex.sendRanks[r]] = (int64_t)ex.sendCounts[r] * (int64_t)eb;
      so[(size_t)ex.sendRanks[r]] = (int64_t)ex.sendOffsets[r] * (int64_t)eb;
    }
    for (size_t r = 0; r < ex.recvRanks.size(); ++r) {
      rc[(size_t)ex.recvRanks[r]] = (int64_t)ex.recvCounts[r] * (int64_t)eb;
      ro[(size_t)ex.recvRanks[r]] = (int64_t)ex.recvOffsets[r] * (int64_t)eb;
    }
    c.alltoallv(hs.data(), sc.data(), so.data(), hr.data(), rc.data(), ro.data());
    if (!hr.empty()) CUDA_CHECK(cudaMemcpyAsync(buf + (size_t)Nhalo * K, hr.data(), hr.size(), cudaMemcpyHostToDevice, cs));
    CUDA_CHECK(cudaStreamSynchronize(cs));
    CUDA_CHECK(cudaEventRecord(o.ev_done, cs));
    return;
  }
  c.group_start();
  for (size_t r = 0; r < ex.recvRanks.size(); ++r)
    c.recv(buf + (size_t)Nhalo * K + (size_t)ex.recvOffsets[r] * K, (size_t)ex.recvCounts[r] * K * sizeof(T),
           ex.recvRanks[r], cs);
  for (size_t r = 0; r < ex.sendRanks.size(); ++r)
    c.send(sendBuf + (size_t)ex.sendOffsets[r] * K, (size_t)ex.sendCounts[r] * K * sizeof(T), ex.sendRanks[r], cs);
  c.group_end();
  CUDA_CHECK(cudaEventRecord(o.ev_done, cs));
}

// Finish: make the caller's stream wait for the exchange, then combine own + received values.
template <typename T>
void exchange_finish(libp_ogs_s& o, T* buf, int K, int opk, int trans, cudaStream_t s) {
  const libp_comm_s& c = *o.comm;
  if (c.size == 1) return;
  const ExchangeLists& ex = (trans == LIBP_NOTRANS) ? o.exN : o.exT;
  CUDA_CHECK(cudaStreamWaitEvent(s, o.ev_done, 0));
  if (ex.Nrecv()) op_gather<T>(o.postmpi, buf, buf, K, opk, trans == LIBP_NOTRANS ? LIBP_NOTRANS : LIBP_TRANS, s);
}

template <typename T>
void ogs_gather_start(libp_ogs_s& o, T* gv, const T* v, int K, int opk, int trans, cudaStream_t s) {
  LIBP_CHECK(o.gather_defined, "Gather operation not well-defined.");
  if (trans == LIBP_TRANS) {
    o.alloc_buffers((size_t)K * sizeof(T));
    T* hb = reinterpret_cast<T*>(o.haloBuf.p);
    op_gather<T>(o.gatherHalo, hb, v, K, opk, LIBP_TRANS, s);
    exchange_start<T>(o, hb, K, LIBP_TRANS, s);
  } else {
    op_gather<T>(o.gatherHalo, gv + (size_t)K * o.NlocalT, v, K, opk, trans, s);
  }
}
template <typename T>
void ogs_gather_finish(libp_ogs_s& o, T* gv, const T* v, int K, int opk, int trans, cudaStream_t s) {
  LIBP_CHECK(o.gather_defined, "Gather operation not well-defined.");
  op_gather<T>(o.gatherLocal, gv, v, K, opk, trans, s);
  if (trans == LIBP_TRANS) {
    T* hb = reinterpret_cast<T*>(o.haloBuf.p);
    exchange_finish<T>(o, hb, K, opk, LIBP_TRANS, s);
    if (o.NhaloP)
      CUDA_CHECK(cudaMemcpyAsync(gv + (size_t)K * o.NlocalT, hb, (size_t)K * o.NhaloP * sizeof(T),
                                 cudaMemcpyDeviceToDevice, s));
  }
}
template <typename T>
void ogs_scatter_start(libp_ogs_s& o, T* v, const T* gv, int K, int trans, cudaStream_t s) {
  LIBP_CHECK(o.gather_defined, "Gather operation not well-defined.");
  (void)v;
  if (trans == LIBP_NOTRANS) {
    o.alloc_buffers((size_t)K * sizeof(T));
    T* hb = reinterpret_cast<T*>(o.haloBuf.p);
    if (o.NhaloP)
      CUDA_CHECK(cudaMemcpyAsync(hb, gv + (size_t)K * o.NlocalT, (size_t)K * o.NhaloP * sizeof(T),
                                 cudaMemcpyDeviceToDevice, s));
    exchange_start<T>(o, hb, K, LIBP_NOTRANS, s);
  }
}
template <typename T>
void ogs_scatter_finish(libp_ogs_s& o, T* v, const T* gv, int K, int trans, cudaStream_t s) {
  LIBP_CHECK(o.gather_defined, "Gather operation not well-defined.");
  op_scatter<T>(o.gatherLocal, v, gv, K, trans, s);
  if (trans == LIBP_NOTRANS) {
    T* hb = reinterpret_cast<T*>(o.haloBuf.p);
    exchange_finish<T>(o, hb, K, LIBP_ADD, LIBP_NOTRANS, s);
    op_scatter<T>(o.gatherHalo, v, hb, K, LIBP_NOTRANS, s);
  } else {
    op_scatter<T>(o.gatherHalo, v, gv + (size_t)K * o.NlocalT, K, trans, s);
  }
}
template <typename T>
void ogs_gs_start(libp_ogs_s& o, T* v, int K, int opk, int trans, cudaStream_t s) {
  o.alloc_buffers((size_t)K * sizeof(T));
  T* hb = reinterpret_cast<T*>(o.haloBuf.p);
  op_gather<T>(o.gatherHalo, hb, v, K, opk, trans, s);
  exchange_start<T>(o, hb, K, trans, s);
}
template <typename T>
void ogs_gs_finish(libp_ogs_s& o, T* v, int K, int opk, int trans, cudaStream_t s) {
  op_gather_scatter<T>(o.gatherLocal, v, K, opk, trans, s);
  T* hb = reinterpret_cast<T*>(o.haloBuf.p);
  exchange_finish<T>(o, hb, K, opk, trans, s);
  op_scatter<T>(o.gatherHalo, v, hb, K, trans, s);
}
template <typename T>
void halo_start(libp_ogs_s& o, T* v, int K, cudaStream_t s) {
  if (o.comm->size == 1) return;
  o.alloc_buffers((size_t)K * sizeof(T));
  T* hb = reinterpret_cast<T*>(o.haloBuf.p);
  if (o.kind == LIBP_HALO) {
    // a halo set up from ids (mesh_t::halo, the trace halo of IPDG; not built from a gathered ogs): the owned copies
    // are collected through gatherHalo and the received ones scattered back through it (ogsHalo.cpp:56-58, 111)
    op_gather<T>(o.gatherHalo, hb, v, K, LIBP_ADD, LIBP_NOTRANS, s);
    exchange_start<T>(o, hb, K, LIBP_NOTRANS, s);
    return;
  }
  if (o.NhaloP)
    CUDA_CHECK(cudaMemcpyAsync(hb, v + (size_t)K * o.NlocalT, (size_t)K * o.NhaloP * sizeof(T),
                               cudaMemcpyDeviceToDevice, s));
  exchange_start<T>(o, hb, K, LIBP_NOTRANS, s);
}
template <typename T>
void halo_finish(libp_ogs_s& o, T* v, int K, cudaStream_t s) {
  if (o.comm->size == 1) return;
  T* hb = reinterpret_cast<T*>(o.haloBuf.p);
  exchange_finish<T>(o, hb, K, LIBP_ADD, LIBP_NOTRANS, s);
  if (o.kind == LIBP_HALO) {
    op_scatter<T>(o.gatherHalo, v, hb, K, LIBP_NOTRANS, s);
    return;
  }
  const dlong Nhalo = o.NhaloT - o.NhaloP;
  if (Nhalo)
    CUDA_CHECK(cudaMemcpyAsync(v + (size_t)K * (o.NlocalT + o.NhaloP), hb + (size_t)K * o.NhaloP,
                               (size_t)K * Nhalo * sizeof(T), cudaMemcpyDeviceToDevice, s));
}

// ---- NVLink peer-window exchange (double, k = 1) ---------------------------------------------------------
// pack + send in one kernel: every packed value is stored straight into the neighbour's receive buffer.
struct P2PPackArgs {
  const double* src;            // halo rows of the local vector (v + NlocalT)
  const dlong* sendIds;
  double* const* dst[2];        // remote addresses per value, per parity
  dlong Nsend;
  int nSendRanks;
  const int* sendRanks;
  unsigned long long* const* peerFlags;
  const unsigned long long* acks;  // my window
  unsigned long long* d_seq;
  unsigned int* d_done;
  int rank, size;
  int* err;
  long long timeout;
};
__global__ void __launch_bounds__(kBlock) p2p_pack_send_kernel(P2PPackArgs a) {
  const unsigned long long seq = *a.d_seq + 1;
  const int par = (int)(seq & 1);
  // the buffer of this parity was last used by exchange seq-2: wait until every destination consumed it
  if (seq > 2 && threadIdx.x < a.nSendRanks)
    wait_ge_sys(&a.acks[par * a.size + a.sendRanks[threadIdx.x]], seq - 2, a.err, a.timeout);
  __syncthreads();
  double* const* dst = a.dst[par];
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < a.Nsend; n += gridDim.x * kBlock)
    st_peer(dst[n], a.src[a.sendIds[n]]);
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(a.d_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < a.nSendRanks) st_release_sys(a.peerFlags[threadIdx.x] + par * a.size + a.rank, seq);
    if (threadIdx.x == 0) { *a.d_done = 0; *a.d_seq = seq; }
  }
}

// wait for the senders' flags, then combine.  kTrans = false: halo rows [rowBegin,rowEnd) take the owner's value
// (halo_t::Exchange); kTrans = true: owned rows add the received partial sums in ascending source rank
// (the post-exchange gather of ogsPairwise.cpp:283-336).  The last block acknowledges to every sharer.
struct P2PUnpackArgs {
  double* out;                  // v + NlocalT
  const double* recv;           // my window, parity 0
  size_t cap;
  const unsigned long long* flags;
  int nRecvRanks;
  const int* recvRanks;
  const dlong* rowStarts;
  const dlong* colIds;
  dlong rowBegin, rowEnd, Nhalo;
  unsigned long long* const* peerAcks;
  int nAckRanks;
  const unsigned long long* d_seq;
  unsigned int* d_done;
  int rank, size;
  int* err;
  long long timeout;
};
template <bool kTrans>
__global__ void __launch_bounds__(kBlock) p2p_wait_unpack_kernel(P2PUnpackArgs a) {
  const unsigned long long seq = *a.d_seq;  // already advanced by this exchange's pack kernel
  const int par = (int)(seq & 1);
  if (threadIdx.x < a.nRecvRanks)
    wait_ge_sys(&a.flags[par * a.size + a.recvRanks[threadIdx.x]], seq, a.err, a.timeout);
  __syncthreads();
  const double* recv = a.recv + (size_t)par * a.cap;
  for (dlong row = a.rowBegin + blockIdx.x * kBlock + threadIdx.x; row < a.rowEnd; row += gridDim.x * kBlock) {
    const dlong s = a.rowStarts[row], e = a.rowStarts[row + 1];
    if (kTrans) {
      double acc = a.out[row];  // own partial first (colIds[s] == row)
      for (dlong g = s + 1; g < e; ++g) acc += ld_peer_written(recv + (a.colIds[g] - a.Nhalo));
      a.out[row] = acc;
    } else {
      if (e > s) a.out[row] = ld_peer_written(recv + (a.colIds[s] - a.Nhalo));
    }
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(a.d_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x < a.nAckRanks) st_release_sys(a.peerAcks[threadIdx.x] + par * a.size + a.rank, seq);
    if (threadIdx.x == 0) *a.d_done = 0;
  }
}

// flavour 0: owners -> sharers (halo of q); flavour 1: all sharers (combine of the partial sums in place)
void p2p_start(libp_ogs_s& o, double* v, int flavour, cudaStream_t s) {
  const libp_comm_s& c = *o.comm;
  P2PExchange& x = o.p2p;
  const ExchangeLists& ex = flavour == 0 ? o.exN : o.exT;
  LIBP_CHECK((int)ex.sendRanks.size() <= kBlock && (int)o.exT.recvRanks.size() <= kBlock, "too many neighbours");
  cudaStream_t cs = c.comm_stream;
  CUDA_CHECK(cudaEventRecord(o.ev_ready, s));
  CUDA_CHECK(cudaStreamWaitEvent(cs, o.ev_ready, 0));
  P2PPackArgs a;
  a.src = v + o.NlocalT;
  a.sendIds = ex.d_sendIds.p;
  a.dst[0] = x.sendDst[flavour][0].p;
  a.dst[1] = x.sendDst[flavour][1].p;
  a.Nsend = ex.Nsend();
  a.nSendRanks = (int)ex.sendRanks.size();
  a.sendRanks = x.sendRanks[flavour].p;
  a.peerFlags = x.peerFlags[flavour].p;
  a.acks = x.acks;
  a.d_seq = x.d_seq;
  a.d_done = x.d_done;
  a.rank = c.rank;
  a.size = c.size;
  a.err = c.d_p2p_err;
  a.timeout = c.p2p_timeout_cycles;
  int g = (int)std::min<size_t>(((size_t)std::max<dlong>(a.Nsend, 1) + kBlock - 1) / kBlock, 64);
  p2p_pack_send_kernel<<<g, kBlock, 0, cs>>>(a);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(o.ev_done, cs));
}
void p2p_finish(libp_ogs_s& o, double* v, int flavour, cudaStream_t s) {
  const libp_comm_s& c = *o.comm;
  P2PExchange& x = o.p2p;
  const ExchangeLists& ex = flavour == 0 ? o.exN : o.exT;
  CUDA_CHECK(cudaStreamWaitEvent(s, o.ev_done, 0));  // pack kernel done => sequence counter advanced
  P2PUnpackArgs a;
  a.out = v + o.NlocalT;
  a.recv = x.recv;
  a.cap = x.cap;
  a.flags = x.flags;
  a.nRecvRanks = (int)ex.recvRanks.size();
  a.recvRanks = x.recvRanks[flavour].p;
  a.rowStarts = flavour == 0 ? o.postmpi.d_rowStartsN.p : o.postmpi.d_rowStartsT.p;
  a.colIds = flavour == 0 ? o.postmpi.d_colIdsN.p : o.postmpi.d_colIdsT.p;
  a.rowBegin = flavour == 0 ? o.NhaloP : 0;
  a.rowEnd = flavour == 0 ? o.NhaloT : o.NhaloP;
  a.Nhalo = o.NhaloT;
  a.peerAcks = x.peerAcks.p;
  a.nAckRanks = x.nAckRanks;
  a.d_seq = x.d_seq;
  a.d_done = x.d_done + 1;
  a.rank = c.rank;
  a.size = c.size;
  a.err = c.d_p2p_err;
  a.timeout = c.p2p_timeout_cycles;
  const dlong nrows = std::max<dlong>(a.rowEnd - a.rowBegin, 1);
  int g = (int)std::min<size_t>(((size_t)nrows + kBlock - 1) / kBlock, 64);
  if (flavour == 0) p2p_wait_unpack_kernel<false><<<g, kBlock, 0, s>>>(a);
  else p2p_wait_unpack_kernel<true><<<g, kBlock, 0, s>>>(a);
  CUDA_CHECK(cudaGetLastError());
}

#define DISPATCH_TYPE(type, CALL)                                             \
  switch (type) {                                                             \
    case LIBP_FLOAT: { typedef float T; CALL; } break;                        \
    case LIBP_DOUBLE: { typedef double T; CALL; } break;                      \
    case LIBP_INT32: { typedef int T; CALL; } break;                          \
    case LIBP_INT64: { typedef long long T; CALL; } break;                    \
    default: throw error("bad ogs type");                                     \
  }

void check(libp_ogs_t o, int k) {
  LIBP_CHECK(o != nullptr, "null ogs handle");
  LIBP_CHECK(k >= 1, "k must be >= 1");
  LIBP_CHECK(o->ev_ready != nullptr, "ogs handle was set up without a CUDA device");
}

}  // namespace

// internal entry points used by elliptic.cu (typed, no dispatch)
namespace libp_b200 {
void ogs_gather_start_f64(libp_ogs_s& o, double* gv, const double* v, int op, int trans, cudaStream_t s) {
  ogs_gather_start<double>(o, gv, v, 1, op, trans, s);
}
void ogs_gather_finish_f64(libp_ogs_s& o, double* gv, const double* v, int op, int trans, cudaStream_t s) {
  ogs_gather_finish<double>(o, gv, v, 1, op, trans, s);
}
void halo_start_f64(libp_ogs_s& o, double* v, cudaStream_t s) {
  if (o.comm->size == 1) return;
  if (o.p2p.enabled) p2p_start(o, v, 0, s);
  else halo_start<double>(o, v, 1, s);
}
void halo_finish_f64(libp_ogs_s& o, double* v, cudaStream_t s) {
  if (o.comm->size == 1) return;
  if (o.p2p.enabled) p2p_finish(o, v, 0, s);
  else halo_finish<double>(o, v, 1, s);
}
// Combine the partial sums of the shared rows, held in place in gv[NlocalT : NlocalT+NhaloT] (fused Ax epilogue),
// across ranks; the owned totals end up in gv[NlocalT : NlocalT+NhaloP].
void halo_combine_start_f64(libp_ogs_s& o, double* gv, cudaStream_t s) {
  if (o.comm->size == 1) return;
  if (o.p2p.enabled) { p2p_start(o, gv, 1, s); return; }
  o.alloc_buffers(sizeof(double));
  CUDA_CHECK(cudaMemcpyAsync(o.haloBuf.p, gv + o.NlocalT, sizeof(double) * (size_t)o.NhaloT, cudaMemcpyDeviceToDevice, s));
  exchange_start<double>(o, reinterpret_cast<double*>(o.haloBuf.p), 1, LIBP_TRANS, s);
}
void halo_combine_finish_f64(libp_ogs_s& o, double* gv, cudaStream_t s) {
  if (o.comm->size == 1) return;
  if (o.p2p.enabled) { p2p_finish(o, gv, 1, s); return; }
  exchange_finish<double>(o, reinterpret_cast<double*>(o.haloBuf.p), 1, LIBP_ADD, LIBP_TRANS, s);
  if (o.NhaloP)
    CUDA_CHECK(cudaMemcpyAsync(gv + o.NlocalT, o.haloBuf.p, sizeof(double) * (size_t)o.NhaloP, cudaMemcpyDeviceToDevice, s));
}
}  // namespace libp_b200

extern "C" {

int libp_ogs_gather_start(libp_ogs_t o, void* gv, const void* v, int k, int type, int op, int trans, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  DISPATCH_TYPE(type, ogs_gather_start<T>(*o, (T*)gv, (const T*)v, k, op, trans, as_stream(stream)));
  LIBP_API_END
}
int libp_ogs_gather_finish(libp_ogs_t o, void* gv, const void* v, int k, int type, int op, int trans, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  DISPATCH_TYPE(type, ogs_gather_finish<T>(*o, (T*)gv, (const T*)v, k, op, trans, as_stream(stream)));
  LIBP_API_END
}
int libp_ogs_gather(libp_ogs_t o, void* gv, const void* v, int k, int type, int op, int trans, void* stream) {
  if (libp_ogs_gather_start(o, gv, v, k, type, op, trans, stream) != LIBP_SUCCESS) return LIBP_ERROR;
  return libp_ogs_gather_finish(o, gv, v, k, type, op, trans, stream);
}
int libp_ogs_scatter_start(libp_ogs_t o, void* v, const void* gv, int k, int type, int trans, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  DISPATCH_TYPE(type, ogs_scatter_start<T>(*o, (T*)v, (const T*)gv, k, trans, as_stream(stream)));
  LIBP_API_END
}
int libp_ogs_scatter_finish(libp_ogs_t o, void* v, const void* gv, int k, int type, int trans, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  DISPATCH_TYPE(type, ogs_scatter_finish<T>(*o, (T*)v, (const T*)gv, k, trans, as_stream(stream)));
  LIBP_API_END
}
int libp_ogs_scatter(libp_ogs_t o, void* v, const void* gv, int k, int type, int trans, void* stream) {
  if (libp_ogs_scatter_start(o, v, gv, k, type, trans, stream) != LIBP_SUCCESS) return LIBP_ERROR;
  return libp_ogs_scatter_finish(o, v, gv, k, type, trans, stream);
}
int libp_ogs_gather_scatter_start(libp_ogs_t o, void* v, int k, int type, int op, int trans, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  DISPATCH_TYPE(type, ogs_gs_start<T>(*o, (T*)v, k, op, trans, as_stream(stream)));
  LIBP_API_END
}
int libp_ogs_gather_scatter_finish(libp_ogs_t o, void* v, int k, int type, int op, int trans, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  DISPATCH_TYPE(type, ogs_gs_finish<T>(*o, (T*)v, k, op, trans, as_stream(stream)));
  LIBP_API_END
}
int libp_ogs_gather_scatter(libp_ogs_t o, void* v, int k, int type, int op, int trans, void* stream) {
  if (libp_ogs_gather_scatter_start(o, v, k, type, op, trans, stream) != LIBP_SUCCESS) return LIBP_ERROR;
  return libp_ogs_gather_scatter_finish(o, v, k, type, op, trans, stream);
}
int libp_halo_exchange_start(libp_ogs_t o, void* v, int k, int type, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  LIBP_CHECK(o->gather_defined, "Gather operation not well-defined.");
  if (type == LIBP_DOUBLE && k == 1 && o->p2p.enabled && o->kind != LIBP_HALO) halo_start_f64(*o, (double*)v, as_stream(stream));
  else DISPATCH_TYPE(type, halo_start<T>(*o, (T*)v, k, as_stream(stream)));
  LIBP_API_END
}
int libp_halo_exchange_finish(libp_ogs_t o, void* v, int k, int type, void* stream) {
  LIBP_API_BEGIN
  check(o, k);
  if (type == LIBP_DOUBLE && k == 1 && o->p2p.enabled && o->kind != LIBP_HALO) halo_finish_f64(*o, (double*)v, as_stream(stream));
  else DISPATCH_TYPE(type, halo_finish<T>(*o, (T*)v, k, as_stream(stream)));
  LIBP_API_END
}
int libp_halo_exchange(libp_ogs_t o, void* v, int k, int type, void* stream) {
  if (libp_halo_exchange_start(o, v, k, type, stream) != LIBP_SUCCESS) return LIBP_ERROR;
  return libp_halo_exchange_finish(o, v, k, type, stream);
}

}  // extern "C"
