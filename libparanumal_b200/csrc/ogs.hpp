// Internal ogs handle layout (host maps + device copies + NCCL pairwise exchange state).
#pragma once
#include "common.hpp"

namespace libp_b200 {

// CSR-without-values gather operator: the Z matrix of the reference
// (include/ogs/ogsOperator.hpp:38-70).  N = owner-copy ("NoTrans") map, T = all-copies map.
struct OgsOperator {
  dlong Ncols = 0, NrowsN = 0, NrowsT = 0;
  std::vector<dlong> rowStartsN, rowStartsT, colIdsN, colIdsT;
  dev_buf<dlong> d_rowStartsN, d_rowStartsT, d_colIdsN, d_colIdsT;
  dlong nnzN() const { return rowStartsN.empty() ? 0 : rowStartsN.back(); }
  dlong nnzT() const { return rowStartsT.empty() ? 0 : rowStartsT.back(); }
  void to_device();
};

// Pairwise exchange lists (libs/ogs/ogsPairwise.cpp:194-415), one set per map flavour.
struct ExchangeLists {
  std::vector<dlong> sendIds;
  dev_buf<dlong> d_sendIds;
  std::vector<int> sendRanks, sendCounts, sendOffsets;  // offsets has NranksSend+1 entries
  std::vector<int> recvRanks, recvCounts, recvOffsets;
  dlong Nsend() const { return (dlong)sendIds.size(); }
  dlong Nrecv() const { return recvOffsets.empty() ? 0 : recvOffsets.back(); }
};

// NVLink peer-window exchange state of one ogs handle (double values, k = 1: the elliptic hot path).
// Senders store packed values straight into the receiver's window buffer `recv[parity]` and then publish the
// exchange sequence number in the receiver's `flags[parity][sender rank]`; receivers acknowledge consumption in
// the sender's `acks[parity][receiver rank]` so a buffer is never overwritten before it was read.
struct P2PExchange {
  bool enabled = false;
  double* recv = nullptr;                 // my window: 2 x cap doubles
  size_t cap = 0;
  unsigned long long* flags = nullptr;    // my window: [2][size]
  unsigned long long* acks = nullptr;     // my window: [2][size]
  unsigned long long* d_seq = nullptr;    // exchange counter (device, private)
  unsigned int* d_done = nullptr;         // [2] block counters (pack, unpack)
  // per flavour: 0 = N (owner -> sharers, halo exchange), 1 = T (all sharers, gather combine)
  dev_buf<double*> sendDst[2][2];         // [flavour][parity][Nsend] remote addresses
  dev_buf<int> sendRanks[2], recvRanks[2];
  dev_buf<unsigned long long*> peerFlags[2];  // [flavour][NranksSend]: peer's flags base
  dev_buf<unsigned long long*> peerAcks;      // [NranksRecv of flavour T]: peer's acks base (acks go to every sharer)
  int nAckRanks = 0;
};

}  // namespace libp_b200

struct libp_ogs_s {
  libp_comm_t comm = nullptr;
  dlong N = 0, Ngather = 0, NlocalT = 0, NlocalP = 0, NhaloT = 0, NhaloP = 0;
  hlong NgatherGlobal = 0;
  int kind = LIBP_SIGNED;
  bool unique = false, gather_defined = false;
  libp_b200::OgsOperator gatherLocal, gatherHalo, postmpi;
  libp_b200::ExchangeLists exN, exT;
  // device workspaces sized on demand: haloBuf holds postmpi.nnzT*k values, sendBuf NsendT*k
  libp_b200::dev_buf<char> haloBuf, sendBuf;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  libp_b200::P2PExchange p2p;
  void alloc_buffers(size_t bytes_per_node);
  ~libp_ogs_s();
};
