// Plan + launcher of the element-chain fused operator kernel (ax_chain.cu).
#pragma once
#include "common.hpp"

namespace libp_b200 {

struct AxChainSegDesc {
  const dlong* list;   // device element list of the launch segment (nullptr = identity)
  dlong first, count;  // entries [first, first + count) of the list
};

// Derived from GlobalToLocal once per operator handle (device kernels): the chain-ordered element sequence of the three
// launch segments of elliptic_t::Operator, compressed connectivity, store/reduce flags and the zero-fill sector mask.
struct AxChainPlan {
  struct Seg { size_t first = 0; dlong count = 0; int nChains = 0; };
  Seg seg[3];
  int Nq = 0, L = 0, stages = 2;
  dlong nRows = 0;          // rows of the accumulator covered by the mask: NlocalT + NhaloT
  size_t nSectors = 0, nPosTotal = 0;
  size_t zeroSectors = 0;   // statistics: sectors that still need the zero-fill
  size_t rawElements = 0;   //             elements whose connectivity did not compress
  bool built = false;
  dev_buf<int> elem, hdr, cid;
  dev_buf<uint16_t> flags;
  dev_buf<uint32_t> zmask;
  alignas(8) unsigned char eo[320];  // even-odd factors of D, passed to the kernel by value
  // ELEMENT MAP = TRILINEAR: element vertices (device, [E][3][8]) + GLL nodes / weights; nullptr = stored factors
  const dfloat* EXYZ = nullptr;
  double gllz[9] = {0}, gllw[9] = {0};
  void build(int Nq, const dfloat* D_host, dlong nRows, dlong NlocalT, const dlong* G2L, int L,
             const AxChainSegDesc (&sd)[3], cudaStream_t s);
  // zero-fills the sectors of Aq[0 : nRows] that are not chain-private (Aq must be 32-byte aligned)
  void zero_fill(dfloat* Aq, const int* doneFlag, cudaStream_t s) const;
  // runs segment k; returns the number of blocks (= dot partials written when dotPartials != nullptr)
  int launch(int k, const dlong* G2L, const dfloat* wJ, const dfloat* ggeo, dfloat lambda, const dfloat* q, dfloat* Aq,
             dfloat* dotPartials, const int* doneFlag, cudaStream_t s) const;
};

int ax_chain_blocks(int Nq, dlong count, int L);

}  // namespace libp_b200
