// Internal linAlg helpers shared by the PCG / preconditioner code.
#pragma once
#include "common.hpp"

namespace libp_b200 {

constexpr int kRedBlock = 256;
constexpr int kRedMaxBlocks = 1024;  // partial sums per reduction (fixed => deterministic order)

inline int red_blocks(dlong N) {
  long b = ((long)N + (long)kRedBlock * 4 - 1) / ((long)kRedBlock * 4);
  if (b > kRedMaxBlocks) b = kRedMaxBlocks;
  if (b < 1) b = 1;
  return (int)b;
}

// Scratch for reductions: partial sums on the device + a pinned host landing slot.
struct RedScratch {
  dev_buf<double> partials;  // kRedMaxBlocks * 4
  dev_buf<double> result;    // 8 doubles
  double* h_result = nullptr;  // pinned
  cudaEvent_t ev = nullptr;
  void ensure();
  ~RedScratch();
};
RedScratch& red_scratch(cudaStream_t s);

// result[slot] = sum_i x[i]*y[i] (y==nullptr -> x[i]); deterministic two-level reduction, stays on device
void dot_to_device(dlong N, const double* x, const double* y, const double* w, double* d_out, cudaStream_t s);

}  // namespace libp_b200
