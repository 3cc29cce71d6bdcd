// Internal handles of the elliptic operator / preconditioners / PCG.
#pragma once
#include <vector>

#include "ax_chain.hpp"
#include "common.hpp"
#include "ogs.hpp"

struct libp_elliptic_s;
namespace libp_b200 {
int ax_hex3d_launch(int Nq, bool fused, const AxD& dc, bool sym, dlong Nelements, const dlong* elementList,
                    const dlong* G2L, const dfloat* wJ, const dfloat* ggeo, dfloat lambda,
                    const dfloat* q, dfloat* Aq, dfloat* dotPartials, const int* doneFlag, cudaStream_t s,
                    const ZeroAhead* za = nullptr);
int ax_hex3d_zero_ahead_epb(int Nq);
bool ax_hex3d_D_is_centro_antisymmetric(int Nq, const dfloat* D_host);
int ax_hex3d_blocks(int Nq, dlong Nelements);
void ogs_gather_start_f64(libp_ogs_s& o, double* gv, const double* v, int op, int trans, cudaStream_t s);
void ogs_gather_finish_f64(libp_ogs_s& o, double* gv, const double* v, int op, int trans, cudaStream_t s);
void halo_start_f64(libp_ogs_s& o, double* v, cudaStream_t s);
void halo_finish_f64(libp_ogs_s& o, double* v, cudaStream_t s);
void halo_combine_start_f64(libp_ogs_s& o, double* gv, cudaStream_t s);
void halo_combine_finish_f64(libp_ogs_s& o, double* gv, cudaStream_t s);
void multigrid_apply(void* impl, const dfloat* r, dfloat* Mr, cudaStream_t s);
// DISCRETIZATION = IPDG (csrc/ipdg.cu): the handle carries its own data and apply
struct IpdgData;
void ipdg_apply(libp_elliptic_s& op, dfloat* q, dfloat* Aq, bool want_dot, const int* doneFlag, cudaStream_t s);
void ipdg_data_free(IpdgData* p);
}  // namespace libp_b200

struct libp_elliptic_s {
  libp_elliptic_desc_t d{};
  int Np = 0;
  bool symD = false;  // D verified centro-antisymmetric at create time -> even-odd contractions
  dlong Ndofs = 0, Nhalo = 0;
  libp_b200::IpdgData* ipdg = nullptr;     // non-null: interior-penalty DG operator (libp_elliptic_create_ipdg)
  libp_b200::dev_buf<dfloat> AqL;          // mode 0 scratch, Nelements*Np
  libp_b200::dev_buf<dfloat> dotPartials;  // one per Ax block (p.Ap partial sums)
  int nDotPartials = 0;
  // ---- slab-wise zero-fill of the fused accumulator (mode 1).  The element lists are cut into pieces of
  // `chunk` elements; before piece k runs, only the part of Aq that piece k is the first to touch is zero-filled
  // ([z0, z1) = ids below the running maximum of the connectivity), so the zero lines are still in the 126 MB L2
  // when the reductions arrive: the accumulator costs one DRAM write per line instead of write + read + write.
  struct AxPiece { int phase; dlong start, count, z0, z1; };
  std::vector<AxPiece> plan;
  dlong chunk = 0;       // elements per piece, 0 = one zero-fill + three launches (reference split)
  bool plan_built = false;
  dlong tail0 = 0;       // [tail0, NlocalT+NhaloT): rows no piece owns (shared rows), zero-filled up front
  void alloc_dot_partials();
  void build_plan(cudaStream_t s);
  bool chunked() const { return d.mode == 1 && chunk > 0; }
  // ---- zero-ahead (mode 1, GLL D): the Ax kernel zero-fills its accumulator itself just ahead of its reductions
  // (protocol in ax_hex3d.cu).  One plan per launch segment of apply(): zoff = running maximum of the connectivity at
  // block granularity, ctr = ticket / prefix / per-group counters (reset before every launch).
  struct ZaSeg {
    libp_b200::dev_buf<dlong> zoff;
    libp_b200::dev_buf<int> ctr;
    int nblocks = 0;
    dlong z_begin = 0, z_pre_end = 0;  // host memset before the launch: rows of the first `delta` blocks
    size_t ctr_count = 0;
  };
  ZaSeg za_seg[3];
  bool za_on = false, za_built = false;
  dlong za_tail0 = 0;  // rows behind every block's range (shared rows): zero-filled up front
  int kZaDelta = 2048, kZaGroup = 64;  // zero-fill distance / counter granularity in blocks (delta >= group)
  bool zero_ahead() const { return d.mode == 1 && symD && za_on && chunk == 0; }
  void build_zero_ahead(cudaStream_t s);
  int zero_ahead_errors();
  // ---- element-chain kernel (mode 1, GLL D; ax_chain.cu): TMA-staged geometric factors, owner-computes stores for
  // chain-private rows, compressed connectivity.  chainL = elements per chain (0 = off -> ax_hex3d_t_kernel).
  libp_b200::AxChainPlan chainPlan;
  int chainL = 0, chainStages = 2;
  double hD[81] = {0};
  libp_b200::AxD axD;          // host copy of D + even-odd factors (kernel parameter of ax_hex3d_t_kernel)
  bool chain_capable() const { return d.mode == 1 && symD && chainL > 0 && chunk == 0 && !za_on; }
  // the sector classification needs a 32-byte aligned accumulator, the bulk copies 16-byte aligned factors
  bool chain_on(const dfloat* Aq) const {
    return chain_capable() && (reinterpret_cast<uintptr_t>(Aq) & 31) == 0 &&
           (reinterpret_cast<uintptr_t>(d.ggeo) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.wJ) & 15) == 0;
  }
  void build_chain_plan(cudaStream_t s);
  // multi-rank: the boundary elements and both exchanges run on this high-priority side stream
  cudaStream_t side = nullptr;
  cudaEvent_t e_fork = nullptr, e_join = nullptr;
  ~libp_elliptic_s();
  cudaEvent_t* tev = nullptr;  // libp_elliptic_operator_timed: [before zero-fill, after zero-fill, end of apply]
  // zero-fill mask for a caller that folds the zero-fill of Aq into its own pass (nullptr: zero everything)
  const uint32_t* chain_zero_mask(const dfloat* Aq, cudaStream_t s) {
    if (!chain_on(Aq)) return nullptr;
    if (!chainPlan.built) build_chain_plan(s);
    return chainPlan.zmask.p;
  }
  // apply; when dot/doneFlag are given the p.Ap partials are produced and the kernels early-exit
  // zeroed: the caller already zero-filled Aq[0 : NlocalT+NhaloT] (PCG folds it into its p-update pass)
  void apply(dfloat* q, dfloat* Aq, bool want_dot, const int* doneFlag, cudaStream_t s, bool zeroed = false);
};

struct libp_precon_s {
  int kind = 0;  // 0 identity, 1 jacobi, 2 multigrid (impl = libp_multigrid_t, not owned)
  dlong N = 0;
  libp_b200::dev_buf<dfloat> invDiag;
  int allNeumann = 0;
  libp_hlong NglobalDofs = 0;
  libp_comm_t comm = nullptr;
  void* impl = nullptr;  // multigrid hierarchy
  void apply(const dfloat* r, dfloat* Mr, cudaStream_t s);
};
