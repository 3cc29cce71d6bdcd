// elliptic_t::Operator, continuous-Galerkin branch (solvers/elliptic/src/ellipticOperator.cpp:31-106):
//   Aq = Z^T A_L Z q on gathered vectors [NlocalT local | NhaloP owned-shared | Nhalo received].
// Same element-list split as the reference so that the halo exchange of q overlaps the interior
// elements and the cross-rank combine of the shared rows overlaps the second half:
//   halo(q) start | Ax(local[0:n/2]) | halo finish | Ax(global) | combine start | Ax(local[n/2:]) | finish
// mode 0 keeps the reference data flow (AqL scratch + ogs gather, left-to-right row sums);
// mode 1 fuses the gather into the Ax epilogue (FP64 reductions straight into Aq).
#include "elliptic.hpp"

#include <algorithm>
#include <cstdlib>
#include <vector>

#ifndef LIBP_AX_ZERO_AHEAD_DEFAULT
#define LIBP_AX_ZERO_AHEAD_DEFAULT false  // in-kernel zero-fill of the fused accumulator; see elliptic.hpp
#endif
#ifndef LIBP_AX_CHAIN_DEFAULT
#define LIBP_AX_CHAIN_DEFAULT -1  // elements per chain of the element-chain kernel (0 = off, -1 = per order); see elliptic.hpp
#endif
#ifndef LIBP_AX_CHUNK_DEFAULT
#define LIBP_AX_CHUNK_DEFAULT 0  // elements per zero-fill piece of the fused operator (0 = off); see elliptic.hpp
#endif

using namespace libp_b200;

namespace {
dlong g_default_chunk = LIBP_AX_CHUNK_DEFAULT;
bool g_default_za = LIBP_AX_ZERO_AHEAD_DEFAULT;
int g_default_chain = LIBP_AX_CHAIN_DEFAULT, g_default_chain_stages = 1;

// largest local gathered id (< limit) touched by each piece of an element list
__global__ void __launch_bounds__(256) piece_max_kernel(const dlong* __restrict__ list, const dlong* __restrict__ G2L,
                                                        int Np, dlong chunk, dlong n, dlong limit, int* __restrict__ mx) {
  const dlong p0 = (dlong)blockIdx.x * chunk;
  const dlong cnt = min(chunk, n - p0);
  int m = -1;
  for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < (size_t)cnt * Np; i += (size_t)gridDim.y * blockDim.x) {
    const dlong e = list[p0 + (dlong)(i / Np)];
    const dlong id = G2L[(size_t)e * Np + (i % Np)];
    if (id < limit && id > m) m = id;
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(&mx[blockIdx.x], m);
}
}  // namespace

void libp_elliptic_s::build_plan(cudaStream_t s) {
  plan.clear();
  plan_built = true;
  if (!chunked()) return;
  libp_ogs_s& ogs = *d.ogsMasked;
  const dlong nL = d.NlocalGatherElements, nG = d.NglobalGatherElements, nL0 = nL / 2;
  const dlong limit = ogs.NlocalT;  // the shared rows behind NlocalT are zero-filled up front (small)
  struct Seg { int phase; const dlong* list; dlong off, n; };
  const Seg segs[3] = {{0, d.localGatherElementList, 0, nL0}, {1, d.globalGatherElementList, 0, nG},
                       {2, d.localGatherElementList, nL0, nL - nL0}};
  dlong hi = 0;
  for (const Seg& sg : segs) {
    if (sg.n <= 0) continue;
    const int np = (int)((sg.n + chunk - 1) / chunk);
    dev_buf<int> mx;
    mx.alloc((size_t)np);
    CUDA_CHECK(cudaMemsetAsync(mx.p, 0xff, sizeof(int) * (size_t)np, s));
    piece_max_kernel<<<dim3(np, 32), 256, 0, s>>>(sg.list + sg.off, d.GlobalToLocal, Np, chunk, sg.n, limit, mx.p);
    CUDA_CHECK(cudaGetLastError());
    std::vector<int> h((size_t)np);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), mx.p, sizeof(int) * (size_t)np, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int k = 0; k < np; ++k) {
      AxPiece pc{sg.phase, sg.off + (dlong)k * chunk, std::min<dlong>(chunk, sg.n - (dlong)k * chunk), hi, hi};
      if (h[(size_t)k] + 1 > hi) hi = h[(size_t)k] + 1;
      pc.z1 = hi;
      plan.push_back(pc);
    }
  }
  tail0 = hi;  // rows behind the running maximum (the shared rows, ids no element touches) are zero-filled up front
}

void libp_elliptic_s::build_zero_ahead(cudaStream_t s) {
  za_built = true;
  if (const char* e = getenv("LIBP_ZA_DELTA")) kZaDelta = std::max(kZaGroup, atoi(e));  // development knob
  const int epb = ax_hex3d_zero_ahead_epb(d.Nq);
  if (epb == 0) { za_on = false; return; }
  libp_ogs_s& ogs = *d.ogsMasked;
  const dlong nL = d.NlocalGatherElements, nG = d.NglobalGatherElements, nL0 = nL / 2;
  const dlong limit = ogs.NlocalT;
  struct Seg { const dlong* list; dlong n; };
  const Seg segs[3] = {{d.localGatherElementList, nL0}, {d.globalGatherElementList, nG},
                       {d.localGatherElementList ? d.localGatherElementList + nL0 : nullptr, nL - nL0}};
  dlong hi = 0;
  for (int k = 0; k < 3; ++k) {
    ZaSeg& z = za_seg[k];
    z.nblocks = 0;
    if (segs[k].n <= 0) continue;
    const int nb = (int)((segs[k].n + epb - 1) / epb);
    dev_buf<int> mx;
    mx.alloc((size_t)nb);
    CUDA_CHECK(cudaMemsetAsync(mx.p, 0xff, sizeof(int) * (size_t)nb, s));
    piece_max_kernel<<<dim3(nb, 1), 64, 0, s>>>(segs[k].list, d.GlobalToLocal, Np, epb, segs[k].n, limit, mx.p);
    CUDA_CHECK(cudaGetLastError());
    std::vector<int> h((size_t)nb);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), mx.p, sizeof(int) * (size_t)nb, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<dlong> zoff((size_t)nb + 1);
    zoff[0] = hi;
    for (int b = 0; b < nb; ++b) {
      if (h[(size_t)b] + 1 > hi) hi = h[(size_t)b] + 1;
      zoff[(size_t)b + 1] = hi;
    }
    z.nblocks = nb;
    z.zoff.upload(zoff);
    z.z_begin = zoff[0];
    z.z_pre_end = zoff[(size_t)std::min(nb, kZaDelta)];
    z.ctr_count = 4 + (size_t)(nb + kZaGroup - 1) / kZaGroup + 1;
    z.ctr.alloc(z.ctr_count);
    CUDA_CHECK(cudaMemsetAsync(z.ctr.p, 0, sizeof(int) * z.ctr_count, s));
  }
  za_tail0 = hi;
}

int libp_elliptic_s::zero_ahead_errors() {
  int bad = 0;
  for (ZaSeg& z : za_seg) {
    if (!z.ctr.p) continue;
    int e = 0;
    CUDA_CHECK(cudaMemcpy(&e, z.ctr.p + 2, sizeof(int), cudaMemcpyDeviceToHost));
    bad += e;
  }
  return bad;
}

void libp_elliptic_s::build_chain_plan(cudaStream_t s) {
  libp_ogs_s& ogs = *d.ogsMasked;
  // two launch segments: every local element (touches no rank-shared row) | the elements on rank boundaries.  The
  // reference cuts the local list in halves around the global one to overlap MPI on ONE stream
  // (ellipticOperator.cpp:40-104); here the global segment and both exchanges run on a side stream beside the local one.
  const AxChainSegDesc sd[3] = {{d.localGatherElementList, 0, d.NlocalGatherElements},
                                {d.globalGatherElementList, 0, d.NglobalGatherElements}, {nullptr, 0, 0}};
  chainPlan.stages = chainStages;
  chainPlan.build(d.Nq, hD, ogs.NlocalT + ogs.NhaloT, ogs.NlocalT, d.GlobalToLocal, chainL, sd, s);
}

libp_elliptic_s::~libp_elliptic_s() {
  if (ipdg) libp_b200::ipdg_data_free(ipdg);
  if (side) cudaStreamDestroy(side);
  if (e_fork) cudaEventDestroy(e_fork);
  if (e_join) cudaEventDestroy(e_join);
}

void libp_elliptic_s::apply(dfloat* q, dfloat* Aq, bool want_dot, const int* doneFlag, cudaStream_t s, bool zeroed) {
  if (ipdg) {
    libp_b200::ipdg_apply(*this, q, Aq, want_dot, doneFlag, s);
    return;
  }
  libp_ogs_s& ogs = *d.ogsMasked;
  const dlong nL = d.NlocalGatherElements, nG = d.NglobalGatherElements;
  const dlong nL0 = nL / 2, nL1 = (nL + 1) / 2;
  const bool fused = d.mode == 1;
  dfloat* out = fused ? Aq : AqL.p;
  dfloat* dp = want_dot ? dotPartials.p : nullptr;
  int doff = 0;
  if (chain_on(Aq)) {
    // element-chain kernel: only the sectors that are not chain-private are zero-filled (`zeroed` callers did that
    // themselves with chain_zero_mask(), or zero-filled everything)
    if (!chainPlan.built) build_chain_plan(s);
    if (tev) CUDA_CHECK(cudaEventRecord(tev[0], s));
    if (!zeroed) chainPlan.zero_fill(Aq, doneFlag, s);
    if (tev) CUDA_CHECK(cudaEventRecord(tev[1], s));
    const int nb0 = ax_chain_blocks(d.Nq, nL, chainPlan.L);
    auto run = [&](int k, int off, cudaStream_t st) {
      return chainPlan.launch(k, d.GlobalToLocal, d.wJ, d.ggeo, d.lambda, q, Aq, dp ? dp + off : nullptr, doneFlag, st);
    };
    int nb1 = 0;
    if (ogs.comm->size > 1) {
      // side stream: halo of q -> boundary elements -> cross-rank combine, concurrent with the local elements
      if (!side) {
        int least = 0, greatest = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUDA_CHECK(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, greatest));
        CUDA_CHECK(cudaEventCreateWithFlags(&e_fork, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&e_join, cudaEventDisableTiming));
      }
      CUDA_CHECK(cudaEventRecord(e_fork, s));
      CUDA_CHECK(cudaStreamWaitEvent(side, e_fork, 0));
      halo_start_f64(ogs, q, side);
      halo_finish_f64(ogs, q, side);
      nb1 = run(1, nb0, side);
      halo_combine_start_f64(ogs, Aq, side);
      halo_combine_finish_f64(ogs, Aq, side);
      CUDA_CHECK(cudaEventRecord(e_join, side));
    }
    const int got0 = run(0, 0, s);
    (void)got0;
    if (ogs.comm->size > 1) CUDA_CHECK(cudaStreamWaitEvent(s, e_join, 0));
    if (tev) CUDA_CHECK(cudaEventRecord(tev[2], s));
    nDotPartials = nb0 + nb1;
    return;
  }
  auto ax = [&](dlong n, const dlong* list, const ZeroAhead* za = nullptr) {
    if (n <= 0) return;
    int nb = ax_hex3d_launch(d.Nq, fused, axD, symD, n, list, d.GlobalToLocal, d.wJ, d.ggeo, d.lambda, q, out,
                             dp ? dp + doff : nullptr, doneFlag, s, za);
    doff += nb;
  };
  if (zero_ahead() && !za_built) build_zero_ahead(s);
  if (zero_ahead()) {
    // the kernel zero-fills its accumulator itself; the host only covers the shared rows, the rows of the first
    // kZaDelta blocks of every launch, and resets the counters (callers never pre-zero in this mode)
    const dlong total = ogs.NlocalT + ogs.NhaloT;
    if (total > za_tail0) CUDA_CHECK(cudaMemsetAsync(Aq + za_tail0, 0, sizeof(dfloat) * (size_t)(total - za_tail0), s));
    auto run = [&](int k, dlong n, const dlong* list) {
      if (n <= 0) return;
      ZaSeg& z = za_seg[k];
      CUDA_CHECK(cudaMemsetAsync(z.ctr.p, 0, sizeof(int) * z.ctr_count, s));
      if (z.z_pre_end > z.z_begin)
        CUDA_CHECK(cudaMemsetAsync(Aq + z.z_begin, 0, sizeof(dfloat) * (size_t)(z.z_pre_end - z.z_begin), s));
      ZeroAhead za;
      za.ctr = z.ctr.p; za.zoff = z.zoff.p; za.nblocks = z.nblocks; za.delta = kZaDelta; za.group = kZaGroup;
      ax(n, list, &za);
    };
    halo_start_f64(ogs, q, s);
    run(0, nL0, d.localGatherElementList);
    halo_finish_f64(ogs, q, s);
    run(1, nG, d.globalGatherElementList);
    halo_combine_start_f64(ogs, Aq, s);
    run(2, nL1, d.localGatherElementList ? d.localGatherElementList + nL0 : nullptr);
    halo_combine_finish_f64(ogs, Aq, s);
    nDotPartials = doff;
    return;
  }
  if (chunked()) {
    // slab-wise zero-fill: see AxPiece.  `zeroed` callers (PCG) skip their own zero-fill when chunked().
    if (!plan_built) build_plan(s);
    const dlong total = ogs.NlocalT + ogs.NhaloT;
    if (!zeroed && total > tail0)
      CUDA_CHECK(cudaMemsetAsync(Aq + tail0, 0, sizeof(dfloat) * (size_t)(total - tail0), s));
    auto run = [&](int phase) {
      for (const AxPiece& pc : plan) {
        if (pc.phase != phase) continue;
        if (!zeroed && pc.z1 > pc.z0)
          CUDA_CHECK(cudaMemsetAsync(Aq + pc.z0, 0, sizeof(dfloat) * (size_t)(pc.z1 - pc.z0), s));
        ax(pc.count, (phase == 1 ? d.globalGatherElementList : d.localGatherElementList) + pc.start);
      }
    };
    halo_start_f64(ogs, q, s);
    run(0);
    halo_finish_f64(ogs, q, s);
    run(1);
    halo_combine_start_f64(ogs, Aq, s);
    run(2);
    halo_combine_finish_f64(ogs, Aq, s);
    nDotPartials = doff;
    return;
  }
  if (tev) CUDA_CHECK(cudaEventRecord(tev[0], s));
  if (fused && !zeroed)
    CUDA_CHECK(cudaMemsetAsync(Aq, 0, sizeof(dfloat) * (size_t)(ogs.NlocalT + ogs.NhaloT), s));
  if (tev) CUDA_CHECK(cudaEventRecord(tev[1], s));
  halo_start_f64(ogs, q, s);
  ax(nL0, d.localGatherElementList);
  halo_finish_f64(ogs, q, s);
  ax(nG, d.globalGatherElementList);
  if (fused) {
    halo_combine_start_f64(ogs, Aq, s);
    ax(nL1, d.localGatherElementList ? d.localGatherElementList + nL0 : nullptr);
    halo_combine_finish_f64(ogs, Aq, s);
  } else {
    ogs_gather_start_f64(ogs, Aq, AqL.p, LIBP_ADD, LIBP_TRANS, s);
    ax(nL1, d.localGatherElementList ? d.localGatherElementList + nL0 : nullptr);
    ogs_gather_finish_f64(ogs, Aq, AqL.p, LIBP_ADD, LIBP_TRANS, s);
  }
  if (tev) CUDA_CHECK(cudaEventRecord(tev[2], s));
  nDotPartials = doff;
}

extern "C" int libp_elliptic_create(const libp_elliptic_desc_t* desc, libp_elliptic_t* op) {
  LIBP_API_BEGIN
  LIBP_CHECK(desc && op, "null argument");
  LIBP_CHECK(desc->Nq >= 2 && desc->Nq <= 9, "Nq must be in [2, 9]");
  LIBP_CHECK(desc->ogsMasked != nullptr, "ogsMasked handle required");
  LIBP_CHECK(desc->mode == 0 || desc->mode == 1, "mode must be 0 or 1");
  LIBP_CHECK(desc->NlocalGatherElements + desc->NglobalGatherElements == desc->Nelements,
             "element lists must cover all elements");
  LIBP_CHECK(desc->GlobalToLocal && desc->wJ && desc->ggeo && desc->D, "null device pointer");
  LIBP_CHECK(desc->NlocalGatherElements == 0 || desc->localGatherElementList, "null local element list");
  LIBP_CHECK(desc->NglobalGatherElements == 0 || desc->globalGatherElementList, "null global element list");
  std::unique_ptr<libp_elliptic_s> e(new libp_elliptic_s());
  e->d = *desc;
  e->Np = desc->Nq * desc->Nq * desc->Nq;
  {
    double hD[81];
    CUDA_CHECK(cudaMemcpy(hD, desc->D, sizeof(double) * desc->Nq * desc->Nq, cudaMemcpyDeviceToHost));
    e->symD = ax_hex3d_D_is_centro_antisymmetric(desc->Nq, hD);
    std::copy(hD, hD + desc->Nq * desc->Nq, e->hD);
    e->axD.set(desc->Nq, hD);
  }
  e->Ndofs = desc->ogsMasked->Ngather;
  e->Nhalo = desc->ogsMasked->NhaloT - desc->ogsMasked->NhaloP;
  if (desc->mode == 0) e->AqL.alloc((size_t)desc->Nelements * e->Np);
  e->chunk = (desc->mode == 1) ? g_default_chunk : 0;
  e->za_on = g_default_za;
  // measured on B200 (profiles/r2_d_degree_sweep_chains.jsonl, r2_c_chain_tune_*): 8 elements per chain up to N = 6,
  // 4 from N = 7 (short chains keep the tail of a launch short; longer ones make more sectors chain-private)
  e->chainL = (desc->mode == 1) ? (g_default_chain >= 0 ? g_default_chain : (desc->Nq >= 8 ? 4 : 8)) : 0;
  e->chainStages = g_default_chain_stages;
  e->alloc_dot_partials();
  *op = e.release();
  LIBP_API_END
}

void libp_elliptic_s::alloc_dot_partials() {
  // one partial per Ax block: every piece of the plan rounds its block count up
  const dlong pieces = chunk > 0 ? d.Nelements / chunk + 4 : 0;
  dotPartials.alloc((size_t)ax_hex3d_blocks(d.Nq, d.NlocalGatherElements / 2) +
                    ax_hex3d_blocks(d.Nq, (d.NlocalGatherElements + 1) / 2) +
                    ax_hex3d_blocks(d.Nq, d.NglobalGatherElements) + 4 + (size_t)pieces);
}

extern "C" int libp_elliptic_set_chunk(libp_elliptic_t op, libp_dlong chunkElements) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && chunkElements >= 0, "bad argument");
  if (op->d.mode != 1) chunkElements = 0;
  op->chunk = chunkElements;
  op->plan_built = false;
  op->plan.clear();
  op->alloc_dot_partials();
  LIBP_API_END
}

extern "C" int libp_elliptic_set_chain(libp_elliptic_t op, int chainElements, int stages) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && chainElements >= 0 && chainElements <= 4096 && (stages >= 1 && stages <= 3), "bad argument");
  op->chainL = (op->d.mode == 1) ? chainElements : 0;
  op->chainStages = stages;
  op->chainPlan.built = false;
  LIBP_API_END
}

extern "C" int libp_elliptic_set_default_chain(int chainElements, int stages) {
  LIBP_API_BEGIN
  LIBP_CHECK(chainElements >= -1 && chainElements <= 4096 && (stages >= 1 && stages <= 3), "bad argument");
  g_default_chain = chainElements;
  g_default_chain_stages = stages;
  LIBP_API_END
}

extern "C" int libp_elliptic_set_trilinear(libp_elliptic_t op, const libp_dfloat* EXYZ, const libp_dfloat* gllz,
                                           const libp_dfloat* gllw) {
  LIBP_API_BEGIN
  LIBP_CHECK(op, "null handle");
  if (EXYZ == nullptr) { op->chainPlan.EXYZ = nullptr; return LIBP_SUCCESS; }
  LIBP_CHECK(gllz && gllw, "GLL nodes and weights (host arrays of Nq entries) are required");
  LIBP_CHECK(op->d.mode == 1 && op->symD, "the trilinear operator needs the fused mode and a GLL derivative matrix");
  LIBP_CHECK(op->chainL > 0, "the trilinear operator runs in the element-chain kernel (libp_elliptic_set_chain)");
  op->chainPlan.EXYZ = EXYZ;
  for (int i = 0; i < op->d.Nq; ++i) { op->chainPlan.gllz[i] = gllz[i]; op->chainPlan.gllw[i] = gllw[i]; }
  LIBP_API_END
}

extern "C" int libp_elliptic_chain_stats(libp_elliptic_t op, libp_dfloat* Aq, long long* stats, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && stats, "null argument");
  for (int i = 0; i < 6; ++i) stats[i] = 0;
  if (!op->chain_on(Aq)) return LIBP_SUCCESS;
  if (!op->chainPlan.built) op->build_chain_plan(as_stream(stream));
  stats[0] = op->chainPlan.L;
  stats[1] = (long long)op->chainPlan.nSectors;
  stats[2] = (long long)op->chainPlan.zeroSectors;
  stats[3] = (long long)op->chainPlan.nPosTotal;
  stats[4] = (long long)op->chainPlan.rawElements;
  stats[5] = op->chainPlan.stages;
  LIBP_API_END
}

extern "C" int libp_elliptic_set_zero_ahead(libp_elliptic_t op, int on) {
  LIBP_API_BEGIN
  LIBP_CHECK(op, "null handle");
  op->za_on = on != 0;
  LIBP_API_END
}

extern "C" int libp_elliptic_set_default_zero_ahead(int on) {
  LIBP_API_BEGIN
  g_default_za = on != 0;
  LIBP_API_END
}

extern "C" int libp_elliptic_zero_ahead_errors(libp_elliptic_t op, int* errors) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && errors, "null argument");
  *errors = op->zero_ahead_errors();
  LIBP_API_END
}

extern "C" int libp_elliptic_set_default_chunk(libp_dlong chunkElements) {
  LIBP_API_BEGIN
  LIBP_CHECK(chunkElements >= 0, "bad argument");
  g_default_chunk = chunkElements;
  LIBP_API_END
}

extern "C" int libp_elliptic_free(libp_elliptic_t op) {
  LIBP_API_BEGIN
  delete op;
  LIBP_API_END
}

// ------------------------------------------------------------------ elliptic_t::Run pre/post steps
// The reference inlines the user's data file (forcing / boundary functions) into these kernels at JIT time; across
// a C ABI the functions arrive evaluated at the nodes.
namespace {
__global__ void __launch_bounds__(256) rhs_forcing_kernel(size_t N, const dfloat* __restrict__ wJ,
                                                          const dfloat* __restrict__ f, dfloat* __restrict__ rhs) {
  for (size_t n = (size_t)blockIdx.x * 256 + threadIdx.x; n < N; n += (size_t)gridDim.x * 256) rhs[n] = wJ[n] * f[n];
}
__global__ void __launch_bounds__(256) rhs_bc_kernel(size_t N, const dfloat* __restrict__ AuD,
                                                     const dfloat* __restrict__ ndq, dfloat* __restrict__ rhs) {
  for (size_t n = (size_t)blockIdx.x * 256 + threadIdx.x; n < N; n += (size_t)gridDim.x * 256)
    rhs[n] += (ndq ? ndq[n] : 0.0) - AuD[n];
}
__global__ void __launch_bounds__(256) add_bc_kernel(size_t N, const int* __restrict__ mapB,
                                                     const dfloat* __restrict__ uD, dfloat* __restrict__ q) {
  for (size_t n = (size_t)blockIdx.x * 256 + threadIdx.x; n < N; n += (size_t)gridDim.x * 256)
    if (mapB[n] == 1) q[n] = uD[n];
}
int ew_grid(size_t N) { return (int)std::min<size_t>((N + 255) / 256, (size_t)sm_count() * 16); }
}  // namespace

extern "C" int libp_elliptic_rhs_forcing_hex3d(libp_dlong Nelements, int Np, const libp_dfloat* wJ, const libp_dfloat* f,
                                               libp_dfloat* rhs, void* stream) {
  LIBP_API_BEGIN
  const size_t N = (size_t)Nelements * Np;
  LIBP_CHECK(N == 0 || (wJ && f && rhs), "null device pointer");
  if (N) rhs_forcing_kernel<<<ew_grid(N), 256, 0, as_stream(stream)>>>(N, wJ, f, rhs);
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_elliptic_rhs_bc_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* wJ, const libp_dfloat* ggeo,
                                          const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* uD,
                                          const libp_dfloat* ndq, libp_dfloat* rhs, void* stream) {
  LIBP_API_BEGIN
  const size_t N = (size_t)Nelements * Nq * Nq * Nq;
  LIBP_CHECK(N == 0 || (wJ && ggeo && D && uD && rhs), "null device pointer");
  if (N == 0) return LIBP_SUCCESS;
  cudaStream_t s = as_stream(stream);
  // scratch for A u_D: stream-ordered allocation, released in stream order (no synchronisation with the host)
  dfloat* AuD = nullptr;
  CUDA_CHECK(cudaMallocAsync(&AuD, sizeof(dfloat) * N, s));
  double hD[81];
  CUDA_CHECK(cudaMemcpyAsync(hD, D, sizeof(double) * Nq * Nq, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));  // D is a setup-time input read back once per call
  AxD dc;
  dc.set(Nq, hD);
  ax_hex3d_launch(Nq, false, dc, ax_hex3d_D_is_centro_antisymmetric(Nq, hD), Nelements, nullptr, nullptr, wJ, ggeo, lambda,
                  uD, AuD, nullptr, nullptr, s);
  rhs_bc_kernel<<<ew_grid(N), 256, 0, s>>>(N, AuD, ndq, rhs);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaFreeAsync(AuD, s));
  LIBP_API_END
}

extern "C" int libp_elliptic_add_bc_hex3d(libp_dlong Nelements, int Np, const int* mapB, const libp_dfloat* uD,
                                          libp_dfloat* q, void* stream) {
  LIBP_API_BEGIN
  const size_t N = (size_t)Nelements * Np;
  LIBP_CHECK(N == 0 || (mapB && uD && q), "null device pointer");
  if (N) add_bc_kernel<<<ew_grid(N), 256, 0, as_stream(stream)>>>(N, mapB, uD, q);
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_mass_matrix_apply_hex3d(libp_dlong Nelements, int Np, const libp_dfloat* wJ, const libp_dfloat* q,
                                            libp_dfloat* Mq, void* stream) {
  LIBP_API_BEGIN
  const size_t N = (size_t)Nelements * Np;
  LIBP_CHECK(N == 0 || (wJ && q && Mq), "null device pointer");
  if (N) rhs_forcing_kernel<<<ew_grid(N), 256, 0, as_stream(stream)>>>(N, wJ, q, Mq);
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_elliptic_operator_timed(libp_elliptic_t op, libp_dfloat* q, libp_dfloat* Aq, void* stream,
                                            double* ms) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && q && Aq && ms, "null argument");
  LIBP_CHECK(!op->chunked() && !op->zero_ahead(), "timed apply: not available with the chunk / zero-ahead knobs");
  cudaEvent_t ev[3];
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  op->tev = ev;
  try {
    op->apply(q, Aq, false, nullptr, as_stream(stream));
  } catch (...) {
    op->tev = nullptr;
    for (auto& e : ev) cudaEventDestroy(e);
    throw;
  }
  op->tev = nullptr;
  CUDA_CHECK(cudaEventSynchronize(ev[2]));
  float a = 0.f, b = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&a, ev[0], ev[1]));
  CUDA_CHECK(cudaEventElapsedTime(&b, ev[1], ev[2]));
  ms[0] = a; ms[1] = b;
  for (auto& e : ev) cudaEventDestroy(e);
  LIBP_API_END
}

extern "C" int libp_elliptic_operator(libp_elliptic_t op, libp_dfloat* q, libp_dfloat* Aq, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && q && Aq, "null argument");
  op->apply(q, Aq, false, nullptr, as_stream(stream));
  LIBP_API_END
}
