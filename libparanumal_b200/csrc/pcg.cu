// Preconditioned conjugate gradients, LinearSolver::pcg (libs/linearSolver/linearSolverPCG.cpp:35-171)
// plus the preconditioner applies that sit on the same path (Identity, Jacobi:
// solvers/elliptic/src/ellipticPreconJacobi.cpp:42-51).
//
// Two drivers:
//  * libp_pcg_solve_cb  - reference control flow verbatim, scalars on the host, operators are
//                         callbacks (any un-replaced operator_t still works);
//  * libp_pcg_solve     - native: the same recurrences, but alpha/beta/rdotr live in device memory,
//                         vector updates are fused with their dot products
//                              x += alpha p ; r -= alpha Ap ; r.r ; z = D^-1 r ; r.z     (one pass)
//                              p  = z + beta p                                            (one pass)
//                         p.Ap is produced by the Ax kernel itself (element-local u.A_e u), the SCALAR
//                         reductions are deterministic two-level sums + NCCL all-reduce (the vectors are not:
//                         in operator mode 1 Ap is accumulated with unordered FP64 red.add, so iterates and
//                         iteration counts are reproducible to rounding only; mode 0 keeps the reference's
//                         left-to-right row sums and is bit-reproducible), and the host only reads
//                         the iteration counter back every `check_every` iterations (a converged solve
//                         turns the remaining queued kernels into no-ops through a device flag).
#include <cmath>

#include "elliptic.hpp"
#include "linalg.hpp"
#include "p2p.cuh"

using namespace libp_b200;

namespace {

constexpr int kBlock = 256;

struct PcgScalars {
  double rdotz1, rdotz2, alpha, beta, pAp, rdotr, TOL, zdotAp;
  double red[4];  // landing slots for reductions (NCCL path): [0]=r.r [1]=r.z [2]=z.Ap [3]=p.Ap
  int iter;       // completed iterations
  int done;       // convergence flag: later kernels become no-ops
  int maxit, flexible;
};

inline int vgrid(dlong N) {
  long b = ((long)N + kBlock * 2 - 1) / (kBlock * 2);
  const long cap = (long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ inline double block_sum(double v, double* s_w) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_w[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? s_w[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
}

// x += alpha p ; r -= alpha Ap ; partial r.r ; z = M r and partial r.z when the preconditioner is diagonal
// (kPrecon 1: Jacobi z = invD r, 2: identity z = r, 0: z is produced by a separate precon apply) ;
// partial z.Ap for the flexible variant.  partials = [r.r | r.z | z.Ap], nparts entries each.
template <int kPrecon, bool kFlex>
__global__ void __launch_bounds__(kBlock) update_kernel(dlong N, const PcgScalars* __restrict__ sc,
                                                        const double* __restrict__ p, const double* __restrict__ Ap,
                                                        const double* __restrict__ invD, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ z,
                                                        double* __restrict__ partials, int nparts) {
  if (sc->done) return;
  __shared__ double s_w[32];
  const double alpha = sc->alpha;
  double rr = 0.0, rz = 0.0, zAp = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    double rn = r[n];
    const double Apn = Ap[n];
    x[n] += alpha * p[n];
    rn -= alpha * Apn;
    rr += rn * rn;
    r[n] = rn;
    if (kPrecon != 0) {
      const double zn = (kPrecon == 1) ? invD[n] * rn : rn;
      z[n] = zn;
      rz += rn * zn;
      if (kFlex) zAp += zn * Apn;
    }
  }
  rr = block_sum(rr, s_w);
  if (threadIdx.x == 0) partials[blockIdx.x] = rr;
  if (kPrecon != 0) {
    rz = block_sum(rz, s_w);
    if (threadIdx.x == 0) partials[nparts + blockIdx.x] = rz;
    if (kFlex) {
      zAp = block_sum(zAp, s_w);
      if (threadIdx.x == 0) partials[2 * nparts + blockIdx.x] = zAp;
    }
  }
}

// z = M r (diagonal preconditioners) ; partial r.z   (before the first iteration)
template <int kPrecon>
__global__ void __launch_bounds__(kBlock) precon_dot_kernel(dlong N, const PcgScalars* __restrict__ sc,
                                                            const double* __restrict__ invD,
                                                            const double* __restrict__ r, double* __restrict__ z,
                                                            double* __restrict__ partials) {
  if (sc->done) return;
  __shared__ double s_w[32];
  double rz = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double rn = r[n], zn = (kPrecon == 1) ? invD[n] * rn : rn;
    z[n] = zn;
    rz += rn * zn;
  }
  rz = block_sum(rz, s_w);
  if (threadIdx.x == 0) partials[blockIdx.x] = rz;
}
// partial a.b (general preconditioners: r.z and z.Ap after the precon apply)
__global__ void __launch_bounds__(kBlock) dot_kernel(dlong N, const PcgScalars* __restrict__ sc,
                                                     const double* __restrict__ a, const double* __restrict__ b,
                                                     double* __restrict__ partials) {
  if (sc->done) return;
  __shared__ double s_w[32];
  double v = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) v += a[n] * b[n];
  v = block_sum(v, s_w);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

// p = z + beta p ; optionally zero-fill Ap (the accumulator of the fused Ax epilogue) in the same pass
// zmask (element-chain operator, ax_chain.cu): one bit per 4-row sector, only flagged sectors need the zero-fill
__global__ void __launch_bounds__(kBlock) pupdate_kernel(dlong N, dlong Nzero, const PcgScalars* __restrict__ sc,
                                                         const double* __restrict__ z, double* __restrict__ p,
                                                         double* __restrict__ Ap, const uint32_t* __restrict__ zmask) {
  if (sc->done) return;
  const double beta = sc->beta;
  const dlong M = N > Nzero ? N : Nzero;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < M; n += gridDim.x * kBlock) {
    if (n < N) p[n] = z[n] + beta * p[n];
    if (n < Nzero && (zmask == nullptr || ((zmask[n >> 7] >> ((n >> 2) & 31)) & 1u))) Ap[n] = 0.0;
  }
}

// deterministic single-block sum of n partials: 4 independent accumulators per thread, fixed tree
__device__ inline double block_sum_array(const double* __restrict__ a, int n, double* s_w) {
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  const int T = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * T < n; i += 4 * T) { v0 += a[i]; v1 += a[i + T]; v2 += a[i + 2 * T]; v3 += a[i + 3 * T]; }
  for (; i < n; i += T) v0 += a[i];
  return block_sum((v0 + v1) + (v2 + v3), s_w);
}

// One block per rank: finish the reductions of a stage, all-reduce them across ranks through the NVLink peer
// window (no NCCL launch, bit-identical result on every rank) and run the scalar recurrences of
// linearSolverPCG.cpp:104-150 on the device.
//   kStage 0: partials = [r.z | z.Ap]      -> rdotz1/rdotz2, beta                         (after a precon apply)
//   kStage 1: partials = [p.Ap]            -> alpha                                       (after the Ax)
//   kStage 2: partials = [r.r | r.z | z.Ap]-> rdotr, iteration count, convergence flag, history (+ beta when
//             the update kernel already produced z = M r)
// phase 0 = everything; phase 1 = local sums only (an NCCL all-reduce of sc->red follows); phase 2 = scalars only.
template <int kStage>
__global__ void __launch_bounds__(1024) pcg_stage_kernel(PcgScalars* sc, const double* __restrict__ partials, int n0,
                                                         int n1, int n2, WinAR w, int phase, int with_beta,
                                                         double* __restrict__ hist, int maxhist) {
  if (sc->done) return;
  __shared__ double s_w[32];
  __shared__ double s_v[kWinMaxVals];
  const int nseg = (n2 > 0) ? 3 : (n1 > 0) ? 2 : 1;
  if (phase != 2) {
    const double a0 = block_sum_array(partials, n0, s_w);
    if (threadIdx.x == 0) s_v[0] = a0;
    if (n1 > 0) {
      const double a1 = block_sum_array(partials + n0, n1, s_w);
      if (threadIdx.x == 0) s_v[1] = a1;
    }
    if (n2 > 0) {
      const double a2 = block_sum_array(partials + n0 + n1, n2, s_w);
      if (threadIdx.x == 0) s_v[2] = a2;
    }
    if (phase == 0) win_allreduce_sum(w, s_v, nseg);
    else __syncthreads();
    if (phase == 1) {
      if (threadIdx.x == 0) {
        if (kStage == 0) { sc->red[1] = s_v[0]; sc->red[2] = (n1 > 0) ? s_v[1] : 0.0; }
        if (kStage == 1) sc->red[3] = s_v[0];
        if (kStage == 2) { sc->red[0] = s_v[0]; sc->red[1] = (n1 > 0) ? s_v[1] : 0.0; sc->red[2] = (n2 > 0) ? s_v[2] : 0.0; }
      }
      return;
    }
  }
  if (threadIdx.x != 0) return;
  double v0, v1 = 0.0, v2 = 0.0;
  if (phase == 2) {
    if (kStage == 0) { v0 = sc->red[1]; v1 = sc->red[2]; }
    else if (kStage == 1) v0 = sc->red[3];
    else { v0 = sc->red[0]; v1 = sc->red[1]; v2 = sc->red[2]; }
  } else {
    v0 = s_v[0];
    if (n1 > 0) v1 = s_v[1];
    if (n2 > 0) v2 = s_v[2];
  }
  if (kStage == 1) {
    sc->pAp = v0;
    sc->alpha = sc->rdotz1 / v0;
    return;
  }
  double rz = v0, zAp = v1;
  if (kStage == 2) {
    sc->rdotr = v0;
    const int it = sc->iter + 1;
    sc->iter = it;
    if (it < maxhist) hist[it] = sqrt(v0);
    if (v0 <= sc->TOL || it >= sc->maxit) sc->done = 1;
    if (!with_beta) return;
    rz = v1;
    zAp = v2;
  }
  // beta for the coming iteration (flexible: -alpha (z.Ap) / rdotz_old)
  sc->rdotz2 = sc->rdotz1;
  sc->rdotz1 = rz;
  if (sc->iter == 0) sc->beta = 0.0;
  else sc->beta = sc->flexible ? -sc->alpha * zAp / sc->rdotz2 : rz / sc->rdotz2;
}

}  // namespace

// ------------------------------------------------------------------ preconditioners
void libp_precon_s::apply(const dfloat* r, dfloat* Mr, cudaStream_t s) {
  if (kind == 0) {
    CUDA_CHECK(cudaMemcpyAsync(Mr, r, sizeof(dfloat) * (size_t)N, cudaMemcpyDeviceToDevice, s));
  } else if (kind == 1) {
    LIBP_CHECK(libp_linalg_amxpy(N, 1.0, invDiag.p, r, 0.0, Mr, s) == LIBP_SUCCESS, libp_last_error());
    if (allNeumann) {
      // ZeroMean (solvers/elliptic/src/ellipticZeroMean.cpp): subtract the global mean
      double sum = 0.0;
      LIBP_CHECK(libp_linalg_sum(N, Mr, comm, s, &sum) == LIBP_SUCCESS, libp_last_error());
      LIBP_CHECK(libp_linalg_add(N, -sum / (double)NglobalDofs, Mr, s) == LIBP_SUCCESS, libp_last_error());
    }
  } else if (kind == 2) {
    // MultiGridPrecon::Operator (solvers/elliptic/src/ellipticPreconMultiGrid.cpp:29-37)
    multigrid_apply(impl, r, Mr, s);
    if (allNeumann) {
      double sum = 0.0;
      LIBP_CHECK(libp_linalg_sum(N, Mr, comm, s, &sum) == LIBP_SUCCESS, libp_last_error());
      LIBP_CHECK(libp_linalg_add(N, -sum / (double)NglobalDofs, Mr, s) == LIBP_SUCCESS, libp_last_error());
    }
  } else {
    throw error("unsupported preconditioner kind");
  }
}

extern "C" int libp_precon_identity_create(libp_dlong N, libp_precon_t* precon) {
  LIBP_API_BEGIN
  LIBP_CHECK(precon && N >= 0, "bad argument");
  auto* p = new libp_precon_s();
  p->kind = 0;
  p->N = N;
  *precon = p;
  LIBP_API_END
}

extern "C" int libp_precon_jacobi_create(libp_dlong Ndofs, const libp_dfloat* invDiagA, int allNeumann,
                                         libp_hlong NglobalDofs, libp_comm_t comm, libp_precon_t* precon) {
  LIBP_API_BEGIN
  LIBP_CHECK(precon && Ndofs >= 0 && (Ndofs == 0 || invDiagA), "bad argument");
  std::unique_ptr<libp_precon_s> p(new libp_precon_s());
  p->kind = 1;
  p->N = Ndofs;
  p->allNeumann = allNeumann;
  p->NglobalDofs = NglobalDofs;
  p->comm = comm;
  p->invDiag.alloc((size_t)Ndofs);
  if (Ndofs) CUDA_CHECK(cudaMemcpy(p->invDiag.p, invDiagA, sizeof(dfloat) * (size_t)Ndofs, cudaMemcpyDeviceToDevice));
  *precon = p.release();
  LIBP_API_END
}

extern "C" int libp_precon_apply(libp_precon_t precon, const libp_dfloat* r, libp_dfloat* Mr, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(precon && r && Mr, "null argument");
  precon->apply(r, Mr, as_stream(stream));
  LIBP_API_END
}

extern "C" int libp_precon_free(libp_precon_t precon) {
  LIBP_API_BEGIN
  delete precon;
  LIBP_API_END
}

// ------------------------------------------------------------------ PCG
struct libp_pcg_s {
  dlong N = 0, Nhalo = 0;
  int flexible = 0, stopping = 0;
  libp_comm_t comm = nullptr;
  dev_buf<dfloat> p, z, Ax, Ap;
  dev_buf<double> partials;   // 2*kRedMaxBlocks
  dev_buf<PcgScalars> sc;
  dev_buf<double> d_hist;
  PcgScalars* h_sc = nullptr;  // pinned
  std::vector<double> hist;
  int check_every = 8;
  // one PCG iteration (p update, operator with its exchanges on the side streams, the stage kernels, x/r update)
  // captured once per solve into a CUDA graph and replayed: one launch per iteration instead of ~12
  int use_graph = 1;
  cudaGraphExec_t iter_graph = nullptr;
  ~libp_pcg_s() {
    if (h_sc) cudaFreeHost(h_sc);
    if (iter_graph) cudaGraphExecDestroy(iter_graph);
  }
};

extern "C" int libp_pcg_create(libp_dlong N, libp_dlong Nhalo, int flexible, int stopping, libp_comm_t comm,
                               libp_pcg_t* pcg) {
  LIBP_API_BEGIN
  LIBP_CHECK(pcg && N >= 0 && Nhalo >= 0, "bad argument");
  LIBP_CHECK(stopping == 0 || stopping == 1, "stopping must be 0 (ABS/REL-INITRESID) or 1 (ABS/REL-RHS-2NORM)");
  std::unique_ptr<libp_pcg_s> s(new libp_pcg_s());
  s->N = N; s->Nhalo = Nhalo; s->flexible = flexible; s->stopping = stopping; s->comm = comm;
  const size_t Ntotal = (size_t)N + Nhalo;
  s->p.alloc(Ntotal); s->z.alloc(Ntotal); s->Ax.alloc(Ntotal); s->Ap.alloc(Ntotal);
  for (dev_buf<dfloat>* b : {&s->p, &s->z, &s->Ax, &s->Ap})
    if (Ntotal) CUDA_CHECK(cudaMemset(b->p, 0, sizeof(dfloat) * Ntotal));
  s->partials.alloc((size_t)4 * kRedMaxBlocks);
  s->sc.alloc(1);
  CUDA_CHECK(cudaMallocHost(&s->h_sc, sizeof(PcgScalars)));
  const char* ce = getenv("LIBP_PCG_CHECK_EVERY");
  if (ce && atoi(ce) > 0) s->check_every = atoi(ce);
  const char* ge = getenv("LIBP_PCG_GRAPH");
  if (ge) s->use_graph = atoi(ge) != 0;
  *pcg = s.release();
  LIBP_API_END
}

extern "C" int libp_pcg_free(libp_pcg_t pcg) {
  LIBP_API_BEGIN
  delete pcg;
  LIBP_API_END
}

extern "C" int libp_pcg_residual_history(libp_pcg_t pcg, const libp_dfloat** hist, int* n) {
  LIBP_API_BEGIN
  LIBP_CHECK(pcg && hist && n, "null argument");
  *hist = pcg->hist.data();
  *n = (int)pcg->hist.size();
  LIBP_API_END
}

// Reference control flow, host scalars (linearSolverPCG.cpp:67-151).
extern "C" int libp_pcg_solve_cb(libp_pcg_t pcg, libp_operator_fn A, void* Actx, libp_operator_fn M, void* Mctx,
                                 libp_dfloat* x, libp_dfloat* r, libp_dfloat tol, int maxit, int verbose,
                                 void* stream, int* iters) {
  LIBP_API_BEGIN
  LIBP_CHECK(pcg && A && M && x && r && iters, "null argument");
  const dlong N = pcg->N;
  libp_comm_t comm = pcg->comm;
  const int rank = comm ? comm->rank : 0;
  double rdotz1 = 0, rdotz2 = 0, alpha = 0, beta = 0, pAp = 0, rdotr0 = 0, TOL = 0;
  auto ok = [](int rc) { LIBP_CHECK(rc == LIBP_SUCCESS, libp_last_error()); };
  if (pcg->stopping == 1) {
    double normb;
    ok(libp_linalg_norm2(N, r, comm, stream, &normb));
    TOL = std::max(tol * tol * normb * normb, tol * tol);
  }
  ok(A(Actx, x, pcg->Ax.p, stream));
  ok(libp_linalg_axpy(N, -1.0, pcg->Ax.p, 1.0, r, stream));
  ok(libp_linalg_norm2(N, r, comm, stream, &rdotr0));
  rdotr0 = rdotr0 * rdotr0;
  if (pcg->stopping == 0) TOL = std::max(tol * tol * rdotr0, tol * tol);
  if (verbose && rank == 0) printf("PCG: initial res norm %12.12f \n", sqrt(rdotr0));
  pcg->hist.clear();
  pcg->hist.push_back(sqrt(rdotr0));
  int iter;
  for (iter = 0; iter < maxit; ++iter) {
    if (((iter == 0) && (rdotr0 == 0.0)) || ((iter > 0) && (rdotr0 <= TOL))) break;
    ok(M(Mctx, r, pcg->z.p, stream));
    rdotz2 = rdotz1;
    ok(libp_linalg_inner_prod(N, r, pcg->z.p, comm, stream, &rdotz1));
    if (pcg->flexible) {
      double zdotAp;
      ok(libp_linalg_inner_prod(N, pcg->z.p, pcg->Ap.p, comm, stream, &zdotAp));
      beta = (iter == 0) ? 0.0 : -alpha * zdotAp / rdotz2;
    } else {
      beta = (iter == 0) ? 0.0 : rdotz1 / rdotz2;
    }
    ok(libp_linalg_axpy(N, 1.0, pcg->z.p, beta, pcg->p.p, stream));
    ok(A(Actx, pcg->p.p, pcg->Ap.p, stream));
    ok(libp_linalg_inner_prod(N, pcg->p.p, pcg->Ap.p, comm, stream, &pAp));
    alpha = rdotz1 / pAp;
    // fused x,r update + r.r (UpdatePCG, linearSolverPCG.cpp:153-171)
    {
      cudaStream_t s = as_stream(stream);
      PcgScalars hs{};
      hs.alpha = alpha;
      hs.maxit = maxit + 1;
      CUDA_CHECK(cudaMemcpyAsync(pcg->sc.p, &hs, sizeof(PcgScalars), cudaMemcpyHostToDevice, s));
      const int nb = std::min(vgrid(N), kRedMaxBlocks);
      update_kernel<0, false><<<nb, kBlock, 0, s>>>(N, pcg->sc.p, pcg->p.p, pcg->Ap.p, nullptr, x, r, nullptr,
                                                    pcg->partials.p, nb);
      pcg_stage_kernel<2><<<1, 1024, 0, s>>>(pcg->sc.p, pcg->partials.p, nb, 0, 0, WinAR{}, 1, 0, nullptr, 0);
      CUDA_CHECK(cudaGetLastError());
      if (comm && comm->size > 1) comm->allreduce_sum_dev(pcg->sc.p->red, 1, s);
      CUDA_CHECK(cudaMemcpyAsync(pcg->h_sc, pcg->sc.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      rdotr0 = pcg->h_sc->red[0];
    }
    pcg->hist.push_back(sqrt(rdotr0));
    if (verbose && rank == 0) {
      if (rdotr0 < 0) printf("WARNING CG: rdotr = %17.15lf\n", rdotr0);
      printf("CG: it %d, r norm %12.12le, alpha = %le \n", iter + 1, sqrt(rdotr0), alpha);
    }
  }
  *iters = iter;
  LIBP_API_END
}

static int elliptic_cb(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream) {
  return libp_elliptic_operator(static_cast<libp_elliptic_t>(ctx), in, out, stream);
}
static int precon_cb(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream) {
  return libp_precon_apply(static_cast<libp_precon_t>(ctx), in, out, stream);
}

extern "C" int libp_pcg_solve(libp_pcg_t pcg, libp_elliptic_t A, libp_precon_t M, libp_dfloat* x, libp_dfloat* r,
                              libp_dfloat tol, int maxit, int verbose, void* stream, int* iters) {
  LIBP_API_BEGIN
  LIBP_CHECK(pcg && A && M && x && r && iters, "null argument");
  // all-Neumann Jacobi needs a host-visible mean per apply: use the reference control flow
  if (M->allNeumann)
    return libp_pcg_solve_cb(pcg, elliptic_cb, A, precon_cb, M, x, r, tol, maxit, verbose, stream, iters);
  cudaStream_t s = as_stream(stream);
  const dlong N = pcg->N;
  libp_comm_t comm = pcg->comm;
  const bool multi = comm && comm->size > 1;
  const int rank = comm ? comm->rank : 0;
  const bool jacobi = (M->kind == 1);
  PcgScalars* sc = pcg->sc.p;
  double* parts = pcg->partials.p;
  auto ok = [](int rc) { LIBP_CHECK(rc == LIBP_SUCCESS, libp_last_error()); };

  // ---- r = r - A x ; rdotr0 ; TOL  (host-visible once, like the reference)
  double TOL = 0.0, rdotr0 = 0.0;
  if (pcg->stopping == 1) {
    double normb;
    ok(libp_linalg_norm2(N, r, comm, stream, &normb));
    TOL = std::max(tol * tol * normb * normb, tol * tol);
  }
  A->apply(x, pcg->Ax.p, false, nullptr, s);
  ok(libp_linalg_axpy(N, -1.0, pcg->Ax.p, 1.0, r, stream));
  ok(libp_linalg_norm2(N, r, comm, stream, &rdotr0));
  rdotr0 = rdotr0 * rdotr0;
  if (pcg->stopping == 0) TOL = std::max(tol * tol * rdotr0, tol * tol);
  if (verbose && rank == 0) printf("PCG: initial res norm %12.12f \n", sqrt(rdotr0));
  pcg->hist.clear();
  pcg->hist.push_back(sqrt(rdotr0));
  if (rdotr0 == 0.0 || maxit <= 0) { *iters = 0; return LIBP_SUCCESS; }

  PcgScalars h{};
  h.TOL = TOL; h.rdotr = rdotr0; h.maxit = maxit; h.flexible = pcg->flexible;
  CUDA_CHECK(cudaMemcpyAsync(sc, &h, sizeof(PcgScalars), cudaMemcpyHostToDevice, s));
  if (pcg->d_hist.n < (size_t)maxit + 2) pcg->d_hist.alloc((size_t)maxit + 2);
  CUDA_CHECK(cudaMemsetAsync(pcg->d_hist.p, 0, sizeof(double) * ((size_t)maxit + 2), s));
  const int* done = &sc->done;
  const int nb = std::min(vgrid(N), kRedMaxBlocks);
  const bool flex = pcg->flexible != 0;
  const int precon = jacobi ? 1 : (M->kind == 0 ? 2 : 0);  // diagonal preconditioners ride the update kernel
  // cross-rank reduction of a stage: through the NVLink peer window inside the stage kernel when the
  // communicator has one, otherwise local sums -> NCCL all-reduce -> scalar kernel
  const bool win = multi && comm->p2p;
  WinAR w{};
  if (win) {
    w.rank = comm->rank; w.size = comm->size; w.peer_win = comm->d_peer_win.p; w.seq = comm->d_ar_seq;
    w.err = comm->d_p2p_err; w.timeout = comm->p2p_timeout_cycles;
  }
  auto stage = [&](int st, const double* partials, int n0, int n1, int n2, int with_beta) {
    double* hist = pcg->d_hist.p;
    const int mh = maxit + 2;
#define STAGE(ST, PH) pcg_stage_kernel<ST><<<1, 1024, 0, s>>>(sc, partials, n0, n1, n2, w, PH, with_beta, hist, mh)
    if (!multi || win) {
      if (st == 0) STAGE(0, 0); else if (st == 1) STAGE(1, 0); else STAGE(2, 0);
    } else {
      if (st == 0) STAGE(0, 1); else if (st == 1) STAGE(1, 1); else STAGE(2, 1);
      // red = [r.r, r.z, z.Ap, p.Ap]: every stage all-reduces exactly the slots it wrote
      if (st == 0) comm->allreduce_sum_dev(sc->red + 1, 2, s);
      else if (st == 1) comm->allreduce_sum_dev(sc->red + 3, 1, s);
      else comm->allreduce_sum_dev(sc->red, 3, s);
      if (st == 0) STAGE(0, 2); else if (st == 1) STAGE(1, 2); else STAGE(2, 2);
    }
#undef STAGE
  };
  // with the slab-wise plan the operator zero-fills its accumulator itself, piece by piece (elliptic.hpp)
  const dlong Nzero = (A->d.mode == 1 && !A->chunked() && !A->zero_ahead()) ? (dlong)(A->d.ogsMasked->NlocalT + A->d.ogsMasked->NhaloT) : 0;

  const uint32_t* zmask = (Nzero > 0) ? A->chain_zero_mask(pcg->Ap.p, s) : nullptr;

  // z = M r ; r.z ; beta = 0 for the first iteration
  auto precon_and_rz = [&]() {
    if (precon == 1) precon_dot_kernel<1><<<nb, kBlock, 0, s>>>(N, sc, M->invDiag.p, r, pcg->z.p, parts);
    else if (precon == 2) precon_dot_kernel<2><<<nb, kBlock, 0, s>>>(N, sc, nullptr, r, pcg->z.p, parts);
    else {
      M->apply(r, pcg->z.p, s);
      dot_kernel<<<nb, kBlock, 0, s>>>(N, sc, r, pcg->z.p, parts);
      if (flex) dot_kernel<<<nb, kBlock, 0, s>>>(N, sc, pcg->z.p, pcg->Ap.p, parts + nb);
    }
    stage(0, parts, nb, (precon == 0 && flex) ? nb : 0, 0, 1);
  };
  precon_and_rz();
  int queued = 0, iter = 0;
  // the part of an iteration that needs no host decision (everything for the diagonal preconditioners; up to the
  // convergence test for a general one)
  auto body = [&]() {
    // p = z + beta p (and the zero-fill of Ap for the fused Ax epilogue)
    pupdate_kernel<<<vgrid(std::max(N, Nzero)), kBlock, 0, s>>>(N, Nzero, sc, pcg->z.p, pcg->p.p, pcg->Ap.p, zmask);
    // Ap = A p with p.Ap partials from the Ax kernel ; alpha
    A->apply(pcg->p.p, pcg->Ap.p, true, done, s, /*zeroed=*/Nzero > 0);
    stage(1, A->dotPartials.p, A->nDotPartials, 0, 0, 0);
    // x, r update + r.r (+ z = M r, r.z, z.Ap for diagonal preconditioners) ; convergence ; beta
    if (precon == 1) {
      if (flex) update_kernel<1, true><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, M->invDiag.p, x, r, pcg->z.p, parts, nb);
      else update_kernel<1, false><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, M->invDiag.p, x, r, pcg->z.p, parts, nb);
    } else if (precon == 2) {
      if (flex) update_kernel<2, true><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, nullptr, x, r, pcg->z.p, parts, nb);
      else update_kernel<2, false><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, nullptr, x, r, pcg->z.p, parts, nb);
    } else {
      update_kernel<0, false><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, nullptr, x, r, nullptr, parts, nb);
    }
    CUDA_CHECK(cudaGetLastError());
    if (precon != 0) stage(2, parts, nb, nb, flex ? nb : 0, 1);
    else stage(2, parts, nb, 0, 0, 0);
  };
  // Graph of one iteration.  Only where every kernel of the body is ours (single rank, or the peer-window transport:
  // no NCCL calls inside the capture); the first iteration runs eagerly (it builds lazily created plans / streams).
  bool graph_ok = pcg->use_graph && (!multi || win) && maxit > 2;
  if (pcg->iter_graph) { cudaGraphExecDestroy(pcg->iter_graph); pcg->iter_graph = nullptr; }
  while (true) {
    if (graph_ok && iter >= 1 && pcg->iter_graph == nullptr) {
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        bool thrown = false;
        try { body(); } catch (...) { thrown = true; }
        e = cudaStreamEndCapture(s, &g);
        if (e == cudaSuccess && !thrown && g) e = cudaGraphInstantiate(&pcg->iter_graph, g, 0);
        else if (e == cudaSuccess) e = cudaErrorUnknown;
        if (g) cudaGraphDestroy(g);
      }
      if (e != cudaSuccess) {  // not capturable here: run eagerly
        cudaGetLastError();
        pcg->iter_graph = nullptr;
        graph_ok = false;
      }
    }
    if (pcg->iter_graph) CUDA_CHECK(cudaGraphLaunch(pcg->iter_graph, s));
    else body();
    queued++;
    iter++;
    if (precon != 0) {
      if (queued >= pcg->check_every || iter >= maxit) {
        CUDA_CHECK(cudaMemcpyAsync(pcg->h_sc, sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        queued = 0;
        if (pcg->h_sc->done) break;
      }
    } else {
      // general preconditioner (a multigrid V-cycle, exchanges included): like the reference
      // (linearSolverPCG.cpp:104-110) test convergence BEFORE applying it - one small read-back per iteration,
      // negligible beside a V-cycle, and no preconditioner work after the converged iteration
      CUDA_CHECK(cudaMemcpyAsync(pcg->h_sc, sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (pcg->h_sc->done) break;
      precon_and_rz();  // separate apply, then r.z (z.Ap) and beta
    }
  }
  LIBP_CHECK(!(win && comm->p2p_error()), "a peer-window wait timed out (a peer rank failed or diverged)");
  const int it = pcg->h_sc->iter;
  std::vector<double> hh((size_t)it + 1);
  CUDA_CHECK(cudaMemcpy(hh.data(), pcg->d_hist.p, sizeof(double) * ((size_t)it + 1), cudaMemcpyDeviceToHost));
  for (int i = 1; i <= it; ++i) {
    pcg->hist.push_back(hh[i]);
    if (verbose && rank == 0) printf("CG: it %d, r norm %12.12le\n", i, hh[i]);
  }
  *iters = it;
  LIBP_API_END
}
