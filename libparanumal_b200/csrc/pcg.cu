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
//                         p.Ap is produced by the Ax kernel itself (element-local u.A_e u), reductions
//                         are deterministic two-level sums + NCCL all-reduce, and the host only reads
//                         the iteration counter back every `check_every` iterations (a converged solve
//                         turns the remaining queued kernels into no-ops through a device flag).
#include <cmath>

#include "elliptic.hpp"
#include "linalg.hpp"

using namespace libp_b200;

namespace {

constexpr int kBlock = 256;

struct PcgScalars {
  double rdotz1, rdotz2, alpha, beta, pAp, rdotr, TOL, zdotAp;
  double red[4];  // landing slots for reductions: [0]=rdotr [1]=rdotz [2]=pAp [3]=zdotAp
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

// x += alpha p ; r -= alpha Ap ; partial r.r ; (jacobi) z = invD r ; partial r.z
template <bool kJacobi>
__global__ void __launch_bounds__(kBlock) update_kernel(dlong N, const PcgScalars* __restrict__ sc,
                                                        const double* __restrict__ p, const double* __restrict__ Ap,
                                                        const double* __restrict__ invD, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ z,
                                                        double* __restrict__ partials, int nparts) {
  if (sc->done) return;
  __shared__ double s_w[32];
  const double alpha = sc->alpha;
  double rr = 0.0, rz = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    double rn = r[n];
    x[n] += alpha * p[n];
    rn -= alpha * Ap[n];
    rr += rn * rn;
    r[n] = rn;
    if (kJacobi) {
      const double zn = invD[n] * rn;
      z[n] = zn;
      rz += rn * zn;
    }
  }
  rr = block_sum(rr, s_w);
  if (threadIdx.x == 0) partials[blockIdx.x] = rr;
  if (kJacobi) {
    rz = block_sum(rz, s_w);
    if (threadIdx.x == 0) partials[nparts + blockIdx.x] = rz;
  }
}

// z = invD r ; partial r.z   (first iteration, Jacobi)   or just partial x.y
__global__ void __launch_bounds__(kBlock) jacobi_dot_kernel(dlong N, const PcgScalars* __restrict__ sc,
                                                            const double* __restrict__ invD,
                                                            const double* __restrict__ r, double* __restrict__ z,
                                                            double* __restrict__ partials) {
  if (sc->done) return;
  __shared__ double s_w[32];
  double rz = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double rn = r[n], zn = invD[n] * rn;
    z[n] = zn;
    rz += rn * zn;
  }
  rz = block_sum(rz, s_w);
  if (threadIdx.x == 0) partials[blockIdx.x] = rz;
}
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

// p = z + beta p
__global__ void __launch_bounds__(kBlock) pupdate_kernel(dlong N, const PcgScalars* __restrict__ sc,
                                                         const double* __restrict__ z, double* __restrict__ p) {
  if (sc->done) return;
  const double beta = sc->beta;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) p[n] = z[n] + beta * p[n];
}

// sum `n` partials (several segments) into sc->red[slot...]; one block, fixed order
__global__ void __launch_bounds__(1024) finish_partials_kernel(PcgScalars* sc, const double* __restrict__ partials,
                                                               int n0, int slot0, int n1, int slot1) {
  if (sc->done) return;
  __shared__ double s_w[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n0; i += blockDim.x) v += partials[i];
  v = block_sum(v, s_w);
  if (threadIdx.x == 0) sc->red[slot0] = v;
  if (n1 > 0) {
    double u = 0.0;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) u += partials[n0 + i];
    u = block_sum(u, s_w);
    if (threadIdx.x == 0) sc->red[slot1] = u;
  }
}

// scalar recurrences (one thread).  stage 0: after r.z (+z.Ap) -> beta ; stage 1: after p.Ap -> alpha ;
// stage 2: after r.r -> convergence test + iteration count
__global__ void scalars_kernel(PcgScalars* sc, int stage) {
  if (sc->done) return;
  if (stage == 0) {
    sc->rdotz2 = sc->rdotz1;
    sc->rdotz1 = sc->red[1];
    if (sc->flexible) sc->beta = (sc->iter == 0) ? 0.0 : -sc->alpha * sc->red[3] / sc->rdotz2;
    else sc->beta = (sc->iter == 0) ? 0.0 : sc->rdotz1 / sc->rdotz2;
  } else if (stage == 1) {
    sc->pAp = sc->red[2];
    sc->alpha = sc->rdotz1 / sc->pAp;
  } else {
    sc->rdotr = sc->red[0];
    sc->iter += 1;
    if (sc->rdotr <= sc->TOL || sc->iter >= sc->maxit) sc->done = 1;
  }
}

__global__ void hist_kernel(const PcgScalars* sc, double* hist, int maxhist) {
  // records sqrt(rdotr) after the iteration that just completed (no-op once converged earlier)
  const int it = sc->iter;
  if (it >= 0 && it < maxhist) hist[it] = sqrt(sc->rdotr);
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
  ~libp_pcg_s() { if (h_sc) cudaFreeHost(h_sc); }
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
  s->partials.alloc((size_t)2 * kRedMaxBlocks);
  s->sc.alloc(1);
  CUDA_CHECK(cudaMallocHost(&s->h_sc, sizeof(PcgScalars)));
  const char* ce = getenv("LIBP_PCG_CHECK_EVERY");
  if (ce && atoi(ce) > 0) s->check_every = atoi(ce);
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
      CUDA_CHECK(cudaMemcpyAsync(pcg->sc.p, &hs, sizeof(PcgScalars), cudaMemcpyHostToDevice, s));
      const int nb = std::min(vgrid(N), kRedMaxBlocks);
      update_kernel<false><<<nb, kBlock, 0, s>>>(N, pcg->sc.p, pcg->p.p, pcg->Ap.p, nullptr, x, r, nullptr,
                                                 pcg->partials.p, nb);
      finish_partials_kernel<<<1, 1024, 0, s>>>(pcg->sc.p, pcg->partials.p, nb, 0, 0, 0);
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
  if (M->kind == 1 && M->allNeumann)
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

  // z = M r ; r.z for the first iteration
  auto precon_and_rz = [&]() {
    if (jacobi) {
      jacobi_dot_kernel<<<nb, kBlock, 0, s>>>(N, sc, M->invDiag.p, r, pcg->z.p, parts);
    } else {
      M->apply(r, pcg->z.p, s);
      dot_kernel<<<nb, kBlock, 0, s>>>(N, sc, r, pcg->z.p, parts);
    }
    finish_partials_kernel<<<1, 1024, 0, s>>>(sc, parts, nb, 1, 0, 0);
    if (multi) comm->allreduce_sum_dev(sc->red + 1, 1, s);
  };
  precon_and_rz();
  int queued = 0, iter = 0;
  while (true) {
    // beta (needs r.z, and z.Ap when flexible)
    if (pcg->flexible) {
      dot_kernel<<<nb, kBlock, 0, s>>>(N, sc, pcg->z.p, pcg->Ap.p, parts);
      finish_partials_kernel<<<1, 1024, 0, s>>>(sc, parts, nb, 3, 0, 0);
      if (multi) comm->allreduce_sum_dev(sc->red + 3, 1, s);
    }
    scalars_kernel<<<1, 1, 0, s>>>(sc, 0);
    pupdate_kernel<<<vgrid(N), kBlock, 0, s>>>(N, sc, pcg->z.p, pcg->p.p);
    // Ap = A p with p.Ap partials from the Ax kernel
    A->apply(pcg->p.p, pcg->Ap.p, true, done, s);
    finish_partials_kernel<<<1, 1024, 0, s>>>(sc, A->dotPartials.p, A->nDotPartials, 2, 0, 0);
    if (multi) comm->allreduce_sum_dev(sc->red + 2, 1, s);
    scalars_kernel<<<1, 1, 0, s>>>(sc, 1);
    // x, r update + r.r (+ z, r.z for Jacobi)
    if (jacobi) {
      update_kernel<true><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, M->invDiag.p, x, r, pcg->z.p, parts, nb);
      finish_partials_kernel<<<1, 1024, 0, s>>>(sc, parts, nb, 0, nb, 1);
      if (multi) comm->allreduce_sum_dev(sc->red, 2, s);  // r.r and r.z travel together
    } else {
      update_kernel<false><<<nb, kBlock, 0, s>>>(N, sc, pcg->p.p, pcg->Ap.p, nullptr, x, r, nullptr, parts, nb);
      finish_partials_kernel<<<1, 1024, 0, s>>>(sc, parts, nb, 0, 0, 0);
      if (multi) comm->allreduce_sum_dev(sc->red, 1, s);
    }
    scalars_kernel<<<1, 1, 0, s>>>(sc, 2);
    hist_kernel<<<1, 1, 0, s>>>(sc, pcg->d_hist.p, maxit + 2);
    // non-Jacobi: z = M r ; r.z for the next iteration (the dot kernels no-op once converged);
    // Jacobi: r.z was produced and reduced together with r.r above
    if (!jacobi) precon_and_rz();
    CUDA_CHECK(cudaGetLastError());
    queued++;
    iter++;
    if (queued >= pcg->check_every || iter >= maxit) {
      CUDA_CHECK(cudaMemcpyAsync(pcg->h_sc, sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      queued = 0;
      if (pcg->h_sc->done) break;
    }
  }
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
