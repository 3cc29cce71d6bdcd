// Non-blocking preconditioned conjugate gradients (Gropp's variant), LinearSolver::nbpcg
// (libs/linearSolver/linearSolverNBPCG.cpp:35-229, kernels okl/linearSolverUpdateNBPCG.okl).
//
// Same recurrences and the same overlap structure as the reference: every reduction is issued (fused update kernel ->
// deterministic block sums -> device all-reduce -> asynchronous copy into pinned memory) and the next operator /
// preconditioner apply is queued behind it before the host waits for the scalars, so the wait is covered by
// device work.  Operators are callbacks (libp_operator_fn), so any operator_t / precon_t works; the wrappers
// libp_nbpcg_solve take the native elliptic / preconditioner handles.
#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include "elliptic.hpp"
#include "linalg.hpp"

using namespace libp_b200;

namespace {
constexpr int kBlock = 256;
constexpr int kMaxBlocks = 512;  // NBPCG_BLOCKSIZE partial sums (linearSolverNBPCG.cpp:32)

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kBlock / 32; ++w) tot += s_red[w];
  __syncthreads();
  return tot;
}

// update1NBPCG: p = z + beta p ; s = Z + beta s ; partial p.s
__global__ void __launch_bounds__(kBlock) nb_update1_kernel(dlong N, const double* __restrict__ z,
                                                            const double* __restrict__ Z, double beta,
                                                            double* __restrict__ p, double* __restrict__ s,
                                                            double* __restrict__ partials) {
  __shared__ double s_red[kBlock / 32];
  double acc = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double pn = z[n] + beta * p[n];
    const double sn = Z[n] + beta * s[n];
    p[n] = pn;
    s[n] = sn;
    acc += pn * sn;
  }
  const double t = block_sum(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// update2NBPCG: r -= alpha s ; z -= alpha S ; partial r.z, z.z, r.r
__global__ void __launch_bounds__(kBlock) nb_update2_kernel(dlong N, const double* __restrict__ s,
                                                            const double* __restrict__ S, double alpha,
                                                            double* __restrict__ r, double* __restrict__ z,
                                                            double* __restrict__ partials, int stride) {
  __shared__ double s_red[kBlock / 32];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double rn = r[n] - alpha * s[n];
    const double zn = z[n] - alpha * S[n];
    r[n] = rn;
    z[n] = zn;
    a0 += rn * zn;
    a1 += zn * zn;
    a2 += rn * rn;
  }
  const double t0 = block_sum(a0, s_red), t1 = block_sum(a1, s_red), t2 = block_sum(a2, s_red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = t0;
    partials[stride + blockIdx.x] = t1;
    partials[2 * stride + blockIdx.x] = t2;
  }
}

// out[k] = sum of the nb partials of quantity k, left to right (the reference sums them on the host in that order)
__global__ void nb_finish_kernel(const double* __restrict__ partials, int nb, int stride, int nq, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k < nq) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += partials[k * stride + b];
    out[k] = t;
  }
}
}  // namespace

struct libp_nbpcg_s {
  dlong N = 0, Nhalo = 0;
  libp_comm_t comm = nullptr;
  dev_buf<double> p, s, S, z, Z, Ax, partials, d_dots;
  double* h_dots = nullptr;  // pinned
  cudaEvent_t ev = nullptr;
  std::vector<double> hist;
  ~libp_nbpcg_s() {
    if (h_dots) cudaFreeHost(h_dots);
    if (ev) cudaEventDestroy(ev);
  }
  int nblocks() const {
    const long nb = ((long)N + kBlock - 1) / kBlock;
    return (int)std::max<long>(1, std::min<long>(nb, kMaxBlocks));
  }
  // issue a reduction: the scalars land in h_dots once `ev` has completed
  void post(int nq, cudaStream_t st) {
    nb_finish_kernel<<<1, 32, 0, st>>>(partials.p, nblocks(), kMaxBlocks, nq, d_dots.p);
    CUDA_CHECK(cudaGetLastError());
    if (comm && comm->size > 1) comm->allreduce_sum_dev(d_dots.p, nq, st);
    CUDA_CHECK(cudaMemcpyAsync(h_dots, d_dots.p, sizeof(double) * nq, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaEventRecord(ev, st));
  }
  void wait() { CUDA_CHECK(cudaEventSynchronize(ev)); }
};

extern "C" int libp_nbpcg_create(libp_dlong N, libp_dlong Nhalo, libp_comm_t comm, libp_nbpcg_t* solver) {
  LIBP_API_BEGIN
  LIBP_CHECK(solver && N >= 0 && Nhalo >= 0, "bad argument");
  std::unique_ptr<libp_nbpcg_s> h(new libp_nbpcg_s());
  h->N = N; h->Nhalo = Nhalo; h->comm = comm;
  const size_t Nt = std::max<size_t>((size_t)N + Nhalo, 1);
  for (dev_buf<double>* b : {&h->p, &h->s, &h->S, &h->z, &h->Z, &h->Ax}) {
    b->alloc(Nt);
    CUDA_CHECK(cudaMemset(b->p, 0, sizeof(double) * Nt));
  }
  h->partials.alloc((size_t)3 * kMaxBlocks);
  h->d_dots.alloc(4);
  CUDA_CHECK(cudaMallocHost(&h->h_dots, sizeof(double) * 4));
  CUDA_CHECK(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
  *solver = h.release();
  LIBP_API_END
}

extern "C" int libp_nbpcg_free(libp_nbpcg_t solver) {
  LIBP_API_BEGIN
  delete solver;
  LIBP_API_END
}

extern "C" int libp_nbpcg_residual_history(libp_nbpcg_t solver, const libp_dfloat** hist, int* n) {
  LIBP_API_BEGIN
  LIBP_CHECK(solver && hist && n, "null argument");
  *hist = solver->hist.data();
  *n = (int)solver->hist.size();
  LIBP_API_END
}

// nbpcg::Solve (linearSolverNBPCG.cpp:68-173)
extern "C" int libp_nbpcg_solve_cb(libp_nbpcg_t h, libp_operator_fn A, void* Actx, libp_operator_fn M, void* Mctx,
                                   libp_dfloat* x, libp_dfloat* r, libp_dfloat tol, int maxit, int verbose,
                                   void* stream, int* iters) {
  LIBP_API_BEGIN
  LIBP_CHECK(h && A && M && x && r && iters, "null argument");
  cudaStream_t st = as_stream(stream);
  const dlong N = h->N;
  const int rank = h->comm ? h->comm->rank : 0;
  const int nb = h->nblocks();
  auto ok = [](int rc) { LIBP_CHECK(rc == LIBP_SUCCESS, libp_last_error()); };
  double alpha = 0, beta = 0, gamma0 = 0, gamma1 = 0, delta0 = 0, zdotz0 = 0, rdotr0 = 0;
  auto update1 = [&](double b) {  // p, s, then p.s on its way to the host
    nb_update1_kernel<<<nb, kBlock, 0, st>>>(N, h->z.p, h->Z.p, b, h->p.p, h->s.p, h->partials.p);
    CUDA_CHECK(cudaGetLastError());
    h->post(1, st);
  };
  auto update2 = [&](double a) {  // r, z, then r.z, z.z, r.r on their way to the host
    nb_update2_kernel<<<nb, kBlock, 0, st>>>(N, h->s.p, h->S.p, a, r, h->z.p, h->partials.p, kMaxBlocks);
    CUDA_CHECK(cudaGetLastError());
    h->post(3, st);
  };
  ok(A(Actx, x, h->Ax.p, stream));
  ok(libp_linalg_axpy(N, -1.0, h->Ax.p, 1.0, r, stream));
  ok(M(Mctx, r, h->z.p, stream));   // z = M r
  update2(0.0);                     // alpha = 0: r.z, z.z, r.r of the initial residual
  ok(A(Actx, h->z.p, h->Z.p, stream));
  h->wait();
  gamma0 = h->h_dots[0]; zdotz0 = h->h_dots[1]; rdotr0 = h->h_dots[2];
  const double TOL = std::max(tol * tol * rdotr0, tol * tol);
  if (verbose && rank == 0) printf("NBPCG: initial res norm %12.12f \n", sqrt(rdotr0));
  h->hist.clear();
  h->hist.push_back(sqrt(rdotr0));
  int iter;
  beta = 0.0;
  for (iter = 0; iter < maxit; ++iter) {
    if (rdotr0 <= TOL) break;
    update1(beta);                          // p = z + beta p ; s = Z + beta s ; delta = p.s
    ok(M(Mctx, h->s.p, h->S.p, stream));    // S = M s, overlapping the reduction
    h->wait();
    delta0 = h->h_dots[0];
    alpha = gamma0 / delta0;
    update2(alpha);                         // r -= alpha s ; z -= alpha S ; r.z, z.z, r.r
    ok(libp_linalg_axpy(N, alpha, h->p.p, 1.0, x, stream));  // x += alpha p (delayed)
    ok(A(Actx, h->z.p, h->Z.p, stream));    // Z = A z, overlapping the reduction
    h->wait();
    gamma1 = gamma0;
    gamma0 = h->h_dots[0]; zdotz0 = h->h_dots[1]; rdotr0 = h->h_dots[2];
    beta = gamma0 / gamma1;
    h->hist.push_back(sqrt(std::max(rdotr0, 0.0)));
    if (verbose && rank == 0) {
      if (rdotr0 < 0) printf("WARNING NBPCG: rdotr = %17.15lf\n", rdotr0);
      printf("NBPCG: it %d, r norm %12.12le, gamma = %le zdotz = %le \n", iter + 1, sqrt(rdotr0), gamma0, zdotz0);
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  *iters = iter;
  LIBP_API_END
}

static int nb_elliptic_cb(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream) {
  return libp_elliptic_operator(static_cast<libp_elliptic_t>(ctx), in, out, stream);
}
static int nb_precon_cb(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream) {
  return libp_precon_apply(static_cast<libp_precon_t>(ctx), in, out, stream);
}

extern "C" int libp_nbpcg_solve(libp_nbpcg_t solver, libp_elliptic_t A, libp_precon_t M, libp_dfloat* x, libp_dfloat* r,
                                libp_dfloat tol, int maxit, int verbose, void* stream, int* iters) {
  return libp_nbpcg_solve_cb(solver, nb_elliptic_cb, A, nb_precon_cb, M, x, r, tol, maxit, verbose, stream, iters);
}
