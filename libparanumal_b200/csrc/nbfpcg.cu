// Non-blocking flexible preconditioned conjugate gradients (Sanan et al.), LinearSolver::nbfpcg
// (libs/linearSolver/linearSolverNBFPCG.cpp:71-246, kernels okl/linearSolverUpdateNBFPCG.okl).
//
// Same recurrences, same overlap structure as the reference: one fused update kernel per iteration
//     x += alpha p ; r -= alpha s ; u -= alpha q ; w -= alpha z ; partial u.r, u.s, u.w, r.r
// posts its four scalars (deterministic block sums -> device all-reduce -> asynchronous copy into pinned memory),
// and the preconditioner + operator applies of the NEXT step are queued behind it before the host waits, so the wait
// is covered by device work.  The four direction updates p, s, q, z (four axpy launches in the reference) are ONE
// pass here.  Operators are callbacks (libp_operator_fn); libp_nbfpcg_solve takes the native handles.
#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include "elliptic.hpp"
#include "linalg.hpp"

using namespace libp_b200;

namespace {
constexpr int kBlock = 256;
constexpr int kMaxBlocks = 512;  // NBFPCG_BLOCKSIZE partial sums (linearSolverNBFPCG.cpp:32)

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

// update0NBFPCG: partial u.r, u.w, r.r
__global__ void __launch_bounds__(kBlock) nbf_update0_kernel(dlong N, const double* __restrict__ u,
                                                             const double* __restrict__ r, const double* __restrict__ w,
                                                             double* __restrict__ partials, int stride) {
  __shared__ double s_red[kBlock / 32];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double un = u[n], rn = r[n], wn = w[n];
    a0 += un * rn;
    a1 += un * wn;
    a2 += rn * rn;
  }
  const double t0 = block_sum(a0, s_red), t1 = block_sum(a1, s_red), t2 = block_sum(a2, s_red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = t0; partials[stride + blockIdx.x] = t1; partials[2 * stride + blockIdx.x] = t2;
  }
}

// update1NBFPCG: x += alpha p ; r -= alpha s ; u -= alpha q ; w -= alpha z ; partial u.r, u.s, u.w, r.r
__global__ void __launch_bounds__(kBlock) nbf_update1_kernel(dlong N, const double* __restrict__ p,
                                                             const double* __restrict__ s, const double* __restrict__ q,
                                                             const double* __restrict__ z, double alpha,
                                                             double* __restrict__ x, double* __restrict__ r,
                                                             double* __restrict__ u, double* __restrict__ w,
                                                             double* __restrict__ partials, int stride) {
  __shared__ double s_red[kBlock / 32];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double sn = s[n];
    const double xn = x[n] + alpha * p[n];
    const double rn = r[n] - alpha * sn;
    const double un = u[n] - alpha * q[n];
    const double wn = w[n] - alpha * z[n];
    a0 += un * rn;
    a1 += un * sn;
    a2 += un * wn;
    a3 += rn * rn;
    x[n] = xn; r[n] = rn; u[n] = un; w[n] = wn;
  }
  const double t0 = block_sum(a0, s_red), t1 = block_sum(a1, s_red), t2 = block_sum(a2, s_red), t3 = block_sum(a3, s_red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = t0; partials[stride + blockIdx.x] = t1;
    partials[2 * stride + blockIdx.x] = t2; partials[3 * stride + blockIdx.x] = t3;
  }
}

// p = u + beta p ; s = w + beta s ; q = m + beta q ; z = n + beta z   (linearSolverNBFPCG.cpp:156-166, one pass)
__global__ void __launch_bounds__(kBlock) nbf_directions_kernel(dlong N, double beta, const double* __restrict__ u,
                                                                const double* __restrict__ w, const double* __restrict__ m,
                                                                const double* __restrict__ nn, double* __restrict__ p,
                                                                double* __restrict__ s, double* __restrict__ q,
                                                                double* __restrict__ z) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    p[n] = u[n] + beta * p[n];
    s[n] = w[n] + beta * s[n];
    q[n] = m[n] + beta * q[n];
    z[n] = nn[n] + beta * z[n];
  }
}

__global__ void nbf_finish_kernel(const double* __restrict__ partials, int nb, int stride, int nq, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k < nq) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += partials[k * stride + b];
    out[k] = t;
  }
}
}  // namespace

struct libp_nbfpcg_s {
  dlong N = 0, Nhalo = 0;
  libp_comm_t comm = nullptr;
  dev_buf<double> u, p, w, n, m, s, z, q, Ax, partials, d_dots;
  double* h_dots = nullptr;  // pinned
  cudaEvent_t ev = nullptr;
  std::vector<double> hist;
  ~libp_nbfpcg_s() {
    if (h_dots) cudaFreeHost(h_dots);
    if (ev) cudaEventDestroy(ev);
  }
  int nblocks() const {
    const long nb = ((long)N + kBlock - 1) / kBlock;
    return (int)std::max<long>(1, std::min<long>(nb, kMaxBlocks));
  }
  void post(int nq, cudaStream_t st) {
    nbf_finish_kernel<<<1, 32, 0, st>>>(partials.p, nblocks(), kMaxBlocks, nq, d_dots.p);
    CUDA_CHECK(cudaGetLastError());
    if (comm && comm->size > 1) comm->allreduce_sum_dev(d_dots.p, nq, st);
    CUDA_CHECK(cudaMemcpyAsync(h_dots, d_dots.p, sizeof(double) * nq, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaEventRecord(ev, st));
  }
  void wait() { CUDA_CHECK(cudaEventSynchronize(ev)); }
};

extern "C" int libp_nbfpcg_create(libp_dlong N, libp_dlong Nhalo, libp_comm_t comm, libp_nbfpcg_t* solver) {
  LIBP_API_BEGIN
  LIBP_CHECK(solver && N >= 0 && Nhalo >= 0, "bad argument");
  std::unique_ptr<libp_nbfpcg_s> h(new libp_nbfpcg_s());
  h->N = N; h->Nhalo = Nhalo; h->comm = comm;
  const size_t Nt = std::max<size_t>((size_t)N + Nhalo, 1);
  for (dev_buf<double>* b : {&h->u, &h->p, &h->w, &h->n, &h->m, &h->s, &h->z, &h->q, &h->Ax}) {
    b->alloc(Nt);
    CUDA_CHECK(cudaMemset(b->p, 0, sizeof(double) * Nt));
  }
  h->partials.alloc((size_t)4 * kMaxBlocks);
  h->d_dots.alloc(4);
  CUDA_CHECK(cudaMallocHost(&h->h_dots, sizeof(double) * 4));
  CUDA_CHECK(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
  *solver = h.release();
  LIBP_API_END
}

extern "C" int libp_nbfpcg_free(libp_nbfpcg_t solver) {
  LIBP_API_BEGIN
  delete solver;
  LIBP_API_END
}

extern "C" int libp_nbfpcg_residual_history(libp_nbfpcg_t solver, const libp_dfloat** hist, int* n) {
  LIBP_API_BEGIN
  LIBP_CHECK(solver && hist && n, "null argument");
  *hist = solver->hist.data();
  *n = (int)solver->hist.size();
  LIBP_API_END
}

// nbfpcg::Solve (linearSolverNBFPCG.cpp:71-188)
extern "C" int libp_nbfpcg_solve_cb(libp_nbfpcg_t h, libp_operator_fn A, void* Actx, libp_operator_fn M, void* Mctx,
                                    libp_dfloat* x, libp_dfloat* r, libp_dfloat tol, int maxit, int verbose,
                                    void* stream, int* iters) {
  LIBP_API_BEGIN
  LIBP_CHECK(h && A && M && x && r && iters, "null argument");
  cudaStream_t st = as_stream(stream);
  const dlong N = h->N;
  const size_t bytes = sizeof(double) * (size_t)N;
  const int rank = h->comm ? h->comm->rank : 0;
  const int nb = h->nblocks();
  auto ok = [](int rc) { LIBP_CHECK(rc == LIBP_SUCCESS, libp_last_error()); };
  auto copy = [&](double* dst, const double* src) {
    if (N) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
  };
  double alpha0 = 0, beta0 = 0, gamma0 = 0, delta0 = 0, eta0 = 0, rdotr0 = 0;
  ok(A(Actx, x, h->Ax.p, stream));                               // A x
  ok(libp_linalg_axpy(N, -1.0, h->Ax.p, 1.0, r, stream));        // r = r - A x
  ok(M(Mctx, r, h->u.p, stream));                                // u = M r
  copy(h->p.p, h->u.p);                                          // p = u
  ok(A(Actx, h->p.p, h->w.p, stream));                           // w = A p
  nbf_update0_kernel<<<nb, kBlock, 0, st>>>(N, h->u.p, r, h->w.p, h->partials.p, kMaxBlocks);  // u.r, u.w, r.r
  CUDA_CHECK(cudaGetLastError());
  h->post(3, st);
  ok(M(Mctx, h->w.p, h->m.p, stream));                           // m = M w     (overlaps the reduction)
  ok(A(Actx, h->m.p, h->n.p, stream));                           // n = A m
  copy(h->s.p, h->w.p);
  copy(h->q.p, h->m.p);
  copy(h->z.p, h->n.p);
  h->wait();
  gamma0 = h->h_dots[0]; delta0 = h->h_dots[1]; rdotr0 = h->h_dots[2];
  eta0 = delta0;
  alpha0 = gamma0 / eta0;
  const double TOL = std::max(tol * tol * rdotr0, tol * tol);
  if (verbose && rank == 0) printf("NBFPCG: initial res norm %12.12f \n", sqrt(rdotr0));
  h->hist.clear();
  h->hist.push_back(sqrt(rdotr0));
  int iter;
  for (iter = 0; iter < maxit; ++iter) {
    if (rdotr0 <= TOL) break;
    nbf_update1_kernel<<<nb, kBlock, 0, st>>>(N, h->p.p, h->s.p, h->q.p, h->z.p, alpha0, x, r, h->u.p, h->w.p,
                                              h->partials.p, kMaxBlocks);
    CUDA_CHECK(cudaGetLastError());
    h->post(4, st);
    ok(libp_linalg_zaxpy(N, 1.0, h->w.p, -1.0, r, h->n.p, stream));   // n = w - r
    ok(M(Mctx, h->n.p, h->m.p, stream));                              // m = M (w - r)
    ok(libp_linalg_axpy(N, 1.0, h->u.p, 1.0, h->m.p, stream));        // m = u + M (w - r)
    ok(A(Actx, h->m.p, h->n.p, stream));                              // n = A m
    h->wait();
    gamma0 = h->h_dots[0];
    beta0 = -h->h_dots[1] / eta0;
    delta0 = h->h_dots[2];
    rdotr0 = h->h_dots[3];
    nbf_directions_kernel<<<nb, kBlock, 0, st>>>(N, beta0, h->u.p, h->w.p, h->m.p, h->n.p, h->p.p, h->s.p, h->q.p, h->z.p);
    CUDA_CHECK(cudaGetLastError());
    eta0 = delta0 - beta0 * beta0 * eta0;
    alpha0 = gamma0 / eta0;
    h->hist.push_back(sqrt(std::max(rdotr0, 0.0)));
    if (verbose && rank == 0) {
      if (rdotr0 < 0) printf("WARNING NBFPCG: rdotr = %17.15lf\n", rdotr0);
      printf("NBFPCG: it %d, r norm %12.12le, alpha = %12.12le beta = %le rdotr = %le gamma = %le delta = %le, eta = %le \n",
             iter + 1, sqrt(rdotr0), alpha0, beta0, rdotr0, gamma0, delta0, eta0);
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  *iters = iter;
  LIBP_API_END
}

static int nbf_elliptic_cb(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream) {
  return libp_elliptic_operator(static_cast<libp_elliptic_t>(ctx), in, out, stream);
}
static int nbf_precon_cb(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream) {
  return libp_precon_apply(static_cast<libp_precon_t>(ctx), in, out, stream);
}

extern "C" int libp_nbfpcg_solve(libp_nbfpcg_t solver, libp_elliptic_t A, libp_precon_t M, libp_dfloat* x, libp_dfloat* r,
                                 libp_dfloat tol, int maxit, int verbose, void* stream, int* iters) {
  return libp_nbfpcg_solve_cb(solver, nbf_elliptic_cb, A, nbf_precon_cb, M, x, r, tol, maxit, verbose, stream, iters);
}
