// Initial-guess strategies of linearSolver_t (libs/linearSolver/initialGuess.cpp, include/initialGuess.hpp):
//   ZERO, CLASSIC (Fischer 1998 projection), QR (rolling-QR projection, Christensen), EXTRAP (polynomial extrapolation
//   of the solution history).  linearSolver_t::Solve brackets every solve with FormInitialGuess / Update
//   (libs/linearSolver/linearSolver.cpp:31-44).
// Device work: the multi-dot  c_m = <x, Q_m>  is ONE pass over x for all m (deterministic block sums, device
// all-reduce), reconstruction / scaling / history update / Givens sweep / extrapolation are single fused passes.
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "common.hpp"

using namespace libp_b200;

extern "C" int libp_linalg_norm2(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* out);

namespace {
constexpr int kBlock = 256;
constexpr int kMaxHist = 32;
constexpr int kMaxBlocks = 1024;

inline int grid_for(dlong N) { return (int)std::max<long>(1, std::min<long>(((long)N + kBlock - 1) / kBlock, kMaxBlocks)); }

// partial c_m = sum_n x[n] * Q[m*N + n] for m < dim (igBasisInnerProducts.okl): partials[m*gridDim + block]
template <int kDim>
__global__ void __launch_bounds__(kBlock) ig_inner_products_kernel(dlong N, int dim, const double* __restrict__ x,
                                                                   const double* __restrict__ Q,
                                                                   double* __restrict__ partials) {
  __shared__ double s_red[kDim][kBlock / 32];
  double acc[kDim];
#pragma unroll
  for (int m = 0; m < kDim; ++m) acc[m] = 0.0;
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double xn = x[n];
#pragma unroll
    for (int m = 0; m < kDim; ++m)
      if (m < dim) acc[m] += xn * Q[(size_t)m * N + n];
  }
#pragma unroll
  for (int m = 0; m < kDim; ++m) {
    double v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[m][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < dim && threadIdx.x < kDim) {
    double t = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) t += s_red[threadIdx.x][w];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void ig_finish_kernel(int dim, int nb, const double* __restrict__ partials, double* __restrict__ out) {
  const int m = threadIdx.x;
  if (m < dim) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += partials[(size_t)m * nb + b];  // left to right, as the reference's host loop
    out[m] = t;
  }
}
// unew = u + a * sum_m c[m] Q_m   (igReconstruct.okl)
__global__ void __launch_bounds__(kBlock) ig_reconstruct_kernel(dlong N, int dim, const double* __restrict__ u, double a,
                                                                const double* __restrict__ c, const double* __restrict__ Q,
                                                                double* __restrict__ unew) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    double t = 0.0;
    for (int m = 0; m < dim; ++m) t += c[m] * Q[(size_t)m * N + n];
    unew[n] = u[n] + a * t;
  }
}
// igUpdate.okl: column curDim of the spaces <- scale * (btilde, xtilde)
__global__ void __launch_bounds__(kBlock) ig_update_kernel(dlong N, int col, double scale, double* __restrict__ bt, int bw,
                                                           double* __restrict__ B, double* __restrict__ xt, int xw,
                                                           double* __restrict__ X) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double b = scale * bt[n], x = scale * xt[n];
    B[(size_t)col * N + n] = b;
    X[(size_t)col * N + n] = x;
    if (bw) bt[n] = b;
    if (xw) xt[n] = x;
  }
}
// igScale.okl
__global__ void __launch_bounds__(kBlock) ig_scale_kernel(dlong N, double a, const double* __restrict__ in,
                                                          double* __restrict__ out) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) out[n] = a * in[n];
}
// igDropQRFirstColumn.okl: drop column 0 of Q = Btilde (and of U = Xtilde) and restore the triangular R by a sweep of
// Givens rotations applied to neighbouring columns; the last column becomes zero.  cs = [c_0, s_0, c_1, s_1, ...]
// (computed on the host from R, the same rotations the host applies to R itself).
__global__ void __launch_bounds__(kBlock) ig_drop_first_column_kernel(dlong N, int dim, const double* __restrict__ cs,
                                                                      double* __restrict__ Q, double* __restrict__ U) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    double qi = Q[n], ui = U[n];
    for (int i = 0; i < dim - 1; ++i) {
      const double c = cs[2 * i], s = cs[2 * i + 1];
      const double qn = Q[(size_t)(i + 1) * N + n], un = U[(size_t)(i + 1) * N + n];
      Q[(size_t)i * N + n] = c * qi + s * qn;
      U[(size_t)i * N + n] = c * ui + s * un;
      qi = -s * qi + c * qn;
      ui = -s * ui + c * un;
    }
    Q[(size_t)(dim - 1) * N + n] = 0.0;
    U[(size_t)(dim - 1) * N + n] = 0.0;
  }
}
// igExtrap.okl
__global__ void __launch_bounds__(kBlock) ig_extrap_kernel(dlong N, int nh, int shift, const double* __restrict__ coeffs,
                                                           const double* __restrict__ uh, double* __restrict__ uex) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    double res = 0.0;
    for (int i = 0; i < nh; ++i) {
      const double ci = coeffs[i];
      if (ci != 0.0) res += ci * uh[(size_t)n + (size_t)((i + shift) % nh) * N];
    }
    uex[n] = res;
  }
}

// ---- host: Legendre Vandermonde (mesh_t::Vandermonde1D = JacobiP(a,0,0,i), orthonormal) and the two small solves
double legendre_orthonormal(double a, int i) {
  // JacobiP(a, 0, 0, i) of the reference = orthonormal Legendre polynomial sqrt((2i+1)/2) P_i(a)
  double p0 = 1.0, p1 = a;
  if (i == 0) return std::sqrt(0.5);
  for (int k = 2; k <= i; ++k) {
    const double pn = ((2.0 * k - 1.0) * a * p1 - (k - 1.0) * p0) / k;
    p0 = p1; p1 = pn;
  }
  return std::sqrt((2.0 * i + 1.0) / 2.0) * p1;
}
// minimum-norm c (length M) with  sum_i c_i V[i][j] = b[j]  (j <= m): Householder QR of V (M x (m+1)), c = Q R^-T b
// (what dgels returns for the underdetermined system, linAlgMatrixRightSolve.cpp:204-229)
void min_norm_solve(int M, int n, const std::vector<double>& V, const std::vector<double>& b, std::vector<double>& c) {
  std::vector<double> A(V);  // row major M x n
  std::vector<double> tau((size_t)n, 0.0);
  for (int k = 0; k < n; ++k) {
    double nrm = 0.0;
    for (int i = k; i < M; ++i) nrm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = A[(size_t)k * n + k];
    const double beta = alpha >= 0 ? -nrm : nrm;
    tau[k] = (beta - alpha) / beta;
    const double scale = 1.0 / (alpha - beta);
    for (int i = k + 1; i < M; ++i) A[(size_t)i * n + k] *= scale;
    A[(size_t)k * n + k] = beta;
    for (int j = k + 1; j < n; ++j) {
      double s = A[(size_t)k * n + j];
      for (int i = k + 1; i < M; ++i) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
      s *= tau[k];
      A[(size_t)k * n + j] -= s;
      for (int i = k + 1; i < M; ++i) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
    }
  }
  // y = R^-T b  (R upper triangular n x n in A)
  std::vector<double> y((size_t)M, 0.0);
  for (int j = 0; j < n; ++j) {
    double s = b[j];
    for (int k = 0; k < j; ++k) s -= A[(size_t)k * n + j] * y[k];
    y[j] = s / A[(size_t)j * n + j];
  }
  // c = Q [y; 0] = H_0 H_1 ... H_{n-1} [y; 0]
  for (int k = n - 1; k >= 0; --k) {
    double s = y[k];
    for (int i = k + 1; i < M; ++i) s += A[(size_t)i * n + k] * y[i];
    s *= tau[k];
    y[k] -= s;
    for (int i = k + 1; i < M; ++i) y[i] -= s * A[(size_t)i * n + k];
  }
  c = y;
}
// basic solution by column-pivoted QR of V^T ((m+1) x M): the pivoting picks m+1 history vectors
// (linAlgMatrixRightSolve.cpp:262-316: dgeqp3, dormqr, dtrsm, permutation)
void cpqr_solve(int M, int n, const std::vector<double>& V, const std::vector<double>& b, std::vector<double>& c) {
  // A = V^T, n x M, column j = history vector j
  std::vector<double> A((size_t)n * M);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < n; ++j) A[(size_t)j * M + i] = V[(size_t)i * n + j];
  std::vector<int> piv((size_t)M);
  for (int j = 0; j < M; ++j) piv[j] = j;
  std::vector<double> rhs(b.begin(), b.begin() + n);
  for (int k = 0; k < n; ++k) {
    int best = k;
    double bn = -1.0;
    for (int j = k; j < M; ++j) {
      double s = 0.0;
      for (int i = k; i < n; ++i) s += A[(size_t)i * M + j] * A[(size_t)i * M + j];
      if (s > bn * (1.0 + 1e-14)) { bn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < n; ++i) std::swap(A[(size_t)i * M + k], A[(size_t)i * M + best]);
      std::swap(piv[k], piv[best]);
    }
    double nrm = 0.0;
    for (int i = k; i < n; ++i) nrm += A[(size_t)i * M + k] * A[(size_t)i * M + k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = A[(size_t)k * M + k];
    const double beta = alpha >= 0 ? -nrm : nrm;
    const double tau = (beta - alpha) / beta;
    const double scale = 1.0 / (alpha - beta);
    std::vector<double> v((size_t)n, 0.0);
    v[k] = 1.0;
    for (int i = k + 1; i < n; ++i) v[i] = A[(size_t)i * M + k] * scale;
    for (int j = k; j < M; ++j) {
      double s = 0.0;
      for (int i = k; i < n; ++i) s += v[i] * A[(size_t)i * M + j];
      s *= tau;
      for (int i = k; i < n; ++i) A[(size_t)i * M + j] -= s * v[i];
    }
    double s = 0.0;
    for (int i = k; i < n; ++i) s += v[i] * rhs[i];
    s *= tau;
    for (int i = k; i < n; ++i) rhs[i] -= s * v[i];
  }
  std::vector<double> y((size_t)n, 0.0);
  for (int k = n - 1; k >= 0; --k) {
    double s = rhs[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * M + j] * y[j];
    y[k] = s / A[(size_t)k * M + k];
  }
  c.assign((size_t)M, 0.0);
  for (int k = 0; k < n; ++k) c[piv[k]] = y[k];
}
// Extrap::extrapCoeffs (initialGuess.cpp:440-466)
void extrap_coeffs(int m, int M, bool cpqr, std::vector<double>& c) {
  LIBP_CHECK(M >= m + 1, "Extrapolation space dimension (" + std::to_string(M) + ") too low for degree (" +
                             std::to_string(m) + ").");
  const double h = 2.0 / (M - 1);
  std::vector<double> V((size_t)M * (m + 1)), b((size_t)m + 1);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j <= m; ++j) V[(size_t)i * (m + 1) + j] = legendre_orthonormal(-1.0 + i * h, j);
  for (int j = 0; j <= m; ++j) b[j] = legendre_orthonormal(1.0 + h, j);
  if (cpqr) cpqr_solve(M, m + 1, V, b, c);
  else min_norm_solve(M, m + 1, V, b, c);
}
}  // namespace

struct libp_ig_s {
  int strategy = 0;  // 0 NONE, 1 ZERO, 2 CLASSIC, 3 QR, 4 EXTRAP
  dlong N = 0, Nall = 0;
  libp_comm_t comm = nullptr;
  int curDim = 0, maxDim = 0;
  dev_buf<double> btilde, xtilde, Btilde, Xtilde, partials, d_alphas, d_cs;
  std::vector<double> alphas, R;
  double* h_pinned = nullptr;
  // EXTRAP
  int Nhistory = 0, shift = 0, entry = 0, degree = 0, cpqr = 0;
  dev_buf<double> xh, d_coeffs;
  ~libp_ig_s() { if (h_pinned) cudaFreeHost(h_pinned); }

  void inner_products(const double* x, cudaStream_t s) {  // alphas[0:curDim] = <x, Btilde_m>, host + device copies
    const int nb = grid_for(N);
    if (curDim <= 8) ig_inner_products_kernel<8><<<nb, kBlock, 0, s>>>(N, curDim, x, Btilde.p, partials.p);
    else if (curDim <= 16) ig_inner_products_kernel<16><<<nb, kBlock, 0, s>>>(N, curDim, x, Btilde.p, partials.p);
    else ig_inner_products_kernel<kMaxHist><<<nb, kBlock, 0, s>>>(N, curDim, x, Btilde.p, partials.p);
    ig_finish_kernel<<<1, kMaxHist, 0, s>>>(curDim, nb, partials.p, d_alphas.p);
    CUDA_CHECK(cudaGetLastError());
    if (comm && comm->size > 1) comm->allreduce_sum_dev(d_alphas.p, curDim, s);
    CUDA_CHECK(cudaMemcpyAsync(h_pinned, d_alphas.p, sizeof(double) * curDim, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int m = 0; m < curDim; ++m) alphas[m] = h_pinned[m];
  }
  void reconstruct(const double* u, double a, const double* Q, double* unew, cudaStream_t s) {
    ig_reconstruct_kernel<<<grid_for(N), kBlock, 0, s>>>(N, curDim, u, a, d_alphas.p, Q, unew);
    CUDA_CHECK(cudaGetLastError());
  }
  double norm2(const double* v, cudaStream_t s) {
    double out = 0.0;
    LIBP_CHECK(libp_linalg_norm2(N, v, comm, s, &out) == LIBP_SUCCESS, libp_last_error());
    return out;
  }
};

extern "C" int libp_ig_create(int strategy, libp_dlong N, libp_dlong Nhalo, int maxDim, int extrapDegree, int cpqr,
                              libp_comm_t comm, libp_ig_t* out) {
  LIBP_API_BEGIN
  LIBP_CHECK(out && N >= 0 && Nhalo >= 0, "bad argument");
  LIBP_CHECK(strategy >= 0 && strategy <= 4, "strategy: 0 NONE, 1 ZERO, 2 CLASSIC, 3 QR, 4 EXTRAP");
  std::unique_ptr<libp_ig_s> g(new libp_ig_s());
  g->strategy = strategy; g->N = N; g->Nall = N + Nhalo; g->comm = comm;
  const size_t Nt = std::max<size_t>((size_t)N + Nhalo, 1);
  if (strategy == 2 || strategy == 3) {
    LIBP_CHECK(maxDim >= 1 && maxDim <= kMaxHist, "INITIAL GUESS HISTORY SPACE DIMENSION must be in [1, 32]");
    g->maxDim = maxDim;
    g->btilde.alloc(Nt); g->xtilde.alloc(Nt);
    g->Btilde.alloc((size_t)std::max<dlong>(N, 1) * maxDim); g->Xtilde.alloc((size_t)std::max<dlong>(N, 1) * maxDim);
    CUDA_CHECK(cudaMemset(g->btilde.p, 0, sizeof(double) * Nt));
    CUDA_CHECK(cudaMemset(g->xtilde.p, 0, sizeof(double) * Nt));
    g->partials.alloc((size_t)kMaxHist * kMaxBlocks);
    g->d_alphas.alloc(kMaxHist);
    g->d_cs.alloc(2 * kMaxHist);
    g->alphas.assign(kMaxHist, 0.0);
    g->R.assign((size_t)maxDim * maxDim, 0.0);
    CUDA_CHECK(cudaMallocHost(&g->h_pinned, sizeof(double) * 2 * kMaxHist));
  } else if (strategy == 4) {
    LIBP_CHECK(maxDim >= 1 && maxDim <= kMaxHist, "INITIAL GUESS HISTORY SPACE DIMENSION must be in [1, 32]");
    g->Nhistory = maxDim; g->degree = extrapDegree; g->cpqr = cpqr;
    std::vector<double> c;
    extrap_coeffs(extrapDegree, maxDim, cpqr != 0, c);  // validates (M >= m + 1), as the reference constructor does
    g->d_coeffs.alloc(kMaxHist);
    g->xh.alloc((size_t)std::max<dlong>(N, 1) * maxDim);
    CUDA_CHECK(cudaMemset(g->xh.p, 0, sizeof(double) * (size_t)std::max<dlong>(N, 1) * maxDim));
  }
  *out = g.release();
  LIBP_API_END
}

extern "C" int libp_ig_free(libp_ig_t ig) {
  LIBP_API_BEGIN
  delete ig;
  LIBP_API_END
}

extern "C" int libp_ig_dimension(libp_ig_t ig, int* curDim) {
  LIBP_API_BEGIN
  LIBP_CHECK(ig && curDim, "null argument");
  *curDim = ig->strategy == 4 ? std::min(ig->entry, ig->Nhistory) : ig->curDim;
  LIBP_API_END
}

extern "C" int libp_ig_extrap_coeffs(int m, int M, int cpqr, double* c) {
  LIBP_API_BEGIN
  LIBP_CHECK(c && M >= 1, "bad argument");
  std::vector<double> v;
  extrap_coeffs(m, M, cpqr != 0, v);
  std::copy(v.begin(), v.end(), c);
  LIBP_API_END
}

// initialGuessStrategy_t::FormInitialGuess
extern "C" int libp_ig_form_initial_guess(libp_ig_t g, libp_dfloat* x, const libp_dfloat* rhs, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(g && x && rhs, "null argument");
  cudaStream_t s = as_stream(stream);
  const dlong N = g->N;
  if (g->strategy == 1) {
    if (N) CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)N, s));
  } else if (g->strategy == 2 || g->strategy == 3) {  // Projection::FormInitialGuess (initialGuess.cpp:117-124)
    if (g->curDim > 0) {
      g->inner_products(rhs, s);
      if (N) CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)N, s));
      g->reconstruct(x, 1.0, g->Xtilde.p, x, s);
    }
  } else if (g->strategy == 4) {  // Extrap::FormInitialGuess (initialGuess.cpp:378-432)
    if (g->entry < g->Nhistory) {
      int M, m;
      if (g->entry == g->Nhistory - 1) { M = g->Nhistory; m = g->degree; }
      else { M = std::max(1, g->entry + 1); m = (int)std::sqrt((double)M); }
      std::vector<double> d((size_t)kMaxHist, 0.0);
      if (M == 1) {
        d[g->Nhistory - 1] = 1.0;
      } else {
        std::vector<double> c;
        extrap_coeffs(m, M, g->cpqr != 0, c);
        for (int i = 0; i < M; ++i) d[g->Nhistory - M + i] = c[i];
      }
      if (g->cpqr)
        for (double& v : d)
          if (std::abs(v) <= 1e-14) v = 0.0;  // the sparse kernel keeps |d| > 1e-14 only
      CUDA_CHECK(cudaMemcpyAsync(g->d_coeffs.p, d.data(), sizeof(double) * kMaxHist, cudaMemcpyHostToDevice, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      ++g->entry;
    }
    ig_extrap_kernel<<<grid_for(N), kBlock, 0, s>>>(N, g->Nhistory, g->shift, g->d_coeffs.p, g->xh.p, x);
    CUDA_CHECK(cudaGetLastError());
  }
  LIBP_API_END
}

// initialGuessStrategy_t::Update
extern "C" int libp_ig_update(libp_ig_t g, libp_operator_fn A, void* Actx, libp_dfloat* x, const libp_dfloat* rhs,
                              void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(g && x, "null argument");
  (void)rhs;
  cudaStream_t s = as_stream(stream);
  const dlong N = g->N;
  const int nbk = grid_for(N);
  if (g->strategy == 4) {  // Extrap::Update
    if (N) CUDA_CHECK(cudaMemcpyAsync(g->xh.p + (size_t)N * g->shift, x, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, s));
    g->shift = (g->shift + 1) % g->Nhistory;
    return LIBP_SUCCESS;
  }
  if (g->strategy != 2 && g->strategy != 3) return LIBP_SUCCESS;
  LIBP_CHECK(A != nullptr, "the projection strategies need the operator");
  LIBP_CHECK(A(Actx, x, g->btilde.p, stream) == LIBP_SUCCESS, libp_last_error());  // btilde = A x
  const int Nreorth = 2;
  if (g->strategy == 2) {  // ClassicProjection::Update (initialGuess.cpp:156-203)
    if (g->curDim >= g->maxDim || g->curDim == 0) {
      const double nb = g->norm2(g->btilde.p, s);
      if (nb > 0) {
        ig_scale_kernel<<<nbk, kBlock, 0, s>>>(N, 1.0 / nb, g->btilde.p, g->Btilde.p);
        ig_scale_kernel<<<nbk, kBlock, 0, s>>>(N, 1.0 / nb, x, g->Xtilde.p);
        CUDA_CHECK(cudaGetLastError());
        g->curDim = 1;
      }
    } else {
      if (N) CUDA_CHECK(cudaMemcpyAsync(g->xtilde.p, x, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, s));
      for (int n = 0; n < Nreorth; ++n) {
        g->inner_products(g->btilde.p, s);
        g->reconstruct(g->btilde.p, -1.0, g->Btilde.p, g->btilde.p, s);
        g->reconstruct(g->xtilde.p, -1.0, g->Xtilde.p, g->xtilde.p, s);
      }
      const double inv = 1.0 / g->norm2(g->btilde.p, s);
      ig_update_kernel<<<nbk, kBlock, 0, s>>>(N, g->curDim, inv, g->btilde.p, 1, g->Btilde.p, g->xtilde.p, 1, g->Xtilde.p);
      CUDA_CHECK(cudaGetLastError());
      g->curDim++;
    }
    return LIBP_SUCCESS;
  }
  // RollingQRProjection::Update (initialGuess.cpp:222-333)
  const int md = g->maxDim;
  std::vector<double>& R = g->R;
  if (g->curDim == md) {
    for (int j = 0; j < md; ++j) {
      for (int i = 0; i < md - 1; ++i) R[(size_t)j * md + i] = R[(size_t)j * md + i + 1];
      R[(size_t)j * md + md - 1] = 0.0;
    }
    std::vector<double> cs((size_t)2 * kMaxHist, 0.0);
    for (int j = 0; j < md - 1; ++j) {
      double c = 1.0, sn = 0.0;
      const double a = R[(size_t)j * md + j], b = R[(size_t)(j + 1) * md + j];
      if (b != 0) {
        const double h = std::hypot(a, b), d = 1.0 / h;
        c = std::abs(a) * d;
        sn = std::copysign(d, a) * b;
      }
      cs[2 * j] = c; cs[2 * j + 1] = sn;
      for (int i = j; i < md; ++i) {
        const double Rji = R[(size_t)j * md + i], Rj1i = R[(size_t)(j + 1) * md + i];
        R[(size_t)j * md + i] = c * Rji + sn * Rj1i;
        R[(size_t)(j + 1) * md + i] = -sn * Rji + c * Rj1i;
      }
    }
    CUDA_CHECK(cudaMemcpyAsync(g->d_cs.p, cs.data(), sizeof(double) * 2 * kMaxHist, cudaMemcpyHostToDevice, s));
    ig_drop_first_column_kernel<<<nbk, kBlock, 0, s>>>(N, md, g->d_cs.p, g->Btilde.p, g->Xtilde.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(s));
    g->curDim--;
  }
  if (g->curDim == 0) {
    const double nb = g->norm2(g->btilde.p, s);
    if (nb > 0) {
      ig_update_kernel<<<nbk, kBlock, 0, s>>>(N, 0, 1.0 / nb, g->btilde.p, 0, g->Btilde.p, x, 0, g->Xtilde.p);
      CUDA_CHECK(cudaGetLastError());
      R[0] = nb;
      g->curDim = 1;
    }
  } else {
    if (N) CUDA_CHECK(cudaMemcpyAsync(g->xtilde.p, x, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, s));
    const double nb = g->norm2(g->btilde.p, s);
    for (int i = 0; i < g->curDim; ++i) R[(size_t)i * md + g->curDim] = 0.0;
    for (int n = 0; n < Nreorth; ++n) {
      g->inner_products(g->btilde.p, s);
      g->reconstruct(g->btilde.p, -1.0, g->Btilde.p, g->btilde.p, s);
      g->reconstruct(g->xtilde.p, -1.0, g->Xtilde.p, g->xtilde.p, s);
      for (int i = 0; i < g->curDim; ++i) R[(size_t)i * md + g->curDim] += g->alphas[i];
    }
    const double nbp = g->norm2(g->btilde.p, s);
    if (nbp / nb > 1.0e-10) {
      ig_update_kernel<<<nbk, kBlock, 0, s>>>(N, g->curDim, 1.0 / nbp, g->btilde.p, 1, g->Btilde.p, g->xtilde.p, 1, g->Xtilde.p);
      CUDA_CHECK(cudaGetLastError());
      R[(size_t)g->curDim * md + g->curDim] = nbp;
      g->curDim++;
    }
  }
  LIBP_API_END
}
