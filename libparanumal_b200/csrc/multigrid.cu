// Multigrid preconditioner apply: matrix-free p-multigrid levels (MGLevel,
// solvers/elliptic/src/ellipticPreconMultiGridLevel.cpp:30-206), parAlmond CSR levels
// (libs/parAlmond/parAlmondAMGLevel.cpp:48-84, parAlmondparCSR.cpp:99-141, parAlmondAMGSmoother.cpp:35-160),
// the dense exact coarse solve (parAlmondCoarseExact.cpp:35-73) and the V-cycle driver
// (parAlmondVcycle.cpp:34-60).  Setup products (degree-raise matrices, inverse diagonals, eigenvalue bounds,
// AMG matrices, the coarse inverse) arrive as data; only the apply path lives here.
//
// B200 mapping
//  * every matrix-free level re-uses the fused Ax kernel (gather folded into its epilogue), so a Chebyshev sweep
//    of degree 2 is 2-3 Ax launches plus ONE fused vector kernel per Ax (x += d ; res -= invD Ad ; d = a d + b res)
//    instead of the reference's 3-4 separate BLAS-1 launches;
//  * coarsen = tensor-product restriction with the pre-weighting and the gather onto the coarse ogs fused in
//    (FP64 reductions into the coarse gathered vector: no element-local scratch, no separate gather pass);
//  * prolongate = tensor-product interpolation whose epilogue adds straight into the fine gathered vector through
//    the owner-copy map (ogs NoTrans semantics: exactly one local copy per DOF writes, so no atomics);
//  * CSR levels: 4 lanes per row (rows have 6-30 entries), fixed shuffle tree => deterministic sums.
#include <algorithm>
#include <cmath>

#include "elliptic.hpp"
#include "linalg.hpp"

using namespace libp_b200;

namespace {

constexpr int kBlock = 256;

inline int vgrid(size_t n, int per = 1) {
  size_t b = (n + (size_t)kBlock * per - 1) / ((size_t)kBlock * per);
  const size_t cap = (size_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ fused smoother vector kernels
// res = invD .* (r - Ax)  (Ax arrives in res; kHaveAx=false: x is zero, res = invD .* r) ; d = invTheta * res
template <bool kHaveAx>
__global__ void __launch_bounds__(kBlock) cheb_start_kernel(dlong N, double invTheta, const double* __restrict__ invD,
                                                            const double* __restrict__ r, double* __restrict__ res,
                                                            double* __restrict__ d) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double v = kHaveAx ? invD[n] * (r[n] - res[n]) : invD[n] * r[n];
    res[n] = v;
    d[n] = invTheta * v;
  }
}
// x (+)= d ; res -= invD .* Ad ; d = b*res + a*d ; on the last sweep also x += d_new
__global__ void __launch_bounds__(kBlock) cheb_iter_kernel(dlong N, int x_is_d, int last, double a, double b,
                                                           const double* __restrict__ invD, const double* __restrict__ Ad,
                                                           double* __restrict__ res, double* __restrict__ d,
                                                           double* __restrict__ x) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double dn = d[n];
    double xn = x_is_d ? dn : dn + x[n];
    const double rn = res[n] - invD[n] * Ad[n];
    const double dnew = b * rn + a * dn;
    res[n] = rn;
    d[n] = dnew;
    if (last) xn = dnew + xn;
    x[n] = xn;
  }
}
// Jacobi: x += invD .* (r - Ax)
__global__ void __launch_bounds__(kBlock) jacobi_update_kernel(dlong N, const double* __restrict__ invD,
                                                               const double* __restrict__ r, const double* __restrict__ Ax,
                                                               double* __restrict__ x) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) x[n] += invD[n] * (r[n] - Ax[n]);
}

// ------------------------------------------------------------------ transfer kernels (one element per block)
// qc = (P^T x P^T x P^T)(w .* qf) gathered (Add) onto the coarse vector.  P is [NqF][NqC].
// (ellipticPartialPreconCoarsenHex3D, okl/ellipticPreconCoarsenHex3D.okl:209-294, + amxpy pre-weight + ogs gather)
__global__ void __launch_bounds__(128) coarsen_kernel(dlong Nelements, const dlong* __restrict__ elementList, int NqF,
                                                      int NqC, const dlong* __restrict__ G2LF,
                                                      const dlong* __restrict__ G2LC, const double* __restrict__ P,
                                                      const double* __restrict__ w, const double* __restrict__ qf,
                                                      double* __restrict__ qc) {
  extern __shared__ double sm[];
  const int NpF = NqF * NqF * NqF, NpC = NqC * NqC * NqC;
  double* s0 = sm;                         // [NqF][NqF][NqF]
  double* s1 = s0 + NpF;                   // [NqC][NqF][NqF]
  double* s2 = s1 + NqC * NqF * NqF;       // [NqC][NqC][NqF]
  double* sP = s2 + NqC * NqC * NqF;       // [NqF][NqC]
  const dlong e = elementList ? elementList[blockIdx.x] : (dlong)blockIdx.x;
  const int t = threadIdx.x, T = blockDim.x;
  for (int n = t; n < NqF * NqC; n += T) sP[n] = P[n];
  for (int n = t; n < NpF; n += T) {
    const dlong id = G2LF[(size_t)e * NpF + n];
    s0[n] = (id >= 0) ? qf[id] * w[id] : 0.0;
  }
  __syncthreads();
  for (int o = t; o < NqC * NqF * NqF; o += T) {
    const int kc = o / (NqF * NqF), ji = o - kc * NqF * NqF;
    double acc = 0.0;
    for (int m = 0; m < NqF; ++m) acc += sP[m * NqC + kc] * s0[m * NqF * NqF + ji];
    s1[o] = acc;
  }
  __syncthreads();
  for (int o = t; o < NqC * NqC * NqF; o += T) {
    const int kc = o / (NqC * NqF), r = o - kc * NqC * NqF, jc = r / NqF, i = r - jc * NqF;
    double acc = 0.0;
    for (int m = 0; m < NqF; ++m) acc += sP[m * NqC + jc] * s1[(kc * NqF + m) * NqF + i];
    s2[o] = acc;
  }
  __syncthreads();
  for (int o = t; o < NpC; o += T) {
    const int kc = o / (NqC * NqC), r = o - kc * NqC * NqC, jc = r / NqC, ic = r - jc * NqC;
    double acc = 0.0;
    for (int m = 0; m < NqF; ++m) acc += sP[m * NqC + ic] * s2[(kc * NqC + jc) * NqF + m];
    const dlong id = G2LC[(size_t)e * NpC + o];
    if (id >= 0) atomicAdd(&qc[id], acc);
  }
}

// qf[owner copy] += (P x P x P) qc   (ellipticPartialPreconProlongateHex3D, okl/...ProlongateHex3D.okl:216-301,
// + ogs gather NoTrans + axpy).  idxN = fine GlobalToLocal restricted to the owner copies (-1 elsewhere).
__global__ void __launch_bounds__(128) prolongate_kernel(dlong Nelements, const dlong* __restrict__ elementList, int NqF,
                                                         int NqC, const dlong* __restrict__ idxNF,
                                                         const dlong* __restrict__ G2LC, const double* __restrict__ P,
                                                         const double* __restrict__ qc, double* __restrict__ qf) {
  extern __shared__ double sm[];
  const int NpF = NqF * NqF * NqF, NpC = NqC * NqC * NqC;
  double* s0 = sm;                         // [NqC][NqC][NqC]
  double* s1 = s0 + NpC;                   // [NqC][NqC][NqF]
  double* s2 = s1 + NqC * NqC * NqF;       // [NqC][NqF][NqF]
  double* sP = s2 + NqC * NqF * NqF;       // [NqF][NqC]
  const dlong e = elementList ? elementList[blockIdx.x] : (dlong)blockIdx.x;
  const int t = threadIdx.x, T = blockDim.x;
  for (int n = t; n < NqF * NqC; n += T) sP[n] = P[n];
  for (int n = t; n < NpC; n += T) {
    const dlong id = G2LC[(size_t)e * NpC + n];
    s0[n] = (id >= 0) ? qc[id] : 0.0;
  }
  __syncthreads();
  for (int o = t; o < NqC * NqC * NqF; o += T) {
    const int kc = o / (NqC * NqF), r = o - kc * NqC * NqF, jc = r / NqF, i = r - jc * NqF;
    double acc = 0.0;
    for (int m = 0; m < NqC; ++m) acc += sP[i * NqC + m] * s0[(kc * NqC + jc) * NqC + m];
    s1[o] = acc;
  }
  __syncthreads();
  for (int o = t; o < NqC * NqF * NqF; o += T) {
    const int kc = o / (NqF * NqF), r = o - kc * NqF * NqF, j = r / NqF, i = r - j * NqF;
    double acc = 0.0;
    for (int m = 0; m < NqC; ++m) acc += sP[j * NqC + m] * s1[(kc * NqC + m) * NqF + i];
    s2[o] = acc;
  }
  __syncthreads();
  for (int o = t; o < NpF; o += T) {
    const int k = o / (NqF * NqF), ji = o - k * NqF * NqF;
    const dlong id = idxNF[(size_t)e * NpF + o];
    if (id < 0) continue;
    double acc = 0.0;
    for (int m = 0; m < NqC; ++m) acc += sP[k * NqC + m] * s2[m * NqF * NqF + ji];
    qf[id] += acc;
  }
}

// ------------------------------------------------------------------ CSR kernels: 4 lanes per row
// kMode 0: z = beta*y + alpha*A x                      (SpMVcsr1/2)
// kMode 1: z = alpha*z + dInv .* (beta*y - A x)        (SmoothChebyshevCSR: z is r, y is b)
// kMode 2: z = alpha * dInv .* (y - A x)               (SmoothJacobiCSR: alpha = lambda, y = r, z = d)
template <int kMode>
__global__ void __launch_bounds__(kBlock) csr_kernel(dlong Nrows, const dlong* __restrict__ rowStarts,
                                                     const dlong* __restrict__ cols, const double* __restrict__ vals,
                                                     double alpha, double beta, const double* __restrict__ dInv,
                                                     const double* __restrict__ x, const double* y, double* z) {
  const int lane = threadIdx.x & 3;  // y may alias z (prolongate: x = x + P xC)
  for (dlong row = (blockIdx.x * kBlock + threadIdx.x) >> 2; row < ((Nrows + 63) & ~63); row += (gridDim.x * kBlock) >> 2) {
    double acc = 0.0;
    if (row < Nrows) {
      const dlong s = rowStarts[row], e = rowStarts[row + 1];
      for (dlong g = s + lane; g < e; g += 4) acc += vals[g] * x[cols[g]];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (row < Nrows && lane == 0) {
      if (kMode == 0) z[row] = ((beta != 0.0) ? beta * y[row] : 0.0) + alpha * acc;
      else if (kMode == 1) z[row] = ((alpha != 0.0) ? alpha * z[row] : 0.0) + dInv[row] * (((beta != 0.0) ? beta * y[row] : 0.0) - acc);
      else z[row] = alpha * dInv[row] * (y[row] - acc);
    }
  }
}
// SmoothChebyshevStart: r = dInv .* b ; d = lambda*r ; x = d
__global__ void __launch_bounds__(kBlock) csr_cheb_start_kernel(dlong N, double lambda, const double* __restrict__ dInv,
                                                                const double* __restrict__ b, double* __restrict__ r,
                                                                double* __restrict__ d, double* __restrict__ x) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double v = dInv[n] * b[n];
    r[n] = v;
    d[n] = lambda * v;
    x[n] = lambda * v;
  }
}
// SmoothChebyshevUpdate: d = alpha*d + beta*r ; x += d
__global__ void __launch_bounds__(kBlock) csr_cheb_update_kernel(dlong N, double alpha, double beta, int last,
                                                                 const double* __restrict__ r, double* __restrict__ d,
                                                                 double* __restrict__ x) {
  for (dlong n = blockIdx.x * kBlock + threadIdx.x; n < N; n += gridDim.x * kBlock) {
    const double dk = (alpha != 0.0) ? d[n] : 0.0;
    const double dn = alpha * dk + beta * r[n];
    if (!last) d[n] = dn;
    x[n] += dn;
  }
}
// dense coarse solve: x[n] = sum_m invAT[n + m*N] * rhs[m]   (one warp per row, fixed shuffle tree)
__global__ void __launch_bounds__(kBlock) dense_gemv_kernel(int N, const double* __restrict__ AT, const double* __restrict__ rhs,
                                                            double* __restrict__ x) {
  const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  double acc = 0.0;
  for (int m = lane; m < N; m += 32) acc += AT[(size_t)warp + (size_t)m * N] * rhs[m];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) x[warp] = acc;
}

}  // namespace

// ===================================================================== handles
struct libp_mglevel_s {
  libp_mglevel_desc_t d{};
  dlong Nrows = 0, Ncols = 0, NrowsC = 0, NcolsC = 0;
  dev_buf<double> wG;     // weightG extended to the halo entries (exchanged once)
  dev_buf<dlong> idxN;    // fine GlobalToLocal restricted to owner copies
  dev_buf<double> res, Ad, dd;  // smoother scratch (o_smootherResidual, o_smootherResidual2, o_smootherUpdate)
  void scratch() {
    if (!res.p) {
      res.alloc((size_t)Ncols); Ad.alloc((size_t)Ncols); dd.alloc((size_t)Ncols);
      for (dev_buf<double>* b : {&res, &Ad, &dd}) CUDA_CHECK(cudaMemset(b->p, 0, sizeof(double) * (size_t)Ncols));
    }
  }
  void op(double* x, double* Ax, cudaStream_t s) { d.fine->apply(x, Ax, false, nullptr, s); }
  void smooth(const double* rhs, double* x, bool x_is_zero, cudaStream_t s);
  void residual(const double* rhs, double* x, double* r, cudaStream_t s);
  void coarsen(double* x, double* Rx, cudaStream_t s);
  void prolongate(double* xC, double* x, cudaStream_t s);
};

struct libp_csr_s {
  dlong Nrows = 0, Ncols = 0, nnz = 0;
  dev_buf<dlong> rowStarts, cols;
  dev_buf<double> vals;
  template <int kMode>
  void run(double alpha, double beta, const double* dInv, const double* x, const double* y, double* z, cudaStream_t s) const {
    if (Nrows == 0) return;
    csr_kernel<kMode><<<vgrid((size_t)Nrows * 4), kBlock, 0, s>>>(Nrows, rowStarts.p, cols.p, vals.p, alpha, beta, dInv, x, y, z);
    CUDA_CHECK(cudaGetLastError());
  }
};

struct libp_amglevel_s {
  libp_csr_t A = nullptr, P = nullptr, R = nullptr;
  dev_buf<double> diagInv, sd, sr;  // scratch d, r (o_scratch + 0*Ncols, + 1*Ncols)
  int smoother = 1, ChebyshevIterations = 2;
  double lambda = 0, lambda0 = 0, lambda1 = 0;
  void smooth(const double* rhs, double* x, bool x_is_zero, cudaStream_t s);
};

struct libp_coarse_s {
  int N = 0;
  dev_buf<double> invAT;
};

struct libp_multigrid_s {
  libp_comm_t comm = nullptr;
  struct Level { int kind; libp_mglevel_t mg; libp_amglevel_t amg; dlong Nrows, Ncols; };
  std::vector<Level> levels;
  libp_coarse_t coarse = nullptr;
  std::vector<std::unique_ptr<dev_buf<double>>> rhs, x;  // per level (index 0 unused: caller's vectors)
  dev_buf<double> scratch;                               // residual of the current level
  dlong coarseN = 0;
  void prepare();
  void vcycle(int k, const double* rhs_k, double* x_k, cudaStream_t s);
};

// ===================================================================== MGLevel
void libp_mglevel_s::smooth(const double* rhs, double* x, bool x_is_zero, cudaStream_t s) {
  scratch();
  const dlong N = Nrows;
  const int g = vgrid((size_t)N);
  if (d.smoother == 1) {  // MGLevel::smoothJacobi (:137-154)
    if (x_is_zero) {
      LIBP_CHECK(libp_linalg_amxpy(N, 1.0, d.invDiagA, rhs, 0.0, x, s) == LIBP_SUCCESS, libp_last_error());
      return;
    }
    op(x, res.p, s);
    jacobi_update_kernel<<<g, kBlock, 0, s>>>(N, d.invDiagA, rhs, res.p, x);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  // MGLevel::smoothChebyshev (:156-206)
  const double theta = 0.5 * (d.lambda1 + d.lambda0), delta = 0.5 * (d.lambda1 - d.lambda0);
  const double invTheta = 1.0 / theta, sigma = theta / delta;
  double rho_n = 1.0 / sigma;
  if (x_is_zero) {
    cheb_start_kernel<false><<<g, kBlock, 0, s>>>(N, invTheta, d.invDiagA, rhs, res.p, dd.p);
  } else {
    op(x, res.p, s);
    cheb_start_kernel<true><<<g, kBlock, 0, s>>>(N, invTheta, d.invDiagA, rhs, res.p, dd.p);
  }
  CUDA_CHECK(cudaGetLastError());
  if (d.ChebyshevIterations == 0) {  // x (+)= d
    LIBP_CHECK(libp_linalg_axpy(N, 1.0, dd.p, x_is_zero ? 0.0 : 1.0, x, s) == LIBP_SUCCESS, libp_last_error());
    return;
  }
  for (int k = 0; k < d.ChebyshevIterations; ++k) {
    op(dd.p, Ad.p, s);
    const double rho_np1 = 1.0 / (2.0 * sigma - rho_n);
    const double rhoDivDelta = 2.0 * rho_np1 / delta;
    cheb_iter_kernel<<<g, kBlock, 0, s>>>(N, (x_is_zero && k == 0) ? 1 : 0, k == d.ChebyshevIterations - 1 ? 1 : 0,
                                          rho_np1 * rho_n, rhoDivDelta, d.invDiagA, Ad.p, res.p, dd.p, x);
    CUDA_CHECK(cudaGetLastError());
    rho_n = rho_np1;
  }
}

void libp_mglevel_s::residual(const double* rhs, double* x, double* r, cudaStream_t s) {
  op(x, r, s);
  LIBP_CHECK(libp_linalg_axpy(Nrows, 1.0, rhs, -1.0, r, s) == LIBP_SUCCESS, libp_last_error());
}

void libp_mglevel_s::coarsen(double* x, double* Rx, cudaStream_t s) {
  libp_elliptic_s& F = *d.fine;
  libp_elliptic_s& C = *d.coarse;
  libp_ogs_s& ogsF = *F.d.ogsMasked;
  libp_ogs_s& ogsC = *C.d.ogsMasked;
  const size_t smem = sizeof(double) * ((size_t)d.NqF * d.NqF * d.NqF + (size_t)d.NqC * d.NqF * d.NqF +
                                        (size_t)d.NqC * d.NqC * d.NqF + (size_t)d.NqF * d.NqC);
  CUDA_CHECK(cudaMemsetAsync(Rx, 0, sizeof(double) * (size_t)(ogsC.NlocalT + ogsC.NhaloT), s));
  auto launch = [&](dlong n, const dlong* list) {
    if (n <= 0) return;
    coarsen_kernel<<<n, 128, smem, s>>>(n, list, d.NqF, d.NqC, F.d.GlobalToLocal, C.d.GlobalToLocal, d.P, wG.p, x, Rx);
    CUDA_CHECK(cudaGetLastError());
  };
  halo_start_f64(ogsF, x, s);
  launch(F.d.NlocalGatherElements, F.d.localGatherElementList);
  halo_finish_f64(ogsF, x, s);
  launch(F.d.NglobalGatherElements, F.d.globalGatherElementList);
  halo_combine_start_f64(ogsC, Rx, s);
  halo_combine_finish_f64(ogsC, Rx, s);
}

void libp_mglevel_s::prolongate(double* xC, double* x, cudaStream_t s) {
  libp_elliptic_s& F = *d.fine;
  libp_elliptic_s& C = *d.coarse;
  libp_ogs_s& ogsC = *C.d.ogsMasked;
  const size_t smem = sizeof(double) * ((size_t)d.NqC * d.NqC * d.NqC + (size_t)d.NqC * d.NqC * d.NqF +
                                        (size_t)d.NqC * d.NqF * d.NqF + (size_t)d.NqF * d.NqC);
  auto launch = [&](dlong n, const dlong* list) {
    if (n <= 0) return;
    prolongate_kernel<<<n, 128, smem, s>>>(n, list, d.NqF, d.NqC, idxN.p, C.d.GlobalToLocal, d.P, xC, x);
    CUDA_CHECK(cudaGetLastError());
  };
  // the coarse mesh has the same elements (and element lists) as the fine one: only the degree differs
  halo_start_f64(ogsC, xC, s);
  launch(F.d.NlocalGatherElements, F.d.localGatherElementList);
  halo_finish_f64(ogsC, xC, s);
  launch(F.d.NglobalGatherElements, F.d.globalGatherElementList);
}

extern "C" int libp_mglevel_create(const libp_mglevel_desc_t* desc, libp_mglevel_t* level) {
  LIBP_API_BEGIN
  LIBP_CHECK(desc && level, "null argument");
  LIBP_CHECK(desc->fine && desc->coarse && desc->P && desc->invDiagA && desc->weightG, "null member");
  LIBP_CHECK(desc->NqF == desc->fine->d.Nq && desc->NqC == desc->coarse->d.Nq && desc->NqC < desc->NqF, "degrees do not match the operators");
  LIBP_CHECK(desc->fine->d.Nelements == desc->coarse->d.Nelements, "fine and coarse operators must share the elements");
  LIBP_CHECK(desc->fine->d.mode == 1, "multigrid levels need the fused-gather operator (mode 1)");
  LIBP_CHECK(desc->smoother == 1 || desc->smoother == 2, "smoother must be 1 (JACOBI) or 2 (CHEBYSHEV)");
  std::unique_ptr<libp_mglevel_s> L(new libp_mglevel_s());
  L->d = *desc;
  libp_ogs_s& ogsF = *desc->fine->d.ogsMasked;
  libp_ogs_s& ogsC = *desc->coarse->d.ogsMasked;
  L->Nrows = ogsF.Ngather;
  L->Ncols = ogsF.NlocalT + ogsF.NhaloT;
  L->NrowsC = ogsC.Ngather;
  L->NcolsC = ogsC.NlocalT + ogsC.NhaloT;
  // weights on the halo entries: one exchange at setup instead of a weighted copy per coarsen
  L->wG.alloc((size_t)L->Ncols);
  CUDA_CHECK(cudaMemset(L->wG.p, 0, sizeof(double) * (size_t)L->Ncols));
  CUDA_CHECK(cudaMemcpy(L->wG.p, desc->weightG, sizeof(double) * (size_t)L->Nrows, cudaMemcpyDeviceToDevice));
  halo_start_f64(ogsF, L->wG.p, nullptr);
  halo_finish_f64(ogsF, L->wG.p, nullptr);
  CUDA_CHECK(cudaDeviceSynchronize());
  // owner-copy map of the fine level (ogs NoTrans gather = only the positive copy contributes)
  const size_t Nloc = (size_t)desc->fine->d.Nelements * desc->NqF * desc->NqF * desc->NqF;
  LIBP_CHECK((size_t)ogsF.N == Nloc, "fine ogs does not match the element count");
  std::vector<dlong> idx(Nloc, -1);
  auto mark = [&](const OgsOperator& op, dlong rowOffset) {
    for (dlong r = 0; r < op.NrowsT; ++r)
      for (dlong g = op.rowStartsN[r]; g < op.rowStartsN[r + 1]; ++g) idx[op.colIdsN[g]] = rowOffset + r;
  };
  mark(ogsF.gatherLocal, 0);
  mark(ogsF.gatherHalo, ogsF.NlocalT);
  // only owned rows are updated (row < Ngather); non-owned halo rows have no N-map columns by construction
  L->idxN.upload(idx);
  const size_t F = desc->NqF, Cc = desc->NqC;
  const size_t smem = sizeof(double) * (F * F * F + Cc * F * F + Cc * Cc * F + Cc * Cc * Cc + F * Cc);
  LIBP_CHECK(smem <= 48 * 1024, "transfer kernels need more shared memory than 48 KB");
  *level = L.release();
  LIBP_API_END
}
extern "C" int libp_mglevel_free(libp_mglevel_t level) {
  LIBP_API_BEGIN
  delete level;
  LIBP_API_END
}
extern "C" int libp_mglevel_operator(libp_mglevel_t L, libp_dfloat* x, libp_dfloat* Ax, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && x && Ax, "null argument");
  L->op(x, Ax, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_mglevel_smooth(libp_mglevel_t L, const libp_dfloat* rhs, libp_dfloat* x, int x_is_zero, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && rhs && x, "null argument");
  L->smooth(rhs, x, x_is_zero != 0, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_mglevel_residual(libp_mglevel_t L, const libp_dfloat* rhs, libp_dfloat* x, libp_dfloat* res, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && rhs && x && res, "null argument");
  L->residual(rhs, x, res, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_mglevel_coarsen(libp_mglevel_t L, libp_dfloat* x, libp_dfloat* Rx, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && x && Rx, "null argument");
  L->coarsen(x, Rx, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_mglevel_prolongate(libp_mglevel_t L, libp_dfloat* xC, libp_dfloat* x, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && xC && x, "null argument");
  L->prolongate(xC, x, as_stream(stream));
  LIBP_API_END
}

// ===================================================================== CSR levels
extern "C" int libp_csr_create(libp_dlong Nrows, libp_dlong Ncols, libp_dlong nnz, const libp_dlong* rowStarts,
                               const libp_dlong* cols, const libp_dfloat* vals, libp_dlong offd_nnz, libp_csr_t* csr) {
  LIBP_API_BEGIN
  LIBP_CHECK(csr && Nrows >= 0 && Ncols >= 0 && nnz >= 0, "bad argument");
  LIBP_CHECK(offd_nnz == 0, "multi-rank CSR levels (non-empty off-diagonal block) are not supported in this round");
  LIBP_CHECK(Nrows == 0 || (rowStarts && (nnz == 0 || (cols && vals))), "null array");
  LIBP_CHECK(Nrows == 0 || rowStarts[Nrows] == nnz, "rowStarts does not match nnz");
  std::unique_ptr<libp_csr_s> A(new libp_csr_s());
  A->Nrows = Nrows; A->Ncols = Ncols; A->nnz = nnz;
  if (Nrows) {
    A->rowStarts.upload(rowStarts, (size_t)Nrows + 1);
    A->cols.upload(cols, (size_t)nnz);
    A->vals.upload(vals, (size_t)nnz);
  }
  *csr = A.release();
  LIBP_API_END
}
extern "C" int libp_csr_free(libp_csr_t csr) {
  LIBP_API_BEGIN
  delete csr;
  LIBP_API_END
}
extern "C" int libp_csr_spmv(libp_csr_t A, libp_dfloat alpha, const libp_dfloat* x, libp_dfloat beta, const libp_dfloat* y,
                             libp_dfloat* z, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(A && x && z && (beta == 0.0 || y), "null argument");
  A->run<0>(alpha, beta, nullptr, x, y, z, as_stream(stream));
  LIBP_API_END
}

void libp_amglevel_s::smooth(const double* rhs, double* x, bool x_is_zero, cudaStream_t s) {
  const dlong N = A->Nrows;
  if (N == 0) return;
  const int g = vgrid((size_t)N);
  if (smoother == 0) {  // parCSR::smoothDampedJacobi (parAlmondAMGSmoother.cpp:35-66)
    if (x_is_zero) {
      LIBP_CHECK(libp_linalg_amxpy(N, lambda, diagInv.p, rhs, 0.0, x, s) == LIBP_SUCCESS, libp_last_error());
      return;
    }
    A->run<2>(lambda, 0.0, diagInv.p, x, rhs, sd.p, s);
    LIBP_CHECK(libp_linalg_axpy(N, 1.0, sd.p, 1.0, x, s) == LIBP_SUCCESS, libp_last_error());
    return;
  }
  // parCSR::smoothChebyshev (parAlmondAMGSmoother.cpp:68-160)
  const double theta = 0.5 * (lambda1 + lambda0), delta = 0.5 * (lambda1 - lambda0);
  const double invTheta = 1.0 / theta, sigma = theta / delta;
  double rho_n = 1.0 / sigma;
  if (x_is_zero) {
    csr_cheb_start_kernel<<<g, kBlock, 0, s>>>(N, invTheta, diagInv.p, rhs, sr.p, sd.p, x);
  } else {
    A->run<1>(0.0, 1.0, diagInv.p, x, rhs, sr.p, s);
    csr_cheb_update_kernel<<<g, kBlock, 0, s>>>(N, 0.0, invTheta, ChebyshevIterations == 0 ? 1 : 0, sr.p, sd.p, x);
  }
  CUDA_CHECK(cudaGetLastError());
  for (int k = 0; k < ChebyshevIterations; ++k) {
    A->run<1>(1.0, 0.0, diagInv.p, sd.p, rhs, sr.p, s);
    const double rho_np1 = 1.0 / (2.0 * sigma - rho_n);
    csr_cheb_update_kernel<<<g, kBlock, 0, s>>>(N, rho_np1 * rho_n, 2.0 * rho_np1 / delta, k == ChebyshevIterations - 1 ? 1 : 0,
                                                sr.p, sd.p, x);
    CUDA_CHECK(cudaGetLastError());
    rho_n = rho_np1;
  }
}

extern "C" int libp_amglevel_create(libp_csr_t A, libp_csr_t P, libp_csr_t R, const libp_dfloat* diagInv, int smoother,
                                    libp_dfloat lambda, libp_dfloat lambda0, libp_dfloat lambda1, int ChebyshevIterations,
                                    libp_amglevel_t* level) {
  LIBP_API_BEGIN
  LIBP_CHECK(A && level && (A->Nrows == 0 || diagInv), "null argument");
  LIBP_CHECK(smoother == 0 || smoother == 1, "smoother must be 0 (DAMPED_JACOBI) or 1 (CHEBYSHEV)");
  LIBP_CHECK(!P || P->Nrows == A->Nrows, "P must have A's rows");
  LIBP_CHECK(!R || R->Ncols >= A->Nrows, "R must have A's rows as columns");
  std::unique_ptr<libp_amglevel_s> L(new libp_amglevel_s());
  L->A = A; L->P = P; L->R = R;
  L->smoother = smoother; L->lambda = lambda; L->lambda0 = lambda0; L->lambda1 = lambda1;
  L->ChebyshevIterations = ChebyshevIterations;
  if (A->Nrows) L->diagInv.upload(diagInv, (size_t)A->Nrows);
  const size_t n = (size_t)std::max(A->Ncols, A->Nrows);
  L->sd.alloc(std::max<size_t>(n, 1)); L->sr.alloc(std::max<size_t>(n, 1));
  CUDA_CHECK(cudaMemset(L->sd.p, 0, sizeof(double) * std::max<size_t>(n, 1)));
  CUDA_CHECK(cudaMemset(L->sr.p, 0, sizeof(double) * std::max<size_t>(n, 1)));
  *level = L.release();
  LIBP_API_END
}
extern "C" int libp_amglevel_free(libp_amglevel_t level) {
  LIBP_API_BEGIN
  delete level;
  LIBP_API_END
}
extern "C" int libp_amglevel_smooth(libp_amglevel_t L, const libp_dfloat* rhs, libp_dfloat* x, int x_is_zero, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && rhs && x, "null argument");
  L->smooth(rhs, x, x_is_zero != 0, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_amglevel_residual(libp_amglevel_t L, const libp_dfloat* rhs, const libp_dfloat* x, libp_dfloat* res, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && rhs && x && res, "null argument");
  L->A->run<0>(-1.0, 1.0, nullptr, x, rhs, res, as_stream(stream));  // A.SpMV(-1, x, 1, rhs, res)
  LIBP_API_END
}
extern "C" int libp_amglevel_coarsen(libp_amglevel_t L, const libp_dfloat* x, libp_dfloat* Rx, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && L->R && x && Rx, "null argument (level has no R)");
  L->R->run<0>(1.0, 0.0, nullptr, x, nullptr, Rx, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_amglevel_prolongate(libp_amglevel_t L, const libp_dfloat* xC, libp_dfloat* x, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(L && L->P && xC && x, "null argument (level has no P)");
  L->P->run<0>(1.0, 1.0, nullptr, xC, x, x, as_stream(stream));  // P.SpMV(1, xC, 1, x)
  LIBP_API_END
}

extern "C" int libp_coarse_exact_create(int N, const libp_dfloat* diagInvAT, libp_coarse_t* coarse) {
  LIBP_API_BEGIN
  LIBP_CHECK(coarse && N >= 0 && (N == 0 || diagInvAT), "bad argument");
  std::unique_ptr<libp_coarse_s> c(new libp_coarse_s());
  c->N = N;
  if (N) c->invAT.upload(diagInvAT, (size_t)N * N);
  *coarse = c.release();
  LIBP_API_END
}
extern "C" int libp_coarse_free(libp_coarse_t coarse) {
  LIBP_API_BEGIN
  delete coarse;
  LIBP_API_END
}
extern "C" int libp_coarse_solve(libp_coarse_t c, const libp_dfloat* rhs, libp_dfloat* x, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(c && (c->N == 0 || (rhs && x)), "null argument");
  if (c->N) {
    dense_gemv_kernel<<<(c->N * 32 + kBlock - 1) / kBlock, kBlock, 0, as_stream(stream)>>>(c->N, c->invAT.p, rhs, x);
    CUDA_CHECK(cudaGetLastError());
  }
  LIBP_API_END
}

// ===================================================================== multigrid_t
void libp_multigrid_s::prepare() {
  if (rhs.size() == levels.size() + 1) return;
  rhs.clear(); x.clear();
  dlong maxCols = 1;
  for (size_t k = 0; k <= levels.size(); ++k) {
    rhs.emplace_back(new dev_buf<double>());
    x.emplace_back(new dev_buf<double>());
    dlong ncols = (k < levels.size()) ? levels[k].Ncols : coarseN;
    if (k > 0 && levels[k - 1].kind == 0) ncols = std::max(ncols, levels[k - 1].mg->NcolsC);  // coarsen target
    ncols = std::max<dlong>(ncols, 1);
    maxCols = std::max(maxCols, ncols);
    if (k > 0) {
      rhs[k]->alloc((size_t)ncols); x[k]->alloc((size_t)ncols);
      CUDA_CHECK(cudaMemset(rhs[k]->p, 0, sizeof(double) * (size_t)ncols));
      CUDA_CHECK(cudaMemset(x[k]->p, 0, sizeof(double) * (size_t)ncols));
    }
  }
  scratch.alloc((size_t)maxCols);
  CUDA_CHECK(cudaMemset(scratch.p, 0, sizeof(double) * (size_t)maxCols));
}

// multigrid_t::vcycle (libs/parAlmond/parAlmondVcycle.cpp:34-60)
void libp_multigrid_s::vcycle(int k, const double* rhs_k, double* x_k, cudaStream_t s) {
  if (k == (int)levels.size()) {
    LIBP_CHECK(coarse != nullptr, "multigrid has no coarse solver");
    if (coarse->N) {
      dense_gemv_kernel<<<(coarse->N * 32 + kBlock - 1) / kBlock, kBlock, 0, s>>>(coarse->N, coarse->invAT.p, rhs_k, x_k);
      CUDA_CHECK(cudaGetLastError());
    }
    return;
  }
  Level& L = levels[k];
  double* rhsC = rhs[k + 1]->p;
  double* xC = x[k + 1]->p;
  double* res = scratch.p;
  if (L.kind == 0) {
    L.mg->smooth(rhs_k, x_k, true, s);
    L.mg->residual(rhs_k, x_k, res, s);
    L.mg->coarsen(res, rhsC, s);
    vcycle(k + 1, rhsC, xC, s);
    L.mg->prolongate(xC, x_k, s);
    L.mg->smooth(rhs_k, x_k, false, s);
  } else {
    libp_amglevel_s& A = *L.amg;
    A.smooth(rhs_k, x_k, true, s);
    A.A->run<0>(-1.0, 1.0, nullptr, x_k, rhs_k, res, s);
    LIBP_CHECK(A.R && A.P, "CSR level without transfer operators above the coarse solver");
    A.R->run<0>(1.0, 0.0, nullptr, res, nullptr, rhsC, s);
    vcycle(k + 1, rhsC, xC, s);
    A.P->run<0>(1.0, 1.0, nullptr, xC, x_k, x_k, s);
    A.smooth(rhs_k, x_k, false, s);
  }
}

extern "C" int libp_multigrid_create(libp_comm_t comm, libp_multigrid_t* mg) {
  LIBP_API_BEGIN
  LIBP_CHECK(mg, "null argument");
  auto* m = new libp_multigrid_s();
  m->comm = comm;
  *mg = m;
  LIBP_API_END
}
extern "C" int libp_multigrid_add_mglevel(libp_multigrid_t mg, libp_mglevel_t level) {
  LIBP_API_BEGIN
  LIBP_CHECK(mg && level, "null argument");
  LIBP_CHECK(mg->levels.empty() || mg->levels.back().kind == 0, "matrix-free levels must come before the CSR levels");
  mg->levels.push_back({0, level, nullptr, level->Nrows, level->Ncols});
  mg->rhs.clear();
  LIBP_API_END
}
extern "C" int libp_multigrid_add_amglevel(libp_multigrid_t mg, libp_amglevel_t level) {
  LIBP_API_BEGIN
  LIBP_CHECK(mg && level, "null argument");
  mg->levels.push_back({1, nullptr, level, level->A->Nrows, std::max(level->A->Ncols, level->A->Nrows)});
  mg->rhs.clear();
  LIBP_API_END
}
extern "C" int libp_multigrid_set_coarse(libp_multigrid_t mg, libp_coarse_t coarse) {
  LIBP_API_BEGIN
  LIBP_CHECK(mg && coarse, "null argument");
  mg->coarse = coarse;
  mg->coarseN = coarse->N;
  mg->rhs.clear();
  LIBP_API_END
}
extern "C" int libp_multigrid_vcycle(libp_multigrid_t mg, const libp_dfloat* rhs, libp_dfloat* x, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(mg && rhs && x, "null argument");
  mg->prepare();
  mg->vcycle(0, rhs, x, as_stream(stream));
  LIBP_API_END
}
extern "C" int libp_multigrid_free(libp_multigrid_t mg) {
  LIBP_API_BEGIN
  delete mg;
  LIBP_API_END
}

extern "C" int libp_precon_multigrid_create(libp_multigrid_t mg, int allNeumann, libp_hlong NglobalDofs, libp_comm_t comm,
                                            libp_precon_t* precon) {
  LIBP_API_BEGIN
  LIBP_CHECK(mg && precon, "null argument");
  LIBP_CHECK(!mg->levels.empty() || mg->coarse, "empty multigrid hierarchy");
  auto* p = new libp_precon_s();
  p->kind = 2;
  p->N = mg->levels.empty() ? mg->coarseN : mg->levels[0].Nrows;
  p->allNeumann = allNeumann;
  p->NglobalDofs = NglobalDofs;
  p->comm = comm;
  p->impl = mg;
  *precon = p;
  LIBP_API_END
}

namespace libp_b200 {
void multigrid_apply(void* impl, const dfloat* r, dfloat* Mr, cudaStream_t s) {
  auto* mg = static_cast<libp_multigrid_s*>(impl);
  mg->prepare();
  mg->vcycle(0, r, Mr, s);
}
}  // namespace libp_b200
