// Device-side setup kernels of the hexahedral elliptic path: what the reference computes on the
// host before the first solve, here in HBM directly (at 64^3 elements of degree 7 these arrays are 9 GB).
//   libp_mesh_physical_nodes_hex3d       libs/mesh/meshPhysicalNodesHex3D.cpp   (trilinear nodes from EX, EY, EZ)
//   libp_mesh_geometric_factors_hex3d    libs/mesh/meshGeometricFactorsHex3D.cpp:94-174   (ggeo, wJ, optional vgeo)
//   libp_elliptic_build_diagonal_hex3d   solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:998-1057
//   libp_ax_trilinear_hex3d              solvers/elliptic/okl/ellipticAxHex3D.okl:440-627 (geometry on the fly)
// Same arithmetic, same summation order as the reference loops (sums over m ascending), so results agree to
// rounding (FMA contraction aside).
#include <algorithm>

#include "common.hpp"

using namespace libp_b200;

namespace {

constexpr int kMaxNq = 9;

// ---------------------------------------------------------------- physical nodes
// x(e, n) = sum_v EX[e][v] * shape_v(r_n, s_n, t_n), vertex order of meshPhysicalNodesHex3D / the OKL kernel
__global__ void __launch_bounds__(128) physical_nodes_kernel(int Nq, dlong Nelements, const dfloat* __restrict__ EX,
                                                             const dfloat* __restrict__ EY, const dfloat* __restrict__ EZ,
                                                             const dfloat* __restrict__ gllz, dfloat* __restrict__ x,
                                                             dfloat* __restrict__ y, dfloat* __restrict__ z) {
  const dlong e = blockIdx.x;
  const int Np = Nq * Nq * Nq;
  __shared__ dfloat s_v[3][8];
  __shared__ dfloat s_z[kMaxNq];
  if (threadIdx.x < 8) {
    s_v[0][threadIdx.x] = EX[(size_t)e * 8 + threadIdx.x];
    s_v[1][threadIdx.x] = EY[(size_t)e * 8 + threadIdx.x];
    s_v[2][threadIdx.x] = EZ[(size_t)e * 8 + threadIdx.x];
  }
  if (threadIdx.x < Nq) s_z[threadIdx.x] = gllz[threadIdx.x];
  __syncthreads();
  for (int n = threadIdx.x; n < Np; n += blockDim.x) {
    const int i = n % Nq, j = (n / Nq) % Nq, k = n / (Nq * Nq);
    const dfloat rn = s_z[i], sn = s_z[j], tn = s_z[k];
    const dfloat w[8] = {0.125 * (1 - rn) * (1 - sn) * (1 - tn), 0.125 * (1 + rn) * (1 - sn) * (1 - tn),
                         0.125 * (1 + rn) * (1 + sn) * (1 - tn), 0.125 * (1 - rn) * (1 + sn) * (1 - tn),
                         0.125 * (1 - rn) * (1 - sn) * (1 + tn), 0.125 * (1 + rn) * (1 - sn) * (1 + tn),
                         0.125 * (1 + rn) * (1 + sn) * (1 + tn), 0.125 * (1 - rn) * (1 + sn) * (1 + tn)};
    dfloat xn = 0, yn = 0, zn = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) { xn += w[v] * s_v[0][v]; yn += w[v] * s_v[1][v]; zn += w[v] * s_v[2][v]; }
    x[(size_t)e * Np + n] = xn;
    y[(size_t)e * Np + n] = yn;
    z[(size_t)e * Np + n] = zn;
  }
}

// ---------------------------------------------------------------- geometric factors
// one block per element; the element's x, y, z in shared memory; thread per node
__global__ void __launch_bounds__(256) geometric_factors_kernel(int Nq, dlong Nelements, const dfloat* __restrict__ x,
                                                                const dfloat* __restrict__ y, const dfloat* __restrict__ z,
                                                                const dfloat* __restrict__ D, const dfloat* __restrict__ gllw,
                                                                dfloat* __restrict__ ggeo, dfloat* __restrict__ wJ,
                                                                dfloat* __restrict__ vgeo, int* __restrict__ badJ) {
  extern __shared__ dfloat sm[];
  const int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  dfloat* s_x = sm; dfloat* s_y = s_x + Np; dfloat* s_z = s_y + Np; dfloat* s_D = s_z + Np; dfloat* s_w = s_D + Nq2;
  const dlong e = blockIdx.x;
  for (int n = threadIdx.x; n < Np; n += blockDim.x) {
    s_x[n] = x[(size_t)e * Np + n]; s_y[n] = y[(size_t)e * Np + n]; s_z[n] = z[(size_t)e * Np + n];
  }
  for (int n = threadIdx.x; n < Nq2; n += blockDim.x) s_D[n] = D[n];
  if (threadIdx.x < Nq) s_w[threadIdx.x] = gllw[threadIdx.x];
  __syncthreads();
  for (int n = threadIdx.x; n < Np; n += blockDim.x) {
    const int i = n % Nq, j = (n / Nq) % Nq, k = n / Nq2;
    dfloat xr = 0, xs = 0, xt = 0, yr = 0, ys = 0, yt = 0, zr = 0, zs = 0, zt = 0;
    for (int m = 0; m < Nq; ++m) {
      const int idr = k * Nq2 + j * Nq + m, ids = k * Nq2 + m * Nq + i, idt = m * Nq2 + j * Nq + i;
      const dfloat Dr = s_D[i * Nq + m], Ds = s_D[j * Nq + m], Dt = s_D[k * Nq + m];
      xr += Dr * s_x[idr]; xs += Ds * s_x[ids]; xt += Dt * s_x[idt];
      yr += Dr * s_y[idr]; ys += Ds * s_y[ids]; yt += Dt * s_y[idt];
      zr += Dr * s_z[idr]; zs += Ds * s_z[ids]; zt += Dt * s_z[idt];
    }
    const dfloat J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt);
    if (J < 1e-12) atomicExch(badJ, (int)(e + 1));  // reference: LIBP_ABORT("Negative J found at element e")
    const dfloat rx = (ys * zt - zs * yt) / J, ry = -(xs * zt - zs * xt) / J, rz = (xs * yt - ys * xt) / J;
    const dfloat sx = -(yr * zt - zr * yt) / J, sy = (xr * zt - zr * xt) / J, sz = -(xr * yt - yr * xt) / J;
    const dfloat tx = (yr * zs - zr * ys) / J, ty = -(xr * zs - zr * xs) / J, tz = (xr * ys - yr * xs) / J;
    const dfloat JW = J * s_w[i] * s_w[j] * s_w[k];
    dfloat* g = ggeo + (size_t)e * 6 * Np + n;
    g[0 * Np] = JW * (rx * rx + ry * ry + rz * rz);
    g[1 * Np] = JW * (rx * sx + ry * sy + rz * sz);
    g[2 * Np] = JW * (rx * tx + ry * ty + rz * tz);
    g[3 * Np] = JW * (sx * sx + sy * sy + sz * sz);
    g[4 * Np] = JW * (sx * tx + sy * ty + sz * tz);
    g[5 * Np] = JW * (tx * tx + ty * ty + tz * tz);
    wJ[(size_t)e * Np + n] = JW;
    if (vgeo) {  // RXID..TZID = 0..8, JID 9, JWID 10, IJWID 11 (meshGeometricFactorsHex3D.cpp:35-58)
      dfloat* v = vgeo + (size_t)e * 12 * Np + n;
      v[0 * Np] = rx; v[1 * Np] = ry; v[2 * Np] = rz; v[3 * Np] = sx; v[4 * Np] = sy; v[5 * Np] = sz;
      v[6 * Np] = tx; v[7 * Np] = ty; v[8 * Np] = tz; v[9 * Np] = J; v[10 * Np] = JW; v[11 * Np] = 1.0 / JW;
    }
  }
}

// ---------------------------------------------------------------- operator diagonal (element-local, before the gather)
__global__ void __launch_bounds__(256) build_diagonal_kernel(int Nq, dlong Nelements, const dfloat* __restrict__ ggeo,
                                                             const dfloat* __restrict__ wJ, const dfloat* __restrict__ D,
                                                             const int* __restrict__ mapB, dfloat lambda, dfloat boost,
                                                             dfloat* __restrict__ A) {
  extern __shared__ dfloat sm[];
  const int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  dfloat* s_D = sm;                 // [Nq][Nq]
  dfloat* s_G = s_D + Nq2;          // G00, G11, G22 of the element: [3][Np]
  const dlong e = blockIdx.x;
  const dfloat* g = ggeo + (size_t)e * 6 * Np;
  for (int n = threadIdx.x; n < Nq2; n += blockDim.x) s_D[n] = D[n];
  for (int n = threadIdx.x; n < Np; n += blockDim.x) {
    s_G[n] = g[0 * Np + n]; s_G[Np + n] = g[3 * Np + n]; s_G[2 * Np + n] = g[5 * Np + n];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < Np; n += blockDim.x) {
    const int nx = n % Nq, ny = (n / Nq) % Nq, nz = n / Nq2;
    dfloat a;
    if (mapB[(size_t)e * Np + n] != 1) {
      a = 0;
      a += 2 * g[1 * Np + n] * s_D[nx + nx * Nq] * s_D[ny + ny * Nq];
      a += 2 * g[2 * Np + n] * s_D[nx + nx * Nq] * s_D[nz + nz * Nq];
      a += 2 * g[4 * Np + n] * s_D[ny + ny * Nq] * s_D[nz + nz * Nq];
      for (int k = 0; k < Nq; ++k) a += s_G[k + ny * Nq + nz * Nq2] * s_D[nx + k * Nq] * s_D[nx + k * Nq];
      for (int k = 0; k < Nq; ++k) a += s_G[Np + nx + k * Nq + nz * Nq2] * s_D[ny + k * Nq] * s_D[ny + k * Nq];
      for (int k = 0; k < Nq; ++k) a += s_G[2 * Np + nx + ny * Nq + k * Nq2] * s_D[nz + k * Nq] * s_D[nz + k * Nq];
      a += wJ[(size_t)e * Np + n] * lambda;
      a += boost;  // allNeumannPenalty * allNeumannScale^2, 0 otherwise
    } else {
      a = 1;  // "just put a 1 so A is invertable"
    }
    A[(size_t)e * Np + n] = a;
  }
}

// ---------------------------------------------------------------- trilinear element map: geometry on the fly
// AqL[e] = A_e q[e] with the geometric factors recomputed at every node from the 8 vertices of the element
// (ELEMENT MAP = TRILINEAR).  Traffic per node: q in, Aq out (+ 4 B of connectivity when gathered) instead of
// 48-56 B of stored factors; the price is ~75 FP64 operations and one division per node.  One block per element,
// thread (i, j) owns the k-pencil; layout and slab loop as the reference kernel, D rows in registers.
template <int Nq, bool kGather>
__global__ void __launch_bounds__(Nq * Nq) ax_trilinear_kernel(dlong Nelements, const dlong* __restrict__ elementList,
                                                              const dlong* __restrict__ G2L, const dfloat* __restrict__ EXYZ,
                                                              const dfloat* __restrict__ gllzw, const dfloat* __restrict__ D,
                                                              dfloat lambda, const dfloat* __restrict__ q,
                                                              dfloat* __restrict__ Aq) {
  constexpr int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  __shared__ dfloat s_D[Nq][Nq + 1];
  __shared__ dfloat s_q[Nq][Nq + 1], s_Gqr[Nq][Nq + 1], s_Gqs[Nq][Nq + 1];
  __shared__ dfloat s_zw[2][Nq];
  __shared__ dfloat s_v[3][8];
  const int t = threadIdx.x, j = t / Nq, i = t - j * Nq;
  const dlong e = elementList ? elementList[blockIdx.x] : (dlong)blockIdx.x;
  s_D[j][i] = D[j * Nq + i];
  if (t < 2 * Nq) s_zw[t / Nq][t % Nq] = gllzw[t];
  for (int n = t; n < 24; n += Nq2) s_v[n / 8][n % 8] = EXYZ[(size_t)e * 24 + n];
  dfloat r_q[Nq], r_Aq[Nq];
  dlong r_id[Nq];
  const size_t base = (size_t)e * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    if (kGather) {
      r_id[k] = G2L[base + k * Nq2];
      r_q[k] = r_id[k] >= 0 ? q[r_id[k]] : 0.0;
    } else {
      r_id[k] = 0;
      r_q[k] = q[base + k * Nq2];
    }
    r_Aq[k] = 0;
  }
  __syncthreads();
  const dfloat rn = s_zw[0][i], sn = s_zw[0][j];
  const dfloat* xe = s_v[0]; const dfloat* ye = s_v[1]; const dfloat* ze = s_v[2];
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const dfloat tn = s_zw[0][k];
    const dfloat xr = 0.125 * ((1 - tn) * (1 - sn) * (xe[1] - xe[0]) + (1 - tn) * (1 + sn) * (xe[2] - xe[3]) + (1 + tn) * (1 - sn) * (xe[5] - xe[4]) + (1 + tn) * (1 + sn) * (xe[6] - xe[7]));
    const dfloat xs = 0.125 * ((1 - tn) * (1 - rn) * (xe[3] - xe[0]) + (1 - tn) * (1 + rn) * (xe[2] - xe[1]) + (1 + tn) * (1 - rn) * (xe[7] - xe[4]) + (1 + tn) * (1 + rn) * (xe[6] - xe[5]));
    const dfloat xt = 0.125 * ((1 - rn) * (1 - sn) * (xe[4] - xe[0]) + (1 + rn) * (1 - sn) * (xe[5] - xe[1]) + (1 + rn) * (1 + sn) * (xe[6] - xe[2]) + (1 - rn) * (1 + sn) * (xe[7] - xe[3]));
    const dfloat yr = 0.125 * ((1 - tn) * (1 - sn) * (ye[1] - ye[0]) + (1 - tn) * (1 + sn) * (ye[2] - ye[3]) + (1 + tn) * (1 - sn) * (ye[5] - ye[4]) + (1 + tn) * (1 + sn) * (ye[6] - ye[7]));
    const dfloat ys = 0.125 * ((1 - tn) * (1 - rn) * (ye[3] - ye[0]) + (1 - tn) * (1 + rn) * (ye[2] - ye[1]) + (1 + tn) * (1 - rn) * (ye[7] - ye[4]) + (1 + tn) * (1 + rn) * (ye[6] - ye[5]));
    const dfloat yt = 0.125 * ((1 - rn) * (1 - sn) * (ye[4] - ye[0]) + (1 + rn) * (1 - sn) * (ye[5] - ye[1]) + (1 + rn) * (1 + sn) * (ye[6] - ye[2]) + (1 - rn) * (1 + sn) * (ye[7] - ye[3]));
    const dfloat zr = 0.125 * ((1 - tn) * (1 - sn) * (ze[1] - ze[0]) + (1 - tn) * (1 + sn) * (ze[2] - ze[3]) + (1 + tn) * (1 - sn) * (ze[5] - ze[4]) + (1 + tn) * (1 + sn) * (ze[6] - ze[7]));
    const dfloat zs = 0.125 * ((1 - tn) * (1 - rn) * (ze[3] - ze[0]) + (1 - tn) * (1 + rn) * (ze[2] - ze[1]) + (1 + tn) * (1 - rn) * (ze[7] - ze[4]) + (1 + tn) * (1 + rn) * (ze[6] - ze[5]));
    const dfloat zt = 0.125 * ((1 - rn) * (1 - sn) * (ze[4] - ze[0]) + (1 + rn) * (1 - sn) * (ze[5] - ze[1]) + (1 + rn) * (1 + sn) * (ze[6] - ze[2]) + (1 - rn) * (1 + sn) * (ze[7] - ze[3]));
    const dfloat J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt);
    // note delayed J scaling
    const dfloat rx = (ys * zt - zs * yt), ry = -(xs * zt - zs * xt), rz = (xs * yt - ys * xt);
    const dfloat sx = -(yr * zt - zr * yt), sy = (xr * zt - zr * xt), sz = -(xr * yt - yr * xt);
    const dfloat tx = (yr * zs - zr * ys), ty = -(xr * zs - zr * xs), tz = (xr * ys - yr * xs);
    const dfloat W = s_zw[1][i] * s_zw[1][j] * s_zw[1][k];
    const dfloat sc = W / J;
    const dfloat G00 = sc * (rx * rx + ry * ry + rz * rz), G01 = sc * (rx * sx + ry * sy + rz * sz);
    const dfloat G02 = sc * (rx * tx + ry * ty + rz * tz), G11 = sc * (sx * sx + sy * sy + sz * sz);
    const dfloat G12 = sc * (sx * tx + sy * ty + sz * tz), G22 = sc * (tx * tx + ty * ty + tz * tz);
    const dfloat GwJ = W * J;

    __syncthreads();
    s_q[j][i] = r_q[k];
    dfloat qt = 0;
#pragma unroll
    for (int m = 0; m < Nq; ++m) qt += s_D[k][m] * r_q[m];
    __syncthreads();
    dfloat qr = 0, qs = 0;
#pragma unroll
    for (int m = 0; m < Nq; ++m) { qr += s_D[i][m] * s_q[j][m]; qs += s_D[j][m] * s_q[m][i]; }
    s_Gqs[j][i] = G01 * qr + G11 * qs + G12 * qt;
    s_Gqr[j][i] = G00 * qr + G01 * qs + G02 * qt;
    const dfloat Gqt = G02 * qr + G12 * qs + G22 * qt;
    dfloat Auk = GwJ * lambda * r_q[k];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < Nq; ++m) {
      Auk += s_D[m][j] * s_Gqs[m][i];
      r_Aq[m] += s_D[k][m] * Gqt;
      Auk += s_D[m][i] * s_Gqr[j][m];
    }
    r_Aq[k] += Auk;
  }
#pragma unroll
  for (int k = 0; k < Nq; ++k) Aq[base + k * Nq2] = r_Aq[k];
}

}  // namespace

extern "C" int libp_mesh_physical_nodes_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* EX, const libp_dfloat* EY,
                                              const libp_dfloat* EZ, const libp_dfloat* gllz, libp_dfloat* x,
                                              libp_dfloat* y, libp_dfloat* z, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq, "Nq must be in [2, 9]");
  LIBP_CHECK(Nelements == 0 || (EX && EY && EZ && gllz && x && y && z), "null device pointer");
  if (Nelements > 0)
    physical_nodes_kernel<<<(unsigned)Nelements, 128, 0, as_stream(stream)>>>(Nq, Nelements, EX, EY, EZ, gllz, x, y, z);
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_mesh_geometric_factors_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* x, const libp_dfloat* y,
                                                 const libp_dfloat* z, const libp_dfloat* D, const libp_dfloat* gllw,
                                                 libp_dfloat* ggeo, libp_dfloat* wJ, libp_dfloat* vgeo, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq, "Nq must be in [2, 9]");
  LIBP_CHECK(Nelements == 0 || (x && y && z && D && gllw && ggeo && wJ), "null device pointer");
  if (Nelements == 0) return LIBP_SUCCESS;
  cudaStream_t s = as_stream(stream);
  const int Np = Nq * Nq * Nq;
  const size_t smem = sizeof(dfloat) * (size_t)(3 * Np + Nq * Nq + Nq);
  dev_buf<int> bad;
  bad.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  geometric_factors_kernel<<<(unsigned)Nelements, std::min(256, ((Np + 31) / 32) * 32), smem, s>>>(
      Nq, Nelements, x, y, z, D, gllw, ggeo, wJ, vgeo, bad.p);
  CUDA_CHECK(cudaGetLastError());
  int h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  LIBP_CHECK(h == 0, "Negative J found at element " + std::to_string(h - 1));
  LIBP_API_END
}

extern "C" int libp_elliptic_build_diagonal_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* ggeo,
                                                  const libp_dfloat* wJ, const libp_dfloat* D, const int* mapB,
                                                  libp_dfloat lambda, libp_dfloat allNeumannBoost, libp_dfloat* diagL,
                                                  void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq, "Nq must be in [2, 9]");
  LIBP_CHECK(Nelements == 0 || (ggeo && wJ && D && mapB && diagL), "null device pointer");
  if (Nelements == 0) return LIBP_SUCCESS;
  const int Np = Nq * Nq * Nq;
  const size_t smem = sizeof(dfloat) * (size_t)(Nq * Nq + 3 * Np);
  build_diagonal_kernel<<<(unsigned)Nelements, std::min(256, ((Np + 31) / 32) * 32), smem, as_stream(stream)>>>(
      Nq, Nelements, ggeo, wJ, D, mapB, lambda, allNeumannBoost, diagL);
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_ax_trilinear_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* elementList,
                                       const libp_dlong* GlobalToLocal, const libp_dfloat* EXYZ, const libp_dfloat* gllzw,
                                       const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* q, libp_dfloat* AqL,
                                       void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq, "Nq must be in [2, 9]");
  LIBP_CHECK(Nelements == 0 || (EXYZ && gllzw && D && q && AqL), "null device pointer");
  if (Nelements == 0) return LIBP_SUCCESS;
  cudaStream_t s = as_stream(stream);
  const unsigned g = (unsigned)Nelements;
#define GO(n)                                                                                                          \
  case n:                                                                                                              \
    if (GlobalToLocal) ax_trilinear_kernel<n, true><<<g, n * n, 0, s>>>(Nelements, elementList, GlobalToLocal, EXYZ,    \
                                                                        gllzw, D, lambda, q, AqL);                      \
    else ax_trilinear_kernel<n, false><<<g, n * n, 0, s>>>(Nelements, elementList, nullptr, EXYZ, gllzw, D, lambda, q,  \
                                                           AqL);                                                        \
    break;
  switch (Nq) { GO(2) GO(3) GO(4) GO(5) GO(6) GO(7) GO(8) GO(9) }
#undef GO
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}
