// Interior-penalty DG operator on hexahedra (DISCRETIZATION = IPDG), SURVEY 8(f)-4:
//   elliptic_t::Operator, IPDG branch        solvers/elliptic/src/ellipticOperator.cpp:108-160
//   ellipticPartialGradientHex3D             solvers/elliptic/okl/ellipticGradientHex3D.okl:97-169
//   ellipticPartialAxIpdgHex3D + surfaceTerms  solvers/elliptic/okl/ellipticAxIpdgHex3D.okl:35-85, 359-642
//   BuildOperatorDiagonalIpdgHex3D           solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:868-996
//   SurfaceGeometricFactorsHex3D             libs/mesh/meshSurfaceGeometricFactorsHex3D.cpp:31-195
// Two passes as in the reference (the neighbour's trace needs the neighbour's complete gradient): a gradient pass that
// stores (dq/dx, dq/dy, dq/dz, q) per node, the trace-halo exchange of that array across ranks, and the element pass.
// Both passes keep a thread per (i,j) column with the k-pencil in registers; the element pass computes the flux /
// penalty terms of all six faces into shared memory in one sweep (one face node per thread and face) instead of three
// face-pair phases, then folds them into the column registers.
// HBM bytes per node (N = 7): gradient pass 8 + 72 + 32, element pass 32 + 80 + 8 + ~0.75 * (40 + 8 + 32 gathered).
#include <memory>

#include "elliptic.hpp"

using namespace libp_b200;

namespace {

constexpr int RX = 0, RY = 1, RZ = 2, SX = 3, SY = 4, SZ = 5, TX = 6, TY = 7, TZ = 8, JWID = 10;
constexpr int kNvgeo = 12, kNsgeo = 8;
constexpr int sNX = 0, sNY = 1, sNZ = 2, sSJ = 3, sIJ = 4, sIH = 5, sWSJ = 6, sWIJ = 7;

struct DMat { double D[81]; };

template <int Nq>
__global__ void __launch_bounds__(Nq * Nq) ipdg_gradient_kernel(dlong Nelements, const dfloat* __restrict__ vgeo,
                                                                const __grid_constant__ DMat dm,
                                                                const dfloat* __restrict__ q, double4* __restrict__ grad) {
  constexpr int Nq2 = Nq * Nq, Np = Nq2 * Nq, P = Nq + 1;
  __shared__ dfloat s_D[Nq][Nq];
  __shared__ dfloat s_q[Nq][Nq][P];
  const int t = threadIdx.x, j = t / Nq, i = t - j * Nq;
  const dlong e = blockIdx.x;
  if (e >= Nelements) return;
  s_D[j][i] = dm.D[j * Nq + i];
  dfloat r_q[Nq];
  const dfloat* qe = q + (size_t)e * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) { r_q[k] = qe[k * Nq2]; s_q[k][j][i] = r_q[k]; }
  __syncthreads();
  const dfloat* v = vgeo + (size_t)e * kNvgeo * Np + t;
  double4* ge = grad + (size_t)e * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    dfloat qr = 0, qs = 0, qt = 0;
#pragma unroll
    for (int n = 0; n < Nq; ++n) {
      qr += s_D[i][n] * s_q[k][j][n];
      qs += s_D[j][n] * s_q[k][n][i];
      qt += s_D[k][n] * r_q[n];
    }
    const dfloat* vk = v + k * Nq2;
    double4 g;
    g.x = vk[RX * Np] * qr + vk[SX * Np] * qs + vk[TX * Np] * qt;
    g.y = vk[RY * Np] * qr + vk[SY * Np] * qs + vk[TY * Np] * qt;
    g.z = vk[RZ * Np] * qr + vk[SZ * Np] * qs + vk[TZ * Np] * qt;
    g.w = r_q[k];
    ge[k * Nq2] = g;
  }
}

// volume node of face node n = (b, a) (a fast) on face f: faces 0..5 = t=-1, s=-1, r=+1, s=+1, r=-1, t=+1
template <int Nq>
__device__ __forceinline__ int face_node(int f, int a, int b) {
  constexpr int N = Nq - 1;
  switch (f) {
    case 0: return a + b * Nq;
    case 1: return a + b * Nq * Nq;
    case 2: return N + a * Nq + b * Nq * Nq;
    case 3: return a + N * Nq + b * Nq * Nq;
    case 4: return a * Nq + b * Nq * Nq;
    default: return a + b * Nq + N * Nq * Nq;
  }
}

template <int Nq, bool kDot>
__global__ void __launch_bounds__(Nq * Nq) ipdg_ax_kernel(dlong Nlist, const dlong* __restrict__ elementList,
                                                          const dlong* __restrict__ vmapM, const dlong* __restrict__ vmapP,
                                                          dfloat lambda, dfloat tau, const dfloat* __restrict__ vgeo,
                                                          const dfloat* __restrict__ sgeo, const int* __restrict__ EToB,
                                                          const __grid_constant__ DMat dm, const double4* __restrict__ grad,
                                                          dfloat* __restrict__ Aq, dfloat* __restrict__ dotPartials,
                                                          const int* __restrict__ doneFlag) {
  if (doneFlag != nullptr && *doneFlag) return;
  constexpr int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  __shared__ dfloat s_D[Nq][Nq];
  __shared__ dfloat s_f[6][4][Nq2];       // per face node: flux into d/dx, d/dy, d/dz and the penalty / average term
  __shared__ dfloat s_Gr[Nq][Nq + 1], s_Gs[Nq][Nq + 1];
  const int t = threadIdx.x, j = t / Nq, i = t - j * Nq;
  const dlong e = elementList ? elementList[blockIdx.x] : (dlong)blockIdx.x;
  (void)Nlist;
  s_D[j][i] = dm.D[j * Nq + i];

  // ---- surface terms of the six faces, one face node per thread and face (surfaceTerms of the reference)
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const size_t sk = ((size_t)e * 6 + f) * Nq2 + t;
    const dlong idM = vmapM[sk], idP = vmapP[sk];
    const dfloat* sg = sgeo + sk * kNsgeo;
    const dfloat nx = sg[sNX], ny = sg[sNY], nz = sg[sNZ], WsJ = sg[sWSJ], hinv = sg[sIH];
    const double4 gM = grad[idM];
    double4 gP = grad[idP];
    const int bc = EToB[(size_t)e * 6 + f];
    if (bc == 1) {         // homogeneous Dirichlet: ghost state 2*(0, grad-) - interior
      gP.x = gM.x; gP.y = gM.y; gP.z = gM.z; gP.w = -gM.w;
    } else if (bc == 2) {  // homogeneous Neumann
      gP.x = -gM.x; gP.y = -gM.y; gP.z = -gM.z; gP.w = gM.w;
    }
    const dfloat dq = gP.w - gM.w;
    s_f[f][0][t] = 0.5 * WsJ * nx * dq;
    s_f[f][1][t] = 0.5 * WsJ * ny * dq;
    s_f[f][2][t] = 0.5 * WsJ * nz * dq;
    s_f[f][3][t] = -0.5 * WsJ * (nx * (gP.x + gM.x) + ny * (gP.y + gM.y) + nz * (gP.z + gM.z) + tau * hinv * dq);
  }

  // ---- volume part of the column: JW * grad q, JW * lambda * q
  dfloat r_gx[Nq], r_gy[Nq], r_gz[Nq], r_Aq[Nq], r_q[Nq];
  const dfloat* v = vgeo + (size_t)e * kNvgeo * Np + t;
  const double4* ge = grad + (size_t)e * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const double4 g = ge[k * Nq2];
    const dfloat JW = v[JWID * Np + k * Nq2];
    r_gx[k] = JW * g.x; r_gy[k] = JW * g.y; r_gz[k] = JW * g.z;
    r_Aq[k] = JW * lambda * g.w;
    r_q[k] = g.w;
  }
  __syncthreads();

  // ---- fold the face terms into the column (face order of the reference: 0 & 5, 1 & 3, 2 & 4)
  auto add_face = [&](int f, int n, int k) {
    r_gx[k] += s_f[f][0][n]; r_gy[k] += s_f[f][1][n]; r_gz[k] += s_f[f][2][n]; r_Aq[k] += s_f[f][3][n];
  };
  add_face(0, t, 0);
  add_face(5, t, Nq - 1);
  if (j == 0) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(1, k * Nq + i, k);
  }
  if (j == Nq - 1) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(3, k * Nq + i, k);
  }
  if (i == Nq - 1) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(2, k * Nq + j, k);
  }
  if (i == 0) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(4, k * Nq + j, k);
  }

  // ---- layer by layer: reference-space fluxes, transposed derivatives
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const dfloat* vk = v + k * Nq2;
    const dfloat gx = r_gx[k], gy = r_gy[k], gz = r_gz[k];
    __syncthreads();
    s_Gr[j][i] = vk[RX * Np] * gx + vk[RY * Np] * gy + vk[RZ * Np] * gz;
    s_Gs[j][i] = vk[SX * Np] * gx + vk[SY * Np] * gy + vk[SZ * Np] * gz;
    const dfloat Gt = vk[TX * Np] * gx + vk[TY * Np] * gy + vk[TZ * Np] * gz;
    __syncthreads();
    dfloat dr = 0, ds = 0;
#pragma unroll
    for (int n = 0; n < Nq; ++n) {
      dr += s_D[n][i] * s_Gr[j][n];
      ds += s_D[n][j] * s_Gs[n][i];
      r_Aq[n] += s_D[k][n] * Gt;
    }
    r_Aq[k] += dr + ds;
  }
  dfloat* out = Aq + (size_t)e * Np + t;
  dfloat dacc = 0;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    out[k * Nq2] = r_Aq[k];
    if (kDot) dacc += r_q[k] * r_Aq[k];
  }
  if (kDot) {
    __shared__ dfloat s_red[(Nq2 + 31) / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
    if ((t & 31) == 0) s_red[t >> 5] = dacc;
    __syncthreads();
    if (t == 0) {
      dfloat tot = 0;
      for (int w = 0; w < (Nq2 + 31) / 32; ++w) tot += s_red[w];
      dotPartials[blockIdx.x] = tot;
    }
  }
}

// boundary data into the right-hand side (ellipticRhsBCIpdgHex3D, solvers/elliptic/okl/ellipticRhsBCIpdgHex3D.okl): the
// element pass with a zero interior state and the ghost state (uD, 0) on Dirichlet faces / (0, gN) on Neumann faces.
// uD, gN: nodal boundary data per face node [Nelements][6*Nq^2] (gN = n . (uxB, uyB, uzB) as the data file's Neumann
// macro returns them), the data-file functions the reference inlines at JIT time.
template <int Nq>
__global__ void __launch_bounds__(Nq * Nq) ipdg_rhs_bc_kernel(dlong Nelements, dfloat tau, const dfloat* __restrict__ vgeo,
                                                              const dfloat* __restrict__ sgeo, const int* __restrict__ EToB,
                                                              const __grid_constant__ DMat dm, const dfloat* __restrict__ uD,
                                                              const dfloat* __restrict__ gN, dfloat* __restrict__ rhs) {
  constexpr int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  __shared__ dfloat s_D[Nq][Nq];
  __shared__ dfloat s_f[6][4][Nq2];
  __shared__ dfloat s_Gr[Nq][Nq + 1], s_Gs[Nq][Nq + 1];
  const int t = threadIdx.x, j = t / Nq, i = t - j * Nq;
  const dlong e = blockIdx.x;
  if (e >= Nelements) return;
  s_D[j][i] = dm.D[j * Nq + i];
  bool any = false;
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const size_t sk = ((size_t)e * 6 + f) * Nq2 + t;
    const dfloat* sg = sgeo + sk * kNsgeo;
    const int bc = EToB[(size_t)e * 6 + f];
    any = any || bc > 0;
    const dfloat dq = (bc == 1 && uD) ? uD[sk] : 0.0;
    const dfloat dn = (bc == 2 && gN) ? gN[sk] : 0.0;
    const dfloat WsJ = sg[sWSJ];
    s_f[f][0][t] = WsJ * sg[sNX] * dq;
    s_f[f][1][t] = WsJ * sg[sNY] * dq;
    s_f[f][2][t] = WsJ * sg[sNZ] * dq;
    s_f[f][3][t] = -WsJ * (dn + tau * sg[sIH] * dq);
  }
  __syncthreads();
  if (!any) return;  // block-uniform: EToB is per element and face
  dfloat r_gx[Nq], r_gy[Nq], r_gz[Nq], r_r[Nq];
#pragma unroll
  for (int k = 0; k < Nq; ++k) r_gx[k] = r_gy[k] = r_gz[k] = r_r[k] = 0.0;
  auto add_face = [&](int f, int n, int k) {
    r_gx[k] += s_f[f][0][n]; r_gy[k] += s_f[f][1][n]; r_gz[k] += s_f[f][2][n]; r_r[k] += s_f[f][3][n];
  };
  add_face(0, t, 0);
  add_face(5, t, Nq - 1);
  if (j == 0) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(1, k * Nq + i, k);
  }
  if (j == Nq - 1) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(3, k * Nq + i, k);
  }
  if (i == Nq - 1) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(2, k * Nq + j, k);
  }
  if (i == 0) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) add_face(4, k * Nq + j, k);
  }
  const dfloat* v = vgeo + (size_t)e * kNvgeo * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const dfloat* vk = v + k * Nq2;
    const dfloat gx = r_gx[k], gy = r_gy[k], gz = r_gz[k];
    __syncthreads();
    s_Gr[j][i] = vk[RX * Np] * gx + vk[RY * Np] * gy + vk[RZ * Np] * gz;
    s_Gs[j][i] = vk[SX * Np] * gx + vk[SY * Np] * gy + vk[SZ * Np] * gz;
    const dfloat Gt = vk[TX * Np] * gx + vk[TY * Np] * gy + vk[TZ * Np] * gz;
    __syncthreads();
    dfloat dr = 0, ds = 0;
#pragma unroll
    for (int n = 0; n < Nq; ++n) {
      dr += s_D[n][i] * s_Gr[j][n];
      ds += s_D[n][j] * s_Gs[n][i];
      r_r[n] += s_D[k][n] * Gt;
    }
    r_r[k] += dr + ds;
  }
  dfloat* out = rhs + (size_t)e * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) out[k * Nq2] -= r_r[k];  // the reference subtracts the boundary functional
}

// diagonal: volume lines through the node + the faces the node lies on
template <int Nq>
__global__ void __launch_bounds__(Nq * Nq) ipdg_diag_kernel(dlong Nelements, const dfloat* __restrict__ vgeo,
                                                            const dfloat* __restrict__ sgeo, const int* __restrict__ EToB,
                                                            const __grid_constant__ DMat dm, dfloat lambda, dfloat tau,
                                                            dfloat* __restrict__ A) {
  constexpr int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  __shared__ dfloat s_D[Nq][Nq];
  __shared__ dfloat s_rr[Nq][Nq][Nq + 1], s_ss[Nq][Nq][Nq + 1], s_tt[Nq][Nq][Nq + 1];
  const int t = threadIdx.x, j = t / Nq, i = t - j * Nq;
  const dlong e = blockIdx.x;
  if (e >= Nelements) return;
  s_D[j][i] = dm.D[j * Nq + i];
  const dfloat* v = vgeo + (size_t)e * kNvgeo * Np + t;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const dfloat* vk = v + k * Nq2;
    const dfloat JW = vk[JWID * Np];
    s_rr[k][j][i] = JW * (vk[RX * Np] * vk[RX * Np] + vk[RY * Np] * vk[RY * Np] + vk[RZ * Np] * vk[RZ * Np]);
    s_ss[k][j][i] = JW * (vk[SX * Np] * vk[SX * Np] + vk[SY * Np] * vk[SY * Np] + vk[SZ * Np] * vk[SZ * Np]);
    s_tt[k][j][i] = JW * (vk[TX * Np] * vk[TX * Np] + vk[TY * Np] * vk[TY * Np] + vk[TZ * Np] * vk[TZ * Np]);
  }
  __syncthreads();
  const int bcs[6] = {EToB[(size_t)e * 6 + 0], EToB[(size_t)e * 6 + 1], EToB[(size_t)e * 6 + 2],
                      EToB[(size_t)e * 6 + 3], EToB[(size_t)e * 6 + 4], EToB[(size_t)e * 6 + 5]};
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const dfloat* vk = v + k * Nq2;
    const dfloat JW = vk[JWID * Np];
    dfloat a = 0;
#pragma unroll
    for (int m = 0; m < Nq; ++m)
      a += s_D[m][i] * s_D[m][i] * s_rr[k][j][m] + s_D[m][j] * s_D[m][j] * s_ss[k][m][i] + s_D[m][k] * s_D[m][k] * s_tt[m][j][i];
    const dfloat rx = vk[RX * Np], ry = vk[RY * Np], rz = vk[RZ * Np], sx = vk[SX * Np], sy = vk[SY * Np], sz = vk[SZ * Np];
    const dfloat tx = vk[TX * Np], ty = vk[TY * Np], tz = vk[TZ * Np];
    const dfloat di = s_D[i][i], dj = s_D[j][j], dk = s_D[k][k];
    a += 2 * JW * (di * dj * (rx * sx + ry * sy + rz * sz) + di * dk * (rx * tx + ry * ty + rz * tz) +
                   dj * dk * (sx * tx + sy * ty + sz * tz));
    a += lambda * JW;
    // gradient of the node's own Lagrange function at the node
    const dfloat lx = rx * di + sx * dj + tx * dk, ly = ry * di + sy * dj + ty * dk, lz = rz * di + sz * dj + tz * dk;
    auto face = [&](int f, int n) {
      const dfloat* sg = sgeo + (((size_t)e * 6 + f) * Nq2 + n) * kNsgeo;
      const int bc = bcs[f];
      const dfloat c = (bc == 1) ? 2.0 : (bc == 2) ? 0.0 : 1.0;  // (1 + bcD)(1 - bcN)
      const dfloat ndg = sg[sNX] * lx + sg[sNY] * ly + sg[sNZ] * lz;
      a += -c * sg[sWSJ] * ndg + 0.5 * c * sg[sWSJ] * tau * sg[sIH];
    };
    if (k == 0) face(0, j * Nq + i);
    if (j == 0) face(1, k * Nq + i);
    if (i == Nq - 1) face(2, k * Nq + j);
    if (j == Nq - 1) face(3, k * Nq + i);
    if (i == 0) face(4, k * Nq + j);
    if (k == Nq - 1) face(5, j * Nq + i);
    A[(size_t)e * Np + k * Nq2 + t] = a;
  }
}

// surface geometric factors; h = sJ / J per face node (the caller exchanges the halo part before the hinv pass)
template <int Nq>
__global__ void __launch_bounds__(Nq * Nq) surface_geofac_kernel(dlong Nelements, const dfloat* __restrict__ x,
                                                                 const dfloat* __restrict__ y, const dfloat* __restrict__ z,
                                                                 const __grid_constant__ DMat dm,
                                                                 const __grid_constant__ DMat gw, dfloat* __restrict__ sgeo,
                                                                 dfloat* __restrict__ h) {
  constexpr int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  __shared__ dfloat s_x[Nq][Nq][Nq + 1], s_y[Nq][Nq][Nq + 1], s_z[Nq][Nq][Nq + 1];
  const int t = threadIdx.x, b = t / Nq, a = t - b * Nq;
  const dlong e = blockIdx.x;
  if (e >= Nelements) return;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const size_t id = (size_t)e * Np + k * Nq2 + t;
    s_x[k][b][a] = x[id]; s_y[k][b][a] = y[id]; s_z[k][b][a] = z[id];
  }
  __syncthreads();
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const int n = face_node<Nq>(f, a, b);
    const int i = n % Nq, j = (n / Nq) % Nq, k = n / Nq2;
    dfloat xr = 0, xs = 0, xt = 0, yr = 0, ys = 0, yt = 0, zr = 0, zs = 0, zt = 0;
    for (int m = 0; m < Nq; ++m) {
      const dfloat Di = dm.D[i * Nq + m], Dj = dm.D[j * Nq + m], Dk = dm.D[k * Nq + m];
      xr += Di * s_x[k][j][m]; xs += Dj * s_x[k][m][i]; xt += Dk * s_x[m][j][i];
      yr += Di * s_y[k][j][m]; ys += Dj * s_y[k][m][i]; yt += Dk * s_y[m][j][i];
      zr += Di * s_z[k][j][m]; zs += Dj * s_z[k][m][i]; zt += Dk * s_z[m][j][i];
    }
    const dfloat J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt);
    const dfloat rx = (ys * zt - zs * yt) / J, ry = -(xs * zt - zs * xt) / J, rz = (xs * yt - ys * xt) / J;
    const dfloat sx = -(yr * zt - zr * yt) / J, sy = (xr * zt - zr * xt) / J, sz = -(xr * yt - yr * xt) / J;
    const dfloat tx = (yr * zs - zr * ys) / J, ty = -(xr * zs - zr * xs) / J, tz = (xr * ys - yr * xs) / J;
    dfloat nx, ny, nz;
    switch (f) {
      case 0: nx = -tx; ny = -ty; nz = -tz; break;
      case 1: nx = -sx; ny = -sy; nz = -sz; break;
      case 2: nx = rx; ny = ry; nz = rz; break;
      case 3: nx = sx; ny = sy; nz = sz; break;
      case 4: nx = -rx; ny = -ry; nz = -rz; break;
      default: nx = tx; ny = ty; nz = tz; break;
    }
    dfloat sJ = sqrt(nx * nx + ny * ny + nz * nz);
    nx /= sJ; ny /= sJ; nz /= sJ;
    sJ *= J;
    const size_t sk = ((size_t)e * 6 + f) * Nq2 + t;
    dfloat* sg = sgeo + sk * kNsgeo;
    sg[sNX] = nx; sg[sNY] = ny; sg[sNZ] = nz; sg[sSJ] = sJ; sg[sIJ] = 1.0 / J;
    sg[sIH] = 0.0;
    sg[sWSJ] = sJ * gw.D[a] * gw.D[b];
    sg[sWIJ] = 1.0 / (J * gw.D[0]);
    h[sk] = sJ / J;
  }
}

__global__ void surface_hinv_kernel(size_t nFaceNodes, const dlong* __restrict__ mapP, const dfloat* __restrict__ h,
                                    dfloat* __restrict__ sgeo) {
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nFaceNodes) return;
  dlong p = mapP[n];
  if (p < 0) p = (dlong)n;
  sgeo[n * kNsgeo + sIH] = fmax(h[n], h[p]);
}

DMat make_D(int Nq, const dfloat* D_dev) {
  DMat m{};
  CUDA_CHECK(cudaMemcpy(m.D, D_dev, sizeof(double) * Nq * Nq, cudaMemcpyDeviceToHost));
  return m;
}

#define NQ_SWITCH(Nq, CALL)                                                                    \
  switch (Nq) {                                                                                \
    case 2: { constexpr int NQ = 2; CALL; } break;                                             \
    case 3: { constexpr int NQ = 3; CALL; } break;                                             \
    case 4: { constexpr int NQ = 4; CALL; } break;                                             \
    case 5: { constexpr int NQ = 5; CALL; } break;                                             \
    case 6: { constexpr int NQ = 6; CALL; } break;                                             \
    case 7: { constexpr int NQ = 7; CALL; } break;                                             \
    case 8: { constexpr int NQ = 8; CALL; } break;                                             \
    case 9: { constexpr int NQ = 9; CALL; } break;                                             \
    default: throw libp_b200::error("Nq must be in [2, 9]");                                   \
  }

}  // namespace

// ---- operator handle ---------------------------------------------------------------------------------------------
struct libp_b200::IpdgData {
  libp_ipdg_desc_t d{};
  DMat dm{};
  dev_buf<dfloat> grad;  // [(Nelements + NhaloElements) * Np][4]
};

void libp_b200::ipdg_data_free(IpdgData* p) { delete p; }

void libp_b200::ipdg_apply(libp_elliptic_s& op, dfloat* q, dfloat* Aq, bool want_dot, const int* doneFlag, cudaStream_t s) {
  IpdgData& I = *op.ipdg;
  const libp_ipdg_desc_t& d = I.d;
  const int Nq = d.Nq;
  double4* grad = reinterpret_cast<double4*>(I.grad.p);
  if (d.Nelements > 0) {
    NQ_SWITCH(Nq, (ipdg_gradient_kernel<NQ><<<(unsigned)d.Nelements, NQ * NQ, 0, s>>>(d.Nelements, d.vgeo, I.dm, q, grad)));
    CUDA_CHECK(cudaGetLastError());
  }
  const bool exchange = d.traceHalo != nullptr && d.NhaloElementsTotal > 0;
  if (exchange && libp_halo_exchange_start(d.traceHalo, grad, 4, LIBP_DOUBLE, s) != LIBP_SUCCESS)
    throw libp_b200::error(std::string("trace halo exchange: ") + libp_last_error());
  int nb = 0;
  auto run = [&](dlong n, const dlong* list) {
    if (n <= 0) return;
    dfloat* dp = want_dot ? op.dotPartials.p + nb : nullptr;
    if (want_dot) {
      NQ_SWITCH(Nq, (ipdg_ax_kernel<NQ, true><<<(unsigned)n, NQ * NQ, 0, s>>>(n, list, d.vmapM, d.vmapP, d.lambda, d.tau, d.vgeo,
                                                                            d.sgeo, d.EToB, I.dm, grad, Aq, dp, doneFlag)));
    } else {
      NQ_SWITCH(Nq, (ipdg_ax_kernel<NQ, false><<<(unsigned)n, NQ * NQ, 0, s>>>(n, list, d.vmapM, d.vmapP, d.lambda, d.tau, d.vgeo,
                                                                             d.sgeo, d.EToB, I.dm, grad, Aq, dp, doneFlag)));
    }
    CUDA_CHECK(cudaGetLastError());
    nb += (int)n;
  };
  if (d.haloElementIds == nullptr && d.NhaloElements == 0 && d.internalElementIds == nullptr) {
    run(d.Nelements, nullptr);  // one rank: every element is internal
  } else {
    run(d.NinternalElements, d.internalElementIds);
  }
  if (exchange && libp_halo_exchange_finish(d.traceHalo, grad, 4, LIBP_DOUBLE, s) != LIBP_SUCCESS)
    throw libp_b200::error(std::string("trace halo exchange: ") + libp_last_error());
  run(d.NhaloElements, d.haloElementIds);
  op.nDotPartials = nb;
}

extern "C" int libp_elliptic_create_ipdg(const libp_ipdg_desc_t* desc, libp_elliptic_t* op) {
  LIBP_API_BEGIN
  LIBP_CHECK(desc && op, "null argument");
  LIBP_CHECK(desc->Nq >= 2 && desc->Nq <= 9, "Nq must be in [2, 9]");
  LIBP_CHECK(desc->Nelements >= 0 && desc->NhaloElementsTotal >= 0, "bad element counts");
  LIBP_CHECK(desc->vmapM && desc->vmapP && desc->vgeo && desc->sgeo && desc->EToB && desc->D, "null device pointer");
  LIBP_CHECK(desc->NinternalElements + desc->NhaloElements == desc->Nelements ||
                 (desc->internalElementIds == nullptr && desc->haloElementIds == nullptr),
             "element lists must cover all elements");
  LIBP_CHECK(desc->NhaloElementsTotal == 0 || desc->traceHalo != nullptr, "trace halo handle required with halo elements");
  std::unique_ptr<libp_elliptic_s> e(new libp_elliptic_s());
  e->d.Nq = desc->Nq;
  e->d.Nelements = desc->Nelements;
  e->d.lambda = desc->lambda;
  e->d.mode = 2;  // neither of the continuous modes: no gather, no accumulator to zero-fill
  e->Np = desc->Nq * desc->Nq * desc->Nq;
  e->Ndofs = desc->Nelements * e->Np;
  e->Nhalo = desc->NhaloElementsTotal * e->Np;
  e->ipdg = new IpdgData();
  e->ipdg->d = *desc;
  e->ipdg->dm = make_D(desc->Nq, desc->D);
  e->ipdg->grad.alloc((size_t)4 * (size_t)(desc->Nelements + desc->NhaloElementsTotal) * e->Np);
  CUDA_CHECK(cudaMemset(e->ipdg->grad.p, 0, sizeof(dfloat) * 4 * (size_t)(desc->Nelements + desc->NhaloElementsTotal) * e->Np));
  e->dotPartials.alloc((size_t)desc->Nelements + 4);
  *op = e.release();
  LIBP_API_END
}

extern "C" int libp_elliptic_ipdg_gradient(libp_elliptic_t op, const libp_dfloat** grad) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && op->ipdg && grad, "not an IPDG operator handle");
  *grad = op->ipdg->grad.p;
  LIBP_API_END
}

extern "C" int libp_elliptic_build_diagonal_ipdg_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* vgeo,
                                                       const libp_dfloat* sgeo, const int* EToB, const libp_dfloat* D,
                                                       libp_dfloat lambda, libp_dfloat tau, libp_dfloat* A, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements >= 0 && vgeo && sgeo && EToB && D && A, "bad argument");
  if (Nelements == 0) return LIBP_SUCCESS;
  const DMat dm = make_D(Nq, D);
  cudaStream_t s = (cudaStream_t)stream;
  NQ_SWITCH(Nq, (ipdg_diag_kernel<NQ><<<(unsigned)Nelements, NQ * NQ, 0, s>>>(Nelements, vgeo, sgeo, EToB, dm, lambda, tau, A)));
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_elliptic_rhs_bc_ipdg_hex3d(int Nq, libp_dlong Nelements, libp_dfloat tau, const libp_dfloat* vgeo,
                                               const libp_dfloat* sgeo, const int* EToB, const libp_dfloat* D,
                                               const libp_dfloat* uD, const libp_dfloat* gN, libp_dfloat* rhs, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements >= 0 && vgeo && sgeo && EToB && D && rhs, "bad argument");
  if (Nelements == 0) return LIBP_SUCCESS;
  const DMat dm = make_D(Nq, D);
  cudaStream_t s = (cudaStream_t)stream;
  NQ_SWITCH(Nq, (ipdg_rhs_bc_kernel<NQ><<<(unsigned)Nelements, NQ * NQ, 0, s>>>(Nelements, tau, vgeo, sgeo, EToB, dm, uD, gN, rhs)));
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_mesh_surface_geometric_factors_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* x,
                                                         const libp_dfloat* y, const libp_dfloat* z, const libp_dfloat* D,
                                                         const libp_dfloat* gllw, libp_dfloat* sgeo, libp_dfloat* h,
                                                         void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements >= 0 && x && y && z && D && gllw && sgeo && h, "bad argument");
  if (Nelements == 0) return LIBP_SUCCESS;
  const DMat dm = make_D(Nq, D);
  DMat gw{};
  CUDA_CHECK(cudaMemcpy(gw.D, gllw, sizeof(double) * Nq, cudaMemcpyDeviceToHost));
  cudaStream_t s = (cudaStream_t)stream;
  NQ_SWITCH(Nq, (surface_geofac_kernel<NQ><<<(unsigned)Nelements, NQ * NQ, 0, s>>>(Nelements, x, y, z, dm, gw, sgeo, h)));
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}

extern "C" int libp_mesh_surface_hinv_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* mapP, const libp_dfloat* h,
                                            libp_dfloat* sgeo, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements >= 0 && mapP && h && sgeo, "bad argument");
  const size_t n = (size_t)Nelements * 6 * Nq * Nq;
  if (n == 0) return LIBP_SUCCESS;
  surface_hinv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, mapP, h, sgeo);
  CUDA_CHECK(cudaGetLastError());
  LIBP_API_END
}
