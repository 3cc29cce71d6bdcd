// Matrix-free screened-Poisson operator on hexahedra (CEED BP5 collocated form).
//
// Replaces ellipticPartialAxHex3D / ellipticAxHex3D
// (solvers/elliptic/okl/ellipticAxHex3D.okl:28-295).  Same mathematics:
//   u      = q[GlobalToLocal[e,:]]            (-1 -> 0: masked Dirichlet node)
//   qr,qs,qt = (D x I x I, I x D x I, I x I x D) u
//   Gq*    = G(e,node) (qr,qs,qt)^T           (6 symmetric factors, ggeo ids G00,G01,G02,G11,G12,G22)
//   Aq     = D^T-applies of Gq* + lambda * wJ * u
//
// sm_100a mapping (HBM-bound kernel: 56 B of geometric factors per node against ~58 FP64 FMAs):
//   * one thread per (i,j) column of an element, k-pencil of u and of the result in registers;
//   * the k-direction contractions use D straight from the constant bank (compile-time indices,
//     DFMA with a c[][] operand - no shared-memory traffic at all);
//   * the in-plane contractions read u / Gqr / Gqs slabs from padded shared memory with
//     128-bit row loads and keep the thread's own rows D[i][:], D[j][:] in registers;
//   * geometric factors are software-prefetched one k-slab ahead (coalesced 256 B per warp);
//   * optional fused epilogue: the ogs gather (Add, Trans) becomes FP64 red.global.add into the
//     gathered vector, removing the AqL round trip (16 B/node) and the separate gather pass.
// FP64 tensor cores are deliberately not used (Nq <= 9 contractions).
#include "common.hpp"

using namespace libp_b200;

namespace {

constexpr int kMaxNq = 9;
__constant__ double c_D[kMaxNq * kMaxNq];  // D[i*Nq+m] = phi'_m(r_i), loaded per launch (D2D async)

template <int Nq>
struct AxCfg {
  static constexpr int Nq2 = Nq * Nq;
  static constexpr int Np = Nq * Nq * Nq;
  // elements per block: keep >= 64 threads per block for small orders
  static constexpr int EPB = (Nq2 >= 64) ? 1 : (64 / Nq2);
  static constexpr int Work = EPB * Nq2;                 // threads that own a column
  static constexpr int Threads = ((Work + 31) / 32) * 32;  // whole warps (shuffle reductions)
  // padded row stride (doubles): even (16-byte rows for 128-bit loads) and != 0 mod 16 banks
  static constexpr int LD = (Nq % 2 == 0) ? Nq + 2 : Nq + 1;
};

// kDot: also emit sum_n u[n]*(A_e u)[n] per block (p.Ap of the PCG iteration computed element-locally:
// p^T Z^T A_L Z p = sum_e u_e^T A_e u_e, so it needs no pass over the gathered result).
template <int Nq, bool kGather, bool kFused, bool kDot>
__global__ void __launch_bounds__(AxCfg<Nq>::Threads)
ax_hex3d_kernel(const dlong Nelements, const dlong* __restrict__ elementList, const dlong* __restrict__ G2L,
                const dfloat* __restrict__ wJ, const dfloat* __restrict__ ggeo, const dfloat lambda,
                const dfloat* __restrict__ q, dfloat* __restrict__ Aq, dfloat* __restrict__ dotPartials,
                const int* __restrict__ doneFlag) {
  if (doneFlag != nullptr && *doneFlag) return;  // converged solver: the iteration body is a no-op
  using C = AxCfg<Nq>;
  constexpr int Nq2 = C::Nq2, Np = C::Np, LD = C::LD;
  __shared__ __align__(16) dfloat s_q[C::EPB][Nq][LD];
  __shared__ __align__(16) dfloat s_Gqr[C::EPB][Nq][LD];
  __shared__ __align__(16) dfloat s_Gqs[C::EPB][Nq][LD];
  __shared__ dfloat s_D[Nq][Nq + 1];

  const int t = threadIdx.x;
  const bool valid = t < C::Work;
  const int es = valid ? t / Nq2 : 0;  // element slot in this block
  const int ij = valid ? t - es * Nq2 : 0;
  const int j = ij / Nq, i = ij - j * Nq;
  const dlong ei = (dlong)blockIdx.x * C::EPB + es;
  const bool active = valid && ei < Nelements;
  const dlong e = active ? (elementList ? elementList[ei] : ei) : 0;

  for (int n = t; n < Nq2; n += C::Threads) s_D[n / Nq][n % Nq] = c_D[n];

  // thread-private rows of D for the in-plane derivatives
  dfloat Di[Nq], Dj[Nq];
#pragma unroll
  for (int m = 0; m < Nq; ++m) { Di[m] = c_D[i * Nq + m]; Dj[m] = c_D[j * Nq + m]; }

  dfloat r_q[Nq], r_Aq[Nq];
  dlong r_id[Nq];
  const size_t ebase = (size_t)e * Np + ij;
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    if (kGather) {
      const dlong id = active ? G2L[ebase + k * Nq2] : -1;
      r_id[k] = id;
      r_q[k] = (id >= 0) ? q[id] : 0.0;
    } else {
      r_id[k] = 0;
      r_q[k] = active ? q[ebase + k * Nq2] : 0.0;
    }
    r_Aq[k] = 0.0;
  }

  const dfloat* __restrict__ gptr = ggeo + (size_t)e * 6 * Np + ij;
  const dfloat* __restrict__ wptr = wJ + ebase;
  dfloat G00, G01, G02, G11, G12, G22, GwJ = 0.0;
  G00 = gptr[0 * Np]; G01 = gptr[1 * Np]; G02 = gptr[2 * Np];
  G11 = gptr[3 * Np]; G12 = gptr[4 * Np]; G22 = gptr[5 * Np];
  if (lambda != 0.0) GwJ = wptr[0];

#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    __syncthreads();  // previous slab's s_Gq* readers are done; s_D visible on k==0
    if (valid) s_q[es][j][i] = r_q[k];
    // prefetch next slab's geometric factors while this slab computes
    dfloat nG00 = 0, nG01 = 0, nG02 = 0, nG11 = 0, nG12 = 0, nG22 = 0, nGwJ = 0;
    if (k + 1 < Nq) {
      const int o = (k + 1) * Nq2;
      nG00 = gptr[o + 0 * Np]; nG01 = gptr[o + 1 * Np]; nG02 = gptr[o + 2 * Np];
      nG11 = gptr[o + 3 * Np]; nG12 = gptr[o + 4 * Np]; nG22 = gptr[o + 5 * Np];
      if (lambda != 0.0) nGwJ = wptr[o];
    }
    dfloat qt = 0.0;
#pragma unroll
    for (int m = 0; m < Nq; ++m) qt += c_D[k * Nq + m] * r_q[m];
    __syncthreads();

    dfloat qr = 0.0, qs = 0.0;
    if (Nq % 2 == 0) {
#pragma unroll
      for (int m = 0; m < Nq; m += 2) {
        const double2 row = *reinterpret_cast<const double2*>(&s_q[es][j][m]);
        qr += Di[m] * row.x;
        qr += Di[m + 1] * row.y;
      }
    } else {
#pragma unroll
      for (int m = 0; m < Nq; ++m) qr += Di[m] * s_q[es][j][m];
    }
#pragma unroll
    for (int m = 0; m < Nq; ++m) qs += Dj[m] * s_q[es][m][i];

    if (valid) {
      s_Gqs[es][j][i] = G01 * qr + G11 * qs + G12 * qt;
      s_Gqr[es][j][i] = G00 * qr + G01 * qs + G02 * qt;
    }
    const dfloat Gqt = G02 * qr + G12 * qs + G22 * qt;
    dfloat Auk = GwJ * lambda * r_q[k];
    __syncthreads();

#pragma unroll
    for (int m = 0; m < Nq; ++m) r_Aq[m] += c_D[k * Nq + m] * Gqt;
    if (Nq % 2 == 0) {
#pragma unroll
      for (int m = 0; m < Nq; m += 2) {
        const double2 row = *reinterpret_cast<const double2*>(&s_Gqr[es][j][m]);
        Auk += s_D[m][i] * row.x;
        Auk += s_D[m + 1][i] * row.y;
      }
    } else {
#pragma unroll
      for (int m = 0; m < Nq; ++m) Auk += s_D[m][i] * s_Gqr[es][j][m];
    }
#pragma unroll
    for (int m = 0; m < Nq; ++m) Auk += s_D[m][j] * s_Gqs[es][m][i];
    r_Aq[k] += Auk;
    G00 = nG00; G01 = nG01; G02 = nG02; G11 = nG11; G12 = nG12; G22 = nG22; GwJ = nGwJ;
  }

  if (kDot) {
    dfloat d = 0.0;
    if (active) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) d += r_q[k] * r_Aq[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    __shared__ dfloat s_dot[(C::Threads + 31) / 32];
    __syncthreads();
    if ((t & 31) == 0) s_dot[t >> 5] = d;
    __syncthreads();
    if (t == 0) {
      dfloat tot = 0.0;
#pragma unroll
      for (int w = 0; w < (C::Threads + 31) / 32; ++w) tot += s_dot[w];
      dotPartials[blockIdx.x] = tot;
    }
  }
  if (!active) return;
  if (kFused) {
#pragma unroll
    for (int k = 0; k < Nq; ++k)
      if (r_id[k] >= 0) atomicAdd(&Aq[r_id[k]], r_Aq[k]);  // red.global.add.f64 (result unused)
  } else {
#pragma unroll
    for (int k = 0; k < Nq; ++k) Aq[ebase + k * Nq2] = r_Aq[k];
  }
}

template <int Nq>
int launch(bool gather, bool fused, dlong Nelements, const dlong* elementList, const dlong* G2L, const dfloat* wJ,
           const dfloat* ggeo, dfloat lambda, const dfloat* q, dfloat* Aq, dfloat* dotPartials, const int* doneFlag,
           cudaStream_t s) {
  using C = AxCfg<Nq>;
  const int grid = (int)((Nelements + C::EPB - 1) / C::EPB);
#define GO(G, F, DOT) ax_hex3d_kernel<Nq, G, F, DOT><<<grid, C::Threads, 0, s>>>(Nelements, elementList, G2L, wJ, ggeo, lambda, q, Aq, dotPartials, doneFlag)
  if (dotPartials) {
    if (fused) GO(true, true, true);
    else if (gather) GO(true, false, true);
    else GO(false, false, true);
  } else {
    if (fused) GO(true, true, false);
    else if (gather) GO(true, false, false);
    else GO(false, false, false);
  }
#undef GO
  CUDA_CHECK(cudaGetLastError());
  return grid;
}

const dfloat* g_cD_owner = nullptr;  // device pointer whose contents currently sit in c_D
int g_cD_nq = 0;

}  // namespace

namespace libp_b200 {

// trusted==true: the caller guarantees D is immutable (operator handles) so the constant bank is
// only reloaded when a different D pointer / order is used.
// Returns the number of blocks launched (= number of dotPartials written when dotPartials != nullptr).
int ax_hex3d_launch(int Nq, bool fused, bool trusted_D, dlong Nelements, const dlong* elementList, const dlong* G2L,
                    const dfloat* wJ, const dfloat* ggeo, const dfloat* D, dfloat lambda, const dfloat* q,
                    dfloat* Aq, dfloat* dotPartials, const int* doneFlag, cudaStream_t s) {
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq, "Nq must be in [2, 9]");
  LIBP_CHECK(!fused || G2L != nullptr, "fused gather needs GlobalToLocal");
  if (Nelements <= 0) return 0;
  if (!(trusted_D && g_cD_owner == D && g_cD_nq == Nq)) {
    CUDA_CHECK(cudaMemcpyToSymbolAsync(c_D, D, sizeof(dfloat) * Nq * Nq, 0, cudaMemcpyDeviceToDevice, s));
    g_cD_owner = trusted_D ? D : nullptr;
    g_cD_nq = Nq;
  }
  const bool gather = G2L != nullptr;
  switch (Nq) {
#define CASE(n) case n: return launch<n>(gather, fused, Nelements, elementList, G2L, wJ, ggeo, lambda, q, Aq, dotPartials, doneFlag, s);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
  }
  return 0;
}
int ax_hex3d_blocks(int Nq, dlong Nelements) {
  const int nq2 = Nq * Nq;
  const int epb = (nq2 >= 64) ? 1 : (64 / nq2);
  return (int)((Nelements + epb - 1) / epb);
}

}  // namespace libp_b200

extern "C" int libp_ax_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* elementList,
                             const libp_dlong* GlobalToLocal, const libp_dfloat* wJ, const libp_dfloat* ggeo,
                             const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* q, libp_dfloat* AqL,
                             void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements == 0 || (wJ && ggeo && D && q && AqL), "null device pointer");
  ax_hex3d_launch(Nq, false, false, Nelements, elementList, GlobalToLocal, wJ, ggeo, D, lambda, q, AqL,
                  nullptr, nullptr, as_stream(stream));
  LIBP_API_END
}

extern "C" int libp_ax_hex3d_gather(int Nq, libp_dlong Nelements, const libp_dlong* elementList,
                                    const libp_dlong* GlobalToLocal, const libp_dfloat* wJ, const libp_dfloat* ggeo,
                                    const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* q, libp_dfloat* Aq,
                                    void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements == 0 || (GlobalToLocal && wJ && ggeo && D && q && Aq), "null device pointer");
  ax_hex3d_launch(Nq, true, false, Nelements, elementList, GlobalToLocal, wJ, ggeo, D, lambda, q, Aq,
                  nullptr, nullptr, as_stream(stream));
  LIBP_API_END
}
