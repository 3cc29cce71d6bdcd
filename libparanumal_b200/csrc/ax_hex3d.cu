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
#include <algorithm>
#include <cmath>

#include "common.hpp"

using namespace libp_b200;

namespace {

// defaults of the transposed-pencil kernel (measured on B200, see profiles/ and DESIGN.md)
#ifndef LIBP_AX_MINB
#define LIBP_AX_MINB 8
#endif
#ifndef LIBP_AX_PF
#define LIBP_AX_PF 1
#endif
#ifndef LIBP_AX_HINT
#define LIBP_AX_HINT true
#endif
constexpr int kMaxNq = 9;
constexpr int kMaxH = kMaxNq / 2;
}  // namespace
namespace libp_b200 {
// Derivative matrix and its even-odd factors, passed to the kernels BY VALUE as a __grid_constant__ parameter (the
// parameter bank serves compile-time indexed DFMA operands exactly like __constant__ memory did, without any
// process-global state: handles of different orders / streams / threads cannot disturb each other).
//   D[i*Nq+m] = phi'_m(r_i)
// Even-odd factors of a centro-antisymmetric D (GLL: D[N-i][N-m] = -D[i][m]), H = Nq/2:
//   De[i*H+m] = (D[i][m] + D[i][N-m])/2,  Do[i*H+m] = (D[i][m] - D[i][N-m])/2      (i,m < H)
//   Dc[i] = D[i][c], Dr[m] = D[c][m] for the centre node c = H of odd Nq.
// They halve the constants a thread keeps in (uniform) registers and cut a pencil contraction from
// Nq^2 DFMAs to Nq^2/2 DFMAs + 2*Nq DADDs.  The same factors serve D^T ((D^T)e = Do^T, (D^T)o = De^T).
void AxD::set(int Nq, const double* Dh) {
  const int H = Nq / 2, N = Nq - 1;
  for (int n = 0; n < Nq * Nq; ++n) D[n] = Dh[n];
  for (int i = 0; i < H; ++i)
    for (int m = 0; m < H; ++m) {
      De[i * H + m] = 0.5 * (Dh[i * Nq + m] + Dh[i * Nq + N - m]);
      Do[i * H + m] = 0.5 * (Dh[i * Nq + m] - Dh[i * Nq + N - m]);
    }
  for (int i = 0; i < H; ++i) {
    Dc[i] = (Nq & 1) ? Dh[i * Nq + H] : 0.0;
    Dr[i] = (Nq & 1) ? Dh[H * Nq + i] : 0.0;
  }
}
}  // namespace libp_b200
namespace {
#define c_D dc.D
#define c_De dc.De
#define c_Do dc.Do
#define c_Dc dc.Dc
#define c_Dr dc.Dr

// o = D v (kT = false) or o = D^T v (kT = true) on a register pencil.
template <int Nq, bool kSym, bool kT>
__device__ __forceinline__ void pencil_apply(const AxD& dc, const dfloat (&v)[Nq], dfloat (&o)[Nq]) {
  if (!kSym) {
#pragma unroll
    for (int i = 0; i < Nq; ++i) {
      dfloat s = 0.0;
#pragma unroll
      for (int m = 0; m < Nq; ++m) s += (kT ? c_D[m * Nq + i] : c_D[i * Nq + m]) * v[m];
      o[i] = s;
    }
  } else {
    constexpr int H = Nq / 2, N = Nq - 1;
    dfloat ve[H > 0 ? H : 1], vo[H > 0 ? H : 1];
#pragma unroll
    for (int m = 0; m < H; ++m) { ve[m] = v[m] + v[N - m]; vo[m] = v[m] - v[N - m]; }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      dfloat E = 0.0, O = 0.0;
      if (Nq & 1) E = (kT ? c_Dr[i] : c_Dc[i]) * v[H];
#pragma unroll
      for (int m = 0; m < H; ++m) {
        E += (kT ? c_Do[m * H + i] : c_De[i * H + m]) * ve[m];
        O += (kT ? c_De[m * H + i] : c_Do[i * H + m]) * vo[m];
      }
      o[i] = E + O;
      o[N - i] = O - E;
    }
    if (Nq & 1) {
      dfloat s = 0.0;
#pragma unroll
      for (int m = 0; m < H; ++m) s += (kT ? c_Dc[m] : c_Dr[m]) * vo[m];
      o[H] = s;
    }
  }
}

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
                const int* __restrict__ doneFlag, const __grid_constant__ AxD dc) {
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


// ---------------------------------------------------------------------------------------------
// Variant 1 ("transposed pencils").  The pencil kernel above reads ~50 doubles of shared memory
// per node (every 1-D contraction output re-reads a row/column of the slab), which on B200 costs
// more SM cycles than the node's 64 B of HBM traffic (128 B/clk/SM of shared-memory bandwidth
// against ~23 B/clk/SM of HBM).  Here every 1-D contraction runs on a pencil held in registers
// against D taken from the constant bank with compile-time indices (DFMA with a c[][] operand),
// and shared memory is used only to re-distribute the element between the three pencil
// orientations: ~15 doubles of shared-memory traffic per node instead of ~50.
//   layout C: thread (i,j) owns the k-pencil   (global loads/stores, geometric factors, t-derivative)
//   layout A: thread (j,k) owns the i-pencil   (r-derivative, 128-bit row accesses)
//   layout B: thread (i,k) owns the j-pencil   (s-derivative, column accesses)
// Row stride LD and slab stride SS are padded so that all three access patterns are
// bank-conflict free for Nq = 8 (LD/2 odd for the 128-bit rows, SS = 8 mod 16 for the columns,
// and layout C pairs rows j and j+4 inside a half-warp: 4*LD = 8 mod 16).
// L2 residency hints: the geometric factors are a pure stream (each byte used once per apply) and
// must not push the re-used q / Aq lines out of the 126 MB L2, so they are loaded evict-first and
// bypass L1; q gathers and the Aq reductions are marked evict-last.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ dfloat ld_stream(const dfloat* p, uint64_t pol) {
  dfloat v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ dlong ld_stream_i(const dlong* p, uint64_t pol) {
  dlong v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ dfloat ld_keep(const dfloat* p, uint64_t pol) {
  dfloat v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_keep(dfloat* p, dfloat v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_keep(dfloat* p, dfloat v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// 128-bit row accesses of a shared-memory row (rows are 16-byte aligned; odd Nq finish with one 64-bit access)
template <int Nq>
__device__ __forceinline__ void load_row(const dfloat* __restrict__ row, dfloat (&v)[Nq]) {
#pragma unroll
  for (int c = 0; c < Nq / 2; ++c) {
    const double2 w = *reinterpret_cast<const double2*>(row + 2 * c);
    v[2 * c] = w.x; v[2 * c + 1] = w.y;
  }
  if (Nq & 1) v[Nq - 1] = row[Nq - 1];
}
template <int Nq>
__device__ __forceinline__ void store_row(dfloat* __restrict__ row, const dfloat (&o)[Nq]) {
#pragma unroll
  for (int c = 0; c < Nq / 2; ++c) *reinterpret_cast<double2*>(row + 2 * c) = make_double2(o[2 * c], o[2 * c + 1]);
  if (Nq & 1) row[Nq - 1] = o[Nq - 1];
}

// Elements per block of the transposed-pencil kernel: chosen so that the Nq^2 columns of the EPB elements fill
// whole warps (idle lanes hold registers but carry no bytes in flight: Nq = 6 with one element per 64-thread
// block leaves 44 % of the lanes empty) while the three staging arrays stay inside 48 KB of static shared memory.
template <int Nq> struct AxEPB { static constexpr int v = 1; };
template <> struct AxEPB<2> { static constexpr int v = 16; };  //  64 /  64 lanes
template <> struct AxEPB<3> { static constexpr int v = 7; };   //  63 /  64
template <> struct AxEPB<4> { static constexpr int v = 4; };   //  64 /  64
template <> struct AxEPB<5> { static constexpr int v = 5; };   // 125 / 128
template <> struct AxEPB<6> { static constexpr int v = 5; };   // 180 / 192
template <> struct AxEPB<7> { static constexpr int v = 3; };   // 147 / 160 (5 elements would need 60 KB)
template <> struct AxEPB<8> { static constexpr int v = 1; };   //  64 /  64
template <> struct AxEPB<9> { static constexpr int v = 1; };   //  81 /  96 (3 elements would need 67 KB)

template <int Nq>
struct AxT {
  static constexpr int Nq2 = Nq * Nq;
  static constexpr int Np = Nq * Nq * Nq;
  static constexpr int EPB = AxEPB<Nq>::v;
  static constexpr int Work = EPB * Nq2;
  static constexpr int Threads = ((Work + 31) / 32) * 32;
  // Shared-memory strides (doubles): row stride LD, slab stride SS, element stride ESS - all even, so rows stay
  // 16-byte aligned for the 128-bit accesses of layout A.  Chosen per order by tools/smem_layout_search.py, a
  // bank model of the three access patterns (32 banks x 4 B, cost = max distinct words per bank): wavefronts over
  // the conflict-free minimum are 1.00 (Nq = 2, 4, 8), 1.12 (6), 1.27 (9), 1.36 (7), 1.42 (5), 1.54 (3); the previous
  // uniform rule (LD/2 odd, SS = 8 mod 16, tuned for Nq = 8 only) cost 1.35x - 3.5x for the other orders, and ncu
  // counted every second shared wavefront of the Nq = 5 kernel as a conflict (profiles/r1_o_ncu_*).
  static constexpr int LD = (Nq == 2) ? 2 : (Nq <= 4) ? 4 : (Nq <= 6) ? 6 : 10;
  static constexpr int SS = (Nq == 2) ? 4 : (Nq == 3) ? 12 : (Nq == 4) ? 18 : (Nq == 5) ? 30 : (Nq == 6) ? 38
                          : (Nq == 7) ? 70 : (Nq == 8) ? 88 : 90;
  static constexpr int ESS = (Nq == 2) ? 10 : (Nq == 3) ? 42 : (Nq == 4) ? 72 : (Nq == 5) ? 150 : (Nq == 6) ? 228
                           : (Nq == 7) ? 496 : Nq * SS;
  static_assert(LD >= Nq + (Nq & 1) && LD % 2 == 0 && SS % 2 == 0 && ESS % 2 == 0 && SS >= Nq * LD && ESS >= Nq * SS,
                "shared-memory strides");
  // resident blocks asked of the compiler: ~512 threads per SM (128 registers per thread), as for Nq = 8.
  // Nq = 9 keeps 3 blocks of 96 threads: capped at 113 registers it measured 16 % slower (profiles/r1_k_*)
  static constexpr int MinBlocks = (Nq == 9 || Nq == 7) ? 3 : (512 + Threads - 1) / Threads;
  // geometric-factor slabs in flight per thread.  A thread of a low-order element has few bytes to ask for
  // (Nq = 4: 4 indices + 6 factors = 64 B with one slab), and ~512 threads per SM then keep only ~32 KB in
  // flight.  Low orders therefore request all (or most of) their slabs up front; Nq = 8 stays at the measured
  // optimum of one slab.  (Measured neutral for N = 3..5, profiles/r1_o_*: those kernels are bound by the LSU /
  // shared-memory pipe, not by bytes in flight.)
  static constexpr int PF = (LIBP_AX_PF != 1) ? LIBP_AX_PF : (Nq <= 4) ? Nq : (Nq == 5) ? 3 : (Nq <= 7) ? 2 : 1;
  static_assert(3 * EPB * ESS * 8 <= 48 * 1024, "staging arrays exceed static shared memory");
};

// PF = geometric-factor slabs in flight per thread, kHint = L2 residency hints on/off
// kZA = zero-ahead (fused mode): the kernel zero-fills its own accumulator just ahead of the reductions, so the
// zero lines are still in L2 when the first red.add arrives (one DRAM write per Aq line instead of memset write +
// fill read + write-back).  Producer: block vb zero-fills the rows that virtual block vb + delta is the first
// to touch, [zoff[vb+delta], zoff[vb+delta+1]) - zoff is the running maximum of the connectivity, so every row a
// block touches lies below its own zoff[b+1] - then counts itself into done[(vb+delta)/group] (fence + atomic).
// Consumer: before its reductions, vb waits until every group up to its own is complete (ld.acquire; `prefix` caches
// the number of complete leading groups).  The producers it waits for are blocks < vb (delta >= group), which were
// dispatched earlier and never wait themselves before producing.  Rows of the first delta blocks are zero-filled
// by the host before the launch.  ctr = [unused, prefix, error, pad, done[0..]].
template <int Nq, bool kGather, bool kFused, bool kDot, int PF, bool kHint, int kMinB, bool kSym, bool kZA = false>
__global__ void __launch_bounds__(AxT<Nq>::Threads, (Nq == 8) ? kMinB : AxT<Nq>::MinBlocks)
ax_hex3d_t_kernel(const dlong Nelements, const dlong* __restrict__ elementList, const dlong* __restrict__ G2L,
                  const dfloat* __restrict__ wJ, const dfloat* __restrict__ ggeo, const dfloat lambda,
                  const dfloat* __restrict__ q, dfloat* __restrict__ Aq, dfloat* __restrict__ dotPartials,
                  const int* __restrict__ doneFlag, const __grid_constant__ AxD dc, const ZeroAhead za = ZeroAhead()) {
  if (doneFlag != nullptr && *doneFlag) return;
  using C = AxT<Nq>;
  // zero-ahead relies on blocks being dispatched in blockIdx order (so that the producers a block waits for are
  // already running); a ticket counter would make that formal but costs an L2 round trip before the first load.
  // The consumer's spin is bounded and reports through ctr[2] (libp_elliptic_zero_ahead_errors).
  const int vb = blockIdx.x;
  constexpr int Nq2 = C::Nq2, Np = C::Np, LD = C::LD, SS = C::SS;
  static_assert(PF >= 1 && PF <= Nq, "prefetch depth");
  constexpr int ESS = C::ESS;
  __shared__ __align__(16) dfloat s_u[C::EPB * ESS];
  __shared__ __align__(16) dfloat s_r[C::EPB * ESS];
  __shared__ __align__(16) dfloat s_s[C::EPB * ESS];

  const int t = threadIdx.x;
  const bool valid = t < C::Work;
  const int es = valid ? t / Nq2 : 0;
  const int ij = valid ? t - es * Nq2 : 0;
  const int b = ij / Nq, a = ij - b * Nq;
  // layout C: i = a, j = jc.  For Nq = 8 a half-warp holds rows jc and jc+4 (conflict-free with LD = 10).
  const int jc = (Nq == 8) ? (4 * (b & 1) + (b >> 1)) : b;
  const int nC = jc * Nq + a;                 // node offset inside a k-slab (global arrays)
  const int sC = es * ESS + jc * LD + a;      // shared offset of (k=0, jc, a); slab k adds k*SS
  const int sA = es * ESS + b * SS + a * LD;  // layout A: row (k=b, j=a, i=0..)
  const int sB = es * ESS + b * SS + a;       // layout B: column (k=b, j=0.., i=a); j adds LD

  const dlong ei = (dlong)vb * C::EPB + es;
  const bool active = valid && ei < Nelements;
  const dlong e = active ? (elementList ? elementList[ei] : ei) : 0;
  const size_t ebase = (size_t)e * Np + nC;

  // ---- issue the connectivity loads first, then the first geometric-factor slabs, then the q gathers
  const uint64_t polS = kHint ? policy_evict_first() : 0, polK = kHint ? policy_evict_last() : 0;
  dlong r_id[Nq];
  if (kGather) {
#pragma unroll
    for (int k = 0; k < Nq; ++k)
      r_id[k] = active ? (kHint ? ld_stream_i(G2L + ebase + k * Nq2, polS) : __ldg(G2L + ebase + k * Nq2)) : -1;
  }
  auto ldgeo = [&](const dfloat* p) -> dfloat { return kHint ? ld_stream(p, polS) : __ldg(p); };
  const dfloat* __restrict__ gptr = ggeo + (size_t)e * 6 * Np + nC;
  const dfloat* __restrict__ wptr = wJ + ebase;
  const bool screened = (lambda != 0.0);
  dfloat g[PF][7];
#pragma unroll
  for (int p = 0; p < PF; ++p) {
#pragma unroll
    for (int c = 0; c < 6; ++c) g[p][c] = active ? ldgeo(gptr + p * Nq2 + c * Np) : 0.0;
    g[p][6] = (active && screened) ? ldgeo(wptr + p * Nq2) : 0.0;
  }
  dfloat r_q[Nq];
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    if (kGather) r_q[k] = (r_id[k] >= 0) ? (kHint ? ld_keep(q + r_id[k], polK) : q[r_id[k]]) : 0.0;
    else r_q[k] = active ? q[ebase + k * Nq2] : 0.0;
  }

  // ---- zero-ahead producer (after this block's own loads are in flight): all threads store, one thread publishes
  // (bar.sync orders the block's stores before thread 0's fence; the fence + counter increment is a cumulative release)
  if (kZA) {
    const int zb = vb + za.delta;
    if (zb < za.nblocks) {
      const dlong z0 = za.zoff[zb], z1 = za.zoff[zb + 1];
      for (dlong i = z0 + t; i < z1; i += C::Threads) st_keep(Aq + i, 0.0, polK);
      __syncthreads();
      if (t == 0) { __threadfence(); atomicAdd(&za.ctr[4 + zb / za.group], 1); }
    }
  }

  // ---- phase 0 (layout C): publish u, t-derivative in registers
  if (valid) {
#pragma unroll
    for (int k = 0; k < Nq; ++k) s_u[sC + k * SS] = r_q[k];
  }
  dfloat r_t[Nq];  // qt, later Gqt
  pencil_apply<Nq, kSym, false>(dc, r_q, r_t);
  __syncthreads();

  // ---- phase 1: r-derivative on i-pencils (layout A), s-derivative on j-pencils (layout B)
  if (valid) {
    dfloat v[Nq], o[Nq];
    load_row<Nq>(&s_u[sA], v);
    pencil_apply<Nq, kSym, false>(dc, v, o);
    store_row<Nq>(&s_r[sA], o);
#pragma unroll
    for (int m = 0; m < Nq; ++m) v[m] = s_u[sB + m * LD];
    pencil_apply<Nq, kSym, false>(dc, v, o);
#pragma unroll
    for (int j = 0; j < Nq; ++j) s_s[sB + j * LD] = o[j];
  }
  __syncthreads();

  // ---- phase 2 (layout C): geometric factors, slab by slab, loads PF slabs ahead
  dfloat r_Aq[Nq];
#pragma unroll
  for (int k = 0; k < Nq; ++k) {
    const dfloat qr = s_r[sC + k * SS], qs = s_s[sC + k * SS], qt = r_t[k];
    const dfloat G00 = g[k % PF][0], G01 = g[k % PF][1], G02 = g[k % PF][2];
    const dfloat G11 = g[k % PF][3], G12 = g[k % PF][4], G22 = g[k % PF][5], GwJ = g[k % PF][6];
    if (k + PF < Nq) {
#pragma unroll
      for (int c = 0; c < 6; ++c) g[k % PF][c] = active ? ldgeo(gptr + (k + PF) * Nq2 + c * Np) : 0.0;
      g[k % PF][6] = (active && screened) ? ldgeo(wptr + (k + PF) * Nq2) : 0.0;
    }
    if (valid) {
      s_r[sC + k * SS] = G00 * qr + G01 * qs + G02 * qt;
      s_s[sC + k * SS] = G01 * qr + G11 * qs + G12 * qt;
    }
    r_t[k] = G02 * qr + G12 * qs + G22 * qt;
    r_Aq[k] = screened ? GwJ * lambda * s_u[sC + k * SS] : 0.0;  // u is re-read: not kept in registers
  }
  {
    dfloat o[Nq];
    pencil_apply<Nq, kSym, true>(dc, r_t, o);
#pragma unroll
    for (int k = 0; k < Nq; ++k) r_Aq[k] += o[k];
  }
  __syncthreads();

  // ---- phase 3: transposed derivatives, in place (each thread rewrites exactly what it read)
  if (valid) {
    dfloat v[Nq], o[Nq];
    load_row<Nq>(&s_r[sA], v);
    pencil_apply<Nq, kSym, true>(dc, v, o);
    store_row<Nq>(&s_r[sA], o);
#pragma unroll
    for (int m = 0; m < Nq; ++m) v[m] = s_s[sB + m * LD];
    pencil_apply<Nq, kSym, true>(dc, v, o);
#pragma unroll
    for (int j = 0; j < Nq; ++j) s_s[sB + j * LD] = o[j];
  }
  if (kZA && t == 0) {  // zero-ahead consumer: every row this block reduces into has been zero-filled
    const int gneed = vb / za.group;
    int g = ld_acquire_gpu(&za.ctr[1]);
    const int g0 = g;
    while (g <= gneed) {
      const int lo = max(g * za.group, za.delta), hi = min((g + 1) * za.group, za.nblocks);
      int spins = 0;
      while (ld_acquire_gpu(&za.ctr[4 + g]) < hi - lo) {
        __nanosleep(64);
        if (++spins > (1 << 22)) { za.ctr[2] = 1; break; }  // never observed; keeps a protocol bug from hanging the GPU
      }
      ++g;
    }
    if (g > g0) { __threadfence(); atomicMax(&za.ctr[1], g); }
  }
  __syncthreads();

  // ---- phase 4 (layout C): collect, optional p.Ap partial, store / scatter-add
#pragma unroll
  for (int k = 0; k < Nq; ++k) r_Aq[k] += s_r[sC + k * SS] + s_s[sC + k * SS];

  if (kDot) {
    dfloat d = 0.0;
    if (active) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) d += s_u[sC + k * SS] * r_Aq[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    __shared__ dfloat s_dot[C::Threads / 32];
    if ((t & 31) == 0) s_dot[t >> 5] = d;
    __syncthreads();
    if (t == 0) {
      dfloat tot = 0.0;
#pragma unroll
      for (int w = 0; w < C::Threads / 32; ++w) tot += s_dot[w];
      dotPartials[blockIdx.x] = tot;
    }
  }
  if (!active) return;
  if (kFused) {
#pragma unroll
    for (int k = 0; k < Nq; ++k)
      if (r_id[k] >= 0) {
        if (kHint) red_keep(&Aq[r_id[k]], r_Aq[k], polK);
        else atomicAdd(&Aq[r_id[k]], r_Aq[k]);
      }
  } else {
#pragma unroll
    for (int k = 0; k < Nq; ++k) Aq[ebase + k * Nq2] = r_Aq[k];
  }
}

int g_variant = 1;                     // 0 = pencil kernel, 1 = transposed-pencil kernel
int g_pf = 2, g_hint = 1, g_minb = 6;  // tuning state of variant 1 (libp_ax_hex3d_tune)

// Default instantiation for every order; for the fused N=7 kernels (the headline path) a small grid of
// (prefetch depth, L2 hints, resident blocks) is compiled so the choice can be measured on the device.
template <int Nq, bool G, bool F, bool DOT, bool SYM>
void launch_t(int grid, dlong Nelements, const dlong* elementList, const dlong* G2L, const dfloat* wJ,
              const dfloat* ggeo, dfloat lambda, const dfloat* q, dfloat* Aq, dfloat* dotPartials,
              const int* doneFlag, const ZeroAhead* za, cudaStream_t s, const AxD& dc) {
  if constexpr (G && F && SYM) {
    if (za != nullptr) {
      ax_hex3d_t_kernel<Nq, G, F, DOT, AxT<Nq>::PF, LIBP_AX_HINT, LIBP_AX_MINB, SYM, true>
          <<<grid, AxT<Nq>::Threads, 0, s>>>(Nelements, elementList, G2L, wJ, ggeo, lambda, q, Aq, dotPartials, doneFlag, dc, *za);
      return;
    }
  }
#define GOT(PF, H, MB)                                                                                        \
  ax_hex3d_t_kernel<Nq, G, F, DOT, PF, H, MB, SYM><<<grid, AxT<Nq>::Threads, 0, s>>>(                         \
      Nelements, elementList, G2L, wJ, ggeo, lambda, q, Aq, dotPartials, doneFlag, dc)
#ifdef LIBP_AX_TUNE_GRID
  if constexpr (Nq == 8 && G && F && SYM && !DOT) {
#define ROW(PF, H)                                        \
    if (g_pf == PF && g_hint == H) {                      \
      if (g_minb == 4) { GOT(PF, H, 4); return; }         \
      if (g_minb == 6) { GOT(PF, H, 6); return; }         \
      if (g_minb == 8) { GOT(PF, H, 8); return; }         \
      if (g_minb == 10) { GOT(PF, H, 10); return; }       \
    }
    ROW(1, true) ROW(2, true) ROW(3, true) ROW(4, true) ROW(2, false)
#undef ROW
  }
#endif
  // a general (non-GLL) D needs all Nq^2 constants: give the N=7 kernel the registers instead of spilling
  GOT(AxT<Nq>::PF, LIBP_AX_HINT, (SYM || Nq < 8) ? LIBP_AX_MINB : 4);
#undef GOT
}

template <int Nq>
int launch(bool gather, bool fused, bool sym, dlong Nelements, const dlong* elementList, const dlong* G2L,
           const dfloat* wJ, const dfloat* ggeo, dfloat lambda, const dfloat* q, dfloat* Aq, dfloat* dotPartials,
           const int* doneFlag, const ZeroAhead* za, cudaStream_t s, const AxD& dc) {
  using C = AxCfg<Nq>;
  const int epb = (g_variant == 1) ? AxT<Nq>::EPB : C::EPB;
  const int grid = (int)((Nelements + epb - 1) / epb);
#define ARGS grid, Nelements, elementList, G2L, wJ, ggeo, lambda, q, Aq, dotPartials, doneFlag, za, s, dc
#define GO(G, F, DOT)                                                                                         \
  do {                                                                                                        \
    if (g_variant == 1) {                                                                                     \
      if (sym) launch_t<Nq, G, F, DOT, true>(ARGS);                                                           \
      else launch_t<Nq, G, F, DOT, false>(ARGS);                                                              \
    } else {                                                                                                  \
      ax_hex3d_kernel<Nq, G, F, DOT><<<grid, C::Threads, 0, s>>>(Nelements, elementList, G2L, wJ, ggeo, lambda, \
                                                                 q, Aq, dotPartials, doneFlag, dc);           \
    }                                                                                                         \
  } while (0)
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
#undef ARGS
  CUDA_CHECK(cudaGetLastError());
  return grid;
}

}  // namespace

namespace libp_b200 {

// dc: host copy of D (+ even-odd factors), passed to the kernel by value.  sym: D was verified centro-antisymmetric
// (libp_elliptic_create does that once), which enables the even-odd contractions.
// Returns the number of blocks launched (= number of dotPartials written when dotPartials != nullptr).
int ax_hex3d_launch(int Nq, bool fused, const AxD& dc, bool sym, dlong Nelements, const dlong* elementList,
                    const dlong* G2L, const dfloat* wJ, const dfloat* ggeo, dfloat lambda,
                    const dfloat* q, dfloat* Aq, dfloat* dotPartials, const int* doneFlag, cudaStream_t s,
                    const ZeroAhead* za) {
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq, "Nq must be in [2, 9]");
  LIBP_CHECK(za == nullptr || (fused && sym && g_variant == 1), "zero-ahead needs the fused even-odd kernel");
  LIBP_CHECK(!fused || G2L != nullptr, "fused gather needs GlobalToLocal");
  if (Nelements <= 0) return 0;
  const bool gather = G2L != nullptr;
  switch (Nq) {
#define CASE(n) case n: return launch<n>(gather, fused, sym, Nelements, elementList, G2L, wJ, ggeo, lambda, q, Aq, dotPartials, doneFlag, za, s, dc);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
  }
  return 0;
}
// Is D centro-antisymmetric (D[N-i][N-m] == -D[i][m]) to rounding?  True for any GLL derivative matrix.
bool ax_hex3d_D_is_centro_antisymmetric(int Nq, const dfloat* D_host) {
  double mx = 0.0, dev = 0.0;
  for (int i = 0; i < Nq; ++i)
    for (int m = 0; m < Nq; ++m) {
      const double a = D_host[i * Nq + m], b = D_host[(Nq - 1 - i) * Nq + (Nq - 1 - m)];
      mx = std::max(mx, std::abs(a));
      dev = std::max(dev, std::abs(a + b));
    }
  return dev <= 1e-13 * mx;
}
// elements per block of the kernel variant in use, or 0 when zero-ahead is not available for it
int ax_hex3d_zero_ahead_epb(int Nq) {
  static const int epbT[10] = {0, 0, AxEPB<2>::v, AxEPB<3>::v, AxEPB<4>::v, AxEPB<5>::v, AxEPB<6>::v, AxEPB<7>::v,
                               AxEPB<8>::v, AxEPB<9>::v};
  return (g_variant == 1 && Nq >= 2 && Nq <= kMaxNq) ? epbT[Nq] : 0;
}
int ax_hex3d_blocks(int Nq, dlong Nelements) {
  // upper bound over both kernel variants (the pencil kernel packs 64 / Nq^2 elements, the transposed one AxEPB)
  const int nq2 = Nq * Nq;
  const int epbP = (nq2 >= 64) ? 1 : (64 / nq2);
  static const int epbT[10] = {1, 1, AxEPB<2>::v, AxEPB<3>::v, AxEPB<4>::v, AxEPB<5>::v, AxEPB<6>::v, AxEPB<7>::v,
                               AxEPB<8>::v, AxEPB<9>::v};
  const int epb = std::min(epbP, epbT[Nq]);
  return (int)((Nelements + epb - 1) / epb);
}

}  // namespace libp_b200

// Registry of derivative matrices the caller promised to keep immutable (libp_ax_hex3d_register_D):
// device pointer -> host copy of D with its even-odd factors and whether it is centro-antisymmetric.  Registered
// matrices skip the per-call device-to-host read of D and, when they are GLL matrices, run the even-odd kernels.
namespace {
struct RegisteredD { const dfloat* ptr; int Nq; bool sym; AxD dc; };
std::vector<RegisteredD> g_registry;
const RegisteredD* find_registered(const dfloat* D, int Nq) {
  for (const auto& r : g_registry)
    if (r.ptr == D && r.Nq == Nq) return &r;
  return nullptr;
}
// raw entry points with an unregistered D: read it back (stream-ordered, then synchronise) - correct but slow
void fetch_D(int Nq, const dfloat* D, cudaStream_t s, AxD& dc) {
  double hD[kMaxNq * kMaxNq];
  CUDA_CHECK(cudaMemcpyAsync(hD, D, sizeof(double) * Nq * Nq, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  dc.set(Nq, hD);
}
}  // namespace

extern "C" int libp_ax_hex3d_register_D(int Nq, const libp_dfloat* D) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nq >= 2 && Nq <= kMaxNq && D, "bad argument");
  double hD[kMaxNq * kMaxNq];
  CUDA_CHECK(cudaMemcpy(hD, D, sizeof(double) * Nq * Nq, cudaMemcpyDeviceToHost));
  RegisteredD r{D, Nq, ax_hex3d_D_is_centro_antisymmetric(Nq, hD), AxD()};
  r.dc.set(Nq, hD);
  for (auto& e : g_registry)
    if (e.ptr == D) { e = r; return LIBP_SUCCESS; }
  g_registry.push_back(r);
  LIBP_API_END
}

extern "C" int libp_ax_hex3d_unregister_D(const libp_dfloat* D) {
  LIBP_API_BEGIN
  for (size_t i = 0; i < g_registry.size(); ++i)
    if (g_registry[i].ptr == D) { g_registry.erase(g_registry.begin() + i); break; }
  LIBP_API_END
}

extern "C" int libp_ax_hex3d_set_variant(int variant) {
  LIBP_API_BEGIN
  LIBP_CHECK(variant == 0 || variant == 1, "variant must be 0 (pencil) or 1 (transposed pencils)");
  g_variant = variant;
  LIBP_API_END
}

extern "C" int libp_ax_hex3d_tune(int prefetch_slabs, int l2_hints, int min_blocks) {
  LIBP_API_BEGIN
#ifndef LIBP_AX_TUNE_GRID
  throw error("library was built without LIBP_AX_TUNE_GRID");
#endif
  LIBP_CHECK(prefetch_slabs >= 1 && prefetch_slabs <= 4, "prefetch_slabs in [1,4]");
  LIBP_CHECK(min_blocks == 4 || min_blocks == 6 || min_blocks == 8 || min_blocks == 10, "min_blocks in {4,6,8,10}");
  g_pf = prefetch_slabs; g_hint = l2_hints ? 1 : 0; g_minb = min_blocks;
  LIBP_API_END
}

extern "C" int libp_ax_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* elementList,
                             const libp_dlong* GlobalToLocal, const libp_dfloat* wJ, const libp_dfloat* ggeo,
                             const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* q, libp_dfloat* AqL,
                             void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements == 0 || (wJ && ggeo && D && q && AqL), "null device pointer");
  if (Nelements == 0) return LIBP_SUCCESS;
  const RegisteredD* r = find_registered(D, Nq);
  AxD tmp;
  if (!r) fetch_D(Nq, D, as_stream(stream), tmp);
  ax_hex3d_launch(Nq, false, r ? r->dc : tmp, r && r->sym, Nelements, elementList, GlobalToLocal, wJ, ggeo, lambda, q,
                  AqL, nullptr, nullptr, as_stream(stream), nullptr);
  LIBP_API_END
}

extern "C" int libp_ax_hex3d_gather(int Nq, libp_dlong Nelements, const libp_dlong* elementList,
                                    const libp_dlong* GlobalToLocal, const libp_dfloat* wJ, const libp_dfloat* ggeo,
                                    const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* q, libp_dfloat* Aq,
                                    void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nelements == 0 || (GlobalToLocal && wJ && ggeo && D && q && Aq), "null device pointer");
  if (Nelements == 0) return LIBP_SUCCESS;
  const RegisteredD* r = find_registered(D, Nq);
  AxD tmp;
  if (!r) fetch_D(Nq, D, as_stream(stream), tmp);
  ax_hex3d_launch(Nq, true, r ? r->dc : tmp, r && r->sym, Nelements, elementList, GlobalToLocal, wJ, ggeo, lambda, q,
                  Aq, nullptr, nullptr, as_stream(stream), nullptr);
  LIBP_API_END
}
