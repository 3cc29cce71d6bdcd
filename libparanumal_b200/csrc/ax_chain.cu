// Fused operator kernel, "element chain" variant (mode 1 of elliptic_t::Operator, GLL D):
//   Aq[GlobalToLocal] (+)= A_e q[GlobalToLocal]     for the elements of one launch segment
// replacing ellipticPartialAxHex3D + ogs Gather(Add, Trans) (solvers/elliptic/okl/ellipticAxHex3D.okl:156-295,
// libs/ogs/ogs.cpp:203-298) like ax_hex3d_t_kernel does, with three changes that cut DRAM traffic and raise the bytes
// in flight:
//
//  1. Bulk-async geometric factors.  A CTA walks a CHAIN of L consecutive elements of the segment.  The contiguous
//     [6 (+1)][Np] block of geometric factors (+ wJ) of element n+1 is copied into shared memory by the TMA unit
//     (cp.async.bulk + mbarrier complete_tx, L2 evict-first) while element n is computed: no registers hold loads in
//     flight, 25-29 KB per CTA are always in flight.
//  2. Owner-computes accumulation.  A row of the gathered vector that is only touched by nodes of ONE chain
//     ("chain-private") needs no atomics across CTAs and no zero-fill: its first touch (in program order of the
//     chain) is a plain store, later touches by the same CTA are reductions that follow it in program order (bar.sync
//     orders them).  This is applied at the granularity of 32-byte DRAM sectors (4 rows): a sector whose 4 rows are
//     all chain-private is never zero-filled and never read for ownership; every other sector is zero-filled by
//     ax_chain_zero_fill() (bit mask) and accumulated with red.global.add as before.
//  3. Compressed connectivity.  Along i the ids of an element line are almost always [x, b, b+1, ..., b+Nq-2]
//     (first-appearance numbering of ogsBase_t::Setup, libs/ogs/ogsSetup.cpp:411-433): such elements store 2 ids per
//     line instead of Nq.  Other elements (Dirichlet faces, rank-shared rows) read the caller's GlobalToLocal.
//
// The classification is derived from GlobalToLocal alone by the plan kernels below (no assumption on the mesh).
// Layout C / A / B = the three pencil orientations of ax_hex3d.cu; contractions run on register pencils with the
// even-odd factors of D passed as a __grid_constant__ kernel parameter (no global constant state).
#include <algorithm>
#include <climits>
#include <cstdint>
#include <vector>

#include "ax_chain.hpp"

using namespace libp_b200;

namespace {

constexpr int kMaxNq = 9, kMaxH = kMaxNq / 2;

// even-odd factors of a centro-antisymmetric D (see ax_hex3d.cu)
struct EoD {
  double De[kMaxH * kMaxH], Do[kMaxH * kMaxH], Dc[kMaxH], Dr[kMaxH];
};

template <int Nq, bool kT>
__device__ __forceinline__ void eo_apply(const EoD& c, const dfloat (&v)[Nq], dfloat (&o)[Nq]) {
  constexpr int H = Nq / 2, N = Nq - 1;
  dfloat ve[H > 0 ? H : 1], vo[H > 0 ? H : 1];
#pragma unroll
  for (int m = 0; m < H; ++m) { ve[m] = v[m] + v[N - m]; vo[m] = v[m] - v[N - m]; }
#pragma unroll
  for (int i = 0; i < H; ++i) {
    dfloat E = 0.0, O = 0.0;
    if (Nq & 1) E = (kT ? c.Dr[i] : c.Dc[i]) * v[H];
#pragma unroll
    for (int m = 0; m < H; ++m) {
      E += (kT ? c.Do[m * H + i] : c.De[i * H + m]) * ve[m];
      O += (kT ? c.De[m * H + i] : c.Do[i * H + m]) * vo[m];
    }
    o[i] = E + O;
    o[N - i] = O - E;
  }
  if (Nq & 1) {
    dfloat s = 0.0;
#pragma unroll
    for (int m = 0; m < H; ++m) s += (kT ? c.Dc[m] : c.Dr[m]) * vo[m];
    o[H] = s;
  }
}

// ---- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pol_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t pol_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk copy by the TMA unit; completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ dfloat ld_keep(const dfloat* p, uint64_t pol) {
  dfloat v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ld_stream_i(const int* p, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ unsigned ld_stream_u16(const uint16_t* p, uint64_t pol) {
  unsigned short v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_keep(dfloat* p, dfloat v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_keep(dfloat* p, dfloat v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// 8-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 writes zeros (masked node)
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, uint32_t src_bytes, uint64_t pol) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2, %3;" ::"r"(dst), "l"(src), "r"(src_bytes),
               "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int Nq>
__device__ __forceinline__ void load_row(const dfloat* __restrict__ row, dfloat (&v)[Nq]) {
  if constexpr (Nq == 5 || Nq == 7 || Nq == 9) {  // dense rows (ChT::kDense) are not 16-byte aligned
#pragma unroll
    for (int c = 0; c < Nq; ++c) v[c] = row[c];
  } else {
#pragma unroll
    for (int c = 0; c < Nq / 2; ++c) {
      const double2 w = *reinterpret_cast<const double2*>(row + 2 * c);
      v[2 * c] = w.x; v[2 * c + 1] = w.y;
    }
    if (Nq & 1) v[Nq - 1] = row[Nq - 1];
  }
}
template <int Nq>
__device__ __forceinline__ void store_row(dfloat* __restrict__ row, const dfloat (&o)[Nq]) {
  if constexpr (Nq == 5 || Nq == 7 || Nq == 9) {
#pragma unroll
    for (int c = 0; c < Nq; ++c) row[c] = o[c];
  } else {
#pragma unroll
    for (int c = 0; c < Nq / 2; ++c) *reinterpret_cast<double2*>(row + 2 * c) = make_double2(o[2 * c], o[2 * c + 1]);
    if (Nq & 1) row[Nq - 1] = o[Nq - 1];
  }
}

// Shared-memory geometry of the staging arrays.  Nq = 8: dense rows (LD = 8), slab stride 66; with layout C on
// natural (j, j+1) half-warps, layout A on (k = a, j = b) and layout B on (i = a, k = 4(b&1) + b/2) all three
// access patterns AND the dense TMA-written geometric factors are bank-conflict free.  Other orders: strides of
// ax_hex3d.cu (tools/smem_layout_search.py).
template <int Nq>
struct ChT {
  static constexpr int Nq2 = Nq * Nq, Np = Nq * Nq * Nq;
  // Odd orders (Nq = 5, 7, 9): "dense" geometry.  A 64-bit shared access is served per half-warp (16 lanes, 16
  // eight-byte banks), so an element's Nq2 columns are padded to a multiple of 16 threads (half-warps never straddle
  // two elements), rows and slabs are stored without row padding (layout C and the TMA-written factors are contiguous
  // per half-warp), and the pencils of layouts A and B are dealt to the lanes by a table (ChainPerm) such that the 16
  // lanes of a half-warp start in 16 different banks.  Slab strides below are the ones for which such a deal exists
  // (every residue of k*SS + j*Nq resp. k*SS + i mod 16 occurs at most TPE/16 times); Nq = 5 has no common stride
  // for A and B, so s_r / s_s get their own and s_u keeps one two-way conflict in layout B.
  static constexpr bool kDense = (Nq == 5 || Nq == 7 || Nq == 9);
  // Even orders with several elements per block (Nq = 4, 6): lanes stay packed (element e owns lanes e*Nq2 ...), but
  // in layouts A and B a lane works on the (element, pencil) pair a table deals to it - any pair of the block - so
  // that the 8 lanes of a quarter-warp (128-bit rows, layout A) / the 16 lanes of a half-warp (layout B) start in
  // different banks although elements straddle those groups.
  static constexpr bool kDeal = (Nq == 4 || Nq == 6);
  static constexpr int TPE = kDense ? ((Nq2 + 15) / 16) * 16 : Nq2;  // threads per element
  static constexpr int EPB = kDense ? (Nq == 5 ? 4 : Nq == 7 ? 2 : 1)
                           : (Nq == 2) ? 16 : (Nq == 3) ? 7 : (Nq == 4) ? 4 : (Nq == 6) ? 5 : 1;
  static constexpr int Work = EPB * TPE;
  static constexpr int Threads = ((Work + 31) / 32) * 32;
  static constexpr int LD = kDense ? Nq : (Nq == 8) ? 8 : (Nq == 2) ? 2 : (Nq <= 4) ? 4 : 6;
  static constexpr int SS = (Nq == 8) ? 66 : (Nq == 2) ? 4 : (Nq == 3) ? 12 : (Nq == 4) ? 18 : (Nq == 6) ? 38 : 0;
  static constexpr int ESS = (Nq == 8) ? 8 * 66 : (Nq == 2) ? 10 : (Nq == 3) ? 42 : (Nq == 4) ? 72 : (Nq == 6) ? 228 : 0;
  // slab stride / element stride of s_u, s_r (layouts C + A), s_s (layouts C + B)
  static constexpr int SSu = kDense ? (Nq == 5 ? 25 : Nq == 7 ? 50 : 82) : SS;
  static constexpr int SSr = kDense ? (Nq == 5 ? 25 : Nq == 7 ? 50 : 82) : SS;
  static constexpr int SSs = kDense ? (Nq == 5 ? 26 : Nq == 7 ? 50 : 82) : SS;
  static constexpr int ESu = kDense ? (((Nq - 1) * SSu + Nq2 + 1) & ~1) : ESS;
  static constexpr int ESr = kDense ? (((Nq - 1) * SSr + Nq2 + 1) & ~1) : ESS;
  static constexpr int ESs = kDense ? (((Nq - 1) * SSs + Nq2 + 1) & ~1) : ESS;
  // wJ can only ride the bulk copy when its per-element block keeps 16-byte alignment
  static constexpr bool kBulkWJ = (Np % 2 == 0);
  // components per stage: 6 geometric factors (+ wJ for the screened operator)
  __host__ __device__ static constexpr int ng(bool scr) { return (scr && kBulkWJ) ? 7 : 6; }
  // bulk-copy destinations must be 16-byte aligned
  // ... and, where lanes of two elements share a half-warp (packed even orders), congruent to Nq2 mod 16 doubles so
  // that layout C reads of the dense factor blocks stay contiguous in bank space across the element boundary
  __host__ __device__ static constexpr int slot_doubles(bool scr) {
    const int base = ng(scr) * Np;
    if ((Nq2 & 1) || EPB == 1 || kDense) return base + (base & 1);
    return base + ((Nq2 - base) % 16 + 16) % 16;
  }
  __host__ __device__ static constexpr int stage_doubles(bool scr) { return EPB * slot_doubles(scr); }
};

// Dense and dealt orders: lane t of the block -> the i-pencil it works on in layout A and the j-pencil it works on in
// layout B, encoded element_slot * Nq2 + (k*Nq + j) resp. (k*Nq + i); kIdle = no work in that phase.  Built on the
// host (chain_perm<Nq>()).
struct ChainPerm {
  static constexpr unsigned short kIdle = 0xffff;
  unsigned short A[192], B[192];
};

// Per-thread shared-memory offsets of the three pencil orientations
template <int Nq>
struct ChIdx {
  int es, ij, a, b;
  bool valid, vA, vB;
  int uC, rC, sC, uA, rA, uB, sB;
  __host__ __device__ __forceinline__ ChIdx(int t, const ChainPerm& pm) {
    using C = ChT<Nq>;
    constexpr int Nq2 = C::Nq2, LD = C::LD;
    const bool inSlot = (C::Work == C::Threads) ? true : t < C::Work;  // folded away when the block has no idle lanes
    const int e0 = t / C::TPE, c = t - e0 * C::TPE;
    es = inSlot ? e0 : 0;
    valid = (C::TPE == Nq2) ? inSlot : (inSlot && c < Nq2);
    ij = valid ? c : 0;
    b = ij / Nq; a = ij - b * Nq;
    uC = es * C::ESu + b * LD + a;
    rC = es * C::ESr + b * LD + a;
    sC = es * C::ESs + b * LD + a;
    int kA, jA, kB, iB, eA = es, eB = es;
    if constexpr (C::kDense || C::kDeal) {
      const int cA = pm.A[t], cB = pm.B[t];
      vA = cA != ChainPerm::kIdle; vB = cB != ChainPerm::kIdle;
      eA = vA ? cA / Nq2 : 0; eB = vB ? cB / Nq2 : 0;
      const int pA = vA ? cA - eA * Nq2 : 0, pB = vB ? cB - eB * Nq2 : 0;
      kA = pA / Nq; jA = pA - kA * Nq;
      kB = pB / Nq; iB = pB - kB * Nq;
    } else {
      vA = vB = valid;
      kA = (Nq == 8) ? a : b; jA = (Nq == 8) ? b : a;      // layout A: i-pencil (row)
      kB = (Nq == 8) ? (4 * (b & 1) + (b >> 1)) : b; iB = a;  // layout B: j-pencil (column)
    }
    uA = eA * C::ESu + kA * C::SSu + jA * LD;
    rA = eA * C::ESr + kA * C::SSr + jA * LD;
    uB = eB * C::ESu + kB * C::SSu + iB;
    sB = eB * C::ESs + kB * C::SSs + iB;
  }
};

template <int Nq, int S, bool kScr>
constexpr size_t chain_smem_bytes() {
  using C = ChT<Nq>;
  return (size_t)8 * (S * C::stage_doubles(kScr) + C::EPB * (C::ESu + C::ESr + C::ESs)) + 8 * S + 4 * C::EPB + 16;
}

struct ChainArgs {
  const int* hdr;         // [nPos]  (element << 1) | raw, -1 = padding
  const int* cid;         // [nPos][2 * Nq2] compressed ids: [k*Nq + j] = id of node i=0, [Nq2 + k*Nq + j] = id of node i=1
  const uint16_t* flags;  // [nPos][Nq2]  column (j*Nq + i): bit k set = plain store (first touch of a private row)
  const dlong* G2L;
  const dfloat* wJ;
  const dfloat* ggeo;
  const dfloat* q;
  dfloat* Aq;
  dfloat* dotPartials;
  const int* doneFlag;
  dfloat lambda;
  int nChains, L;
  dlong count;  // elements of the segment (the last chain may be shorter than L)
};

template <int Nq, int S, bool kDot, bool kScr, int kMinB>
__global__ void __launch_bounds__(ChT<Nq>::Threads, kMinB)
ax_hex3d_chain_kernel(const ChainArgs A, const __grid_constant__ EoD eo, const __grid_constant__ ChainPerm pm) {
  if (A.doneFlag != nullptr && *A.doneFlag) return;
  using C = ChT<Nq>;
  constexpr int Nq2 = C::Nq2, Np = C::Np, LD = C::LD, SSu = C::SSu, SSr = C::SSr, SSs = C::SSs, EPB = C::EPB;
  constexpr int StageDoubles = C::stage_doubles(kScr), SlotDoubles = C::slot_doubles(kScr);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  dfloat* s_g = reinterpret_cast<dfloat*>(smem_raw);                    // [S][EPB][NG][Np]
  dfloat* s_u = s_g + S * StageDoubles;
  dfloat* s_r = s_u + EPB * C::ESu;
  dfloat* s_s = s_r + EPB * C::ESr;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_s + EPB * C::ESs);    // [S]
  int* s_hdr = reinterpret_cast<int*>(s_bar + S);                       // [EPB] header of the step to be issued next

  const int t = threadIdx.x;
  const ChIdx<Nq> X(t, pm);                        // layout C: i = a, j = b (natural); A: i-pencils; B: j-pencils
  const bool valid = X.valid;
  const int es = X.es, ij = X.ij, a = X.a, b = X.b;
  const int nC = ij;

  constexpr bool screened = kScr;  // lambda != 0
  constexpr bool bulkW = screened && C::kBulkWJ;
  const uint64_t polS = pol_evict_first(), polK = pol_evict_last();
  const int chain = blockIdx.x * EPB + es;         // one chain per element slot
  const bool chainOK = valid && chain < A.nChains;
  const size_t p0 = (size_t)chain * A.L;
  const int L = A.L;

  if (t == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&s_bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // element header of step n of this slot's chain (-1: nothing to do)
  auto header = [&](int n) -> int { return (chainOK && n < L) ? __ldg(A.hdr + p0 + n) : -1; };
  // thread 0 arms stage (n % S) and issues the copies of step n for every slot of the block.  The headers come from
  // shared memory (each slot's first thread publishes the one it prefetched), not from a dependent global load.
  auto issue = [&](int n) {
    const uint32_t bar = smem_u32(&s_bar[n % S]);
    uint32_t bytes = 0;
#pragma unroll
    for (int x = 0; x < EPB; ++x)
      if (s_hdr[x] >= 0) bytes += (uint32_t)(8 * Np * (bulkW ? 7 : 6));
    if (bytes == 0) return;
    mbar_expect_tx(bar, bytes);
#pragma unroll
    for (int x = 0; x < EPB; ++x) {
      const int h = s_hdr[x];
      if (h < 0) continue;
      dfloat* dst = s_g + (n % S) * StageDoubles + x * SlotDoubles;
      bulk_g2s(smem_u32(dst), A.ggeo + (size_t)(h >> 1) * 6 * Np, 8 * 6 * Np, bar, polS);
      if (bulkW) bulk_g2s(smem_u32(dst + 6 * Np), A.wJ + (size_t)(h >> 1) * Np, 8 * Np, bar, polS);
    }
  };
  const bool slotLead = valid && ij == 0;
  for (int s = 0; s < S; ++s) {  // prologue: the first S steps
    if (t < EPB) s_hdr[t] = -1;
    __syncthreads();
    if (slotLead) s_hdr[es] = header(s);
    __syncthreads();
    if (t == 0) issue(s);
    __syncthreads();
  }

  // connectivity of step n: this thread's k-pencil of ids (still encoded) + store/reduce flags.  Nothing loaded here
  // is looked at before the end of the iteration that issued it (decode_ids): an early use would stall the warp on
  // the scoreboard it shares with the q gathers.
  auto load_ids = [&](int n, int h, dlong (&id)[Nq], unsigned& fl) {
    if (h < 0) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = -1;
      fl = 0;
      return;
    }
    const size_t p = p0 + n;
    fl = ld_stream_u16(A.flags + p * Nq2 + ij, polS);
    if (h & 1) {
      const dlong* g = A.G2L + (size_t)(h >> 1) * Np + nC;
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = ld_stream_i(g + k * Nq2, polS);
    } else {
      const int* c = A.cid + p * (2 * Nq2) + (a == 0 ? 0 : Nq2) + b;
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = ld_stream_i(c + k * Nq, polS);
    }
  };
  auto decode_ids = [&](int h, dlong (&id)[Nq]) {
    if (h >= 0 && !(h & 1) && a != 0) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = (id[k] < 0) ? id[k] : id[k] + (a - 1);
    }
  };
  // q of a step is gathered straight into s_u (layout C slots of this thread) by 8-byte asynchronous copies: no
  // registers are held while the gathers are in flight
  const uint32_t su_base = smem_u32(s_u + X.uC);
  auto gather_q_async = [&](const dlong (&id)[Nq]) {
    if (valid) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) {
        const bool on = id[k] >= 0;
        cp_async8(su_base + 8u * (uint32_t)(k * SSu), A.q + (on ? id[k] : 0), on ? 8u : 0u, polK);
      }
    }
    cp_async_commit();
  };

  // software pipeline: ids two steps ahead (registers), q one step ahead (cp.async into s_u after phase 1 of the
  // previous step), geometric factors S steps ahead (TMA)
  int h_cur = header(0), h_nxt = header(1), h_nn = header(2);
  dlong id_cur[Nq], id_nxt[Nq];
  unsigned fl_cur, fl_nxt;
  load_ids(0, h_cur, id_cur, fl_cur);
  load_ids(1, h_nxt, id_nxt, fl_nxt);
  decode_ids(h_cur, id_cur);
  decode_ids(h_nxt, id_nxt);
  gather_q_async(id_cur);
  dfloat dacc = 0.0;

  // block-uniform trip count: the first slot owns the longest chain of the block (only the last chain is short)
  const long long left = (long long)A.count - (long long)blockIdx.x * EPB * L;
  const int nsteps = (int)(left < (long long)L ? (left < 0 ? 0 : left) : (long long)L);
  for (int n = 0; n < nsteps; ++n) {
    const bool active = h_cur >= 0;
    const dfloat* __restrict__ sg = s_g + (n % S) * StageDoubles + es * SlotDoubles + nC;

    // ---- prefetch: ids of step n+2 (not looked at before the end of this iteration)
    dlong id_nn[Nq];
    unsigned fl_nn;
    load_ids(n + 2, h_nn, id_nn, fl_nn);
    const int h_nnn = header(n + 3);
    // header of step n + S for the copy issued after phase 2 (S <= 2: already in registers)
    if (slotLead) s_hdr[es] = (S == 1) ? h_nxt : (S == 2) ? h_nn : h_nnn;
    dfloat r_w[Nq];
    if (screened && !C::kBulkWJ) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) r_w[k] = active ? __ldg(A.wJ + (size_t)(h_cur >> 1) * Np + nC + k * Nq2) : 0.0;
    }

    // ---- phase 0 (layout C): u arrived in s_u (own slots), t-derivative in registers
    cp_async_wait_all();
    dfloat q_cur[Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) q_cur[k] = valid ? s_u[X.uC + k * SSu] : 0.0;  // idle lanes of a padded element would only add bank conflicts
    dfloat r_t[Nq];
    eo_apply<Nq, false>(eo, q_cur, r_t);
    __syncthreads();

    // ---- phase 1: r-derivative on i-pencils (layout A), s-derivative on j-pencils (layout B)
    if (X.vA) {
      dfloat v[Nq], o[Nq];
      load_row<Nq>(&s_u[X.uA], v);
      eo_apply<Nq, false>(eo, v, o);
      store_row<Nq>(&s_r[X.rA], o);
    }
    if (X.vB) {
      dfloat v[Nq], o[Nq];
#pragma unroll
      for (int m = 0; m < Nq; ++m) v[m] = s_u[X.uB + m * LD];
      eo_apply<Nq, false>(eo, v, o);
#pragma unroll
      for (int j = 0; j < Nq; ++j) s_s[X.sB + j * LD] = o[j];
    }
    __syncthreads();
    // s_u is free: gather q of the next step into it while phases 2-4 run
    gather_q_async(id_nxt);

    // ---- phase 2 (layout C): geometric factors from the TMA-filled stage
    mbar_wait(smem_u32(&s_bar[n % S]), (uint32_t)((n / S) & 1));
    dfloat r_Aq[Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      const dfloat qr = valid ? s_r[X.rC + k * SSr] : 0.0, qs = valid ? s_s[X.sC + k * SSs] : 0.0, qt = r_t[k];
      dfloat G00 = 0, G01 = 0, G02 = 0, G11 = 0, G12 = 0, G22 = 0, GwJ = 0;
      if (active) {
        G00 = sg[0 * Np + k * Nq2]; G01 = sg[1 * Np + k * Nq2]; G02 = sg[2 * Np + k * Nq2];
        G11 = sg[3 * Np + k * Nq2]; G12 = sg[4 * Np + k * Nq2]; G22 = sg[5 * Np + k * Nq2];
        if (screened) GwJ = C::kBulkWJ ? sg[6 * Np + k * Nq2] : r_w[k];
      }
      if (valid) {
        s_r[X.rC + k * SSr] = G00 * qr + G01 * qs + G02 * qt;
        s_s[X.sC + k * SSs] = G01 * qr + G11 * qs + G12 * qt;
      }
      r_t[k] = G02 * qr + G12 * qs + G22 * qt;
      r_Aq[k] = screened ? GwJ * A.lambda * q_cur[k] : 0.0;
    }
    {
      dfloat o[Nq];
      eo_apply<Nq, true>(eo, r_t, o);
#pragma unroll
      for (int k = 0; k < Nq; ++k) r_Aq[k] += o[k];
    }
    __syncthreads();  // the stage has been consumed by every thread; Gq* published
    if (t == 0) issue(n + S);

    // ---- phase 3: transposed derivatives, in place
    if (X.vA) {
      dfloat v[Nq], o[Nq];
      load_row<Nq>(&s_r[X.rA], v);
      eo_apply<Nq, true>(eo, v, o);
      store_row<Nq>(&s_r[X.rA], o);
    }
    if (X.vB) {
      dfloat v[Nq], o[Nq];
#pragma unroll
      for (int m = 0; m < Nq; ++m) v[m] = s_s[X.sB + m * LD];
      eo_apply<Nq, true>(eo, v, o);
#pragma unroll
      for (int j = 0; j < Nq; ++j) s_s[X.sB + j * LD] = o[j];
    }
    __syncthreads();

    // ---- phase 4 (layout C): collect, p.Ap partial, plain store (first touch of a private row) or reduction
#pragma unroll
    for (int k = 0; k < Nq; ++k) r_Aq[k] += valid ? s_r[X.rC + k * SSr] + s_s[X.sC + k * SSs] : 0.0;
    if (kDot && active) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) dacc += q_cur[k] * r_Aq[k];
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) {
        if (id_cur[k] >= 0) {
          if ((fl_cur >> k) & 1u) st_keep(A.Aq + id_cur[k], r_Aq[k], polK);
          else red_keep(A.Aq + id_cur[k], r_Aq[k], polK);
        }
      }
    }
    // rotate the pipeline registers (first use of this iteration's id loads)
    decode_ids(h_nn, id_nn);
    h_cur = h_nxt; h_nxt = h_nn; h_nn = h_nnn;
    fl_cur = fl_nxt; fl_nxt = fl_nn;
#pragma unroll
    for (int k = 0; k < Nq; ++k) { id_cur[k] = id_nxt[k]; id_nxt[k] = id_nn[k]; }
    // phase 4 reads of s_r / s_s must finish before the next phase 1 overwrites them: the barrier after the next
    // phase 0 orders that
  }
  cp_async_wait_all();

  if (kDot) {
    dfloat d = dacc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    __shared__ dfloat s_dot[C::Threads / 32];
    if ((t & 31) == 0) s_dot[t >> 5] = d;
    __syncthreads();
    if (t == 0) {
      dfloat tot = 0.0;
#pragma unroll
      for (int w = 0; w < C::Threads / 32; ++w) tot += s_dot[w];
      A.dotPartials[blockIdx.x] = tot;
    }
  }
}

// ------------------------------------------------------------------------------------------------ trilinear variant
// ELEMENT MAP = TRILINEAR (ellipticPartialAxTrilinearHex3D, solvers/elliptic/okl/ellipticAxHex3D.okl:440-627): the
// geometric factors are not streamed but recomputed from the 8 vertices of the element (EXYZ [E][3][8], 192 bytes per
// element = 0.4 B per node at N = 7).  Same chain structure, pencil orientations, accumulation and connectivity as
// ax_hex3d_chain_kernel; no TMA stage (there is nothing to stream).  The Jacobian of a trilinear map is linear in each
// reference coordinate, so a thread (fixed r_i, s_j) keeps two coefficients per entry and evaluates a k-level with one
// FMA each; an element whose edges are parallel (affine map: every box element) has constant J and cofactors, which
// are then computed once per element and only scaled by the quadrature weight per node.
struct TriConst {
  double z[kMaxNq], w[kMaxNq];  // GLL nodes and weights
};

template <int Nq, bool kDot, bool kScr, int kMinB>
__global__ void __launch_bounds__(ChT<Nq>::Threads, kMinB)
ax_hex3d_chain_tri_kernel(const ChainArgs A, const dfloat* __restrict__ EXYZ, const __grid_constant__ EoD eo,
                          const __grid_constant__ TriConst gl, const __grid_constant__ ChainPerm pm) {
  if (A.doneFlag != nullptr && *A.doneFlag) return;
  using C = ChT<Nq>;
  constexpr int Nq2 = C::Nq2, Np = C::Np, LD = C::LD, SSu = C::SSu, SSr = C::SSr, SSs = C::SSs, EPB = C::EPB;
  constexpr int NV = (24 + Nq2 - 1) / Nq2;  // vertex coordinates each thread of a slot fetches
  __shared__ __align__(16) dfloat s_u[EPB * C::ESu];
  __shared__ __align__(16) dfloat s_r[EPB * C::ESr];
  __shared__ __align__(16) dfloat s_s[EPB * C::ESs];
  __shared__ dfloat s_v[EPB][24];

  const int t = threadIdx.x;
  const ChIdx<Nq> X(t, pm);                        // layout C: i = a, j = b (natural); A: i-pencils; B: j-pencils
  const bool valid = X.valid;
  const int es = X.es, ij = X.ij, a = X.a, b = X.b;
  const int nC = ij;
  constexpr bool screened = kScr;
  const uint64_t polS = pol_evict_first(), polK = pol_evict_last();
  const int chain = blockIdx.x * EPB + es;
  const bool chainOK = valid && chain < A.nChains;
  const size_t p0 = (size_t)chain * A.L;
  const int L = A.L;
  const dfloat rn = gl.z[a], sn = gl.z[b], wij = gl.w[a] * gl.w[b];

  auto header = [&](int n) -> int { return (chainOK && n < L) ? __ldg(A.hdr + p0 + n) : -1; };
  auto load_ids = [&](int n, int h, dlong (&id)[Nq], unsigned& fl) {
    if (h < 0) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = -1;
      fl = 0;
      return;
    }
    const size_t p = p0 + n;
    fl = ld_stream_u16(A.flags + p * Nq2 + ij, polS);
    if (h & 1) {
      const dlong* g = A.G2L + (size_t)(h >> 1) * Np + nC;
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = ld_stream_i(g + k * Nq2, polS);
    } else {
      const int* c = A.cid + p * (2 * Nq2) + (a == 0 ? 0 : Nq2) + b;
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = ld_stream_i(c + k * Nq, polS);
    }
  };
  auto decode_ids = [&](int h, dlong (&id)[Nq]) {
    if (h >= 0 && !(h & 1) && a != 0) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) id[k] = (id[k] < 0) ? id[k] : id[k] + (a - 1);
    }
  };
  auto load_vertices = [&](int h, dfloat (&v)[NV]) {
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      const int c = ij + m * Nq2;
      v[m] = (h >= 0 && c < 24) ? __ldg(EXYZ + (size_t)(h >> 1) * 24 + c) : 0.0;
    }
  };
  const uint32_t su_base = smem_u32(s_u + X.uC);
  auto gather_q_async = [&](const dlong (&id)[Nq]) {
    if (valid) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) {
        const bool on = id[k] >= 0;
        cp_async8(su_base + 8u * (uint32_t)(k * SSu), A.q + (on ? id[k] : 0), on ? 8u : 0u, polK);
      }
    }
    cp_async_commit();
  };

  int h_cur = header(0), h_nxt = header(1), h_nn = header(2);
  dlong id_cur[Nq], id_nxt[Nq];
  unsigned fl_cur, fl_nxt;
  dfloat v_cur[NV], v_nxt[NV];
  load_ids(0, h_cur, id_cur, fl_cur);
  load_ids(1, h_nxt, id_nxt, fl_nxt);
  load_vertices(h_cur, v_cur);
  load_vertices(h_nxt, v_nxt);
  decode_ids(h_cur, id_cur);
  decode_ids(h_nxt, id_nxt);
  gather_q_async(id_cur);
  dfloat dacc = 0.0;

  const long long left = (long long)A.count - (long long)blockIdx.x * EPB * L;
  const int nsteps = (int)(left < (long long)L ? (left < 0 ? 0 : left) : (long long)L);
  for (int n = 0; n < nsteps; ++n) {
    const bool active = h_cur >= 0;
    dlong id_nn[Nq];
    unsigned fl_nn;
    dfloat v_nn[NV];
    load_ids(n + 2, h_nn, id_nn, fl_nn);
    load_vertices(h_nn, v_nn);
    const int h_nnn = header(n + 3);

    // ---- phase 0: vertices of this element to shared memory, u from the asynchronous gather, t-derivative
    if (valid) {
#pragma unroll
      for (int m = 0; m < NV; ++m)
        if (ij + m * Nq2 < 24) s_v[es][ij + m * Nq2] = v_cur[m];
    }
    cp_async_wait_all();
    dfloat q_cur[Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) q_cur[k] = valid ? s_u[X.uC + k * SSu] : 0.0;  // idle lanes of a padded element would only add bank conflicts
    dfloat r_t[Nq];
    eo_apply<Nq, false>(eo, q_cur, r_t);
    __syncthreads();

    // ---- phase 1
    if (X.vA) {
      dfloat v[Nq], o[Nq];
      load_row<Nq>(&s_u[X.uA], v);
      eo_apply<Nq, false>(eo, v, o);
      store_row<Nq>(&s_r[X.rA], o);
    }
    if (X.vB) {
      dfloat v[Nq], o[Nq];
#pragma unroll
      for (int m = 0; m < Nq; ++m) v[m] = s_u[X.uB + m * LD];
      eo_apply<Nq, false>(eo, v, o);
#pragma unroll
      for (int j = 0; j < Nq; ++j) s_s[X.sB + j * LD] = o[j];
    }
    // Jacobian coefficients of this thread's (r_i, s_j): d/dr and d/ds are linear in t, d/dt does not depend on t
    dfloat ar[3], br[3], as_[3], bs[3], ct[3];
    bool affine = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const dfloat* e = s_v[es] + 8 * d;
      const dfloat e10 = e[1] - e[0], e23 = e[2] - e[3], e54 = e[5] - e[4], e67 = e[6] - e[7];
      const dfloat e30 = e[3] - e[0], e21 = e[2] - e[1], e74 = e[7] - e[4], e65 = e[6] - e[5];
      const dfloat e40 = e[4] - e[0], e51 = e[5] - e[1], e62 = e[6] - e[2], e73 = e[7] - e[3];
      ar[d] = (1 - sn) * e10 + (1 + sn) * e23;  br[d] = (1 - sn) * e54 + (1 + sn) * e67;
      as_[d] = (1 - rn) * e30 + (1 + rn) * e21; bs[d] = (1 - rn) * e74 + (1 + rn) * e65;
      ct[d] = 0.125 * ((1 - rn) * (1 - sn) * e40 + (1 + rn) * (1 - sn) * e51 + (1 + rn) * (1 + sn) * e62 + (1 - rn) * (1 + sn) * e73);
      affine = affine && e10 == e23 && e10 == e54 && e10 == e67 && e30 == e21 && e30 == e74 && e30 == e65 &&
               e40 == e51 && e40 == e62 && e40 == e73;
    }
    __syncthreads();
    gather_q_async(id_nxt);

    // ---- phase 2: geometric factors on the fly
    dfloat r_Aq[Nq];
    dfloat cG[7];  // affine element: cofactor products / J and J, the node only contributes its weight
    if (affine) {
      const dfloat xr = 0.25 * ar[0], yr = 0.25 * ar[1], zr = 0.25 * ar[2];
      const dfloat xs = 0.25 * as_[0], ys = 0.25 * as_[1], zs = 0.25 * as_[2];
      const dfloat xt = ct[0], yt = ct[1], zt = ct[2];
      const dfloat rx = (ys * zt - zs * yt), ry = -(xs * zt - zs * xt), rz = (xs * yt - ys * xt);
      const dfloat sx = -(yr * zt - zr * yt), sy = (xr * zt - zr * xt), sz = -(xr * yt - yr * xt);
      const dfloat tx = (yr * zs - zr * ys), ty = -(xr * zs - zr * xs), tz = (xr * ys - yr * xs);
      const dfloat J = xr * rx + yr * ry + zr * rz, iJ = 1.0 / J;
      cG[0] = iJ * (rx * rx + ry * ry + rz * rz); cG[1] = iJ * (rx * sx + ry * sy + rz * sz);
      cG[2] = iJ * (rx * tx + ry * ty + rz * tz); cG[3] = iJ * (sx * sx + sy * sy + sz * sz);
      cG[4] = iJ * (sx * tx + sy * ty + sz * tz); cG[5] = iJ * (tx * tx + ty * ty + tz * tz);
      cG[6] = J;
    }
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      const dfloat qr = valid ? s_r[X.rC + k * SSr] : 0.0, qs = valid ? s_s[X.sC + k * SSs] : 0.0, qt = r_t[k];
      const dfloat W = wij * gl.w[k];
      dfloat G00, G01, G02, G11, G12, G22, GwJ;
      if (affine) {
        G00 = W * cG[0]; G01 = W * cG[1]; G02 = W * cG[2]; G11 = W * cG[3]; G12 = W * cG[4]; G22 = W * cG[5];
        GwJ = W * cG[6];
      } else {
        const dfloat c0 = 0.125 * (1 - gl.z[k]), c1 = 0.125 * (1 + gl.z[k]);
        const dfloat xr = c0 * ar[0] + c1 * br[0], yr = c0 * ar[1] + c1 * br[1], zr = c0 * ar[2] + c1 * br[2];
        const dfloat xs = c0 * as_[0] + c1 * bs[0], ys = c0 * as_[1] + c1 * bs[1], zs = c0 * as_[2] + c1 * bs[2];
        const dfloat xt = ct[0], yt = ct[1], zt = ct[2];
        const dfloat rx = (ys * zt - zs * yt), ry = -(xs * zt - zs * xt), rz = (xs * yt - ys * xt);
        const dfloat sx = -(yr * zt - zr * yt), sy = (xr * zt - zr * xt), sz = -(xr * yt - yr * xt);
        const dfloat tx = (yr * zs - zr * ys), ty = -(xr * zs - zr * xs), tz = (xr * ys - yr * xs);
        const dfloat J = xr * rx + yr * ry + zr * rz;
        const dfloat sc = W / J;  // delayed J scaling, as the reference kernel
        G00 = sc * (rx * rx + ry * ry + rz * rz); G01 = sc * (rx * sx + ry * sy + rz * sz);
        G02 = sc * (rx * tx + ry * ty + rz * tz); G11 = sc * (sx * sx + sy * sy + sz * sz);
        G12 = sc * (sx * tx + sy * ty + sz * tz); G22 = sc * (tx * tx + ty * ty + tz * tz);
        GwJ = W * J;
      }
      if (valid) {
        s_r[X.rC + k * SSr] = G00 * qr + G01 * qs + G02 * qt;
        s_s[X.sC + k * SSs] = G01 * qr + G11 * qs + G12 * qt;
      }
      r_t[k] = G02 * qr + G12 * qs + G22 * qt;
      r_Aq[k] = screened ? GwJ * A.lambda * q_cur[k] : 0.0;
    }
    {
      dfloat o[Nq];
      eo_apply<Nq, true>(eo, r_t, o);
#pragma unroll
      for (int k = 0; k < Nq; ++k) r_Aq[k] += o[k];
    }
    __syncthreads();

    // ---- phase 3
    if (X.vA) {
      dfloat v[Nq], o[Nq];
      load_row<Nq>(&s_r[X.rA], v);
      eo_apply<Nq, true>(eo, v, o);
      store_row<Nq>(&s_r[X.rA], o);
    }
    if (X.vB) {
      dfloat v[Nq], o[Nq];
#pragma unroll
      for (int m = 0; m < Nq; ++m) v[m] = s_s[X.sB + m * LD];
      eo_apply<Nq, true>(eo, v, o);
#pragma unroll
      for (int j = 0; j < Nq; ++j) s_s[X.sB + j * LD] = o[j];
    }
    __syncthreads();

    // ---- phase 4
#pragma unroll
    for (int k = 0; k < Nq; ++k) r_Aq[k] += valid ? s_r[X.rC + k * SSr] + s_s[X.sC + k * SSs] : 0.0;
    if (kDot && active) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) dacc += q_cur[k] * r_Aq[k];
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < Nq; ++k) {
        if (id_cur[k] >= 0) {
          if ((fl_cur >> k) & 1u) st_keep(A.Aq + id_cur[k], r_Aq[k], polK);
          else red_keep(A.Aq + id_cur[k], r_Aq[k], polK);
        }
      }
    }
    decode_ids(h_nn, id_nn);
    h_cur = h_nxt; h_nxt = h_nn; h_nn = h_nnn;
    fl_cur = fl_nxt; fl_nxt = fl_nn;
#pragma unroll
    for (int k = 0; k < Nq; ++k) { id_cur[k] = id_nxt[k]; id_nxt[k] = id_nn[k]; }
#pragma unroll
    for (int m = 0; m < NV; ++m) { v_cur[m] = v_nxt[m]; v_nxt[m] = v_nn[m]; }
  }
  cp_async_wait_all();

  if (kDot) {
    dfloat d = dacc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    __shared__ dfloat s_dot[C::Threads / 32];
    if ((t & 31) == 0) s_dot[t >> 5] = d;
    __syncthreads();
    if (t == 0) {
      dfloat tot = 0.0;
#pragma unroll
      for (int w = 0; w < C::Threads / 32; ++w) tot += s_dot[w];
      A.dotPartials[blockIdx.x] = tot;
    }
  }
}

// ------------------------------------------------------------------------------------------------ plan kernels
// position p of the concatenated, chain-padded element sequence -> element (or -1)
__global__ void __launch_bounds__(256) plan_touch_kernel(size_t nPos, int Np, const int* __restrict__ elem,
                                                         const dlong* __restrict__ G2L, dlong NlocalT,
                                                         int* __restrict__ firstPos, int* __restrict__ lastPos) {
  const size_t total = nPos * (size_t)Np;
  for (size_t x = (size_t)blockIdx.x * 256 + threadIdx.x; x < total; x += (size_t)gridDim.x * 256) {
    const size_t p = x / Np;
    const int e = elem[p];
    if (e < 0) continue;
    const dlong id = G2L[(size_t)e * Np + (x - p * Np)];
    if (id < 0 || id >= NlocalT) continue;
    atomicMin(&firstPos[id], (int)p);
    atomicMax(&lastPos[id], (int)p);
  }
}
__global__ void __launch_bounds__(256) plan_count_kernel(size_t nPos, int Np, const int* __restrict__ elem,
                                                         const dlong* __restrict__ G2L, dlong NlocalT,
                                                         const int* __restrict__ firstPos, int* __restrict__ cnt) {
  const size_t total = nPos * (size_t)Np;
  for (size_t x = (size_t)blockIdx.x * 256 + threadIdx.x; x < total; x += (size_t)gridDim.x * 256) {
    const size_t p = x / Np;
    const int e = elem[p];
    if (e < 0) continue;
    const dlong id = G2L[(size_t)e * Np + (x - p * Np)];
    if (id < 0 || id >= NlocalT) continue;
    if (firstPos[id] == (int)p) atomicAdd(&cnt[id], 1);
  }
}
// one thread per 32 sectors: bit = 1 -> the sector (4 rows) must be zero-filled before the operator runs
__global__ void __launch_bounds__(256) plan_sector_kernel(size_t nWords, dlong nRows, dlong NlocalT, int L,
                                                          const int* __restrict__ firstPos,
                                                          const int* __restrict__ lastPos, const int* __restrict__ cnt,
                                                          uint32_t* __restrict__ zmask) {
  const size_t w = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (w >= nWords) return;
  uint32_t bits = 0;
  for (int s = 0; s < 32; ++s) {
    const size_t r0 = (w * 32 + s) * 4;
    if (r0 >= (size_t)nRows) break;
    bool clean = true;
    for (int r = 0; r < 4; ++r) {
      const size_t row = r0 + r;
      if (row >= (size_t)nRows || row >= (size_t)NlocalT) { clean = false; break; }
      const int f = firstPos[row], l = lastPos[row];
      if (l < 0 || f / L != l / L || cnt[row] != 1) { clean = false; break; }
    }
    if (!clean) bits |= (1u << s);
  }
  zmask[w] = bits;
}
// one block per position: store/reduce flags, compressed ids, header
__global__ void plan_element_kernel(int Nq, const int* __restrict__ elem, const dlong* __restrict__ G2L, dlong NlocalT,
                                    const int* __restrict__ firstPos, const uint32_t* __restrict__ zmask,
                                    int* __restrict__ hdr, int* __restrict__ cid, uint16_t* __restrict__ flags) {
  const size_t p = blockIdx.x;
  const int Nq2 = Nq * Nq, Np = Nq2 * Nq;
  const int e = elem[p];
  const int t = threadIdx.x;  // Nq2 threads
  if (e < 0) {
    if (t == 0) hdr[p] = -1;
    if (t < Nq2) { flags[p * Nq2 + t] = 0; cid[p * 2 * Nq2 + t] = -1; cid[p * 2 * Nq2 + Nq2 + t] = -1; }
    return;
  }
  const dlong* g = G2L + (size_t)e * Np;
  int ok = 1;
  if (t < Nq2) {
    // as a column (j*Nq + i = t): flags over k
    unsigned f = 0;
    for (int k = 0; k < Nq; ++k) {
      const dlong id = g[k * Nq2 + t];
      if (id >= 0 && id < NlocalT && firstPos[id] == (int)p) {
        const size_t sec = (size_t)id >> 2;
        if (!((zmask[sec >> 5] >> (sec & 31)) & 1u)) f |= (1u << k);
      }
    }
    flags[p * Nq2 + t] = (uint16_t)f;
    // as a line (k*Nq + j = t): is it [x, b, b+1, ...] or [x, -1, -1, ...]?
    const dlong* ln = g + (size_t)t * Nq;
    const dlong id1 = ln[1];
    for (int i = 2; i < Nq; ++i) {
      const dlong want = (id1 < 0) ? -1 : id1 + (i - 1);
      if (ln[i] != want) ok = 0;
    }
    cid[p * 2 * Nq2 + t] = ln[0];
    cid[p * 2 * Nq2 + Nq2 + t] = id1;
  }
  ok = __syncthreads_and(ok);
  if (t == 0) hdr[p] = (e << 1) | (ok ? 0 : 1);
}
__global__ void __launch_bounds__(256) plan_fill_elem_kernel(size_t nPos, int L, const dlong* __restrict__ list,
                                                             dlong first, dlong count, int* __restrict__ elem) {
  // positions of one segment: chains of L consecutive list entries, the last chain padded with -1
  const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= nPos) return;
  elem[p] = (p < (size_t)count) ? (list ? list[first + p] : (dlong)(first + p)) : -1;
}
__global__ void __launch_bounds__(256) plan_stats_kernel(size_t nWords, size_t nSectors, const uint32_t* __restrict__ zmask,
                                                         unsigned long long* __restrict__ out) {
  size_t w = (size_t)blockIdx.x * 256 + threadIdx.x;
  unsigned c = 0;
  if (w < nWords) {
    uint32_t m = zmask[w];
    if ((w + 1) * 32 > nSectors) m &= (nSectors - w * 32 >= 32) ? 0xffffffffu : ((1u << (nSectors - w * 32)) - 1u);
    c = __popc(m);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}
__global__ void __launch_bounds__(256) plan_count_raw_kernel(size_t nPos, const int* __restrict__ hdr,
                                                             unsigned long long* __restrict__ out) {
  size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
  unsigned c = (p < nPos && hdr[p] >= 0 && (hdr[p] & 1)) ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// zero-fill of the sectors the plan marks (4 rows = 32 bytes each); one thread per sector, four sectors of the
// grid-stride loop in flight per thread (the loop is latency-bound on the mask words otherwise)
__device__ __forceinline__ void zero_sector(size_t s, dlong nRows, dfloat* __restrict__ Aq) {
  const size_t r0 = s * 4;
  if (r0 + 4 <= (size_t)nRows) {
    double2* d = reinterpret_cast<double2*>(Aq + r0);
    d[0] = make_double2(0.0, 0.0);
    d[1] = make_double2(0.0, 0.0);
  } else {
    for (size_t r = r0; r < (size_t)nRows; ++r) Aq[r] = 0.0;
  }
}
__global__ void __launch_bounds__(256) zero_fill_kernel(size_t nSectors, dlong nRows, const uint32_t* __restrict__ zmask,
                                                        dfloat* __restrict__ Aq, const int* __restrict__ doneFlag) {
  if (doneFlag != nullptr && *doneFlag) return;
  const size_t stride = (size_t)gridDim.x * 256;
  size_t s = (size_t)blockIdx.x * 256 + threadIdx.x;
  for (; s + 3 * stride < nSectors; s += 4 * stride) {
    uint32_t m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) m[u] = __ldg(zmask + ((s + u * stride) >> 5));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if ((m[u] >> ((s + u * stride) & 31)) & 1u) zero_sector(s + u * stride, nRows, Aq);
  }
  for (; s < nSectors; s += stride)
    if ((__ldg(zmask + (s >> 5)) >> (s & 31)) & 1u) zero_sector(s, nRows, Aq);
}

int grid_for(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 32); }

// Deal the pencils of layouts A and B to the lanes of an element (dense orders, see ChT): the pencils whose first
// word falls into the same 8-byte bank go to different half-warps.
template <int Nq>
const ChainPerm& chain_perm() {
  static const ChainPerm pm = [] {
    using C = ChT<Nq>;
    constexpr int Nq2 = C::Nq2;
    ChainPerm p;
    for (int i = 0; i < 192; ++i) p.A[i] = p.B[i] = ChainPerm::kIdle;
    // deal `n` pencils (code, bank residue) to lane groups of `gs` lanes starting at lane `lane0`: the m-th pencil of
    // a residue goes to the m-th group that does not hold that residue yet
    auto deal = [&](unsigned short* out, int lane0, int lanes, int gs, int nres, int n, auto code, auto residue) {
      const int G = lanes / gs;
      std::vector<int> used(G, 0);
      std::vector<std::vector<char>> has(G, std::vector<char>(nres, 0));
      std::vector<int> next(nres, 0);
      for (int q = 0; q < n; ++q) {
        const int r = residue(q);
        int g = next[r];
        while (g < G && (has[g][r] || used[g] >= gs)) ++g;
        if (g >= G) {  // more copies of a residue than groups: least loaded group, one conflict
          g = 0;
          for (int h = 1; h < G; ++h)
            if (used[h] < used[g]) g = h;
        } else {
          next[r] = g + 1;
        }
        has[g][r] = 1;
        out[lane0 + gs * g + used[g]++] = (unsigned short)code(q);
      }
    };
    if (C::kDense) {  // per element: its own TPE lanes, 64-bit accesses (half-warps, 16 eight-byte banks)
      for (int e = 0; e < C::EPB; ++e) {
        deal(p.A, e * C::TPE, C::TPE, 16, 16, Nq2, [&](int q) { return e * Nq2 + q; },
             [&](int q) { return ((q / Nq) * C::SSr + (q % Nq) * Nq) & 15; });
        deal(p.B, e * C::TPE, C::TPE, 16, 16, Nq2, [&](int q) { return e * Nq2 + q; },
             [&](int q) { return ((q / Nq) * C::SSs + (q % Nq)) & 15; });
      }
    } else if (C::kDeal) {  // whole block; layout A reads 128-bit row pieces (quarter-warps, 8 sixteen-byte banks)
      const int n = C::EPB * Nq2;
      deal(p.A, 0, C::Threads, 8, 8, n, [&](int q) { return q; },
           [&](int q) { const int e = q / Nq2, pc = q % Nq2; return ((e * C::ESr + (pc / Nq) * C::SSr + (pc % Nq) * C::LD) / 2) & 7; });
      deal(p.B, 0, C::Threads, 16, 16, n, [&](int q) { return q; },
           [&](int q) { const int e = q / Nq2, pc = q % Nq2; return (e * C::ESs + (pc / Nq) * C::SSs + (pc % Nq)) & 15; });
    }
    return p;
  }();
  return pm;
}

template <int Nq, int S, bool kDot, bool kScr>
void launch_chain_t(const ChainArgs& A, const EoD& eo, cudaStream_t s) {
  using C = ChT<Nq>;
  constexpr size_t smem = chain_smem_bytes<Nq, S, kScr>();
  // resident blocks by shared memory (227 KB per SM, 1 KB reserved per block)
  constexpr int fit = (int)(232448 / (smem + 1024));
  constexpr int cap = (512 + C::Threads - 1) / C::Threads;  // ~512 threads per SM keeps >= 128 registers per thread
  constexpr int minb = fit < 1 ? 1 : (fit > cap ? cap : fit);
  auto kern = ax_hex3d_chain_kernel<Nq, S, kDot, kScr, minb>;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = (A.nChains + C::EPB - 1) / C::EPB;
  kern<<<grid, C::Threads, smem, s>>>(A, eo, chain_perm<Nq>());
  CUDA_CHECK(cudaGetLastError());
}

template <int Nq, bool kDot, bool kScr>
void launch_chain_s(const ChainArgs& A, const EoD& eo, int stages, cudaStream_t s) {
  if (stages == 3) launch_chain_t<Nq, 3, kDot, kScr>(A, eo, s);
  else if (stages == 2) launch_chain_t<Nq, 2, kDot, kScr>(A, eo, s);
  else launch_chain_t<Nq, 1, kDot, kScr>(A, eo, s);
}

template <int Nq>
int launch_chain(const ChainArgs& A, const EoD& eo, int stages, cudaStream_t s) {
  using C = ChT<Nq>;
  const bool scr = A.lambda != 0.0;
  if (A.dotPartials) {
    if (scr) launch_chain_s<Nq, true, true>(A, eo, stages, s); else launch_chain_s<Nq, true, false>(A, eo, stages, s);
  } else {
    if (scr) launch_chain_s<Nq, false, true>(A, eo, stages, s); else launch_chain_s<Nq, false, false>(A, eo, stages, s);
  }
  return (A.nChains + C::EPB - 1) / C::EPB;
}

template <int Nq>
int launch_chain_tri(const ChainArgs& A, const dfloat* EXYZ, const EoD& eo, const TriConst& gl, cudaStream_t s) {
  using C = ChT<Nq>;
  constexpr int minb = (384 + C::Threads - 1) / C::Threads;  // ~170 registers per thread: no spills at Nq = 8
  const int grid = (A.nChains + C::EPB - 1) / C::EPB;
  const bool scr = A.lambda != 0.0;
#define GO(DOT, SCR) ax_hex3d_chain_tri_kernel<Nq, DOT, SCR, minb><<<grid, C::Threads, 0, s>>>(A, EXYZ, eo, gl, chain_perm<Nq>())
  if (A.dotPartials) { if (scr) GO(true, true); else GO(true, false); }
  else { if (scr) GO(false, true); else GO(false, false); }
#undef GO
  CUDA_CHECK(cudaGetLastError());
  return grid;
}

// Host-side audit of the compiled-in shared-memory geometry of order Nq (no GPU needed): every pencil of every element
// slot is owned by exactly one lane in layouts A and B, all offsets stay inside the arrays, and the number of shared
// wavefronts each access pattern costs under the bank model of DESIGN.md 4.1c (64-bit accesses per half-warp over 16
// eight-byte banks, 128-bit accesses per quarter-warp over 8 sixteen-byte banks) next to the ideal number.
template <int Nq>
void layout_audit(int* ok, int* wf) {
  using C = ChT<Nq>;
  constexpr int Nq2 = C::Nq2, LD = C::LD;
  const ChainPerm& pm = chain_perm<Nq>();
  std::vector<ChIdx<Nq>> X;
  for (int t = 0; t < C::Threads; ++t) X.emplace_back(t, pm);
  bool good = true;
  // ownership: (element slot, pencil) of layouts A / B exactly once
  std::vector<int> ownA(C::EPB * Nq2, 0), ownB(C::EPB * Nq2, 0), ownC(C::EPB * Nq2, 0);
  for (const auto& x : X) {
    if (x.vA) {
      const int e = x.rA / C::ESr, rem = x.rA - e * C::ESr, k = rem / C::SSr, j = (rem - k * C::SSr) / LD;
      good = good && e < C::EPB && k < Nq && j < Nq && (rem - k * C::SSr) % LD == 0;
      good = good && x.uA == e * C::ESu + k * C::SSu + j * LD;
      if (good) ownA[e * Nq2 + k * Nq + j]++;
    }
    if (x.vB) {
      const int e = x.sB / C::ESs, rem = x.sB - e * C::ESs, k = rem / C::SSs, i = rem - k * C::SSs;
      good = good && e < C::EPB && k < Nq && i < Nq;
      good = good && x.uB == e * C::ESu + k * C::SSu + i;
      if (good) ownB[e * Nq2 + k * Nq + i]++;
    }
    if (x.valid) {
      good = good && x.uC == x.es * C::ESu + x.b * LD + x.a && x.es < C::EPB;
      if (good) ownC[x.es * Nq2 + x.ij]++;
    }
  }
  for (int v : ownA) good = good && v == 1;
  for (int v : ownB) good = good && v == 1;
  for (int v : ownC) good = good && v == 1;
  // bounds
  good = good && (Nq - 1) * C::SSu + (Nq - 1) * LD + Nq - 1 < C::ESu && (Nq - 1) * C::SSr + (Nq - 1) * LD + Nq - 1 < C::ESr &&
         (Nq - 1) * C::SSs + (Nq - 1) * LD + Nq - 1 < C::ESs;
  auto cost = [&](int group, int unit_doubles, auto addr, auto on, int& actual, int& ideal) {
    // one instruction: lanes of a `group` share a wavefront unless two of them hit the same bank with different words
    const int nb = 128 / (8 * unit_doubles);
    for (int g0 = 0; g0 < C::Threads; g0 += group) {
      std::vector<std::vector<int>> words(nb);
      bool any = false;
      for (int t = g0; t < g0 + group; ++t) {
        if (!on(X[t])) continue;
        any = true;
        const int w = addr(X[t]) / unit_doubles;
        auto& b = words[w % nb];
        if (std::find(b.begin(), b.end(), w) == b.end()) b.push_back(w);
      }
      if (!any) continue;
      size_t mx = 1;
      for (auto& b : words) mx = std::max(mx, b.size());
      actual += (int)mx;
      ideal += 1;
    }
  };
  for (int i = 0; i < 6; ++i) wf[i] = 0;
  // layout C: one 64-bit access per k on each of s_u, s_r, s_s and the factor slot (same pattern for every k)
  cost(16, 1, [](const ChIdx<Nq>& x) { return x.uC; }, [](const ChIdx<Nq>& x) { return x.valid; }, wf[0], wf[1]);
  cost(16, 1, [](const ChIdx<Nq>& x) { return x.rC; }, [](const ChIdx<Nq>& x) { return x.valid; }, wf[0], wf[1]);
  cost(16, 1, [](const ChIdx<Nq>& x) { return x.sC; }, [](const ChIdx<Nq>& x) { return x.valid; }, wf[0], wf[1]);
  cost(16, 1, [](const ChIdx<Nq>& x) { return x.es * C::slot_doubles(false) + x.ij; },
       [](const ChIdx<Nq>& x) { return x.valid; }, wf[0], wf[1]);
  // layout A: rows of s_u and s_r (128-bit pieces where load_row vectorises, 64-bit otherwise)
  constexpr bool vec = !(Nq == 5 || Nq == 7 || Nq == 9);
  for (int c = 0; c < (vec ? Nq / 2 : Nq); ++c) {
    const int off = vec ? 2 * c : c;
    cost(vec ? 8 : 16, vec ? 2 : 1, [off](const ChIdx<Nq>& x) { return x.uA + off; }, [](const ChIdx<Nq>& x) { return x.vA; }, wf[2], wf[3]);
    cost(vec ? 8 : 16, vec ? 2 : 1, [off](const ChIdx<Nq>& x) { return x.rA + off; }, [](const ChIdx<Nq>& x) { return x.vA; }, wf[2], wf[3]);
  }
  if (vec && (Nq & 1)) {
    cost(16, 1, [](const ChIdx<Nq>& x) { return x.uA + Nq - 1; }, [](const ChIdx<Nq>& x) { return x.vA; }, wf[2], wf[3]);
    cost(16, 1, [](const ChIdx<Nq>& x) { return x.rA + Nq - 1; }, [](const ChIdx<Nq>& x) { return x.vA; }, wf[2], wf[3]);
  }
  // layout B: columns of s_u and s_s
  for (int m = 0; m < Nq; ++m) {
    cost(16, 1, [m](const ChIdx<Nq>& x) { return x.uB + m * C::LD; }, [](const ChIdx<Nq>& x) { return x.vB; }, wf[4], wf[5]);
    cost(16, 1, [m](const ChIdx<Nq>& x) { return x.sB + m * C::LD; }, [](const ChIdx<Nq>& x) { return x.vB; }, wf[4], wf[5]);
  }
  *ok = good ? 1 : 0;
}

}  // namespace

extern "C" int libp_ax_chain_layout_selftest(int Nq, int* ok, int* wavefronts) {
  LIBP_API_BEGIN
  LIBP_CHECK(Nq >= 2 && Nq <= 9 && ok && wavefronts, "bad argument");
  switch (Nq) {
#define CASE(n) case n: layout_audit<n>(ok, wavefronts); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
  }
  LIBP_API_END
}

namespace libp_b200 {

int ax_chain_blocks(int Nq, dlong count, int L) {
  if (count <= 0) return 0;
  const int chains = (int)((count + L - 1) / L);
  static const int epb[10] = {1, 1, ChT<2>::EPB, ChT<3>::EPB, ChT<4>::EPB, ChT<5>::EPB, ChT<6>::EPB, ChT<7>::EPB,
                              ChT<8>::EPB, ChT<9>::EPB};
  return (chains + epb[Nq] - 1) / epb[Nq];
}

void AxChainPlan::build(int Nq_, const dfloat* D_host, dlong nRows_, dlong NlocalT, const dlong* G2L, int L_,
                        const AxChainSegDesc (&sd)[3], cudaStream_t s) {
  Nq = Nq_; L = L_; nRows = nRows_;
  const int Nq2 = Nq * Nq, Np = Nq2 * Nq, H = Nq / 2, N = Nq - 1;
  // even-odd factors of D (host, once)
  EoD e{};
  for (int i = 0; i < H; ++i)
    for (int m = 0; m < H; ++m) {
      e.De[i * H + m] = 0.5 * (D_host[i * Nq + m] + D_host[i * Nq + N - m]);
      e.Do[i * H + m] = 0.5 * (D_host[i * Nq + m] - D_host[i * Nq + N - m]);
    }
  for (int i = 0; i < H; ++i) {
    e.Dc[i] = (Nq & 1) ? D_host[i * Nq + H] : 0.0;
    e.Dr[i] = (Nq & 1) ? D_host[H * Nq + i] : 0.0;
  }
  static_assert(sizeof(EoD) == sizeof(eo), "EoD storage");
  std::memcpy(eo, &e, sizeof(EoD));

  size_t nPos = 0;
  for (int k = 0; k < 3; ++k) {
    seg[k].first = nPos;
    seg[k].count = sd[k].count;
    seg[k].nChains = (int)((sd[k].count + L - 1) / L);
    nPos += (size_t)seg[k].nChains * L;
  }
  nPosTotal = nPos;
  LIBP_CHECK(nPos < (size_t)INT_MAX, "too many elements for the chain plan");
  elem.alloc(nPos);
  hdr.alloc(nPos);
  cid.alloc(nPos * 2 * Nq2);
  flags.alloc(nPos * Nq2);
  nSectors = ((size_t)nRows + 3) / 4;
  const size_t nWords = (nSectors + 31) / 32;
  zmask.alloc(nWords);
  for (int k = 0; k < 3; ++k) {
    const size_t n = (size_t)seg[k].nChains * L;
    if (n) plan_fill_elem_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(n, L, sd[k].list, sd[k].first, sd[k].count,
                                                                       elem.p + seg[k].first);
  }
  CUDA_CHECK(cudaGetLastError());
  dev_buf<int> firstPos, lastPos, cnt;
  const size_t nr = (size_t)std::max<dlong>(nRows, 1);
  firstPos.alloc(nr); lastPos.alloc(nr); cnt.alloc(nr);
  CUDA_CHECK(cudaMemsetAsync(firstPos.p, 0x7f, sizeof(int) * nr, s));  // 0x7f7f7f7f: larger than any position
  CUDA_CHECK(cudaMemsetAsync(lastPos.p, 0xff, sizeof(int) * nr, s));   // -1
  CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, sizeof(int) * nr, s));
  if (nPos) {
    plan_touch_kernel<<<grid_for(nPos * Np), 256, 0, s>>>(nPos, Np, elem.p, G2L, NlocalT, firstPos.p, lastPos.p);
    plan_count_kernel<<<grid_for(nPos * Np), 256, 0, s>>>(nPos, Np, elem.p, G2L, NlocalT, firstPos.p, cnt.p);
  }
  plan_sector_kernel<<<(int)((nWords + 255) / 256), 256, 0, s>>>(nWords, nRows, NlocalT, L, firstPos.p, lastPos.p, cnt.p,
                                                                 zmask.p);
  if (nPos) plan_element_kernel<<<(unsigned)nPos, ((Nq2 + 31) / 32) * 32, 0, s>>>(Nq, elem.p, G2L, NlocalT, firstPos.p,
                                                                                   zmask.p, hdr.p, cid.p, flags.p);
  CUDA_CHECK(cudaGetLastError());
  // statistics (how much of the accumulator still needs the zero-fill; how many elements are uncompressed)
  dev_buf<unsigned long long> st;
  st.alloc(2);
  CUDA_CHECK(cudaMemsetAsync(st.p, 0, 16, s));
  plan_stats_kernel<<<(int)((nWords + 255) / 256), 256, 0, s>>>(nWords, nSectors, zmask.p, st.p);
  if (nPos) plan_count_raw_kernel<<<(int)((nPos + 255) / 256), 256, 0, s>>>(nPos, hdr.p, st.p + 1);
  unsigned long long h[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(h, st.p, 16, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  zeroSectors = (size_t)h[0];
  rawElements = (size_t)h[1];
  built = true;
}

void AxChainPlan::zero_fill(dfloat* Aq, const int* doneFlag, cudaStream_t s) const {
  if (nSectors == 0) return;
  zero_fill_kernel<<<grid_for(nSectors), 256, 0, s>>>(nSectors, nRows, zmask.p, Aq, doneFlag);
  CUDA_CHECK(cudaGetLastError());
}

int AxChainPlan::launch(int k, const dlong* G2L, const dfloat* wJ, const dfloat* ggeo, dfloat lambda, const dfloat* q,
                        dfloat* Aq, dfloat* dotPartials, const int* doneFlag, cudaStream_t s) const {
  const Seg& sg = seg[k];
  if (sg.nChains == 0) return 0;
  ChainArgs A;
  A.hdr = hdr.p + sg.first;
  A.cid = cid.p + sg.first * 2 * Nq * Nq;
  A.flags = flags.p + sg.first * Nq * Nq;
  A.G2L = G2L; A.wJ = wJ; A.ggeo = ggeo; A.q = q; A.Aq = Aq; A.dotPartials = dotPartials; A.doneFlag = doneFlag;
  A.lambda = lambda; A.nChains = sg.nChains; A.L = L; A.count = sg.count;
  EoD e;
  std::memcpy(&e, eo, sizeof(EoD));
  if (EXYZ != nullptr) {  // trilinear element map: geometry on the fly
    TriConst gl;
    for (int i = 0; i < kMaxNq; ++i) { gl.z[i] = i < Nq ? gllz[i] : 0.0; gl.w[i] = i < Nq ? gllw[i] : 0.0; }
    switch (Nq) {
#define CASET(n) case n: return launch_chain_tri<n>(A, EXYZ, e, gl, s);
      CASET(2) CASET(3) CASET(4) CASET(5) CASET(6) CASET(7) CASET(8) CASET(9)
#undef CASET
    }
    return 0;
  }
  switch (Nq) {
#define CASE(n) case n: return launch_chain<n>(A, e, stages, s);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
  }
  return 0;
}

}  // namespace libp_b200
