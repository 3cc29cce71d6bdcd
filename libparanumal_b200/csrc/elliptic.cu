// elliptic_t::Operator, continuous-Galerkin branch (solvers/elliptic/src/ellipticOperator.cpp:31-106):
//   Aq = Z^T A_L Z q on gathered vectors [NlocalT local | NhaloP owned-shared | Nhalo received].
// Same element-list split as the reference so that the halo exchange of q overlaps the interior
// elements and the cross-rank combine of the shared rows overlaps the second half:
//   halo(q) start | Ax(local[0:n/2]) | halo finish | Ax(global) | combine start | Ax(local[n/2:]) | finish
// mode 0 keeps the reference data flow (AqL scratch + ogs gather, left-to-right row sums);
// mode 1 fuses the gather into the Ax epilogue (FP64 reductions straight into Aq).
#include "elliptic.hpp"

using namespace libp_b200;

void libp_elliptic_s::apply(dfloat* q, dfloat* Aq, bool want_dot, const int* doneFlag, cudaStream_t s, bool zeroed) {
  libp_ogs_s& ogs = *d.ogsMasked;
  const dlong nL = d.NlocalGatherElements, nG = d.NglobalGatherElements;
  const dlong nL0 = nL / 2, nL1 = (nL + 1) / 2;
  const bool fused = d.mode == 1;
  dfloat* out = fused ? Aq : AqL.p;
  dfloat* dp = want_dot ? dotPartials.p : nullptr;
  int doff = 0;
  auto ax = [&](dlong n, const dlong* list) {
    if (n <= 0) return;
    int nb = ax_hex3d_launch(d.Nq, fused, true, symD, n, list, d.GlobalToLocal, d.wJ, d.ggeo, d.D, d.lambda, q, out,
                             dp ? dp + doff : nullptr, doneFlag, s);
    doff += nb;
  };
  if (fused && !zeroed)
    CUDA_CHECK(cudaMemsetAsync(Aq, 0, sizeof(dfloat) * (size_t)(ogs.NlocalT + ogs.NhaloT), s));
  halo_start_f64(ogs, q, s);
  ax(nL0, d.localGatherElementList);
  halo_finish_f64(ogs, q, s);
  ax(nG, d.globalGatherElementList);
  if (fused) {
    halo_combine_start_f64(ogs, Aq, s);
    ax(nL1, d.localGatherElementList ? d.localGatherElementList + nL0 : nullptr);
    halo_combine_finish_f64(ogs, Aq, s);
  } else {
    ogs_gather_start_f64(ogs, Aq, AqL.p, LIBP_ADD, LIBP_TRANS, s);
    ax(nL1, d.localGatherElementList ? d.localGatherElementList + nL0 : nullptr);
    ogs_gather_finish_f64(ogs, Aq, AqL.p, LIBP_ADD, LIBP_TRANS, s);
  }
  nDotPartials = doff;
}

extern "C" int libp_elliptic_create(const libp_elliptic_desc_t* desc, libp_elliptic_t* op) {
  LIBP_API_BEGIN
  LIBP_CHECK(desc && op, "null argument");
  LIBP_CHECK(desc->Nq >= 2 && desc->Nq <= 9, "Nq must be in [2, 9]");
  LIBP_CHECK(desc->ogsMasked != nullptr, "ogsMasked handle required");
  LIBP_CHECK(desc->mode == 0 || desc->mode == 1, "mode must be 0 or 1");
  LIBP_CHECK(desc->NlocalGatherElements + desc->NglobalGatherElements == desc->Nelements,
             "element lists must cover all elements");
  LIBP_CHECK(desc->GlobalToLocal && desc->wJ && desc->ggeo && desc->D, "null device pointer");
  LIBP_CHECK(desc->NlocalGatherElements == 0 || desc->localGatherElementList, "null local element list");
  LIBP_CHECK(desc->NglobalGatherElements == 0 || desc->globalGatherElementList, "null global element list");
  std::unique_ptr<libp_elliptic_s> e(new libp_elliptic_s());
  e->d = *desc;
  e->Np = desc->Nq * desc->Nq * desc->Nq;
  {
    double hD[81];
    CUDA_CHECK(cudaMemcpy(hD, desc->D, sizeof(double) * desc->Nq * desc->Nq, cudaMemcpyDeviceToHost));
    e->symD = ax_hex3d_D_is_centro_antisymmetric(desc->Nq, hD);
  }
  e->Ndofs = desc->ogsMasked->Ngather;
  e->Nhalo = desc->ogsMasked->NhaloT - desc->ogsMasked->NhaloP;
  if (desc->mode == 0) e->AqL.alloc((size_t)desc->Nelements * e->Np);
  e->dotPartials.alloc((size_t)ax_hex3d_blocks(desc->Nq, desc->NlocalGatherElements / 2) +
                       ax_hex3d_blocks(desc->Nq, (desc->NlocalGatherElements + 1) / 2) +
                       ax_hex3d_blocks(desc->Nq, desc->NglobalGatherElements) + 4);
  *op = e.release();
  LIBP_API_END
}

extern "C" int libp_elliptic_free(libp_elliptic_t op) {
  LIBP_API_BEGIN
  delete op;
  LIBP_API_END
}

extern "C" int libp_elliptic_operator(libp_elliptic_t op, libp_dfloat* q, libp_dfloat* Aq, void* stream) {
  LIBP_API_BEGIN
  LIBP_CHECK(op && q && Aq, "null argument");
  op->apply(q, Aq, false, nullptr, as_stream(stream));
  LIBP_API_END
}
