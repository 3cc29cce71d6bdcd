/* Golden-fixture dump driver for the MULTIGRID preconditioner path (test infrastructure; runs only in the
 * build container, links against the UNMODIFIED reference built by build_ref.sh).
 *
 * Dumps what the reference's p-multigrid + parAlmond hierarchy hands to its apply path, so that the CUDA
 * V-cycle can be pinned against the reference's own numbers:
 *   per matrix-free MGLevel (solvers/elliptic/src/ellipticPreconMultiGridLevel.cpp): degree, P, invDiagA,
 *       Chebyshev bounds, weightG, GlobalToLocal / maskedGlobalIds, D, ggeo, wJ of that degree
 *   per amgLevel (libs/parAlmond/parAlmondAMGLevel.cpp): A, P, R as CSR(+MCSR), diagInv, smoother parameters
 *   coarse exactSolver_t (libs/parAlmond/parAlmondCoarseExact.cpp): invA^T
 *   z = M r for a seeded r (one V-cycle), and the MULTIGRID-PCG solve (iterations, history on stdout)
 *
 * usage: dump_mg_driver setup.rc outdir
 */
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <memory>
#include <map>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <functional>
// the hierarchy lives in private members of precon_t / MultiGridPrecon / parAlmond_t: this driver (our code)
// only needs to READ them, so it opens the access specifiers for the reference headers it includes.
#define private public
#define protected public
#include "elliptic.hpp"
#include "ellipticPrecon.hpp"
#include "parAlmond.hpp"
#include "parAlmond/parAlmondAMGLevel.hpp"
#include "parAlmond/parAlmondCoarseSolver.hpp"
#undef private
#undef protected

using namespace libp;

static std::string g_out;

template <typename T>
static void dump(const std::string& name, const char* dtype, const T* p, size_t n) {
  std::string fn = g_out + "/" + name + "." + dtype + ".bin";
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) { perror(fn.c_str()); exit(1); }
  if (n) fwrite(p, sizeof(T), n, f);
  fclose(f);
}

static inline double splitmix_uniform(uint64_t seed, uint64_t n) {
  uint64_t z = seed + (n + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

static void dumpCSR(const std::string& pre, parAlmond::parCSR& M) {
  long long meta[6] = {M.Nrows, M.Ncols, M.diag.nnz, M.offd.nnz, M.offd.nzRows, M.NlocalCols};
  dump(pre + "_meta", "i64", meta, 6);
  dump(pre + "_rowStarts", "i32", M.diag.rowStarts.ptr(), (size_t)M.Nrows + 1);
  dump(pre + "_cols", "i32", M.diag.cols.ptr(), (size_t)M.diag.nnz);
  dump(pre + "_vals", "f64", M.diag.vals.ptr(), (size_t)M.diag.nnz);
  if (M.offd.nnz) {
    dump(pre + "_offd_mRowStarts", "i32", M.offd.mRowStarts.ptr(), (size_t)M.offd.nzRows + 1);
    dump(pre + "_offd_rows", "i32", M.offd.rows.ptr(), (size_t)M.offd.nzRows);
    dump(pre + "_offd_cols", "i32", M.offd.cols.ptr(), (size_t)M.offd.nnz);
    dump(pre + "_offd_vals", "f64", M.offd.vals.ptr(), (size_t)M.offd.nnz);
  }
}

static void dumpElliptic(const std::string& pre, elliptic_t& e) {
  mesh_t& m = e.mesh;
  const size_t Ntot = (size_t)m.Np * m.Nelements;
  dump(pre + "_GlobalToLocal", "i32", e.GlobalToLocal.ptr(), Ntot);
  dump(pre + "_maskedGlobalIds", "i64", e.maskedGlobalIds.ptr(), Ntot);
  dump(pre + "_weightG", "f64", e.weightG.ptr(), (size_t)e.ogsMasked.Ngather);
  dump(pre + "_D", "f64", m.D.ptr(), (size_t)m.Nq * m.Nq);
  dump(pre + "_ggeo", "f64", m.ggeo.ptr(), Ntot * m.Nggeo);
  dump(pre + "_wJ", "f64", m.wJ.ptr(), Ntot);
  long long cnt[4] = {m.N, (long long)m.Nelements, e.ogsMasked.Ngather, e.gHalo.Nhalo};
  dump(pre + "_counts", "i64", cnt, 4);
}

int main(int argc, char** argv) {
  Comm::Init(argc, argv);
  LIBP_ABORT("Usage: ./dump_mg_driver setupfile outdir", argc != 3);
  g_out = argv[2];
  {
    comm_t comm(Comm::World().Dup());
    platformSettings_t platformSettings(comm);
    meshSettings_t meshSettings(comm);
    ellipticSettings_t ellipticSettings(comm);
    ellipticAddRunSettings(ellipticSettings);
    ellipticSettings.parseFromFile(platformSettings, meshSettings, argv[1]);

    platform_t platform(platformSettings);
    mesh_t mesh(platform, meshSettings, comm);
    dfloat lambda = 0.0;
    ellipticSettings.getSetting("LAMBDA", lambda);
    memory<int> BCType(3);
    BCType[0] = 0; BCType[1] = 1; BCType[2] = 2;
    elliptic_t elliptic(platform, mesh, ellipticSettings, lambda, 3, BCType);
    mesh_t& m = elliptic.mesh;
    const size_t Ntot = (size_t)m.Np * m.Nelements;
    const dlong Ndofs = elliptic.Ndofs, Nhalo = elliptic.Nhalo;

    MultiGridPrecon* mgp = dynamic_cast<MultiGridPrecon*>(elliptic.precon.precon.get());
    LIBP_ABORT("PRECONDITIONER must be MULTIGRID", mgp == nullptr);
    parAlmond::multigrid_t& mg = *mgp->parAlmond.multigrid;

    int meta[8] = {m.N, m.Nq, m.Np, (int)m.Nelements, mg.numLevels, mg.baseLevel, elliptic.allNeumann, (int)mg.ctype};
    dump("meta", "i32", meta, 8);
    double dmeta[1] = {lambda};
    dump("dmeta", "f64", dmeta, 1);
    dumpElliptic("fine", elliptic);

    std::vector<int> kinds;  // 0 = matrix-free MGLevel, 1 = amgLevel
    for (int l = 0; l < mg.numLevels - 1; ++l) {
      std::string pre = "L" + std::to_string(l);
      if (MGLevel* L = dynamic_cast<MGLevel*>(mg.levels[l].get())) {
        kinds.push_back(0);
        long long lm[8] = {L->mesh.N, L->Nrows, L->Ncols, (long long)L->stype, L->ChebyshevIterations, L->meshC.N,
                           (long long)L->mesh.Nelements, 0};
        dump(pre + "_meta", "i64", lm, 8);
        double ld[2] = {L->lambda0, L->lambda1};
        dump(pre + "_lambda", "f64", ld, 2);
        dump(pre + "_P", "f64", L->P.ptr(), (size_t)L->mesh.Nq * L->meshC.Nq);
        memory<dfloat> inv(L->Nrows);
        L->o_invDiagA.copyTo(inv);
        dump(pre + "_invDiagA", "f64", inv.ptr(), (size_t)L->Nrows);
        dumpElliptic(pre + "_F", L->elliptic);
        dumpElliptic(pre + "_C", L->ellipticC);
      } else if (parAlmond::amgLevel* A = dynamic_cast<parAlmond::amgLevel*>(mg.levels[l].get())) {
        kinds.push_back(1);
        long long lm[4] = {A->Nrows, A->Ncols, (long long)A->stype, A->ChebyshevIterations};
        dump(pre + "_meta", "i64", lm, 4);
        double ld[3] = {A->lambda0, A->lambda1, A->lambda};
        dump(pre + "_lambda", "f64", ld, 3);
        dumpCSR(pre + "_A", A->A);
        dumpCSR(pre + "_P", A->P);
        dumpCSR(pre + "_R", A->R);
        dump(pre + "_diagInv", "f64", A->A.diagInv.ptr(), (size_t)A->A.Nrows);
      } else {
        LIBP_FORCE_ABORT("unknown level type");
      }
    }
    dump("level_kinds", "i32", kinds.data(), kinds.size());
    {
      parAlmond::exactSolver_t* cs = dynamic_cast<parAlmond::exactSolver_t*>(mg.coarseSolver.get());
      LIBP_ABORT("coarse solver must be the exact solver", cs == nullptr);
      long long cm[3] = {cs->N, cs->coarseTotal, cs->offdTotal};
      dump("coarse_meta", "i64", cm, 3);
      dump("coarse_diagInvAT", "f64", cs->diagInvAT.ptr(), (size_t)cs->N * cs->N);
      dumpCSR("coarse_A", cs->A);
    }

    // ---- one preconditioner apply on a seeded vector ----
    memory<dfloat> r(Ndofs + Nhalo, 0.0), z(Ndofs + Nhalo, 0.0);
    for (dlong n = 0; n < Ndofs; ++n) r[n] = splitmix_uniform(4321, (uint64_t)n);
    deviceMemory<dfloat> o_r = platform.malloc<dfloat>(r);
    deviceMemory<dfloat> o_z = platform.malloc<dfloat>(z);
    elliptic.precon.Operator(o_r, o_z);
    o_z.copyTo(z);
    dump("vc_r", "f64", r.ptr(), (size_t)Ndofs);
    dump("vc_z", "f64", z.ptr(), (size_t)Ndofs);

    // ---- level-0 pieces on their own (smooth with zero guess, residual, coarsen, prolongate) ----
    if (MGLevel* L0 = dynamic_cast<MGLevel*>(mg.levels[0].get())) {
      memory<dfloat> x(L0->Ncols, 0.0), res(L0->Ncols, 0.0);
      deviceMemory<dfloat> o_x = platform.malloc<dfloat>(x);
      deviceMemory<dfloat> o_res = platform.malloc<dfloat>(res);
      L0->smooth(o_r, o_x, true);
      o_x.copyTo(x);
      dump("l0_smooth0", "f64", x.ptr(), (size_t)Ndofs);
      L0->residual(o_r, o_x, o_res);
      o_res.copyTo(res);
      dump("l0_residual", "f64", res.ptr(), (size_t)Ndofs);
      L0->smooth(o_r, o_x, false);
      o_x.copyTo(x);
      dump("l0_smooth1", "f64", x.ptr(), (size_t)Ndofs);
      const dlong NrowsC = L0->ellipticC.ogsMasked.Ngather, NcolsC = NrowsC + L0->ellipticC.gHalo.Nhalo;
      memory<dfloat> rc(NcolsC, 0.0);
      deviceMemory<dfloat> o_rc = platform.malloc<dfloat>(rc);
      L0->coarsen(o_res, o_rc);
      o_rc.copyTo(rc);
      dump("l0_coarsen", "f64", rc.ptr(), (size_t)NrowsC);
      memory<dfloat> px(L0->Ncols, 0.0);
      deviceMemory<dfloat> o_px = platform.malloc<dfloat>(px);
      L0->prolongate(o_rc, o_px);
      o_px.copyTo(px);
      dump("l0_prolongate", "f64", px.ptr(), (size_t)Ndofs);
    }

    // ---- the MULTIGRID-PCG solve on the reference right-hand side ----
    properties_t kernelInfo = m.props;
    std::string dataFileName;
    ellipticSettings.getSetting("DATA FILE", dataFileName);
    kernelInfo["includes"] += dataFileName;
    kernelInfo["includes"] += std::string(DELLIPTIC "/data/ellipticBoundary3D.h");
    kernelInfo["defines/" "p_Nmax"] = std::max(m.Np, m.Nfaces * m.Nfp);
    kernelInfo["defines/" "p_Nfields"] = 1;
    kernel_t forcingKernel = platform.buildKernel(DELLIPTIC "/okl/ellipticRhsHex3D.okl", "ellipticRhsHex3D", kernelInfo);
    kernel_t rhsBCKernel = platform.buildKernel(DELLIPTIC "/okl/ellipticRhsBCHex3D.okl", "ellipticRhsBCHex3D", kernelInfo);
    memory<dfloat> rL(Ntot, 0.0), xL(Ntot, 0.0);
    deviceMemory<dfloat> o_rL = platform.malloc<dfloat>(rL);
    deviceMemory<dfloat> o_xL = platform.malloc<dfloat>(xL);
    deviceMemory<dfloat> o_rhs = platform.malloc<dfloat>(Ndofs + Nhalo);
    deviceMemory<dfloat> o_x = platform.malloc<dfloat>(Ndofs + Nhalo);
    forcingKernel(m.Nelements, m.o_wJ, m.o_MM, m.o_x, m.o_y, m.o_z, lambda, o_rL);
    rhsBCKernel(m.Nelements, m.o_wJ, m.o_ggeo, m.o_sgeo, m.o_D, m.o_S, m.o_MM, m.o_vmapM, m.o_sM,
                lambda, m.o_x, m.o_y, m.o_z, elliptic.o_mapB, o_rL);
    elliptic.ogsMasked.Gather(o_rhs, o_rL, 1, ogs::Add, ogs::Trans);
    elliptic.ogsMasked.Gather(o_x, o_xL, 1, ogs::Add, ogs::NoTrans);
    memory<dfloat> rhs(Ndofs);
    o_rhs.copyTo(rhs, Ndofs);
    dump("r", "f64", rhs.ptr(), (size_t)Ndofs);
    linearSolver_t linearSolver;
    linearSolver.Setup<LinearSolver::pcg>(Ndofs, Nhalo, platform, ellipticSettings, comm);
    int iter = elliptic.Solve(linearSolver, o_x, o_rhs, 1.0e-8, 5000, 1);
    memory<dfloat> x(Ndofs);
    o_x.copyTo(x, Ndofs);
    dump("xsol", "f64", x.ptr(), (size_t)Ndofs);
    int imeta[1] = {iter};
    dump("iterations", "i32", imeta, 1);
    printf("ITERATIONS = %d\n", iter);
  }
  Comm::Finalize();
  return 0;
}
