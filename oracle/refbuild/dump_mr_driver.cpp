/* Multi-rank golden-fixture dump driver (test infrastructure; runs only in the build container, under
 * oracle/refbuild/mpirun_stub.sh with the multi-process MPI stand-in).
 *
 * Links against the UNMODIFIED reference (libelliptic.a + libs) and writes, PER RANK, what the multi-rank half of the
 * hot path is pinned against:
 *   mesh:      D, ggeo, wJ, globalIds, element lists (box decomposition of libs/mesh/meshSetupBoxHex3D.cpp)
 *   ogs:       maskedGlobalIds (post-setup, signed: owner choice across ranks), GlobalToLocal, counters
 *              (N, Ngather, NlocalT/P, NhaloT/P, NgatherGlobal, Nhalo), gatherLocal / gatherHalo CSR maps
 *   pairwise:  the ogsPairwise_t the reference builds for these ids (libs/ogs/ogsPairwise.cpp:194-415): send lists
 *              sendIdsN/T, send/recv ranks, counts, offsets, and the postmpi combine operator
 *   operator:  Aq = elliptic.Operator(q) on every rank for a seeded q (halo exchange + Ax + cross-rank combine)
 *   solve:     Jacobi/None PCG iteration count and the gathered solution
 *
 * usage: mpirun_stub.sh P dump_mr_driver setup.rc outdir        (writes outdir/r<rank>/...)
 */
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <sys/stat.h>

// the pairwise lists are private members of ogsPairwise_t; the driver is our own translation unit, the reference
// itself is compiled unmodified (access specifiers do not change the layout GCC gives these classes)
#include "elliptic.hpp"
#include "ogs/ogsOperator.hpp"
#define private public
#define protected public
#include "ogs/ogsExchange.hpp"  // everything it includes is already in (guards): only its own classes open up
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

struct OgsPeek : public ogs::ogs_t {
  ogs::ogsOperator_t& local() { return *gatherLocal; }
  ogs::ogsOperator_t& halo() { return *gatherHalo; }
  ogs::ogsExchange_t* ex() { return exchange.get(); }
};

static void dumpOp(const std::string& pre, ogs::ogsOperator_t& op) {
  dlong cnt[5] = {op.Ncols, op.NrowsN, op.NrowsT, op.nnzN, op.nnzT};
  dump(pre + "_counts", "i32", cnt, 5);
  dump(pre + "_rowStartsN", "i32", op.rowStartsN.ptr(), (size_t)op.NrowsT + 1);
  dump(pre + "_rowStartsT", "i32", op.rowStartsT.ptr(), (size_t)op.NrowsT + 1);
  dump(pre + "_colIdsN", "i32", op.colIdsN.ptr(), (size_t)op.nnzN);
  dump(pre + "_colIdsT", "i32", op.colIdsT.ptr(), (size_t)op.nnzT);
}

static inline double splitmix_uniform(uint64_t seed, uint64_t n) {
  uint64_t z = seed + (n + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

int main(int argc, char** argv) {
  Comm::Init(argc, argv);
  LIBP_ABORT("Usage: ./dump_mr_driver setupfile outdir", argc != 3);
  {
    comm_t comm(Comm::World().Dup());
    const int rank = comm.rank(), size = comm.size();
    g_out = std::string(argv[2]) + "/r" + std::to_string(rank);
    mkdir(argv[2], 0777);
    mkdir(g_out.c_str(), 0777);

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
    int meta[10] = {m.N, m.Nq, m.Np, (int)m.Nelements, (int)m.NlocalGatherElements, (int)m.NglobalGatherElements,
                    elliptic.allNeumann, (int)elliptic.Nmasked, rank, size};
    dump("meta", "i32", meta, 10);
    double dmeta[1] = {lambda};
    dump("dmeta", "f64", dmeta, 1);
    dump("D", "f64", m.D.ptr(), (size_t)m.Nq * m.Nq);
    dump("ggeo", "f64", m.ggeo.ptr(), Ntot * m.Nggeo);
    dump("wJ", "f64", m.wJ.ptr(), Ntot);
    dump("globalIds", "i64", m.globalIds.ptr(), Ntot);
    dump("localGatherElementList", "i32", m.localGatherElementList.ptr(), (size_t)m.NlocalGatherElements);
    dump("globalGatherElementList", "i32", m.globalGatherElementList.ptr(), (size_t)m.NglobalGatherElements);
    dump("mapB", "i32", elliptic.mapB.ptr(), Ntot);
    dump("maskedGlobalIds", "i64", elliptic.maskedGlobalIds.ptr(), Ntot);
    dump("GlobalToLocal", "i32", elliptic.GlobalToLocal.ptr(), Ntot);
    ogs::ogs_t& o = elliptic.ogsMasked;
    long long cnt[8] = {o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.NgatherGlobal, elliptic.gHalo.Nhalo};
    dump("ogs_counts", "i64", cnt, 8);
    OgsPeek& peek = static_cast<OgsPeek&>(o);
    dumpOp("gatherLocal", peek.local());
    dumpOp("gatherHalo", peek.halo());

    // ---- the pairwise exchange of these ids: a second setup on the already signed ids (unique = false: no rand(),
    // same maps) with Method = Pairwise, so that the object is an ogsPairwise_t whatever AutoSetup timed fastest
    {
      memory<hlong> ids(Ntot);
      ids.copyFrom(elliptic.maskedGlobalIds);
      ogs::ogs_t o2;
      o2.Setup((dlong)Ntot, ids, comm, ogs::Signed, ogs::Pairwise, false, false, platform);
      OgsPeek& p2 = static_cast<OgsPeek&>(o2);
      LIBP_ABORT("second setup changed the counters",
                 o2.Ngather != o.Ngather || o2.NlocalT != o.NlocalT || o2.NhaloT != o.NhaloT || o2.NhaloP != o.NhaloP);
      bool same = p2.halo().nnzT == peek.halo().nnzT && p2.halo().nnzN == peek.halo().nnzN;
      for (dlong i = 0; same && i < p2.halo().nnzT; ++i) same = p2.halo().colIdsT[i] == peek.halo().colIdsT[i];
      for (dlong i = 0; same && i < p2.halo().nnzN; ++i) same = p2.halo().colIdsN[i] == peek.halo().colIdsN[i];
      LIBP_ABORT("second setup changed the halo gather maps", !same);
      ogs::ogsPairwise_t* pw = dynamic_cast<ogs::ogsPairwise_t*>(p2.ex());
      LIBP_ABORT("exchange is not an ogsPairwise_t", pw == nullptr);
      int pc[8] = {(int)pw->NsendN, (int)pw->NsendT, pw->NranksSendN, pw->NranksRecvN, pw->NranksSendT, pw->NranksRecvT,
                   (int)pw->Nhalo, (int)pw->NhaloP};
      dump("pw_counts", "i32", pc, 8);
      dump("pw_sendIdsN", "i32", pw->sendIdsN.ptr(), (size_t)pw->NsendN);
      dump("pw_sendIdsT", "i32", pw->sendIdsT.ptr(), (size_t)pw->NsendT);
      dump("pw_sendRanksN", "i32", pw->sendRanksN.ptr(), (size_t)pw->NranksSendN);
      dump("pw_recvRanksN", "i32", pw->recvRanksN.ptr(), (size_t)pw->NranksRecvN);
      dump("pw_sendRanksT", "i32", pw->sendRanksT.ptr(), (size_t)pw->NranksSendT);
      dump("pw_recvRanksT", "i32", pw->recvRanksT.ptr(), (size_t)pw->NranksRecvT);
      dump("pw_sendCountsN", "i32", pw->sendCountsN.ptr(), (size_t)pw->NranksSendN);
      dump("pw_recvCountsN", "i32", pw->recvCountsN.ptr(), (size_t)pw->NranksRecvN);
      dump("pw_sendCountsT", "i32", pw->sendCountsT.ptr(), (size_t)pw->NranksSendT);
      dump("pw_recvCountsT", "i32", pw->recvCountsT.ptr(), (size_t)pw->NranksRecvT);
      dump("pw_sendOffsetsN", "i32", pw->sendOffsetsN.ptr(), (size_t)pw->NranksSendN + 1);
      dump("pw_recvOffsetsN", "i32", pw->recvOffsetsN.ptr(), (size_t)pw->NranksRecvN + 1);
      dump("pw_sendOffsetsT", "i32", pw->sendOffsetsT.ptr(), (size_t)pw->NranksSendT + 1);
      dump("pw_recvOffsetsT", "i32", pw->recvOffsetsT.ptr(), (size_t)pw->NranksRecvT + 1);
      dumpOp("pw_postmpi", pw->postmpi);
    }

    const dlong Ndofs = elliptic.Ndofs, Nhalo = elliptic.Nhalo;
    memory<dfloat> diagA(Ndofs);
    elliptic.BuildOperatorDiagonal(diagA);
    dump("diagA", "f64", diagA.ptr(), (size_t)Ndofs);

    // ---- operator apply on a seeded vector (seed depends on the rank: q is a function of (rank, gathered index))
    memory<dfloat> q(Ndofs + Nhalo, 0.0), Aq(Ndofs + Nhalo, 0.0);
    for (dlong n = 0; n < Ndofs; ++n) q[n] = splitmix_uniform(1234 + 7919 * (uint64_t)rank, (uint64_t)n);
    deviceMemory<dfloat> o_q = platform.malloc<dfloat>(q);
    deviceMemory<dfloat> o_Aq = platform.malloc<dfloat>(Aq);
    elliptic.Operator(o_q, o_Aq);
    o_Aq.copyTo(Aq);
    o_q.copyTo(q);
    dump("q", "f64", q.ptr(), (size_t)Ndofs);
    dump("q_after_halo", "f64", q.ptr(), (size_t)(Ndofs + Nhalo));  // the exchange fills the tail of the input
    dump("Aq", "f64", Aq.ptr(), (size_t)Ndofs);

    // ---- rhs as elliptic_t::Run builds it, then the PCG solve
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
    deviceMemory<dfloat> o_r = platform.malloc<dfloat>(Ndofs + Nhalo);
    deviceMemory<dfloat> o_x = platform.malloc<dfloat>(Ndofs + Nhalo);
    forcingKernel(m.Nelements, m.o_wJ, m.o_MM, m.o_x, m.o_y, m.o_z, lambda, o_rL);
    rhsBCKernel(m.Nelements, m.o_wJ, m.o_ggeo, m.o_sgeo, m.o_D, m.o_S, m.o_MM, m.o_vmapM, m.o_sM,
                lambda, m.o_x, m.o_y, m.o_z, elliptic.o_mapB, o_rL);
    elliptic.ogsMasked.Gather(o_r, o_rL, 1, ogs::Add, ogs::Trans);
    elliptic.ogsMasked.Gather(o_x, o_xL, 1, ogs::Add, ogs::NoTrans);
    memory<dfloat> r(Ndofs);
    o_r.copyTo(r, Ndofs);
    dump("r", "f64", r.ptr(), (size_t)Ndofs);
    linearSolver_t linearSolver;
    linearSolver.Setup<LinearSolver::pcg>(Ndofs, Nhalo, platform, ellipticSettings, comm);
    int iter = elliptic.Solve(linearSolver, o_x, o_r, 1.0e-8, 5000, 1);
    memory<dfloat> x(Ndofs);
    o_x.copyTo(x, Ndofs);
    dump("xsol", "f64", x.ptr(), (size_t)Ndofs);
    int imeta[1] = {iter};
    dump("iterations", "i32", imeta, 1);
    if (rank == 0) printf("ITERATIONS = %d\n", iter);
  }
  Comm::Finalize();
  return 0;
}
