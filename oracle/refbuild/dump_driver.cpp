/* Golden-fixture dump driver (test infrastructure; runs only in the build container).
 *
 * Links against the UNMODIFIED reference (libelliptic.a + libs, built by build_ref.sh) and
 * writes the arrays on the elliptic hot path to raw binary files so that oracle/ and the
 * CUDA path can be pinned against the reference's own output:
 *   mesh:      D, gllz, gllw, x,y,z, ggeo, wJ, globalIds, mapB, element lists
 *   elliptic:  maskedGlobalIds (post-setup, signed), GlobalToLocal, ogsMasked counters and
 *              the gatherLocal / gatherHalo CSR maps, weightG, diagA
 *   operator:  Aq = elliptic.Operator(q) for a splitmix64-seeded q
 *   solve:     gathered rhs r, solution x, iteration count (residual history is on stdout)
 *
 * usage: dump_driver setup.rc outdir
 */
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include "elliptic.hpp"
#include "ogs/ogsOperator.hpp"
#include "timer.hpp"

using namespace libp;

static std::string g_out;

template <typename T>
static void dump(const char* name, const char* dtype, const T* p, size_t n) {
  std::string fn = g_out + "/" + name + "." + dtype + ".bin";
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) { perror(fn.c_str()); exit(1); }
  if (n) fwrite(p, sizeof(T), n, f);
  fclose(f);
}

struct OgsPeek : public ogs::ogs_t {
  ogs::ogsOperator_t& local() { return *gatherLocal; }
  ogs::ogsOperator_t& halo() { return *gatherHalo; }
};

static void dumpOp(const std::string& pre, ogs::ogsOperator_t& op) {
  dlong cnt[5] = {op.Ncols, op.NrowsN, op.NrowsT, op.nnzN, op.nnzT};
  dump((pre + "_counts").c_str(), "i32", cnt, 5);
  dump((pre + "_rowStartsN").c_str(), "i32", op.rowStartsN.ptr(), (size_t)op.NrowsT + 1);
  dump((pre + "_rowStartsT").c_str(), "i32", op.rowStartsT.ptr(), (size_t)op.NrowsT + 1);
  dump((pre + "_colIdsN").c_str(), "i32", op.colIdsN.ptr(), (size_t)op.nnzN);
  dump((pre + "_colIdsT").c_str(), "i32", op.colIdsT.ptr(), (size_t)op.nnzT);
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
  LIBP_ABORT("Usage: ./dump_driver setupfile outdir", argc != 3);
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
    int meta[8] = {m.N, m.Nq, m.Np, (int)m.Nelements, (int)m.NlocalGatherElements,
                   (int)m.NglobalGatherElements, elliptic.allNeumann, (int)elliptic.Nmasked};
    dump("meta", "i32", meta, 8);
    double dmeta[2] = {lambda, elliptic.allNeumannScale};
    dump("dmeta", "f64", dmeta, 2);
    dump("D", "f64", m.D.ptr(), (size_t)m.Nq * m.Nq);
    dump("gllz", "f64", m.gllz.ptr(), (size_t)m.Nq);
    dump("gllw", "f64", m.gllw.ptr(), (size_t)m.Nq);
    dump("x", "f64", m.x.ptr(), Ntot);
    dump("y", "f64", m.y.ptr(), Ntot);
    dump("z", "f64", m.z.ptr(), Ntot);
    dump("ggeo", "f64", m.ggeo.ptr(), Ntot * m.Nggeo);
    dump("wJ", "f64", m.wJ.ptr(), Ntot);
    dump("globalIds", "i64", m.globalIds.ptr(), Ntot);
    dump("meshMapB", "i32", m.mapB.ptr(), Ntot);
    dump("localGatherElementList", "i32", m.localGatherElementList.ptr(), (size_t)m.NlocalGatherElements);
    dump("globalGatherElementList", "i32", m.globalGatherElementList.ptr(), (size_t)m.NglobalGatherElements);

    dump("mapB", "i32", elliptic.mapB.ptr(), Ntot);
    dump("maskedGlobalIds", "i64", elliptic.maskedGlobalIds.ptr(), Ntot);
    dump("GlobalToLocal", "i32", elliptic.GlobalToLocal.ptr(), Ntot);
    ogs::ogs_t& o = elliptic.ogsMasked;
    long long cnt[8] = {o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.NgatherGlobal,
                        elliptic.gHalo.Nhalo};
    dump("ogs_counts", "i64", cnt, 8);
    OgsPeek& peek = static_cast<OgsPeek&>(o);
    dumpOp("gatherLocal", peek.local());
    dumpOp("gatherHalo", peek.halo());
    dump("weightG", "f64", elliptic.weightG.ptr(), (size_t)o.Ngather);

    const dlong Ndofs = elliptic.Ndofs, Nhalo = elliptic.Nhalo;
    memory<dfloat> diagA(Ndofs);
    elliptic.BuildOperatorDiagonal(diagA);
    dump("diagA", "f64", diagA.ptr(), (size_t)Ndofs);

    // ---- operator apply on a seeded vector ----
    memory<dfloat> q(Ndofs + Nhalo, 0.0), Aq(Ndofs + Nhalo, 0.0);
    for (dlong n = 0; n < Ndofs; ++n) q[n] = splitmix_uniform(1234, (uint64_t)n);
    deviceMemory<dfloat> o_q = platform.malloc<dfloat>(q);
    deviceMemory<dfloat> o_Aq = platform.malloc<dfloat>(Aq);
    elliptic.Operator(o_q, o_Aq);
    o_Aq.copyTo(Aq);
    dump("q", "f64", q.ptr(), (size_t)Ndofs);
    dump("Aq", "f64", Aq.ptr(), (size_t)Ndofs);

    // ---- timing of the operator (CPU baseline figure) ----
    {
      int nrep = 20;
      for (int i = 0; i < 3; ++i) elliptic.Operator(o_q, o_Aq);
      platform.finish();
      timePoint_t t0 = GlobalPlatformTime(platform);
      for (int i = 0; i < nrep; ++i) elliptic.Operator(o_q, o_Aq);
      timePoint_t t1 = GlobalPlatformTime(platform);
      double el = ElapsedTime(t0, t1) / nrep;
      printf("AXTIME: Ndofs=%d sec_per_apply=%g GDOF/s=%g\n", (int)Ndofs, el, Ndofs / el / 1e9);
    }

    // ---- rhs as elliptic_t::Run builds it (solvers/elliptic/src/ellipticRun.cpp:84-185) ----
    properties_t kernelInfo = m.props;
    std::string dataFileName;
    ellipticSettings.getSetting("DATA FILE", dataFileName);
    kernelInfo["includes"] += dataFileName;
    kernelInfo["includes"] += std::string(DELLIPTIC "/data/ellipticBoundary3D.h");
    kernelInfo["defines/" "p_Nmax"] = std::max(m.Np, m.Nfaces * m.Nfp);
    kernelInfo["defines/" "p_Nfields"] = 1;
    kernel_t forcingKernel = platform.buildKernel(DELLIPTIC "/okl/ellipticRhsHex3D.okl", "ellipticRhsHex3D", kernelInfo);
    kernel_t rhsBCKernel = platform.buildKernel(DELLIPTIC "/okl/ellipticRhsBCHex3D.okl", "ellipticRhsBCHex3D", kernelInfo);
    kernel_t addBCKernel = platform.buildKernel(DELLIPTIC "/okl/ellipticAddBCHex3D.okl", "ellipticAddBCHex3D", kernelInfo);

    memory<dfloat> rL(Ntot, 0.0), xL(Ntot, 0.0);
    deviceMemory<dfloat> o_rL = platform.malloc<dfloat>(rL);
    deviceMemory<dfloat> o_xL = platform.malloc<dfloat>(xL);
    deviceMemory<dfloat> o_r = platform.malloc<dfloat>(Ndofs + Nhalo);
    deviceMemory<dfloat> o_x = platform.malloc<dfloat>(Ndofs + Nhalo);
    forcingKernel(m.Nelements, m.o_wJ, m.o_MM, m.o_x, m.o_y, m.o_z, lambda, o_rL);
    rhsBCKernel(m.Nelements, m.o_wJ, m.o_ggeo, m.o_sgeo, m.o_D, m.o_S, m.o_MM, m.o_vmapM, m.o_sM,
                lambda, m.o_x, m.o_y, m.o_z, elliptic.o_mapB, o_rL);
    o_rL.copyTo(rL);
    dump("rL", "f64", rL.ptr(), Ntot);
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
    elliptic.ogsMasked.Scatter(o_xL, o_x, 1, ogs::NoTrans);
    addBCKernel(m.Nelements, m.o_x, m.o_y, m.o_z, elliptic.o_mapB, o_xL);
    deviceMemory<dfloat> o_MxL = platform.malloc<dfloat>(xL);
    m.MassMatrixKernelSetup(1);
    m.MassMatrixApply(o_xL, o_MxL);
    dfloat norm2 = sqrt(platform.linAlg().innerProd((dlong)Ntot, o_xL, o_MxL, m.comm));
    int imeta[1] = {iter};
    dump("iterations", "i32", imeta, 1);
    double nmeta[1] = {norm2};
    dump("solnorm", "f64", nmeta, 1);
    printf("ITERATIONS = %d\n", iter);
    printf("Solution norm = %17.15lg\n", norm2);
  }
  Comm::Finalize();
  return 0;
}
