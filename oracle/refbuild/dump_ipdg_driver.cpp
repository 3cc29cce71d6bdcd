/* Golden-fixture dump driver for DISCRETIZATION = IPDG on hexahedra (test infrastructure; runs only in the build
 * container, single rank or P ranks under mpirun_stub.sh).
 *
 * Links against the UNMODIFIED reference (libelliptic.a + libs, built by build_ref.sh) and writes, per rank, what the
 * IPDG operator path reads and produces:
 *   mesh:      D, gllw, x,y,z, vgeo, sgeo, vmapM, vmapP, mapP, EToE, EToF, EToP, EToB (mesh), element halo lists
 *              (haloElementIds / internalElementIds, totalHaloPairs), the trace-halo global ids as
 *              mesh_t::HaloTraceSetup forms them
 *   elliptic:  EToB after the boundary-type translation, tau, lambda
 *   operator:  grad(q) as the gradient kernel + trace halo exchange leave it, Aq = elliptic.Operator(q)
 *   solve:     diagonal (BuildOperatorDiagonal), right-hand side of elliptic_t::Run, solution, iteration count
 *
 * usage: dump_ipdg_driver setup.rc outdir
 */
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include "elliptic.hpp"
#include "timer.hpp"

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

int main(int argc, char** argv) {
  Comm::Init(argc, argv);
  LIBP_ABORT("Usage: ./dump_ipdg_driver setupfile outdir", argc != 3);
  {
    comm_t comm(Comm::World().Dup());
    const int rank = comm.rank(), size = comm.size();
    g_out = std::string(argv[2]) + "/r" + std::to_string(rank);
    std::string cmd = "mkdir -p " + g_out;
    if (system(cmd.c_str())) return 1;
    platformSettings_t platformSettings(comm);
    meshSettings_t meshSettings(comm);
    ellipticSettings_t ellipticSettings(comm);
    ellipticAddRunSettings(ellipticSettings);
    ellipticSettings.parseFromFile(platformSettings, meshSettings, argv[1]);
    LIBP_ABORT("this driver is for DISCRETIZATION = IPDG", !ellipticSettings.compareSetting("DISCRETIZATION", "IPDG"));

    platform_t platform(platformSettings);
    mesh_t mesh(platform, meshSettings, comm);
    dfloat lambda = 0.0;
    ellipticSettings.getSetting("LAMBDA", lambda);
    memory<int> BCType(3);
    BCType[0] = 0; BCType[1] = 1; BCType[2] = 2;
    elliptic_t elliptic(platform, mesh, ellipticSettings, lambda, 3, BCType);
    mesh_t& m = elliptic.mesh;

    const size_t E = m.Nelements, Eh = m.totalHaloPairs, Np = m.Np;
    const size_t Ntot = Np * E, Nf = (size_t)m.Nfaces * m.Nfp * E;
    int meta[10] = {m.N, m.Nq, m.Np, (int)E, (int)Eh, (int)m.NinternalElements, (int)m.NhaloElements, m.Nvgeo, m.Nsgeo,
                    elliptic.allNeumann};
    dump("meta", "i32", meta, 10);
    int ranks[2] = {rank, size};
    dump("ranks", "i32", ranks, 2);
    double dmeta[3] = {lambda, elliptic.tau, elliptic.allNeumannScale};
    dump("dmeta", "f64", dmeta, 3);
    dump("D", "f64", m.D.ptr(), (size_t)m.Nq * m.Nq);
    dump("gllz", "f64", m.gllz.ptr(), (size_t)m.Nq);
    dump("gllw", "f64", m.gllw.ptr(), (size_t)m.Nq);
    dump("x", "f64", m.x.ptr(), Ntot);
    dump("y", "f64", m.y.ptr(), Ntot);
    dump("z", "f64", m.z.ptr(), Ntot);
    dump("vgeo", "f64", m.vgeo.ptr(), Ntot * m.Nvgeo);
    dump("sgeo", "f64", m.sgeo.ptr(), Nf * m.Nsgeo);
    dump("vmapM", "i32", m.vmapM.ptr(), Nf);
    dump("vmapP", "i32", m.vmapP.ptr(), Nf);
    dump("mapP", "i32", m.mapP.ptr(), Nf);
    dump("EToE", "i32", m.EToE.ptr(), E * m.Nfaces);
    dump("EToF", "i32", m.EToF.ptr(), E * m.Nfaces);
    dump("EToP", "i32", m.EToP.ptr(), E * m.Nfaces);
    dump("meshEToB", "i32", m.EToB.ptr(), E * m.Nfaces);
    dump("EToB", "i32", elliptic.EToB.ptr(), E * m.Nfaces);
    dump("internalElementIds", "i32", m.internalElementIds.ptr(), (size_t)m.NinternalElements);
    dump("haloElementIds", "i32", m.haloElementIds.ptr(), (size_t)m.NhaloElements);

    // the ids mesh_t::HaloTraceSetup(1) hands to halo_t::Setup (libs/mesh/meshHaloTraceSetup.cpp:38-83), recomputed
    // here with the mesh's own element halo so the harness can be checked against them
    {
      hlong localNelements = m.Nelements, globalOffset = m.Nelements;
      comm.Scan(localNelements, globalOffset);
      globalOffset -= localNelements;
      memory<hlong> gids((E + Eh) * Np, 0);
      for (size_t e = 0; e < E; ++e)
        for (size_t n = 0; n < Np; ++n) gids[e * Np + n] = (hlong)(e + globalOffset) * Np + n + 1;
      m.halo.Exchange(gids, (int)Np);
      for (size_t id = 0; id < Nf; ++id) {
        const dlong idP = m.vmapP[id];
        if ((size_t)(idP / (dlong)Np) >= E) gids[idP] = -std::abs(gids[idP]);
      }
      for (size_t n = Ntot; n < (E + Eh) * Np; ++n)
        if (gids[n] > 0) gids[n] = 0;
      dump("traceGlobalIds", "i64", gids.ptr(), (E + Eh) * Np);
      hlong off[1] = {globalOffset};
      dump("elementOffset", "i64", off, 1);
    }

    const dlong Ndofs = elliptic.Ndofs, Nhalo = elliptic.Nhalo;
    memory<dfloat> diagA(Ndofs);
    elliptic.BuildOperatorDiagonal(diagA);
    dump("diagA", "f64", diagA.ptr(), (size_t)Ndofs);

    // ---- operator apply on a seeded vector (seeded by GLOBAL node number so any partition sees the same field)
    hlong eoff = 0;
    {
      hlong l = m.Nelements, o = m.Nelements;
      comm.Scan(l, o);
      eoff = o - l;
    }
    memory<dfloat> q(Ndofs + Nhalo, 0.0), Aq(Ndofs + Nhalo, 0.0);
    for (dlong n = 0; n < Ndofs; ++n) q[n] = splitmix_uniform(1234, (uint64_t)(eoff * Np + n));
    deviceMemory<dfloat> o_q = platform.malloc<dfloat>(q);
    deviceMemory<dfloat> o_Aq = platform.malloc<dfloat>(Aq);
    elliptic.Operator(o_q, o_Aq);
    o_Aq.copyTo(Aq);
    dump("q", "f64", q.ptr(), (size_t)Ndofs);
    dump("Aq", "f64", Aq.ptr(), (size_t)Ndofs);
    {
      memory<dfloat> g((E + Eh) * Np * 4);
      elliptic.o_grad.copyTo(g);
      dump("grad", "f64", g.ptr(), (E + Eh) * Np * 4);
    }

    // ---- rhs as elliptic_t::Run builds it for IPDG (solvers/elliptic/src/ellipticRun.cpp:84-160) and the solve
    properties_t kernelInfo = m.props;
    std::string dataFileName;
    ellipticSettings.getSetting("DATA FILE", dataFileName);
    kernelInfo["includes"] += dataFileName;
    kernelInfo["includes"] += std::string(DELLIPTIC "/data/ellipticBoundary3D.h");
    kernelInfo["defines/" "p_Nmax"] = std::max(m.Np, m.Nfaces * m.Nfp);
    kernelInfo["defines/" "p_Nfields"] = 1;
    kernel_t forcingKernel = platform.buildKernel(DELLIPTIC "/okl/ellipticRhsHex3D.okl", "ellipticRhsHex3D", kernelInfo);
    kernel_t rhsBCKernel =
        platform.buildKernel(DELLIPTIC "/okl/ellipticRhsBCIpdgHex3D.okl", "ellipticRhsBCIpdgHex3D", kernelInfo);
    memory<dfloat> rL(Ntot, 0.0);
    deviceMemory<dfloat> o_r = platform.malloc<dfloat>(Ndofs + Nhalo);
    deviceMemory<dfloat> o_x = platform.malloc<dfloat>(q.length());
    {
      memory<dfloat> zeros(Ndofs + Nhalo, 0.0);
      o_x.copyFrom(zeros);
      o_r.copyFrom(zeros);
    }
    forcingKernel(m.Nelements, m.o_wJ, m.o_MM, m.o_x, m.o_y, m.o_z, lambda, o_r);
    rhsBCKernel(m.Nelements, m.o_vmapM, elliptic.tau, m.o_x, m.o_y, m.o_z, m.o_vgeo, m.o_sgeo, elliptic.o_EToB, m.o_D,
                m.o_LIFT, m.o_MM, o_r);
    memory<dfloat> r(Ndofs);
    o_r.copyTo(r, Ndofs);
    dump("r", "f64", r.ptr(), (size_t)Ndofs);
    linearSolver_t linearSolver;
    linearSolver.Setup<LinearSolver::pcg>(Ndofs, Nhalo, platform, ellipticSettings, comm);
    int iter = elliptic.Solve(linearSolver, o_x, o_r, 1.0e-8, 5000, 1);
    memory<dfloat> x(Ndofs);
    o_x.copyTo(x, Ndofs);
    dump("xsol", "f64", x.ptr(), (size_t)Ndofs);
    deviceMemory<dfloat> o_Mx = platform.malloc<dfloat>(Ntot);
    m.MassMatrixKernelSetup(1);
    m.MassMatrixApply(o_x, o_Mx);
    dfloat norm2 = sqrt(platform.linAlg().innerProd((dlong)Ntot, o_x, o_Mx, m.comm));
    int imeta[1] = {iter};
    dump("iterations", "i32", imeta, 1);
    double nmeta[1] = {norm2};
    dump("solnorm", "f64", nmeta, 1);
    if (rank == 0) {
      printf("ITERATIONS = %d\n", iter);
      printf("Solution norm = %17.15lg\n", norm2);
    }
  }
  Comm::Finalize();
  return 0;
}
