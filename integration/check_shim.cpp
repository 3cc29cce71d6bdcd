// Translation unit that instantiates the shim against the reference headers (compile check only).
#include "libp_b200_shim.hpp"

using namespace libp;

int instantiate(platform_t& platform, elliptic_t& elliptic, settings_t& settings, comm_t comm) {
  b200::runtime_t rt;
  rt.Setup(platform, comm, 0);
  b200::ogsB200_t ogs;
  ogs.Setup(elliptic.mesh.Nelements * elliptic.mesh.Np, elliptic.maskedGlobalIds, rt, ogs::Signed, true, false);
  b200::ellipticOperatorB200_t A;
  A.Setup(elliptic, ogs, rt);
  b200::JacobiPreconB200 M(elliptic, rt, ogs.NgatherGlobal);
  linearSolver_t solver;
  solver.Setup<b200::pcgB200>(ogs.Ngather, ogs.Nhalo, platform, settings, comm, rt);
  linearSolver_t nb;
  nb.Setup<b200::nbpcgB200>(ogs.Ngather, ogs.Nhalo, platform, settings, comm, rt);
  deviceMemory<dfloat> o_x = platform.malloc<dfloat>(ogs.Ngather + ogs.Nhalo);
  deviceMemory<dfloat> o_r = platform.malloc<dfloat>(ogs.Ngather + ogs.Nhalo);
  ogs.Gather(o_r, o_x, 1, ogs::Add, ogs::Trans);
  ogs.ExchangeStart(o_x, 1);
  ogs.ExchangeFinish(o_x, 1);
  A.UseTrilinearMap(platform, elliptic.mesh);
  return solver.Solve(A, M, o_x, o_r, 1e-8, 100, 0) + nb.Solve(A, M, o_x, o_r, 1e-8, 100, 0);
}

// DISCRETIZATION = IPDG: the operator handle from the reference's DG arrays, solved with the same pcg class
int instantiate_ipdg(platform_t& platform, elliptic_t& elliptic, settings_t& settings, comm_t comm, b200::runtime_t& rt) {
  b200::ellipticIpdgOperatorB200_t A;
  A.Setup(elliptic, rt);
  linearSolver_t solver;
  solver.Setup<b200::pcgB200>(elliptic.Ndofs, elliptic.Nhalo, platform, settings, comm, rt);
  deviceMemory<dfloat> o_x = platform.malloc<dfloat>(elliptic.Ndofs + elliptic.Nhalo);
  deviceMemory<dfloat> o_r = platform.malloc<dfloat>(elliptic.Ndofs + elliptic.Nhalo);
  return solver.Solve(A, elliptic.precon, o_x, o_r, 1e-8, 100, 0);
}

// what MultiGridPrecon::MultiGridPrecon does after parAlmond.AMGSetup(...): hand every level to the library
void instantiate_multigrid(b200::runtime_t& rt, parAlmond::parAlmond_t& parAlmond, int NpMGlevels,
                           parAlmond::exactSolver_t& exact, b200::ellipticOperatorB200_t* ops, elliptic_t& elliptic,
                           deviceMemory<dfloat>& o_r, deviceMemory<dfloat>& o_Mr) {
  b200::MultiGridPreconB200 M(rt);
  // the last amgLevel is the base level: its matrix belongs to the exact coarse solver (parAlmondAMGSetup.cpp:128-131)
  const int Nlevels = parAlmond.NumLevels() - 1;
  for (int l = 0; l < Nlevels; ++l) {
    if (l < NpMGlevels)
      M.AddLevel(b200::make_mglevel(parAlmond.GetLevel<MGLevel>(l), ops[l], ops[l + 1]));
    else
      M.AddLevel(b200::make_amglevel(rt, parAlmond.GetLevel<parAlmond::amgLevel>(l)));
  }
  M.Finish(b200::make_coarse(rt, exact), elliptic.allNeumann, elliptic.ogsMasked.NgatherGlobal);
  M.Operator(o_r, o_Mr);
}
