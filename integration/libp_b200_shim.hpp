// Reference-side binding of the B200 hot path: what a libParanumal maintainer adds next to
// solvers/elliptic/elliptic.hpp so that elliptic_t::Operator, the Jacobi / multigrid preconditioners and the
// PCG / NBPCG solvers run through include/libp_b200.h instead of the OCCA kernels.  This header is OUR code; it
// includes the reference's headers only to subclass its interfaces (operator_t, include/operator.hpp:35-40;
// linearSolverBase_t, include/linearSolver.hpp:78-97).  integration/check_shim.sh compiles it against the
// reference tree (syntax + type check; running it needs libParanumal in THREAD MODEL = CUDA on a B200).
//
// Conventions: libParanumal runs in CUDA mode, so deviceMemory<T>::ptr() (include/memory.hpp:322-327) is a CUDA
// device address; the shim creates the CUDA stream itself and hands it to OCCA (device.wrapStream,
// occa/include/occa/core/device.hpp:407), so both sides queue work on the same stream.
#pragma once
#include <cuda_runtime_api.h>

#include <cstring>
#include <memory>
#include <vector>

#include "elliptic.hpp"
#include "ellipticPrecon.hpp"
#include "parAlmond.hpp"
#include "parAlmond/parAlmondAMGLevel.hpp"
#include "parAlmond/parAlmondCoarseSolver.hpp"
#include "libp_b200.h"

namespace libp {
namespace b200 {

#define B200_CHECK(call)                                           \
  do {                                                             \
    if ((call) != LIBP_SUCCESS) LIBP_FORCE_ABORT(libp_last_error()); \
  } while (0)

// ---------------------------------------------------------------------------------- runtime: device, stream, comm
struct runtime_t {
  cudaStream_t stream = nullptr;
  libp_comm_t comm = nullptr;
  MPI_Comm mpicomm;
  platform_t* plat = nullptr;
  bool shared_stream = false;
  // Ordering between OCCA's stream and the library's: either both sides use ONE stream (share_stream: the library's
  // stream is handed to OCCA with device.wrapStream), or every call into the library drains OCCA's queue first and
  // the library's stream afterwards (default: robust with any OCCA version, costs two host waits per call).
  void enter() { if (!shared_stream) plat->finish(); }
  void leave() { if (!shared_stream) cudaStreamSynchronize(stream); }

  static int a2a(void* c, const void* s, void* r, size_t n) {
    return MPI_Alltoall(const_cast<void*>(s), (int)n, MPI_CHAR, r, (int)n, MPI_CHAR, *static_cast<MPI_Comm*>(c));
  }
  static int a2av(void* c, const void* s, const int64_t* sc, const int64_t* so, void* r, const int64_t* rc,
                  const int64_t* ro) {
    int size;
    MPI_Comm_size(*static_cast<MPI_Comm*>(c), &size);
    std::vector<int> isc(size), iso(size), irc(size), iro(size);
    for (int i = 0; i < size; ++i) { isc[i] = (int)sc[i]; iso[i] = (int)so[i]; irc[i] = (int)rc[i]; iro[i] = (int)ro[i]; }
    return MPI_Alltoallv(const_cast<void*>(s), isc.data(), iso.data(), MPI_CHAR, r, irc.data(), iro.data(), MPI_CHAR,
                         *static_cast<MPI_Comm*>(c));
  }
  static MPI_Op mpi_op(int op) { return op == LIBP_ADD ? MPI_SUM : op == LIBP_MAX ? MPI_MAX : MPI_MIN; }
  static int ar_i64(void* c, int64_t* b, int n, int op) {
    return MPI_Allreduce(MPI_IN_PLACE, b, n, MPI_LONG_LONG_INT, mpi_op(op), *static_cast<MPI_Comm*>(c));
  }
  static int ar_f64(void* c, double* b, int n, int op) {
    return MPI_Allreduce(MPI_IN_PLACE, b, n, MPI_DOUBLE, mpi_op(op), *static_cast<MPI_Comm*>(c));
  }

  // after platform_t selected its device (libs/core/platformDeviceConfig.cpp:33-179)
  void Setup(platform_t& platform, comm_t c, int device_id, bool share_stream = false) {
    B200_CHECK(libp_b200_init(device_id));
    plat = &platform;
    shared_stream = share_stream;
    cudaStreamCreate(&stream);
    if (share_stream) platform.setStream(platform.device.wrapStream(stream));   // OCCA and the library share one stream
    mpicomm = c.comm();
    libp_host_collectives_t host{&mpicomm, &a2a, &a2av, &ar_i64, &ar_f64};
    B200_CHECK(libp_comm_create(c.rank(), c.size(), &host, &comm));
    if (c.size() > 1) {
      char uid[128];
      std::memset(uid, 0, sizeof(uid));
      if (c.rank() == 0) B200_CHECK(libp_comm_nccl_unique_id(uid));
      MPI_Bcast(uid, 128, MPI_CHAR, 0, mpicomm);
      B200_CHECK(libp_comm_nccl_init(comm, uid));
      libp_comm_p2p_init(comm, 0);                             // NVLink peer window when every peer is mappable
    }
  }
};

template <typename T> constexpr libp_type_t type_of();
template <> constexpr libp_type_t type_of<float>() { return LIBP_FLOAT; }
template <> constexpr libp_type_t type_of<double>() { return LIBP_DOUBLE; }
template <> constexpr libp_type_t type_of<int>() { return LIBP_INT32; }
template <> constexpr libp_type_t type_of<long long int>() { return LIBP_INT64; }

// ---------------------------------------------------------------------------------- ogs::ogs_t + gathered halo_t
// (include/ogs.hpp:216-397).  The enums share their numeric values with ogs::Type/Op/Transpose/Kind (:178-200).
class ogsB200_t {
 public:
  libp_ogs_t h = nullptr;
  runtime_t* rt = nullptr;
  dlong N = 0, Ngather = 0, Nhalo = 0;
  hlong NgatherGlobal = 0;

  void Setup(const dlong N_, memory<hlong> ids, runtime_t& rt_, const ogs::Kind kind, const bool unique,
             const bool verbose) {
    rt = &rt_;
    static_assert(sizeof(hlong) == sizeof(libp_hlong), "global ids are 64-bit on both sides");
    B200_CHECK(libp_ogs_setup(N_, reinterpret_cast<libp_hlong*>(ids.ptr()), rt->comm, (int)kind, unique, verbose, &h));  // signed in place when unique
    libp_ogs_info_t i;
    B200_CHECK(libp_ogs_info(h, &i));
    N = i.N; Ngather = i.Ngather; Nhalo = i.Nhalo; NgatherGlobal = i.NgatherGlobal;
  }
  void SetupGlobalToLocalMapping(memory<dlong> GlobalToLocal) { B200_CHECK(libp_ogs_global_to_local(h, GlobalToLocal.ptr())); }

  template <typename T>
  void Gather(deviceMemory<T> o_gv, deviceMemory<T> o_v, const int k, const ogs::Op op, const ogs::Transpose trans) {
    B200_CHECK(libp_ogs_gather(h, o_gv.ptr(), o_v.ptr(), k, type_of<T>(), (int)op, (int)trans, rt->stream));
  }
  template <typename T>
  void GatherStart(deviceMemory<T> o_gv, deviceMemory<T> o_v, const int k, const ogs::Op op, const ogs::Transpose trans) {
    B200_CHECK(libp_ogs_gather_start(h, o_gv.ptr(), o_v.ptr(), k, type_of<T>(), (int)op, (int)trans, rt->stream));
  }
  template <typename T>
  void GatherFinish(deviceMemory<T> o_gv, deviceMemory<T> o_v, const int k, const ogs::Op op, const ogs::Transpose trans) {
    B200_CHECK(libp_ogs_gather_finish(h, o_gv.ptr(), o_v.ptr(), k, type_of<T>(), (int)op, (int)trans, rt->stream));
  }
  template <typename T>
  void Scatter(deviceMemory<T> o_v, deviceMemory<T> o_gv, const int k, const ogs::Transpose trans) {
    B200_CHECK(libp_ogs_scatter(h, o_v.ptr(), o_gv.ptr(), k, type_of<T>(), (int)trans, rt->stream));
  }
  template <typename T>
  void GatherScatter(deviceMemory<T> o_v, const int k, const ogs::Op op, const ogs::Transpose trans) {
    B200_CHECK(libp_ogs_gather_scatter(h, o_v.ptr(), k, type_of<T>(), (int)op, (int)trans, rt->stream));
  }
  // halo_t built by SetupFromGather (libs/ogs/ogsSetup.cpp:888-916)
  template <typename T> void ExchangeStart(deviceMemory<T> o_v, const int k) {
    B200_CHECK(libp_halo_exchange_start(h, o_v.ptr(), k, type_of<T>(), rt->stream));
  }
  template <typename T> void ExchangeFinish(deviceMemory<T> o_v, const int k) {
    B200_CHECK(libp_halo_exchange_finish(h, o_v.ptr(), k, type_of<T>(), rt->stream));
  }
  ~ogsB200_t() { if (h) libp_ogs_free(h); }
};

// ---------------------------------------------------------------------------------- elliptic_t::Operator (C0 hex)
// solvers/elliptic/src/ellipticOperator.cpp:31-106.  Built after elliptic_t::Setup / BoundarySetup produced
// maskedGlobalIds (unsigned copy) and the mesh arrays; `ogs` is the B200 twin of elliptic.ogsMasked.
class ellipticOperatorB200_t : public operator_t {
 public:
  libp_elliptic_t h = nullptr;
  runtime_t* rt = nullptr;
  deviceMemory<dlong> o_GlobalToLocal;

  void Setup(elliptic_t& e, ogsB200_t& ogs, runtime_t& rt_, int mode = 1) {
    rt = &rt_;
    mesh_t& mesh = e.mesh;
    memory<dlong> G2L(mesh.Nelements * mesh.Np);
    ogs.SetupGlobalToLocalMapping(G2L);
    o_GlobalToLocal = e.platform.malloc<dlong>(G2L);
    libp_elliptic_desc_t d{};
    d.Nq = mesh.Nq;
    d.Nelements = mesh.Nelements;
    d.NlocalGatherElements = mesh.NlocalGatherElements;
    d.NglobalGatherElements = mesh.NglobalGatherElements;
    // an element list can be empty (no rank-shared element on one rank): its deviceMemory is then uninitialised
    d.localGatherElementList = mesh.NlocalGatherElements ? mesh.o_localGatherElementList.ptr() : nullptr;
    d.globalGatherElementList = mesh.NglobalGatherElements ? mesh.o_globalGatherElementList.ptr() : nullptr;
    d.GlobalToLocal = o_GlobalToLocal.ptr();
    d.wJ = mesh.o_wJ.ptr();
    d.ggeo = mesh.o_ggeo.ptr();
    d.D = mesh.o_D.ptr();
    d.lambda = e.lambda;
    d.ogsMasked = ogs.h;
    d.mode = mode;  // 1: fused gather epilogue, 0: reference data flow (AqL + ogs gather)
    B200_CHECK(libp_elliptic_create(&d, &h));
  }
  // ELEMENT MAP = TRILINEAR (ellipticSetup.cpp:131-134; the kernel's o_EXYZ argument, ellipticAxHex3D.okl:440-447):
  // geometry from the element vertices instead of the stored factors.  This version of mesh_t keeps the vertices on
  // the host (EX, EY, EZ, include/mesh.hpp:74-76), so the [Nelements][3][8] device array is packed here.
  deviceMemory<dfloat> o_EXYZ;
  void UseTrilinearMap(platform_t& platform, mesh_t& mesh) {
    memory<dfloat> EXYZ(mesh.Nelements * 24);
    for (dlong el = 0; el < mesh.Nelements; el++)
      for (int v = 0; v < 8; v++) {
        EXYZ[el * 24 + v] = mesh.EX[el * mesh.Nverts + v];
        EXYZ[el * 24 + 8 + v] = mesh.EY[el * mesh.Nverts + v];
        EXYZ[el * 24 + 16 + v] = mesh.EZ[el * mesh.Nverts + v];
      }
    o_EXYZ = platform.malloc<dfloat>(EXYZ);
    B200_CHECK(libp_elliptic_set_trilinear(h, o_EXYZ.ptr(), mesh.gllz.ptr(), mesh.gllw.ptr()));
  }
  void Operator(deviceMemory<dfloat>& o_q, deviceMemory<dfloat>& o_Aq) override {
    rt->enter();
    B200_CHECK(libp_elliptic_operator(h, o_q.ptr(), o_Aq.ptr(), rt->stream));
    rt->leave();
  }
  ~ellipticOperatorB200_t() { if (h) libp_elliptic_free(h); }
};

// ---------------------------------------------------------------------------------- elliptic_t::Operator (IPDG hex)
// solvers/elliptic/src/ellipticOperator.cpp:108-160.  Every array is the reference's own device allocation; the trace
// halo is set up through the C ABI from the ids mesh_t::HaloTraceSetup builds (libs/mesh/meshHaloTraceSetup.cpp:38-83),
// repeated here because the reference keeps them local to that function.
class ellipticIpdgOperatorB200_t : public operator_t {
 public:
  libp_elliptic_t h = nullptr;
  runtime_t* rt = nullptr;
  ogsB200_t traceHalo;

  void Setup(elliptic_t& e, runtime_t& rt_) {
    rt = &rt_;
    mesh_t& mesh = e.mesh;
    const dlong Np = mesh.Np;
    if (mesh.totalHaloPairs > 0) {
      hlong localNelements = mesh.Nelements, globalOffset = mesh.Nelements;
      mesh.comm.Scan(localNelements, globalOffset);
      globalOffset -= localNelements;
      memory<hlong> ids((mesh.Nelements + mesh.totalHaloPairs) * Np, 0);
      for (dlong el = 0; el < mesh.Nelements; el++)
        for (int n = 0; n < Np; n++) ids[el * Np + n] = (el + globalOffset) * Np + n + 1;
      mesh.halo.Exchange(ids, Np);
      for (dlong id = 0; id < mesh.Nelements * mesh.Nfp * mesh.Nfaces; id++) {
        const dlong idP = mesh.vmapP[id];
        if (idP / Np >= mesh.Nelements && ids[idP] > 0) ids[idP] *= -1;  // flag the trace ids we need
      }
      for (dlong n = mesh.Nelements * Np; n < (mesh.Nelements + mesh.totalHaloPairs) * Np; n++)
        if (ids[n] > 0) ids[n] = 0;
      traceHalo.Setup((mesh.Nelements + mesh.totalHaloPairs) * Np, ids, rt_, ogs::Halo, false, false);
    }
    libp_ipdg_desc_t d{};
    d.Nq = mesh.Nq;
    d.Nelements = mesh.Nelements;
    d.NhaloElementsTotal = mesh.totalHaloPairs;
    d.NinternalElements = mesh.NinternalElements;
    d.NhaloElements = mesh.NhaloElements;
    d.internalElementIds = mesh.NinternalElements ? mesh.o_internalElementIds.ptr() : nullptr;
    d.haloElementIds = mesh.NhaloElements ? mesh.o_haloElementIds.ptr() : nullptr;
    d.vmapM = mesh.o_vmapM.ptr();
    d.vmapP = mesh.o_vmapP.ptr();
    d.vgeo = mesh.o_vgeo.ptr();
    d.sgeo = mesh.o_sgeo.ptr();
    d.EToB = e.o_EToB.ptr();
    d.D = mesh.o_D.ptr();
    d.lambda = e.lambda;
    d.tau = e.tau;
    d.traceHalo = traceHalo.h;
    B200_CHECK(libp_elliptic_create_ipdg(&d, &h));
  }
  void Operator(deviceMemory<dfloat>& o_q, deviceMemory<dfloat>& o_Aq) override {
    rt->enter();
    B200_CHECK(libp_elliptic_operator(h, o_q.ptr(), o_Aq.ptr(), rt->stream));
    rt->leave();
  }
  ~ellipticIpdgOperatorB200_t() { if (h) libp_elliptic_free(h); }
};

// ---------------------------------------------------------------------------------- JacobiPrecon
// solvers/elliptic/src/ellipticPreconJacobi.cpp:30-51 (the diagonal is still built by the reference on the host)
class JacobiPreconB200 : public operator_t {
 public:
  libp_precon_t h = nullptr;
  runtime_t* rt;
  JacobiPreconB200(elliptic_t& e, runtime_t& rt_, hlong NglobalDofs) : rt(&rt_) {
    memory<dfloat> diagA(e.Ndofs), invDiagA(e.Ndofs);
    e.BuildOperatorDiagonal(diagA);
    for (dlong n = 0; n < e.Ndofs; n++) invDiagA[n] = 1.0 / diagA[n];
    deviceMemory<dfloat> o_invDiagA = e.platform.malloc<dfloat>(invDiagA);
    B200_CHECK(libp_precon_jacobi_create(e.Ndofs, o_invDiagA.ptr(), e.allNeumann, NglobalDofs, rt->comm, &h));
  }
  void Operator(deviceMemory<dfloat>& o_r, deviceMemory<dfloat>& o_Mr) override {
    rt->enter();
    B200_CHECK(libp_precon_apply(h, o_r.ptr(), o_Mr.ptr(), rt->stream));
    rt->leave();
  }
  ~JacobiPreconB200() { if (h) libp_precon_free(h); }
};

// ---------------------------------------------------------------------------------- LinearSolver::pcg / nbpcg
// include/linearSolver.hpp:99-119.  Native handles take the fused device-resident iteration; any other operator_t
// (an un-replaced preconditioner, another solver's operator) goes through callbacks.
inline int op_trampoline(void* ctx, libp_dfloat* in, libp_dfloat* out, void* /*stream*/) {
  struct ctx_t { operator_t* op; platform_t* platform; dlong Ntotal; runtime_t* rt; };
  ctx_t* c = static_cast<ctx_t*>(ctx);
  try {
    deviceMemory<dfloat> o_in(c->platform->device.wrapMemory<dfloat>(in, c->Ntotal));
    deviceMemory<dfloat> o_out(c->platform->device.wrapMemory<dfloat>(out, c->Ntotal));
    if (!c->rt->shared_stream) cudaStreamSynchronize(c->rt->stream);  // the library's queue, then OCCA's
    c->op->Operator(o_in, o_out);
    if (!c->rt->shared_stream) c->platform->finish();
    return LIBP_SUCCESS;
  } catch (...) {
    return LIBP_ERROR;
  }
}

class pcgB200 : public LinearSolver::linearSolverBase_t {
  libp_pcg_t h = nullptr;
  runtime_t* rt;

 public:
  pcgB200(dlong _N, dlong _Nhalo, platform_t& _platform, settings_t& _settings, comm_t _comm, runtime_t& rt_)
      : linearSolverBase_t(_N, _Nhalo, _platform, _settings, _comm), rt(&rt_) {
    const int flexible = settings.compareSetting("LINEAR SOLVER", "FPCG");
    const int stopping = settings.compareSetting("LINEAR SOLVER STOPPING CRITERION", "ABS/REL-RHS-2NORM");
    B200_CHECK(libp_pcg_create(N, Nhalo, flexible, stopping, rt->comm, &h));
  }
  int Solve(operator_t& A, operator_t& M, deviceMemory<dfloat>& o_x, deviceMemory<dfloat>& o_r, const dfloat tol,
            const int MAXIT, const int verbose) override {
    int iters = 0;
    rt->enter();
    auto* a = dynamic_cast<ellipticOperatorB200_t*>(&A);
    auto* j = dynamic_cast<JacobiPreconB200*>(&M);
    if (a && j) {
      B200_CHECK(libp_pcg_solve(h, a->h, j->h, o_x.ptr(), o_r.ptr(), tol, MAXIT, verbose, rt->stream, &iters));
    } else {
      struct ctx_t { operator_t* op; platform_t* platform; dlong Ntotal; runtime_t* rt; };
      ctx_t ca{&A, &platform, N + Nhalo, rt}, cm{&M, &platform, N + Nhalo, rt};
      B200_CHECK(libp_pcg_solve_cb(h, &op_trampoline, &ca, &op_trampoline, &cm, o_x.ptr(), o_r.ptr(), tol, MAXIT, verbose,
                                   rt->stream, &iters));
    }
    rt->leave();
    return iters;
  }
  ~pcgB200() { if (h) libp_pcg_free(h); }
};

class nbpcgB200 : public LinearSolver::linearSolverBase_t {
  libp_nbpcg_t h = nullptr;
  runtime_t* rt;

 public:
  nbpcgB200(dlong _N, dlong _Nhalo, platform_t& _platform, settings_t& _settings, comm_t _comm, runtime_t& rt_)
      : linearSolverBase_t(_N, _Nhalo, _platform, _settings, _comm), rt(&rt_) {
    B200_CHECK(libp_nbpcg_create(N, Nhalo, rt->comm, &h));
  }
  int Solve(operator_t& A, operator_t& M, deviceMemory<dfloat>& o_x, deviceMemory<dfloat>& o_r, const dfloat tol,
            const int MAXIT, const int verbose) override {
    int iters = 0;
    struct ctx_t { operator_t* op; platform_t* platform; dlong Ntotal; runtime_t* rt; };
    ctx_t ca{&A, &platform, N + Nhalo, rt}, cm{&M, &platform, N + Nhalo, rt};
    B200_CHECK(libp_nbpcg_solve_cb(h, &op_trampoline, &ca, &op_trampoline, &cm, o_x.ptr(), o_r.ptr(), tol, MAXIT, verbose,
                                   rt->stream, &iters));
    return iters;
  }
  ~nbpcgB200() { if (h) libp_nbpcg_free(h); }
};

// ---------------------------------------------------------------------------------- MultiGridPrecon / parAlmond
// The reference keeps its whole setup (degree ladder, SetupSmoother, BuildOperatorMatrixContinuous, AMGSetup) and hands
// the products over level by level.  These helpers are what MultiGridPrecon::MultiGridPrecon gains after
// parAlmond.AMGSetup(...) (solvers/elliptic/src/ellipticPreconMultiGrid.cpp:150): it walks
// parAlmond.GetLevel<MGLevel>(l) / GetLevel<parAlmond::amgLevel>(l) (include/parAlmond.hpp:205-212) and the exact solver.
static_assert(sizeof(pfloat) == sizeof(libp_dfloat), "CSR values are FP64 on both sides");

// parCSR (include/parAlmond/parAlmondparCSR.hpp:35-100) -> libp_parcsr_create; after parCSR::haloSetup the non-local
// part of colMap holds ascending global column ids (parAlmondparCSR.cpp:307-317)
inline libp_csr_t make_parcsr(runtime_t& rt, parAlmond::parCSR& M) {
  libp_parcsr_desc_t d{};
  d.Nrows = M.Nrows;
  d.NlocalCols = M.NlocalCols;
  d.diag_nnz = M.diag.nnz;
  d.diag_rowStarts = M.diag.rowStarts.ptr();
  d.diag_cols = M.diag.cols.ptr();
  d.diag_vals = M.diag.vals.ptr();
  d.offd_nnz = M.offd.nnz;
  d.offd_nzRows = M.offd.nzRows;
  d.offd_rows = M.offd.rows.ptr();
  d.offd_mRowStarts = M.offd.mRowStarts.ptr();
  d.offd_cols = M.offd.cols.ptr();
  d.offd_vals = M.offd.vals.ptr();
  d.Noffdcols = M.Ncols - M.NlocalCols;
  d.offd_colIds = reinterpret_cast<const libp_hlong*>(M.colMap.ptr()) + M.NlocalCols;
  d.globalColStarts = reinterpret_cast<const libp_hlong*>(M.globalColStarts.ptr());
  libp_csr_t h = nullptr;
  B200_CHECK(libp_parcsr_create(rt.comm, &d, &h));
  return h;
}

// parAlmond::amgLevel (include/parAlmond/parAlmondAMGLevel.hpp:37-64)
inline libp_amglevel_t make_amglevel(runtime_t& rt, parAlmond::amgLevel& L) {
  libp_csr_t A = make_parcsr(rt, L.A);
  libp_csr_t P = make_parcsr(rt, L.P);
  libp_csr_t R = make_parcsr(rt, L.R);
  libp_amglevel_t h = nullptr;
  B200_CHECK(libp_amglevel_create(A, P, R, L.A.diagInv.ptr(), L.stype == parAlmond::CHEBYSHEV ? 1 : 0, L.lambda, L.lambda0,
                                  L.lambda1, L.ChebyshevIterations, &h));
  return h;
}

// MGLevel (solvers/elliptic/ellipticPrecon.hpp:125-185): fine / coarse are the B200 operators of L.elliptic / L.ellipticC
inline libp_mglevel_t make_mglevel(MGLevel& L, ellipticOperatorB200_t& fine, ellipticOperatorB200_t& coarse) {
  libp_mglevel_desc_t m{};
  m.fine = fine.h;
  m.coarse = coarse.h;
  m.NqF = L.mesh.Nq;
  m.NqC = L.meshC.Nq;
  m.P = L.o_P.ptr();
  m.invDiagA = L.o_invDiagA.ptr();
  m.weightG = L.elliptic.o_weightG.ptr();
  m.smoother = (int)L.stype;  // JACOBI = 1, CHEBYSHEV = 2 on both sides
  m.lambda0 = L.lambda0;
  m.lambda1 = L.lambda1;
  m.ChebyshevIterations = L.ChebyshevIterations;
  libp_mglevel_t h = nullptr;
  B200_CHECK(libp_mglevel_create(&m, &h));
  return h;
}

// parAlmond::exactSolver_t (include/parAlmond/parAlmondCoarseSolver.hpp:70-105) after setup()
inline libp_coarse_t make_coarse(runtime_t& rt, parAlmond::exactSolver_t& E) {
  libp_coarse_t h = nullptr;
  B200_CHECK(libp_coarse_exact_create_par(rt.comm, E.N, reinterpret_cast<const libp_hlong*>(E.A.globalRowStarts.ptr()),
                                          E.diagInvAT.ptr(), E.offdInvAT.ptr(), &h));
  return h;
}

// MultiGridPrecon::Operator (ellipticPreconMultiGrid.cpp:29-37): one V-cycle (+ ZeroMean when allNeumann)
class MultiGridPreconB200 : public operator_t {
 public:
  libp_multigrid_t mg = nullptr;
  libp_precon_t h = nullptr;
  runtime_t* rt;
  explicit MultiGridPreconB200(runtime_t& rt_) : rt(&rt_) { B200_CHECK(libp_multigrid_create(rt->comm, &mg)); }
  void AddLevel(libp_mglevel_t l) { B200_CHECK(libp_multigrid_add_mglevel(mg, l)); }
  void AddLevel(libp_amglevel_t l) { B200_CHECK(libp_multigrid_add_amglevel(mg, l)); }
  void Finish(libp_coarse_t coarse, int allNeumann, hlong NglobalDofs) {
    B200_CHECK(libp_multigrid_set_coarse(mg, coarse));
    B200_CHECK(libp_precon_multigrid_create(mg, allNeumann, NglobalDofs, rt->comm, &h));
  }
  void Operator(deviceMemory<dfloat>& o_r, deviceMemory<dfloat>& o_Mr) override {
    B200_CHECK(libp_precon_apply(h, o_r.ptr(), o_Mr.ptr(), rt->stream));
  }
};

}  // namespace b200
}  // namespace libp
