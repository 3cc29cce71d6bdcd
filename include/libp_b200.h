/*
 * libp_b200.h - C ABI of the B200-native elliptic hot path for libParanumal.
 *
 * Every entry point replaces one reference interface on the path
 *   elliptic_t::Operator -> ogs_t / halo_t -> linAlg_t -> LinearSolver::pcg -> precon_t
 * (file:line citations are relative to the libParanumal 0.5.0 tree).  The reference reaches
 * these through C++ virtuals and occa::kernel functors; the shim a maintainer adds on the
 * reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are CUDA device addresses
 *     (deviceMemory<T>::ptr(), include/memory.hpp:322-327)
 *   - dlong = int32, hlong = int64, dfloat = pfloat = double   (include/types.h:31-64)
 *   - every function returns LIBP_SUCCESS (0) or LIBP_ERROR (-1)  (include/utils.hpp:48-49);
 *     libp_last_error() returns the thread-local message; the C++ shim turns non-zero into
 *     LIBP_FORCE_ABORT so the reference's throw-libp::exception behaviour is kept
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - one caller thread per GPU, Start/Finish pairs in order on one handle (the reference's
 *     own threading contract)
 */
#ifndef LIBP_B200_H
#define LIBP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same values as the reference's own macros (include/utils.hpp:48-51); guarded so both headers can be included */
#ifndef LIBP_SUCCESS
#define LIBP_SUCCESS 0
#endif
#ifndef LIBP_ERROR
#define LIBP_ERROR (-1)
#endif

typedef int32_t libp_dlong;
typedef int64_t libp_hlong;
typedef double libp_dfloat;

/* ogs enums: include/ogs.hpp:178-199 (same numeric values) */
typedef enum { LIBP_FLOAT = 0, LIBP_DOUBLE = 1, LIBP_INT32 = 2, LIBP_INT64 = 3 } libp_type_t;
typedef enum { LIBP_ADD = 0, LIBP_MUL = 1, LIBP_MAX = 2, LIBP_MIN = 3 } libp_op_t;
typedef enum { LIBP_SYM = 0, LIBP_NOTRANS = 1, LIBP_TRANS = 2 } libp_transpose_t;
typedef enum { LIBP_UNSIGNED = 0, LIBP_SIGNED = 1, LIBP_HALO = 2 } libp_kind_t;

typedef struct libp_comm_s* libp_comm_t;
typedef struct libp_ogs_s* libp_ogs_t;
typedef struct libp_elliptic_s* libp_elliptic_t;
typedef struct libp_pcg_s* libp_pcg_t;
typedef struct libp_nbpcg_s* libp_nbpcg_t;
typedef struct libp_nbfpcg_s* libp_nbfpcg_t;
typedef struct libp_ig_s* libp_ig_t;
typedef struct libp_precon_s* libp_precon_t;
typedef struct libp_mglevel_s* libp_mglevel_t;
typedef struct libp_csr_s* libp_csr_t;
typedef struct libp_amglevel_s* libp_amglevel_t;
typedef struct libp_coarse_s* libp_coarse_t;
typedef struct libp_multigrid_s* libp_multigrid_t;

/* ------------------------------------------------------------------ runtime / errors */
const char* libp_last_error(void);
/* Library build info: "sm_100a;<git describe or date>" */
const char* libp_b200_version(void);
/* Select the CUDA device of this process (platform_t device selection,
 * libs/core/platformDeviceConfig.cpp:33-179: device_id = local rank). */
int libp_b200_init(int device_id);
/* Blocks until `stream` has drained (platform_t::finish, include/platform.hpp). */
int libp_b200_finish(void* stream);

/* ------------------------------------------------------------------ communicator
 * Replaces comm_t (include/comm.hpp:95-...) on this path.
 * Host-side collectives are needed only at setup (ogsBase_t::Setup uses Alltoall / Alltoallv /
 * Allreduce / Scan); they are supplied by the embedding program as callbacks so the library
 * has no MPI dependency: the reference shim fills them with MPI_*, the Python harness with
 * torch.distributed (gloo).  With size==1 they may be NULL.
 * Data-path collectives (halo exchange, dot products) run on NCCL over NVLink.           */
typedef struct {
  void* ctx;
  /* every rank sends `bytes_per_rank` bytes to each rank (MPI_Alltoall of bytes) */
  int (*alltoall)(void* ctx, const void* send, void* recv, size_t bytes_per_rank);
  /* MPI_Alltoallv of bytes: counts/offsets are in bytes, arrays of length size */
  int (*alltoallv)(void* ctx, const void* send, const int64_t* send_counts, const int64_t* send_offsets,
                   void* recv, const int64_t* recv_counts, const int64_t* recv_offsets);
  /* in-place allreduce of n int64 values; op: LIBP_ADD / LIBP_MAX / LIBP_MIN */
  int (*allreduce_i64)(void* ctx, int64_t* inout, int n, int op);
  /* in-place allreduce of n doubles (host); used by setup-time eigenvalue estimates */
  int (*allreduce_f64)(void* ctx, double* inout, int n, int op);
} libp_host_collectives_t;

int libp_comm_create(int rank, int size, const libp_host_collectives_t* host, libp_comm_t* comm);
int libp_comm_free(libp_comm_t comm);
int libp_comm_rank(libp_comm_t comm, int* rank, int* size);
/* NCCL bootstrap: rank 0 calls libp_comm_nccl_unique_id (128 bytes), the embedding program
 * broadcasts it (MPI_Bcast / torch.distributed), every rank calls libp_comm_nccl_init. */
int libp_comm_nccl_unique_id(void* uid128);
int libp_comm_nccl_init(libp_comm_t comm, const void* uid128);

/* NVLink peer window (replaces the reference's GPU-aware-MPI exchange, libs/ogs/ogsPairwise.cpp:116-183, and
 * the 8-byte MPI_Allreduce of linAlg.cpp:149-187 on the PCG path): every rank allocates `window_bytes`
 * (0 = LIBP_P2P_WINDOW_MB or 256 MB) of device memory and maps every peer's window through CUDA IPC, using the
 * host all-to-all callback to trade handles.  Afterwards halo packing kernels store straight into the
 * neighbours' receive buffers over NVLink and reduction kernels all-reduce their scalars through mailboxes,
 * with release/acquire flags instead of NCCL launches.  Collective; all ranks of one node.  If it is not
 * called (or fails on any rank: LIBP_ERROR on every rank) the NCCL path is used.                          */
int libp_comm_p2p_init(libp_comm_t comm, size_t window_bytes);
int libp_comm_p2p_enabled(libp_comm_t comm, int* enabled);
/* Every in-kernel wait on a peer's flag is bounded (LIBP_P2P_TIMEOUT_MS, default 30000): when one times out - a peer
 * threw between collectives, or ranks diverged in the exchanges they issue - the kernels stop waiting and raise an
 * error word instead of hanging the GPU.  libp_comm_p2p_status reads it (non-zero: results computed through this
 * communicator since are invalid; libp_pcg_solve checks it itself); libp_comm_p2p_reset clears it. */
int libp_comm_p2p_status(libp_comm_t comm, int* timed_out);
int libp_comm_p2p_reset(libp_comm_t comm);

/* ------------------------------------------------------------------ ogs
 * ogs::ogs_t::Setup (include/ogs.hpp:216-226; libs/ogs/ogsSetup.cpp:43-190).
 * `ids` (host, length N): 0 = ignored, sign = flag.  When `unique` the array is rewritten
 * with the owner copy positive (ogsSetup.cpp:138-141), using glibc rand() exactly like the
 * reference (ogsSetup.cpp:268-274) so maps are bit-identical in the same call sequence.   */
int libp_ogs_setup(libp_dlong N, libp_hlong* ids, libp_comm_t comm, int kind, int unique, int verbose,
                   libp_ogs_t* ogs);
/* Test hook: ogsBase_t::Setup depends on the tie order of libstdc++'s std::sort (ogsSetup.cpp:245-275); the library
 * runs the same introsort as parallel tasks.  Compares the two on n records with nkeys distinct keys. */
int libp_ogs_sort_selftest(libp_dlong n, libp_dlong nkeys, unsigned int seed, int* same);
/* Self-test of the bulk rand() draw of the setup (one draw per id group, ogsSetup.cpp:262): srand(seed), n draws taken
 * in bulk from glibc's generator state followed by 100 rand() calls must equal n + 100 rand() calls.  Reseeds rand(). */
int libp_ogs_rand_selftest(unsigned int seed, libp_dlong n, int* same);
int libp_ogs_free(libp_ogs_t ogs);

typedef struct {
  libp_dlong N, Ngather, NlocalT, NlocalP, NhaloT, NhaloP, Nhalo;
  libp_hlong NgatherGlobal;
  int gather_defined;
  /* pairwise exchange shape (libs/ogs/ogsPairwise.cpp:194-415) */
  int NranksSendN, NranksSendT, NranksRecvN, NranksRecvT;
  libp_dlong NsendN, NsendT, NrecvN, NrecvT;
} libp_ogs_info_t;
int libp_ogs_info(libp_ogs_t ogs, libp_ogs_info_t* info);

/* Host copies of the maps for the bit-exact check.  which: 0 = gatherLocal, 1 = gatherHalo,
 * 2 = exchange postmpi (ogsPairwise.cpp:283-336).  Pointers stay owned by the handle.
 * nrows = NrowsT of that operator; rowStarts* have nrows+1 entries.                        */
int libp_ogs_maps(libp_ogs_t ogs, int which, libp_dlong* NrowsN, libp_dlong* NrowsT,
                  const libp_dlong** rowStartsN, const libp_dlong** rowStartsT,
                  const libp_dlong** colIdsN, const libp_dlong** colIdsT);
/* Pairwise send lists / per-neighbour counts (host).  trans selects N (NOTRANS) or T lists. */
int libp_ogs_exchange_lists(libp_ogs_t ogs, int trans, libp_dlong* Nsend, const libp_dlong** sendIds,
                            int* NranksSend, const int** sendRanks, const int** sendCounts, const int** sendOffsets,
                            int* NranksRecv, const int** recvRanks, const int** recvCounts, const int** recvOffsets);
/* ogs_t::SetupGlobalToLocalMapping (libs/ogs/ogsSetup.cpp:862-886): host array of N dlong. */
int libp_ogs_global_to_local(libp_ogs_t ogs, libp_dlong* GlobalToLocal);

/* Device apply (include/ogs.hpp:228-344; libs/ogs/ogs.cpp:39-488).  gv is the gathered vector
 * [NlocalT | NhaloP | ...], v the local vector of N*k entries, k interleaved values per node. */
int libp_ogs_gather(libp_ogs_t ogs, void* gv, const void* v, int k, int type, int op, int trans, void* stream);
int libp_ogs_gather_start(libp_ogs_t ogs, void* gv, const void* v, int k, int type, int op, int trans, void* stream);
int libp_ogs_gather_finish(libp_ogs_t ogs, void* gv, const void* v, int k, int type, int op, int trans, void* stream);
int libp_ogs_scatter(libp_ogs_t ogs, void* v, const void* gv, int k, int type, int trans, void* stream);
int libp_ogs_scatter_start(libp_ogs_t ogs, void* v, const void* gv, int k, int type, int trans, void* stream);
int libp_ogs_scatter_finish(libp_ogs_t ogs, void* v, const void* gv, int k, int type, int trans, void* stream);
int libp_ogs_gather_scatter(libp_ogs_t ogs, void* v, int k, int type, int op, int trans, void* stream);
int libp_ogs_gather_scatter_start(libp_ogs_t ogs, void* v, int k, int type, int op, int trans, void* stream);
int libp_ogs_gather_scatter_finish(libp_ogs_t ogs, void* v, int k, int type, int op, int trans, void* stream);

/* halo_t built by SetupFromGather (libs/ogs/ogsSetup.cpp:888-916) on the same handle:
 * ExchangeStart/Finish (libs/ogs/ogsHalo.cpp:46-143) send v[NlocalT : NlocalT+NhaloP] and fill
 * v[NlocalT+NhaloP : NlocalT+NhaloT].                                                     */
int libp_halo_exchange_start(libp_ogs_t ogs, void* v, int k, int type, void* stream);
int libp_halo_exchange_finish(libp_ogs_t ogs, void* v, int k, int type, void* stream);
int libp_halo_exchange(libp_ogs_t ogs, void* v, int k, int type, void* stream);

/* ------------------------------------------------------------------ ellipticAx Hex3D
 * ellipticPartialAxHex3D (solvers/elliptic/okl/ellipticAxHex3D.okl:156-295), same argument
 * meaning and order (S, MM are unused by the reference kernel and dropped):
 *   AqL[e,:] = A_e * q[GlobalToLocal[e,:]]   for e in elementList[0:Nelements]
 * elementList==NULL means e = 0..Nelements-1; GlobalToLocal==NULL gives the element-local twin
 * ellipticAxHex3D (:28-152).  Nq = N+1 in [2, 9].                                         */
int libp_ax_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* elementList, const libp_dlong* GlobalToLocal,
                  const libp_dfloat* wJ, const libp_dfloat* ggeo, const libp_dfloat* D, libp_dfloat lambda,
                  const libp_dfloat* q, libp_dfloat* AqL, void* stream);
/* Fused variant: the gather (ogs Add, Trans) is folded into the epilogue, Aq[GlobalToLocal]
 * += A_e q.  Aq must be zeroed by the caller (libp_elliptic_operator does).               */
int libp_ax_hex3d_gather(int Nq, libp_dlong Nelements, const libp_dlong* elementList, const libp_dlong* GlobalToLocal,
                         const libp_dfloat* wJ, const libp_dfloat* ggeo, const libp_dfloat* D, libp_dfloat lambda,
                         const libp_dfloat* q, libp_dfloat* Aq, void* stream);

/* Optional promise that the Nq x Nq device array D stays immutable until unregistered (mesh.o_D does):
 * the library then classifies it once (a GLL derivative matrix is centro-antisymmetric, which enables the
 * even-odd form of the 1-D contractions) and stops reloading it into constant memory on every call.
 * libp_elliptic_create does the same internally.  Unregistered pointers always take the general path. */
int libp_ax_hex3d_register_D(int Nq, const libp_dfloat* D);
int libp_ax_hex3d_unregister_D(const libp_dfloat* D);

/* Tuning knob (not a reference interface): 0 = one-thread-per-column "pencil" kernel that mirrors the
 * reference's thread layout, 1 = transposed-pencil kernel (default; all contractions in registers). */
int libp_ax_hex3d_set_variant(int variant);
/* Development knob, only in builds with -DLIBP_AX_TUNE_GRID (returns LIBP_ERROR otherwise): selects the
 * prefetch depth / L2 hint / resident-block instantiation of the fused N=7 kernel. */
int libp_ax_hex3d_tune(int prefetch_slabs, int l2_hints, int min_blocks);

/* ------------------------------------------------------------------ elliptic_t::Operator (C0)
 * solvers/elliptic/src/ellipticOperator.cpp:31-106.  All pointers are device pointers that
 * stay owned by the caller (mesh.o_*, o_GlobalToLocal).                                    */
typedef struct {
  int Nq;
  libp_dlong Nelements;
  libp_dlong NlocalGatherElements, NglobalGatherElements;
  const libp_dlong* localGatherElementList;  /* device */
  const libp_dlong* globalGatherElementList; /* device */
  const libp_dlong* GlobalToLocal;           /* device, Nelements*Np */
  const libp_dfloat* wJ;                     /* device, Nelements*Np */
  const libp_dfloat* ggeo;                   /* device, Nelements*6*Np */
  const libp_dfloat* D;                      /* device, Nq*Nq row-major D[i*Nq+m] */
  libp_dfloat lambda;
  libp_ogs_t ogsMasked;
  /* 0: reference data flow  (partial Ax -> AqL -> ogs gather), bit-reproducible summation order
   * 1: fused Ax+gather epilogue (FP64 atomics, no AqL round trip)                           */
  int mode;
} libp_elliptic_desc_t;
int libp_elliptic_create(const libp_elliptic_desc_t* desc, libp_elliptic_t* op);
int libp_elliptic_free(libp_elliptic_t op);

/* ---- DISCRETIZATION = IPDG on hexahedra -------------------------------------------------------------
 * elliptic_t::Operator, IPDG branch (solvers/elliptic/src/ellipticOperator.cpp:108-160): ellipticPartialGradientHex3D
 * (okl/ellipticGradientHex3D.okl:97-169), traceHalo.Exchange of the (dq/dx, dq/dy, dq/dz, q) array with 4 entries per
 * node, ellipticPartialAxIpdgHex3D on the internal and then the halo elements (okl/ellipticAxIpdgHex3D.okl:359-642).
 * All arrays are the reference's (device pointers): vgeo [Nelements][12][Np] (mesh_t::vgeo, RX..IJW), sgeo
 * [Nelements][6*Nq^2][8] (NX,NY,NZ,SJ,IJ,IH,WSJ,WIJ), vmapM / vmapP [Nelements][6*Nq^2] node numbers in the
 * (Nelements + NhaloElementsTotal) * Np node array, EToB [Nelements][6] boundary TYPE after the BCType translation of
 * ellipticBoundarySetup.cpp:37-48 (1 Dirichlet, 2 Neumann, <= 0 none), tau = elliptic_t::tau (ellipticSetup.cpp:66-75).
 * traceHalo = handle of the halo set up from mesh_t::HaloTraceSetup's ids (NULL on one rank); internalElementIds /
 * haloElementIds = mesh_t::internalElementIds / haloElementIds (both NULL: every element is internal).
 * The handle is a libp_elliptic_t: libp_elliptic_operator, libp_pcg_solve, ... take it; vectors are
 * [Nelements*Np | NhaloElementsTotal*Np] as elliptic_t::Ndofs / Nhalo (ellipticSetup.cpp:162-163). */
typedef struct {
  int Nq;
  libp_dlong Nelements, NhaloElementsTotal;           /* mesh.Nelements, mesh.totalHaloPairs */
  libp_dlong NinternalElements, NhaloElements;        /* lengths of the two element lists */
  const libp_dlong* internalElementIds;
  const libp_dlong* haloElementIds;
  const libp_dlong* vmapM;
  const libp_dlong* vmapP;
  const libp_dfloat* vgeo;
  const libp_dfloat* sgeo;
  const int* EToB;
  const libp_dfloat* D;
  libp_dfloat lambda, tau;
  libp_ogs_t traceHalo;
} libp_ipdg_desc_t;
int libp_elliptic_create_ipdg(const libp_ipdg_desc_t* desc, libp_elliptic_t* op);
/* elliptic_t::o_grad of an IPDG handle (device, [(Nelements + NhaloElementsTotal)*Np][4]) after the last apply */
int libp_elliptic_ipdg_gradient(libp_elliptic_t op, const libp_dfloat** grad);
/* BuildOperatorDiagonalIpdgHex3D (solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:868-996): A [Nelements*Np];
 * EToB here is the translated boundary TYPE as above */
int libp_elliptic_build_diagonal_ipdg_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* vgeo,
                                            const libp_dfloat* sgeo, const int* EToB, const libp_dfloat* D,
                                            libp_dfloat lambda, libp_dfloat tau, libp_dfloat* A, void* stream);
/* ellipticRhsBCIpdgHex3D (solvers/elliptic/okl/ellipticRhsBCIpdgHex3D.okl, called from ellipticRun.cpp:150-163):
 * rhs -= boundary functional of the IPDG form.  uD / gN: nodal Dirichlet values / Neumann data n.(uxB,uyB,uzB) per face
 * node [Nelements][6*Nq^2] (the data-file macros the reference inlines at JIT time); either may be NULL (= 0). */
int libp_elliptic_rhs_bc_ipdg_hex3d(int Nq, libp_dlong Nelements, libp_dfloat tau, const libp_dfloat* vgeo,
                                    const libp_dfloat* sgeo, const int* EToB, const libp_dfloat* D, const libp_dfloat* uD,
                                    const libp_dfloat* gN, libp_dfloat* rhs, void* stream);
/* mesh_t::SurfaceGeometricFactorsHex3D (libs/mesh/meshSurfaceGeometricFactorsHex3D.cpp:31-195) in two steps, because
 * the penalty length IHID = max(sJ/J of both sides) needs the neighbours' values across ranks in between:
 *   1. sgeo (all entries but IHID) and h[Nelements*6*Nq^2] = sJ/J from the physical nodes x, y, z;
 *   2. IHID from h (own + halo part, exchanged by the caller with mesh_t::halo) through mapP (face-node index of the
 *      neighbour in h, < 0: none). */
int libp_mesh_surface_geometric_factors_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* x, const libp_dfloat* y,
                                              const libp_dfloat* z, const libp_dfloat* D, const libp_dfloat* gllw,
                                              libp_dfloat* sgeo, libp_dfloat* h, void* stream);
int libp_mesh_surface_hinv_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* mapP, const libp_dfloat* h,
                                 libp_dfloat* sgeo, void* stream);
/* Fused mode only (no reference counterpart; tuning of this implementation): the element lists are run in
 * pieces of `chunkElements` elements and the accumulator o_Aq is zero-filled piece by piece, just ahead of
 * the reductions that land in it, so the zero lines are still in L2 (saves one DRAM read + one write of the
 * result vector per apply).  0 = one zero-fill and three launches, the reference split.  The default applies
 * to handles created afterwards. */
/* Fused mode, GLL derivative matrix (tuning of this implementation): the Ax kernel zero-fills o_Aq itself, a few
 * thousand elements ahead of its own reductions, behind a ticket / per-group counter protocol (csrc/ax_hex3d.cu),
 * so the accumulator costs one DRAM pass instead of three.  libp_elliptic_zero_ahead_errors reports protocol
 * time-outs (0 in every run so far; a non-zero value means results of that handle are not to be trusted). */
int libp_elliptic_set_zero_ahead(libp_elliptic_t op, int on);
int libp_elliptic_set_default_zero_ahead(int on);
int libp_elliptic_zero_ahead_errors(libp_elliptic_t op, int* errors);
/* Fused mode, GLL derivative matrix (tuning of this implementation, no reference counterpart): the element-chain
 * kernel (csrc/ax_chain.cu).  One CTA walks `chainElements` consecutive elements of a launch segment of
 * elliptic_t::Operator (ellipticOperator.cpp:40-104); geometric factors arrive in shared memory by bulk-async copies
 * (TMA), rows of o_Aq touched by a single chain are written with plain stores and never zero-filled, connectivity is
 * run-length compressed.  0 = off (ax_hex3d_t_kernel).  `stages` = geometric-factor blocks in flight per CTA (2 or 3).
 * Needs o_Aq 32-byte aligned and ggeo / wJ 16-byte aligned; other pointers take the previous kernel.
 * libp_elliptic_chain_stats: {chainElements, sectors, sectors still zero-filled, positions, uncompressed elements,
 * stages} of the plan derived from GlobalToLocal for this handle (zeros when the chain kernel is not in use). */
int libp_elliptic_set_chain(libp_elliptic_t op, int chainElements, int stages);
int libp_elliptic_set_default_chain(int chainElements, int stages);
int libp_elliptic_chain_stats(libp_elliptic_t op, libp_dfloat* Aq, long long* stats, void* stream);
/* Host-side audit (no GPU needed) of the shared-memory geometry compiled into the element-chain kernel of order Nq:
 * ok = every pencil of layouts A / B and every column of layout C is owned by exactly one lane and all offsets are in
 * bounds; wavefronts[6] = {actual, ideal} shared wavefronts of one element step for layouts C, A, B under the bank model
 * of DESIGN.md 4.1c (64-bit accesses per half-warp, 128-bit per quarter-warp). */
int libp_ax_chain_layout_selftest(int Nq, int* ok, int* wavefronts);
/* ELEMENT MAP = TRILINEAR (ellipticSetup.cpp:131-134 selects ellipticPartialAxTrilinearHex3D,
 * solvers/elliptic/okl/ellipticAxHex3D.okl:440-627): the operator recomputes the geometric factors from the element
 * vertices EXYZ (device, [Nelements][3][8], reference vertex order) and the GLL nodes / weights (host, Nq entries)
 * instead of streaming ggeo / wJ: 16 B per DOF + 0.4 B per node of geometry.  Fused mode, GLL D, chain kernel.
 * Parallelepiped elements (every box element) take a constant-Jacobian shortcut.  EXYZ = NULL switches back. */
int libp_elliptic_set_trilinear(libp_elliptic_t op, const libp_dfloat* EXYZ, const libp_dfloat* gllz,
                                const libp_dfloat* gllw);
int libp_elliptic_set_chunk(libp_elliptic_t op, libp_dlong chunkElements);
int libp_elliptic_set_default_chunk(libp_dlong chunkElements);
/* o_q and o_Aq are gathered vectors of Ndofs+Nhalo entries; the Nhalo tail of o_q is
 * overwritten by the halo exchange exactly like the reference.         */
int libp_elliptic_operator(libp_elliptic_t op, libp_dfloat* q, libp_dfloat* Aq, void* stream);
/* The same apply with CUDA events recorded on `stream` around its two parts (measurement aid for the roofline line
 * of bench.py): ms[0] = zero-fill of the accumulator, ms[1] = halo exchange + Ax launches + combine.  Synchronises. */
int libp_elliptic_operator_timed(libp_elliptic_t op, libp_dfloat* q, libp_dfloat* Aq, void* stream, double* ms);

/* ------------------------------------------------------------------ device-side setup
 * What the reference computes on the host before the first solve, computed in HBM directly.
 * libp_mesh_physical_nodes_hex3d: libs/mesh/meshPhysicalNodesHex3D.cpp - trilinear nodes x, y, z [Nelements*Np] from the
 *   element vertices EX, EY, EZ [Nelements*8] (vertex order of the reference) and the GLL nodes gllz [Nq].
 * libp_mesh_geometric_factors_hex3d: libs/mesh/meshGeometricFactorsHex3D.cpp:94-174 - ggeo [Nelements][6][Np]
 *   (G00,G01,G02,G11,G12,G22), wJ [Nelements*Np] and, when vgeo != NULL, vgeo [Nelements][12][Np]
 *   (rx..tz, J, JW, 1/JW) from the nodal coordinates and D; a non-positive Jacobian is an error, as in the reference.
 * libp_elliptic_build_diagonal_hex3d: BuildOperatorDiagonalContinuousHex3D
 *   (solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:998-1057) - element-local diagonal [Nelements*Np]
 *   (masked nodes 1), allNeumannBoost = allNeumannPenalty * allNeumannScale^2 (0 unless all-Neumann); the caller
 *   gathers it (ogs Add, Trans).
 * libp_ax_trilinear_hex3d: ellipticPartialAxTrilinearHex3D (solvers/elliptic/okl/ellipticAxHex3D.okl:440-627,
 *   ELEMENT MAP = TRILINEAR, selected at ellipticSetup.cpp:131) - element-local A_e q with the geometric factors
 *   recomputed at every node from EXYZ [Nelements][3][8] and gllzw = [GLL nodes | GLL weights] (2*Nq);
 *   GlobalToLocal != NULL reads q through the connectivity (-1 -> 0) like ellipticPartialAx*, else q is element-local. */
int libp_mesh_physical_nodes_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* EX, const libp_dfloat* EY,
                                   const libp_dfloat* EZ, const libp_dfloat* gllz, libp_dfloat* x, libp_dfloat* y,
                                   libp_dfloat* z, void* stream);
int libp_mesh_geometric_factors_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* x, const libp_dfloat* y,
                                      const libp_dfloat* z, const libp_dfloat* D, const libp_dfloat* gllw,
                                      libp_dfloat* ggeo, libp_dfloat* wJ, libp_dfloat* vgeo, void* stream);
int libp_elliptic_build_diagonal_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* ggeo, const libp_dfloat* wJ,
                                       const libp_dfloat* D, const int* mapB, libp_dfloat lambda,
                                       libp_dfloat allNeumannBoost, libp_dfloat* diagL, void* stream);
int libp_ax_trilinear_hex3d(int Nq, libp_dlong Nelements, const libp_dlong* elementList,
                            const libp_dlong* GlobalToLocal, const libp_dfloat* EXYZ, const libp_dfloat* gllzw,
                            const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* q, libp_dfloat* AqL,
                            void* stream);

/* ------------------------------------------------------------------ elliptic_t::Run pre/post steps (Hex3D, C0)
 * solvers/elliptic/src/ellipticRun.cpp:139-246.  The reference JIT-inlines the user's data file (forcing and
 * boundary functions, e.g. data/ellipticSine3D.h) into these kernels; across a C ABI they arrive evaluated at the
 * element-local nodes (device arrays of Nelements*Np entries).
 *  rhs_forcing: okl/ellipticRhsHex3D.okl:28-50        rhs = wJ .* f
 *  rhs_bc:      okl/ellipticRhsBCHex3D.okl:59-300     rhs += ndq - A_L uD   (uD = Dirichlet data on the mapB==1
 *               nodes, 0 elsewhere; ndq = summed Neumann face fluxes -WsJ n.grad u per node, may be NULL)
 *  add_bc:      okl/ellipticAddBCHex3D.okl:28-48      q[mapB==1] = uD
 *  mass_matrix: mesh_t::MassMatrixApply (collocated GLL hex mass matrix)  Mq = wJ .* q                      */
int libp_elliptic_rhs_forcing_hex3d(libp_dlong Nelements, int Np, const libp_dfloat* wJ, const libp_dfloat* f,
                                    libp_dfloat* rhs, void* stream);
int libp_elliptic_rhs_bc_hex3d(int Nq, libp_dlong Nelements, const libp_dfloat* wJ, const libp_dfloat* ggeo,
                               const libp_dfloat* D, libp_dfloat lambda, const libp_dfloat* uD, const libp_dfloat* ndq,
                               libp_dfloat* rhs, void* stream);
int libp_elliptic_add_bc_hex3d(libp_dlong Nelements, int Np, const int* mapB, const libp_dfloat* uD, libp_dfloat* q,
                               void* stream);
int libp_mass_matrix_apply_hex3d(libp_dlong Nelements, int Np, const libp_dfloat* wJ, const libp_dfloat* q,
                                 libp_dfloat* Mq, void* stream);

/* ------------------------------------------------------------------ linAlg_t
 * include/linAlg.hpp:52-120; kernels libs/linAlg/okl/linAlg*.okl.  beta==0 variants never read y. */
int libp_linalg_set(libp_dlong N, libp_dfloat alpha, libp_dfloat* a, void* stream);
int libp_linalg_add(libp_dlong N, libp_dfloat alpha, libp_dfloat* a, void* stream);
int libp_linalg_scale(libp_dlong N, libp_dfloat alpha, libp_dfloat* a, void* stream);
int libp_linalg_axpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* x, libp_dfloat beta, libp_dfloat* y, void* stream);
int libp_linalg_zaxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* x, libp_dfloat beta, const libp_dfloat* y,
                      libp_dfloat* z, void* stream);
int libp_linalg_amx(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, libp_dfloat* x, void* stream);
int libp_linalg_amxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                      libp_dfloat* y, void* stream);
int libp_linalg_zamxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                       const libp_dfloat* y, libp_dfloat* z, void* stream);
int libp_linalg_adx(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, libp_dfloat* x, void* stream);
int libp_linalg_adxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                      libp_dfloat* y, void* stream);
int libp_linalg_zadxpy(libp_dlong N, libp_dfloat alpha, const libp_dfloat* a, const libp_dfloat* x, libp_dfloat beta,
                       const libp_dfloat* y, libp_dfloat* z, void* stream);
/* Reductions (libs/linAlg/linAlg.cpp:104-224): deterministic two-level device reduction, then
 * NCCL all-reduce when comm has size>1; the result is returned to the host (blocking), as the
 * reference does.  comm may be NULL (single rank).                                          */
int libp_linalg_min(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result);
int libp_linalg_max(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result);
int libp_linalg_sum(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result);
int libp_linalg_norm2(libp_dlong N, const libp_dfloat* a, libp_comm_t comm, void* stream, libp_dfloat* result);
int libp_linalg_inner_prod(libp_dlong N, const libp_dfloat* x, const libp_dfloat* y, libp_comm_t comm, void* stream,
                           libp_dfloat* result);
int libp_linalg_weighted_norm2(libp_dlong N, const libp_dfloat* w, const libp_dfloat* a, libp_comm_t comm, void* stream,
                               libp_dfloat* result);
int libp_linalg_weighted_inner_prod(libp_dlong N, const libp_dfloat* w, const libp_dfloat* x, const libp_dfloat* y,
                                    libp_comm_t comm, void* stream, libp_dfloat* result);

/* ------------------------------------------------------------------ precon_t
 * operator_t::Operator(o_r, o_Mr) implementations (include/precon.hpp:190-210).            */
/* IdentityPrecon (include/precon.hpp) */
int libp_precon_identity_create(libp_dlong N, libp_precon_t* precon);
/* JacobiPrecon (solvers/elliptic/src/ellipticPreconJacobi.cpp:30-51): invDiagA is a device
 * array of Ndofs entries, copied.  allNeumann -> ZeroMean on the output (ellipticZeroMean.cpp). */
int libp_precon_jacobi_create(libp_dlong Ndofs, const libp_dfloat* invDiagA, int allNeumann,
                              libp_hlong NglobalDofs, libp_comm_t comm, libp_precon_t* precon);
int libp_precon_apply(libp_precon_t precon, const libp_dfloat* r, libp_dfloat* Mr, void* stream);
int libp_precon_free(libp_precon_t precon);

/* ------------------------------------------------------------------ p-multigrid level (matrix-free)
 * MGLevel (solvers/elliptic/ellipticPrecon.hpp:66-116; src/ellipticPreconMultiGridLevel.cpp:30-206).  Every
 * level owns the elliptic_t of its degree (`fine`) and refers to the elliptic_t of the next degree (`coarse`,
 * MGLevel::ellipticC): coarsen gathers onto coarse.ogsMasked, prolongate reads through coarse.GlobalToLocal.
 * Setup products (P, invDiagA, Chebyshev bounds from the Arnoldi estimate) cross the boundary as data.   */
typedef struct {
  libp_elliptic_t fine, coarse;
  int NqF, NqC;
  const libp_dfloat* P;        /* device, [NqF][NqC] 1-D degree-raise matrix (MGLevel::o_P) */
  const libp_dfloat* invDiagA; /* device, fine Ndofs (already times lambda0 for the Jacobi smoother, :374-380) */
  const libp_dfloat* weightG;  /* device, fine Ndofs (elliptic_t::o_weightG) */
  int smoother;                /* 1 = JACOBI, 2 = CHEBYSHEV (MGLevel::SmootherType) */
  libp_dfloat lambda0, lambda1;
  int ChebyshevIterations;
} libp_mglevel_desc_t;
int libp_mglevel_create(const libp_mglevel_desc_t* desc, libp_mglevel_t* level);
int libp_mglevel_free(libp_mglevel_t level);
/* multigridLevel interface (include/parAlmond.hpp:70-91): vectors are gathered vectors of Ndofs+Nhalo entries */
int libp_mglevel_operator(libp_mglevel_t level, libp_dfloat* x, libp_dfloat* Ax, void* stream);
int libp_mglevel_smooth(libp_mglevel_t level, const libp_dfloat* rhs, libp_dfloat* x, int x_is_zero, void* stream);
int libp_mglevel_residual(libp_mglevel_t level, const libp_dfloat* rhs, libp_dfloat* x, libp_dfloat* res, void* stream);
int libp_mglevel_coarsen(libp_mglevel_t level, libp_dfloat* x, libp_dfloat* Rx, void* stream);
int libp_mglevel_prolongate(libp_mglevel_t level, libp_dfloat* xC, libp_dfloat* x, void* stream); /* x += P xC */

/* ------------------------------------------------------------------ parAlmond CSR levels (V-cycle apply)
 * parCSR (include/parAlmond/parAlmondparCSR.hpp:35-100): local CSR block; arrays are HOST arrays produced by
 * the reference's AMGSetup (out of scope to rebuild) and are copied to the device.  libp_csr_create is the
 * single-rank form (empty off-diagonal block); libp_parcsr_create below takes the distributed matrix.
 * Vectors multiplied by a handle need Ncols entries and are not const: the column halo is received in place. */
int libp_csr_create(libp_dlong Nrows, libp_dlong Ncols, libp_dlong nnz, const libp_dlong* rowStarts,
                    const libp_dlong* cols, const libp_dfloat* vals, libp_dlong offd_nnz, libp_csr_t* csr);
/* Distributed parCSR (multi-rank AMG levels): local `diag` CSR block + off-rank `offd` block in the reference's
 * compressed-row MCSR form (parAlmondparCSR.hpp:62-79), whose column indices address the non-local columns
 * [NlocalCols, NlocalCols+Noffdcols) of the vector.  offd_colIds are the global ids of those columns, ascending
 * (colMap[NlocalCols:] decoded as -(colMap)-1, parAlmondparCSR.cpp:307-310), globalColStarts the column partition
 * (size+1 entries).  parCSR::haloSetup's halo_t is replaced by a pack kernel + grouped NCCL send/recv straight
 * into the tail of the input vector, which therefore needs NlocalCols+Noffdcols entries and is written by every
 * product (as in the reference, where SpMV takes a non-const o_x).  Collective over `comm`.                  */
typedef struct {
  libp_dlong Nrows, NlocalCols;
  libp_dlong diag_nnz;
  const libp_dlong* diag_rowStarts; /* Nrows+1 */
  const libp_dlong* diag_cols;
  const libp_dfloat* diag_vals;
  libp_dlong offd_nnz, offd_nzRows;
  const libp_dlong* offd_rows;       /* offd_nzRows: local row of every compressed row */
  const libp_dlong* offd_mRowStarts; /* offd_nzRows+1 */
  const libp_dlong* offd_cols;       /* in [NlocalCols, NlocalCols+Noffdcols) */
  const libp_dfloat* offd_vals;
  libp_dlong Noffdcols;
  const libp_hlong* offd_colIds;     /* Noffdcols global column ids, strictly ascending */
  const libp_hlong* globalColStarts; /* comm size + 1 */
} libp_parcsr_desc_t;
int libp_parcsr_create(libp_comm_t comm, const libp_parcsr_desc_t* desc, libp_csr_t* csr);
/* Shape and exchange lists of a CSR handle (host; for the setup checks): Ncols = NlocalCols + Noffdcols */
int libp_csr_info(libp_csr_t csr, libp_dlong* Nrows, libp_dlong* NlocalCols, libp_dlong* Ncols, libp_dlong* Nsend,
                  const libp_dlong** sendIds, int* NranksSend, int* NranksRecv);
int libp_csr_free(libp_csr_t csr);
/* parCSR::SpMV (libs/parAlmond/parAlmondparCSR.cpp:99-141): z = beta*y + alpha*A*x  (z may alias y) */
int libp_csr_spmv(libp_csr_t A, libp_dfloat alpha, libp_dfloat* x, libp_dfloat beta, const libp_dfloat* y,
                  libp_dfloat* z, void* stream);
/* amgLevel (libs/parAlmond/parAlmondAMGLevel.cpp:48-84): A, P (may be NULL on the last level), R (may be NULL),
 * diagInv (host, Nrows), smoother 0 = DAMPED_JACOBI (lambda) / 1 = CHEBYSHEV (lambda0, lambda1)         */
int libp_amglevel_create(libp_csr_t A, libp_csr_t P, libp_csr_t R, const libp_dfloat* diagInv, int smoother,
                         libp_dfloat lambda, libp_dfloat lambda0, libp_dfloat lambda1, int ChebyshevIterations,
                         libp_amglevel_t* level);
int libp_amglevel_free(libp_amglevel_t level);
int libp_amglevel_smooth(libp_amglevel_t level, const libp_dfloat* rhs, libp_dfloat* x, int x_is_zero, void* stream);
int libp_amglevel_residual(libp_amglevel_t level, const libp_dfloat* rhs, libp_dfloat* x, libp_dfloat* res, void* stream);
int libp_amglevel_coarsen(libp_amglevel_t level, libp_dfloat* x, libp_dfloat* Rx, void* stream);
int libp_amglevel_prolongate(libp_amglevel_t level, libp_dfloat* xC, libp_dfloat* x, void* stream);
/* exactSolver_t::solve (libs/parAlmond/parAlmondCoarseExact.cpp:35-73): x = invA * rhs with the dense inverse
 * stored transposed, invAT[n + m*N] (host array, N*N; single rank: offdTotal == 0)                      */
int libp_coarse_exact_create(int N, const libp_dfloat* diagInvAT, libp_coarse_t* coarse);
/* Multi-rank exact solver: rank r owns coarse rows [coarseOffsets[r], coarseOffsets[r+1]) (= A.globalRowStarts,
 * parAlmondCoarseExact.cpp:88-97).  diagInvAT[n + m*N] as above; offdInvAT[n + m*N] for the offdTotal =
 * coarseTotal-N right-hand-side entries of the other ranks in ascending rank order (:120-170).  The Alltoallv of
 * the right-hand side through the host (:60-61) becomes one grouped NCCL exchange on the device.            */
int libp_coarse_exact_create_par(libp_comm_t comm, int N, const libp_hlong* coarseOffsets, const libp_dfloat* diagInvAT,
                                 const libp_dfloat* offdInvAT, libp_coarse_t* coarse);
int libp_coarse_free(libp_coarse_t coarse);
int libp_coarse_solve(libp_coarse_t coarse, const libp_dfloat* rhs, libp_dfloat* x, void* stream);

/* multigrid_t (include/parAlmond.hpp:100-190; vcycle libs/parAlmond/parAlmondVcycle.cpp:34-60): levels are
 * appended finest first (matrix-free p-MG levels, then CSR levels), the coarse solver closes the hierarchy. */
int libp_multigrid_create(libp_comm_t comm, libp_multigrid_t* mg);
int libp_multigrid_add_mglevel(libp_multigrid_t mg, libp_mglevel_t level);
int libp_multigrid_add_amglevel(libp_multigrid_t mg, libp_amglevel_t level);
int libp_multigrid_set_coarse(libp_multigrid_t mg, libp_coarse_t coarse);
int libp_multigrid_vcycle(libp_multigrid_t mg, const libp_dfloat* rhs, libp_dfloat* x, void* stream);
/* PARALMOND CYCLE: kcycle = 0 VCYCLE (parAlmondVcycle.cpp:34-60), 1 KCYCLE (multigrid_t::kcycle,
 * libs/parAlmond/parAlmondKcycle.cpp:34-249: two inner Krylov steps on the first NUMKCYCLES = 3 coarse levels with
 * KCYCLETOL = 0.2, fused reductions of okl/kcycleCombinedOp.okl / vectorAddInnerProd.okl); nonsym = 1 selects the
 * GMRES-type inner products (PARALMOND CYCLE = NONSYM).  libp_multigrid_cycle runs the selected cycle; the
 * preconditioner made by libp_precon_multigrid_create uses it too. */
int libp_multigrid_set_cycle(libp_multigrid_t mg, int kcycle, int nonsym);
int libp_multigrid_cycle(libp_multigrid_t mg, const libp_dfloat* rhs, libp_dfloat* x, void* stream);
int libp_multigrid_free(libp_multigrid_t mg);
/* MultiGridPrecon (solvers/elliptic/src/ellipticPreconMultiGrid.cpp:29-37): one V-cycle (+ZeroMean when allNeumann) */
int libp_precon_multigrid_create(libp_multigrid_t mg, int allNeumann, libp_hlong NglobalDofs, libp_comm_t comm,
                                 libp_precon_t* precon);

/* ------------------------------------------------------------------ LinearSolver::pcg
 * linearSolverBase_t ctor (include/linearSolver.hpp:88-95) + pcg (libs/linearSolver/
 * linearSolverPCG.cpp:35-171).  stopping: 0 = ABS/REL-INITRESID (default), 1 = ABS/REL-RHS-2NORM. */
typedef int (*libp_operator_fn)(void* ctx, libp_dfloat* in, libp_dfloat* out, void* stream);

int libp_pcg_create(libp_dlong N, libp_dlong Nhalo, int flexible, int stopping, libp_comm_t comm, libp_pcg_t* pcg);
int libp_pcg_free(libp_pcg_t pcg);
/* Generic path: A and M are callbacks (an un-replaced reference operator still works).
 * Returns the iteration count in *iters.  Same control flow, scalars on the host.          */
int libp_pcg_solve_cb(libp_pcg_t pcg, libp_operator_fn A, void* Actx, libp_operator_fn M, void* Mctx,
                      libp_dfloat* x, libp_dfloat* r, libp_dfloat tol, int maxit, int verbose, void* stream, int* iters);
/* Native fast path: both operators are library handles; vector updates are fused with their
 * dot products and alpha/beta stay on the device (no host sync inside the iteration other
 * than the convergence read-back every `check_every` iterations; 1 = reference behaviour). */
int libp_pcg_solve(libp_pcg_t pcg, libp_elliptic_t A, libp_precon_t M, libp_dfloat* x, libp_dfloat* r,
                   libp_dfloat tol, int maxit, int verbose, void* stream, int* iters);
/* Residual norms sqrt(r.r) recorded by the last solve: [0] initial, [i] after iteration i. */
int libp_pcg_residual_history(libp_pcg_t pcg, const libp_dfloat** hist, int* n);

/* ------------------------------------------------------------------ LinearSolver::nbpcg (LINEAR SOLVER = NBPCG)
 * libs/linearSolver/linearSolverNBPCG.cpp:35-229: Gropp's non-blocking PCG.  One fused kernel per update
 * (update1: p, s, p.s ; update2: r, z, r.z, z.z, r.r), every reduction is in flight while the next operator /
 * preconditioner apply runs.  Same constructor arguments, stopping rule (ABS/REL-INITRESID), verbose lines and
 * return value (iterations) as the reference. */
int libp_nbpcg_create(libp_dlong N, libp_dlong Nhalo, libp_comm_t comm, libp_nbpcg_t* solver);
int libp_nbpcg_free(libp_nbpcg_t solver);
int libp_nbpcg_solve_cb(libp_nbpcg_t solver, libp_operator_fn A, void* Actx, libp_operator_fn M, void* Mctx,
                        libp_dfloat* x, libp_dfloat* r, libp_dfloat tol, int maxit, int verbose, void* stream,
                        int* iters);
int libp_nbpcg_solve(libp_nbpcg_t solver, libp_elliptic_t A, libp_precon_t M, libp_dfloat* x, libp_dfloat* r,
                     libp_dfloat tol, int maxit, int verbose, void* stream, int* iters);
int libp_nbpcg_residual_history(libp_nbpcg_t solver, const libp_dfloat** hist, int* n);

/* ------------------------------------------------------------------ LinearSolver::nbfpcg (LINEAR SOLVER = NBFPCG)
 * libs/linearSolver/linearSolverNBFPCG.cpp:71-246: non-blocking FLEXIBLE PCG (one fused update per iteration posts
 * u.r, u.s, u.w, r.r; the next preconditioner and operator applies are queued before the host waits for them).
 * Same calling convention as nbpcg; returns the reference's iteration counter. */
int libp_nbfpcg_create(libp_dlong N, libp_dlong Nhalo, libp_comm_t comm, libp_nbfpcg_t* solver);
int libp_nbfpcg_free(libp_nbfpcg_t solver);
int libp_nbfpcg_solve_cb(libp_nbfpcg_t solver, libp_operator_fn A, void* Actx, libp_operator_fn M, void* Mctx,
                         libp_dfloat* x, libp_dfloat* r, libp_dfloat tol, int maxit, int verbose, void* stream,
                         int* iters);
int libp_nbfpcg_solve(libp_nbfpcg_t solver, libp_elliptic_t A, libp_precon_t M, libp_dfloat* x, libp_dfloat* r,
                      libp_dfloat tol, int maxit, int verbose, void* stream, int* iters);
int libp_nbfpcg_residual_history(libp_nbfpcg_t solver, const libp_dfloat** hist, int* n);

/* ------------------------------------------------------------------ initial-guess strategies of linearSolver_t
 * libs/linearSolver/initialGuess.cpp, include/initialGuess.hpp; linearSolver_t::Solve brackets every solve with
 * FormInitialGuess / Update (libs/linearSolver/linearSolver.cpp:31-44).  strategy = INITIAL GUESS STRATEGY:
 * 0 NONE, 1 ZERO, 2 CLASSIC (Fischer projection), 3 QR (rolling-QR projection), 4 EXTRAP.
 * maxDim = INITIAL GUESS HISTORY SPACE DIMENSION, extrapDegree = INITIAL GUESS EXTRAP DEGREE,
 * cpqr = INITIAL GUESS EXTRAP COEFFS METHOD (0 MINNORM, 1 CPQR).  Vectors have N (+ Nhalo for the operator) entries.
 * libp_ig_extrap_coeffs = Extrap::extrapCoeffs (coefficients c[0:M] for degree m). */
int libp_ig_create(int strategy, libp_dlong N, libp_dlong Nhalo, int maxDim, int extrapDegree, int cpqr,
                   libp_comm_t comm, libp_ig_t* ig);
int libp_ig_free(libp_ig_t ig);
int libp_ig_dimension(libp_ig_t ig, int* curDim);
int libp_ig_form_initial_guess(libp_ig_t ig, libp_dfloat* x, const libp_dfloat* rhs, void* stream);
int libp_ig_update(libp_ig_t ig, libp_operator_fn A, void* Actx, libp_dfloat* x, const libp_dfloat* rhs, void* stream);
int libp_ig_extrap_coeffs(int m, int M, int cpqr, double* c);

#ifdef __cplusplus
}
#endif
#endif /* LIBP_B200_H */
