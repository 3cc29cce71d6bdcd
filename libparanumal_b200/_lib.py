"""ctypes binding of the C ABI declared in include/libp_b200.h.

The shared library is hand-written CUDA (sm_100a) + C++; there is NO CPU fallback: if the
library is missing the import fails loudly, and compute entry points fail with LIBP_ERROR when
no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_VAR = os.environ.get("LIBP_B200_VARIANT", "")  # development builds (see build.py); empty = the product library
LIB_PATH = os.path.join(HERE, "lib", "libparanumal_b200" + ("_" + _VAR if _VAR else "") + ".so")

SUCCESS = 0
FLOAT, DOUBLE, INT32, INT64 = 0, 1, 2, 3
ADD, MUL, MAX, MIN = 0, 1, 2, 3
SYM, NOTRANS, TRANS = 0, 1, 2
UNSIGNED, SIGNED, HALO = 0, 1, 2

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f64 = C.c_double
P = C.POINTER


class HostCollectives(C.Structure):
    ALLTOALL = C.CFUNCTYPE(i32, vp, vp, vp, C.c_size_t)
    ALLTOALLV = C.CFUNCTYPE(i32, vp, vp, P(i64), P(i64), vp, P(i64), P(i64))
    ALLREDUCE_I64 = C.CFUNCTYPE(i32, vp, P(i64), i32, i32)
    ALLREDUCE_F64 = C.CFUNCTYPE(i32, vp, P(f64), i32, i32)
    _fields_ = [("ctx", vp), ("alltoall", ALLTOALL), ("alltoallv", ALLTOALLV),
                ("allreduce_i64", ALLREDUCE_I64), ("allreduce_f64", ALLREDUCE_F64)]


class OgsInfo(C.Structure):
    _fields_ = [("N", i32), ("Ngather", i32), ("NlocalT", i32), ("NlocalP", i32), ("NhaloT", i32), ("NhaloP", i32),
                ("Nhalo", i32), ("NgatherGlobal", i64), ("gather_defined", i32),
                ("NranksSendN", i32), ("NranksSendT", i32), ("NranksRecvN", i32), ("NranksRecvT", i32),
                ("NsendN", i32), ("NsendT", i32), ("NrecvN", i32), ("NrecvT", i32)]


class EllipticDesc(C.Structure):
    _fields_ = [("Nq", i32), ("Nelements", i32), ("NlocalGatherElements", i32), ("NglobalGatherElements", i32),
                ("localGatherElementList", vp), ("globalGatherElementList", vp), ("GlobalToLocal", vp),
                ("wJ", vp), ("ggeo", vp), ("D", vp), ("lambda_", f64), ("ogsMasked", vp), ("mode", i32)]


class IpdgDesc(C.Structure):
    _fields_ = [("Nq", i32), ("Nelements", i32), ("NhaloElementsTotal", i32), ("NinternalElements", i32),
                ("NhaloElements", i32), ("internalElementIds", vp), ("haloElementIds", vp), ("vmapM", vp), ("vmapP", vp),
                ("vgeo", vp), ("sgeo", vp), ("EToB", vp), ("D", vp), ("lambda_", f64), ("tau", f64), ("traceHalo", vp)]


class MGLevelDesc(C.Structure):
    _fields_ = [("fine", vp), ("coarse", vp), ("NqF", i32), ("NqC", i32), ("P", vp), ("invDiagA", vp), ("weightG", vp),
                ("smoother", i32), ("lambda0", f64), ("lambda1", f64), ("ChebyshevIterations", i32)]


class ParCsrDesc(C.Structure):
    _fields_ = [("Nrows", i32), ("NlocalCols", i32), ("diag_nnz", i32), ("diag_rowStarts", vp), ("diag_cols", vp),
                ("diag_vals", vp), ("offd_nnz", i32), ("offd_nzRows", i32), ("offd_rows", vp), ("offd_mRowStarts", vp),
                ("offd_cols", vp), ("offd_vals", vp), ("Noffdcols", i32), ("offd_colIds", vp), ("globalColStarts", vp)]


OPERATOR_FN = C.CFUNCTYPE(i32, vp, vp, vp, vp)

# name -> (restype, argtypes); every symbol include/libp_b200.h declares
SIGNATURES = {
    "libp_last_error": (C.c_char_p, []),
    "libp_b200_version": (C.c_char_p, []),
    "libp_b200_init": (i32, [i32]),
    "libp_b200_finish": (i32, [vp]),
    "libp_comm_create": (i32, [i32, i32, P(HostCollectives), P(vp)]),
    "libp_comm_free": (i32, [vp]),
    "libp_comm_rank": (i32, [vp, P(i32), P(i32)]),
    "libp_comm_nccl_unique_id": (i32, [vp]),
    "libp_comm_nccl_init": (i32, [vp, vp]),
    "libp_comm_p2p_init": (i32, [vp, C.c_size_t]),
    "libp_comm_p2p_enabled": (i32, [vp, vp]),
    "libp_comm_p2p_status": (i32, [vp, P(i32)]),
    "libp_comm_p2p_reset": (i32, [vp]),
    "libp_ogs_setup": (i32, [i32, vp, vp, i32, i32, i32, P(vp)]),
    "libp_ogs_free": (i32, [vp]),
    "libp_ogs_sort_selftest": (i32, [i32, i32, C.c_uint, P(i32)]),
    "libp_ogs_rand_selftest": (i32, [C.c_uint, i32, P(i32)]),
    "libp_ax_chain_layout_selftest": (i32, [i32, P(i32), P(i32)]),
    "libp_ogs_info": (i32, [vp, P(OgsInfo)]),
    "libp_ogs_maps": (i32, [vp, i32, P(i32), P(i32), P(vp), P(vp), P(vp), P(vp)]),
    "libp_ogs_exchange_lists": (i32, [vp, i32, P(i32), P(vp), P(i32), P(vp), P(vp), P(vp), P(i32), P(vp), P(vp), P(vp)]),
    "libp_ogs_global_to_local": (i32, [vp, vp]),
    "libp_ogs_gather": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "libp_ogs_gather_start": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "libp_ogs_gather_finish": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "libp_ogs_scatter": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "libp_ogs_scatter_start": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "libp_ogs_scatter_finish": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "libp_ogs_gather_scatter": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "libp_ogs_gather_scatter_start": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "libp_ogs_gather_scatter_finish": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "libp_halo_exchange_start": (i32, [vp, vp, i32, i32, vp]),
    "libp_halo_exchange_finish": (i32, [vp, vp, i32, i32, vp]),
    "libp_halo_exchange": (i32, [vp, vp, i32, i32, vp]),
    "libp_ax_hex3d": (i32, [i32, i32, vp, vp, vp, vp, vp, f64, vp, vp, vp]),
    "libp_ax_hex3d_gather": (i32, [i32, i32, vp, vp, vp, vp, vp, f64, vp, vp, vp]),
    "libp_ax_hex3d_set_variant": (i32, [i32]),
    "libp_ax_hex3d_register_D": (i32, [i32, vp]),
    "libp_ax_hex3d_unregister_D": (i32, [vp]),
    "libp_ax_hex3d_tune": (i32, [i32, i32, i32]),
    "libp_mesh_physical_nodes_hex3d": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "libp_mesh_geometric_factors_hex3d": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "libp_elliptic_build_diagonal_hex3d": (i32, [i32, i32, vp, vp, vp, vp, f64, f64, vp, vp]),
    "libp_ax_trilinear_hex3d": (i32, [i32, i32, vp, vp, vp, vp, vp, f64, vp, vp, vp]),
    "libp_elliptic_create": (i32, [P(EllipticDesc), P(vp)]),
    "libp_elliptic_create_ipdg": (i32, [P(IpdgDesc), P(vp)]),
    "libp_elliptic_ipdg_gradient": (i32, [vp, P(vp)]),
    "libp_elliptic_build_diagonal_ipdg_hex3d": (i32, [i32, i32, vp, vp, vp, vp, f64, f64, vp, vp]),
    "libp_mesh_surface_geometric_factors_hex3d": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "libp_mesh_surface_hinv_hex3d": (i32, [i32, i32, vp, vp, vp, vp]),
    "libp_elliptic_rhs_bc_ipdg_hex3d": (i32, [i32, i32, f64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "libp_elliptic_rhs_forcing_hex3d": (i32, [i32, i32, vp, vp, vp, vp]),
    "libp_elliptic_rhs_bc_hex3d": (i32, [i32, i32, vp, vp, vp, f64, vp, vp, vp, vp]),
    "libp_elliptic_add_bc_hex3d": (i32, [i32, i32, vp, vp, vp, vp]),
    "libp_mass_matrix_apply_hex3d": (i32, [i32, i32, vp, vp, vp, vp]),
    "libp_elliptic_set_zero_ahead": (i32, [vp, i32]),
    "libp_elliptic_set_default_zero_ahead": (i32, [i32]),
    "libp_elliptic_zero_ahead_errors": (i32, [vp, P(i32)]),
    "libp_elliptic_set_chain": (i32, [vp, i32, i32]),
    "libp_elliptic_set_default_chain": (i32, [i32, i32]),
    "libp_elliptic_chain_stats": (i32, [vp, vp, P(C.c_longlong), vp]),
    "libp_elliptic_set_trilinear": (i32, [vp, vp, vp, vp]),
    "libp_elliptic_set_chunk": (i32, [vp, i32]),
    "libp_elliptic_set_default_chunk": (i32, [i32]),
    "libp_elliptic_free": (i32, [vp]),
    "libp_elliptic_operator": (i32, [vp, vp, vp, vp]),
    "libp_elliptic_operator_timed": (i32, [vp, vp, vp, vp, P(f64)]),
    "libp_linalg_set": (i32, [i32, f64, vp, vp]),
    "libp_linalg_add": (i32, [i32, f64, vp, vp]),
    "libp_linalg_scale": (i32, [i32, f64, vp, vp]),
    "libp_linalg_axpy": (i32, [i32, f64, vp, f64, vp, vp]),
    "libp_linalg_zaxpy": (i32, [i32, f64, vp, f64, vp, vp, vp]),
    "libp_linalg_amx": (i32, [i32, f64, vp, vp, vp]),
    "libp_linalg_amxpy": (i32, [i32, f64, vp, vp, f64, vp, vp]),
    "libp_linalg_zamxpy": (i32, [i32, f64, vp, vp, f64, vp, vp, vp]),
    "libp_linalg_adx": (i32, [i32, f64, vp, vp, vp]),
    "libp_linalg_adxpy": (i32, [i32, f64, vp, vp, f64, vp, vp]),
    "libp_linalg_zadxpy": (i32, [i32, f64, vp, vp, f64, vp, vp, vp]),
    "libp_linalg_min": (i32, [i32, vp, vp, vp, P(f64)]),
    "libp_linalg_max": (i32, [i32, vp, vp, vp, P(f64)]),
    "libp_linalg_sum": (i32, [i32, vp, vp, vp, P(f64)]),
    "libp_linalg_norm2": (i32, [i32, vp, vp, vp, P(f64)]),
    "libp_linalg_inner_prod": (i32, [i32, vp, vp, vp, vp, P(f64)]),
    "libp_linalg_weighted_norm2": (i32, [i32, vp, vp, vp, vp, P(f64)]),
    "libp_linalg_weighted_inner_prod": (i32, [i32, vp, vp, vp, vp, vp, P(f64)]),
    "libp_precon_identity_create": (i32, [i32, P(vp)]),
    "libp_precon_jacobi_create": (i32, [i32, vp, i32, i64, vp, P(vp)]),
    "libp_precon_apply": (i32, [vp, vp, vp, vp]),
    "libp_precon_free": (i32, [vp]),
    "libp_mglevel_create": (i32, [P(MGLevelDesc), P(vp)]),
    "libp_mglevel_free": (i32, [vp]),
    "libp_mglevel_operator": (i32, [vp, vp, vp, vp]),
    "libp_mglevel_smooth": (i32, [vp, vp, vp, i32, vp]),
    "libp_mglevel_residual": (i32, [vp, vp, vp, vp, vp]),
    "libp_mglevel_coarsen": (i32, [vp, vp, vp, vp]),
    "libp_mglevel_prolongate": (i32, [vp, vp, vp, vp]),
    "libp_csr_create": (i32, [i32, i32, i32, vp, vp, vp, i32, P(vp)]),
    "libp_parcsr_create": (i32, [vp, P(ParCsrDesc), P(vp)]),
    "libp_csr_info": (i32, [vp, P(i32), P(i32), P(i32), P(i32), P(vp), P(i32), P(i32)]),
    "libp_csr_free": (i32, [vp]),
    "libp_csr_spmv": (i32, [vp, f64, vp, f64, vp, vp, vp]),
    "libp_amglevel_create": (i32, [vp, vp, vp, vp, i32, f64, f64, f64, i32, P(vp)]),
    "libp_amglevel_free": (i32, [vp]),
    "libp_amglevel_smooth": (i32, [vp, vp, vp, i32, vp]),
    "libp_amglevel_residual": (i32, [vp, vp, vp, vp, vp]),
    "libp_amglevel_coarsen": (i32, [vp, vp, vp, vp]),
    "libp_amglevel_prolongate": (i32, [vp, vp, vp, vp]),
    "libp_coarse_exact_create": (i32, [i32, vp, P(vp)]),
    "libp_coarse_exact_create_par": (i32, [vp, i32, vp, vp, vp, P(vp)]),
    "libp_coarse_free": (i32, [vp]),
    "libp_coarse_solve": (i32, [vp, vp, vp, vp]),
    "libp_multigrid_create": (i32, [vp, P(vp)]),
    "libp_multigrid_add_mglevel": (i32, [vp, vp]),
    "libp_multigrid_add_amglevel": (i32, [vp, vp]),
    "libp_multigrid_set_coarse": (i32, [vp, vp]),
    "libp_multigrid_vcycle": (i32, [vp, vp, vp, vp]),
    "libp_multigrid_set_cycle": (i32, [vp, i32, i32]),
    "libp_multigrid_cycle": (i32, [vp, vp, vp, vp]),
    "libp_multigrid_free": (i32, [vp]),
    "libp_precon_multigrid_create": (i32, [vp, i32, i64, vp, P(vp)]),
    "libp_pcg_create": (i32, [i32, i32, i32, i32, vp, P(vp)]),
    "libp_pcg_free": (i32, [vp]),
    "libp_pcg_solve_cb": (i32, [vp, OPERATOR_FN, vp, OPERATOR_FN, vp, vp, vp, f64, i32, i32, vp, P(i32)]),
    "libp_pcg_solve": (i32, [vp, vp, vp, vp, vp, f64, i32, i32, vp, P(i32)]),
    "libp_pcg_residual_history": (i32, [vp, P(vp), P(i32)]),
    "libp_nbpcg_create": (i32, [i32, i32, vp, P(vp)]),
    "libp_nbpcg_free": (i32, [vp]),
    "libp_nbpcg_solve_cb": (i32, [vp, OPERATOR_FN, vp, OPERATOR_FN, vp, vp, vp, f64, i32, i32, vp, P(i32)]),
    "libp_nbpcg_solve": (i32, [vp, vp, vp, vp, vp, f64, i32, i32, vp, P(i32)]),
    "libp_nbpcg_residual_history": (i32, [vp, P(vp), P(i32)]),
    "libp_ig_create": (i32, [i32, i32, i32, i32, i32, i32, vp, P(vp)]),
    "libp_ig_free": (i32, [vp]),
    "libp_ig_dimension": (i32, [vp, P(i32)]),
    "libp_ig_form_initial_guess": (i32, [vp, vp, vp, vp]),
    "libp_ig_update": (i32, [vp, OPERATOR_FN, vp, vp, vp, vp]),
    "libp_ig_extrap_coeffs": (i32, [i32, i32, i32, P(f64)]),
    "libp_nbfpcg_create": (i32, [i32, i32, vp, P(vp)]),
    "libp_nbfpcg_free": (i32, [vp]),
    "libp_nbfpcg_solve_cb": (i32, [vp, OPERATOR_FN, vp, OPERATOR_FN, vp, vp, vp, f64, i32, i32, vp, P(i32)]),
    "libp_nbfpcg_solve": (i32, [vp, vp, vp, vp, vp, f64, i32, i32, vp, P(i32)]),
    "libp_nbfpcg_residual_history": (i32, [vp, P(vp), P(i32)]),
}

_lib = None


class LibpError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m libparanumal_b200.build` "
                "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the hot path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != SUCCESS:
        raise LibpError(load().libp_last_error().decode())
