"""Host-side mirror of the reference interfaces on the elliptic hot path, over the C ABI.

Names and argument meaning follow libParanumal (comm_t, ogs::ogs_t / halo_t, linAlg_t,
elliptic_t::Operator, LinearSolver::pcg, precon_t) so the tests read like the reference's use of
them.  torch is used only for device memory (tensors hand their data_ptr() to the library), the
current CUDA stream, and torch.distributed for the setup-time host collectives / NCCL bootstrap.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._lib import (ADD, DOUBLE, FLOAT, HALO, INT32, INT64, MAX, MIN, MUL, NOTRANS, SIGNED, SYM, TRANS,
                   UNSIGNED, check)

_TYPE_OF = {torch.float32: FLOAT, torch.float64: DOUBLE, torch.int32: INT32, torch.int64: INT64}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream) if torch.cuda.is_available() else C.c_void_p(0)


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(int(t))


def init(device_id=0):
    """platform_t device selection."""
    check(L.load().libp_b200_init(int(device_id)))
    torch.cuda.set_device(device_id)


def version():
    return L.load().libp_b200_version().decode()


# --------------------------------------------------------------------------- comm_t
class Comm:
    """comm_t: rank/size + setup-time host collectives (torch.distributed, any backend that works on
    CPU tensors, e.g. gloo) + NCCL for the data path."""

    def __init__(self, rank=0, size=1, group=None):
        self.rank, self.size, self.group = rank, size, group
        self._h = C.c_void_p()
        self._cb = None
        host = None
        if size > 1:
            import torch.distributed as dist

            def a2a(ctx, send, recv, nbytes):
                try:
                    s = torch.from_numpy(np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), (nbytes * size,)).copy())
                    r = torch.empty_like(s)
                    self._gloo_a2a(r, s, [nbytes] * size, [nbytes] * size)
                    C.memmove(recv, r.numpy().ctypes.data, nbytes * size)
                    return 0
                except Exception as e:  # pragma: no cover
                    print("host alltoall failed:", e)
                    return 1

            def a2av(ctx, send, sc, so, recv, rc, ro):
                try:
                    scl = [int(sc[i]) for i in range(size)]
                    sol = [int(so[i]) for i in range(size)]
                    rcl = [int(rc[i]) for i in range(size)]
                    rol = [int(ro[i]) for i in range(size)]
                    stot = max((o + c for o, c in zip(sol, scl)), default=0)
                    rtot = max((o + c for o, c in zip(rol, rcl)), default=0)
                    if stot:
                        sarr = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), (stot,))
                        chunks = [torch.from_numpy(sarr[o:o + c].copy()) for o, c in zip(sol, scl)]
                    else:
                        chunks = [torch.empty(0, dtype=torch.uint8) for _ in range(size)]
                    outs = self._gloo_a2a_lists(chunks, rcl)
                    if rtot:
                        rarr = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), (rtot,))
                        for o, c, t in zip(rol, rcl, outs):
                            if c:
                                rarr[o:o + c] = t.numpy()
                    return 0
                except Exception as e:  # pragma: no cover
                    print("host alltoallv failed:", e)
                    return 1

            def ar_i64(ctx, buf, n, op):
                try:
                    arr = np.ctypeslib.as_array(buf, (n,))
                    t = torch.from_numpy(arr.copy())
                    dist.all_reduce(t, op={ADD: dist.ReduceOp.SUM, MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN}[op],
                                    group=self.group)
                    arr[:] = t.numpy()
                    return 0
                except Exception as e:  # pragma: no cover
                    print("host allreduce failed:", e)
                    return 1

            def ar_f64(ctx, buf, n, op):
                try:
                    arr = np.ctypeslib.as_array(buf, (n,))
                    t = torch.from_numpy(arr.copy())
                    dist.all_reduce(t, op={ADD: dist.ReduceOp.SUM, MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN}[op],
                                    group=self.group)
                    arr[:] = t.numpy()
                    return 0
                except Exception as e:  # pragma: no cover
                    print("host allreduce failed:", e)
                    return 1

            self._cb = (L.HostCollectives.ALLTOALL(a2a), L.HostCollectives.ALLTOALLV(a2av),
                        L.HostCollectives.ALLREDUCE_I64(ar_i64), L.HostCollectives.ALLREDUCE_F64(ar_f64))
            host = L.HostCollectives(None, *self._cb)
        check(L.load().libp_comm_create(rank, size, C.byref(host) if host is not None else None, C.byref(self._h)))

    # gloo has no all_to_all: emulate with all_gather of sizes + broadcast-free pairwise send/recv
    def _gloo_a2a_lists(self, send_list, recv_counts):
        import torch.distributed as dist
        size, rank = self.size, self.rank
        outs = [torch.empty(c, dtype=torch.uint8) for c in recv_counts]
        outs[rank] = send_list[rank].clone()
        reqs = []
        for peer in range(size):
            if peer == rank:
                continue
            if recv_counts[peer]:
                reqs.append(dist.irecv(outs[peer], src=dist.get_global_rank(self.group, peer) if self.group else peer,
                                       group=self.group))
        for peer in range(size):
            if peer == rank:
                continue
            if send_list[peer].numel():
                reqs.append(dist.isend(send_list[peer], dst=dist.get_global_rank(self.group, peer) if self.group else peer,
                                       group=self.group))
        for r in reqs:
            r.wait()
        return outs

    def _gloo_a2a(self, r, s, scounts, rcounts):
        chunks = list(torch.split(s, scounts))
        outs = self._gloo_a2a_lists(chunks, rcounts)
        r.copy_(torch.cat(outs))

    def allgather_object(self, obj):
        """setup-time host collective (comm_t::Allgatherv of setup arrays): list of every rank's object"""
        if self.size == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def allreduce_sum(self, x):
        """setup-time host all-reduce of a python float / numpy array"""
        if self.size == 1:
            return x
        import torch.distributed as dist
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)).clone().reshape(-1)
        dist.all_reduce(t, group=self.group)
        return float(t[0]) if np.ndim(x) == 0 else t.numpy().reshape(np.shape(x))

    def init_nccl(self):
        """NCCL bootstrap: rank 0 creates the unique id, torch.distributed broadcasts it."""
        if self.size == 1:
            return
        import torch.distributed as dist
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            check(L.load().libp_comm_nccl_unique_id(_ptr(uid)))
        t = torch.from_numpy(uid)
        if dist.get_backend(self.group) == "nccl":
            tc = t.cuda()
            dist.broadcast(tc, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
            t = tc.cpu()
        else:
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        uid = t.numpy().copy()
        check(L.load().libp_comm_nccl_init(self._h, _ptr(uid)))

    def init_p2p(self, window_bytes=0, required=False):
        """NVLink peer window (CUDA IPC): halo exchange and PCG scalar all-reduces run inside our own kernels.
        Returns True when every rank mapped every peer; otherwise the NCCL path stays in use."""
        if self.size == 1:
            return False
        rc = L.load().libp_comm_p2p_init(self._h, int(window_bytes))
        if rc != 0:
            if required:
                check(rc)
            if self.rank == 0:
                print("peer window unavailable, using NCCL:", L.load().libp_last_error().decode())
            return False
        return True

    @property
    def p2p(self):
        e = C.c_int(0)
        check(L.load().libp_comm_p2p_enabled(self._h, C.byref(e)))
        return bool(e.value)

    def p2p_timed_out(self):
        """True when an in-kernel wait on a peer exceeded its bound (libp_comm_p2p_status)"""
        e = C.c_int(0)
        check(L.load().libp_comm_p2p_status(self._h, C.byref(e)))
        return bool(e.value)

    @property
    def handle(self):
        return self._h

    def free(self):
        if self._h:
            L.load().libp_comm_free(self._h)
            self._h = C.c_void_p()


# --------------------------------------------------------------------------- ogs_t / halo_t
class Ogs:
    """ogs::ogs_t (include/ogs.hpp:216-344) + the gathered halo_t built from it."""

    def __init__(self):
        self._h = C.c_void_p()
        self.comm = None

    def Setup(self, N, ids, comm, kind=SIGNED, unique=False, verbose=False):
        """ids: numpy int64 array of N entries; rewritten in place when unique (like the reference)."""
        assert isinstance(ids, np.ndarray) and ids.dtype == np.int64 and ids.flags["C_CONTIGUOUS"]
        assert ids.size == N
        self.comm = comm
        check(L.load().libp_ogs_setup(int(N), _ptr(ids), comm.handle, int(kind), int(bool(unique)), int(verbose),
                                      C.byref(self._h)))
        info = L.OgsInfo()
        check(L.load().libp_ogs_info(self._h, C.byref(info)))
        self.info = info
        for f, _ in L.OgsInfo._fields_:
            setattr(self, f, getattr(info, f))
        return self

    @property
    def handle(self):
        return self._h

    def maps(self, which):
        """which: 'local' | 'halo' | 'postmpi' -> dict of numpy copies of the CSR maps."""
        w = {"local": 0, "halo": 1, "postmpi": 2}[which]
        nN, nT = C.c_int(), C.c_int()
        ps = [C.c_void_p() for _ in range(4)]
        check(L.load().libp_ogs_maps(self._h, w, C.byref(nN), C.byref(nT), *[C.byref(p) for p in ps]))

        def arr(p, n):
            if n == 0 or not p.value:
                return np.zeros(n, dtype=np.int32)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), (n,)).copy()

        rsN = arr(ps[0], nT.value + 1)
        rsT = arr(ps[1], nT.value + 1)
        return dict(NrowsN=nN.value, NrowsT=nT.value, rowStartsN=rsN, rowStartsT=rsT,
                    colIdsN=arr(ps[2], int(rsN[-1])), colIdsT=arr(ps[3], int(rsT[-1])))

    def exchange_lists(self, trans):
        ns = C.c_int()
        nrs, nrr = C.c_int(), C.c_int()
        p = [C.c_void_p() for _ in range(7)]
        check(L.load().libp_ogs_exchange_lists(self._h, int(trans), C.byref(ns), C.byref(p[0]), C.byref(nrs),
                                               C.byref(p[1]), C.byref(p[2]), C.byref(p[3]), C.byref(nrr),
                                               C.byref(p[4]), C.byref(p[5]), C.byref(p[6])))

        def arr(pp, n):
            if n == 0 or not pp.value:
                return np.zeros(n, dtype=np.int32)
            return np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_int32)), (n,)).copy()

        return dict(sendIds=arr(p[0], ns.value), sendRanks=arr(p[1], nrs.value), sendCounts=arr(p[2], nrs.value),
                    sendOffsets=arr(p[3], nrs.value + 1), recvRanks=arr(p[4], nrr.value),
                    recvCounts=arr(p[5], nrr.value), recvOffsets=arr(p[6], nrr.value + 1))

    def SetupGlobalToLocalMapping(self):
        out = np.empty(self.N, dtype=np.int32)
        check(L.load().libp_ogs_global_to_local(self._h, _ptr(out)))
        return out

    # device apply on torch tensors
    def Gather(self, gv, v, k=1, op=ADD, trans=TRANS):
        check(L.load().libp_ogs_gather(self._h, _ptr(gv), _ptr(v), k, _TYPE_OF[v.dtype], op, trans, _stream()))

    def GatherStart(self, gv, v, k=1, op=ADD, trans=TRANS):
        check(L.load().libp_ogs_gather_start(self._h, _ptr(gv), _ptr(v), k, _TYPE_OF[v.dtype], op, trans, _stream()))

    def GatherFinish(self, gv, v, k=1, op=ADD, trans=TRANS):
        check(L.load().libp_ogs_gather_finish(self._h, _ptr(gv), _ptr(v), k, _TYPE_OF[v.dtype], op, trans, _stream()))

    def Scatter(self, v, gv, k=1, trans=NOTRANS):
        check(L.load().libp_ogs_scatter(self._h, _ptr(v), _ptr(gv), k, _TYPE_OF[v.dtype], trans, _stream()))

    def GatherScatter(self, v, k=1, op=ADD, trans=SYM):
        check(L.load().libp_ogs_gather_scatter(self._h, _ptr(v), k, _TYPE_OF[v.dtype], op, trans, _stream()))

    def ExchangeStart(self, v, k=1):
        check(L.load().libp_halo_exchange_start(self._h, _ptr(v), k, _TYPE_OF[v.dtype], _stream()))

    def ExchangeFinish(self, v, k=1):
        check(L.load().libp_halo_exchange_finish(self._h, _ptr(v), k, _TYPE_OF[v.dtype], _stream()))

    def Exchange(self, v, k=1):
        check(L.load().libp_halo_exchange(self._h, _ptr(v), k, _TYPE_OF[v.dtype], _stream()))

    def Free(self):
        if self._h:
            L.load().libp_ogs_free(self._h)
            self._h = C.c_void_p()


# --------------------------------------------------------------------------- kernels
def ax_hex3d(Nq, Nelements, elementList, GlobalToLocal, wJ, ggeo, D, lam, q, AqL):
    """ellipticPartialAxHex3D argument order (S, MM dropped)."""
    check(L.load().libp_ax_hex3d(Nq, Nelements, _ptr(elementList), _ptr(GlobalToLocal), _ptr(wJ), _ptr(ggeo),
                                 _ptr(D), float(lam), _ptr(q), _ptr(AqL), _stream()))


def rhs_forcing_hex3d(Nelements, Np, wJ, f, rhs):
    check(L.load().libp_elliptic_rhs_forcing_hex3d(Nelements, Np, _ptr(wJ), _ptr(f), _ptr(rhs), _stream()))


def rhs_bc_hex3d(Nq, Nelements, wJ, ggeo, D, lam, uD, ndq, rhs):
    check(L.load().libp_elliptic_rhs_bc_hex3d(Nq, Nelements, _ptr(wJ), _ptr(ggeo), _ptr(D), float(lam), _ptr(uD), _ptr(ndq),
                                              _ptr(rhs), _stream()))


def add_bc_hex3d(Nelements, Np, mapB, uD, q):
    check(L.load().libp_elliptic_add_bc_hex3d(Nelements, Np, _ptr(mapB), _ptr(uD), _ptr(q), _stream()))


def mass_matrix_apply_hex3d(Nelements, Np, wJ, q, Mq):
    check(L.load().libp_mass_matrix_apply_hex3d(Nelements, Np, _ptr(wJ), _ptr(q), _ptr(Mq), _stream()))


def register_D(Nq, D):
    """Promise that the device array D is immutable: enables the even-odd kernels for GLL matrices."""
    check(L.load().libp_ax_hex3d_register_D(int(Nq), _ptr(D)))


def unregister_D(D):
    check(L.load().libp_ax_hex3d_unregister_D(_ptr(D)))


def ax_hex3d_gather(Nq, Nelements, elementList, GlobalToLocal, wJ, ggeo, D, lam, q, Aq):
    check(L.load().libp_ax_hex3d_gather(Nq, Nelements, _ptr(elementList), _ptr(GlobalToLocal), _ptr(wJ), _ptr(ggeo),
                                        _ptr(D), float(lam), _ptr(q), _ptr(Aq), _stream()))


def mesh_physical_nodes_hex3d(Nq, Nelements, EX, EY, EZ, gllz, x, y, z):
    check(L.load().libp_mesh_physical_nodes_hex3d(Nq, Nelements, _ptr(EX), _ptr(EY), _ptr(EZ), _ptr(gllz), _ptr(x), _ptr(y),
                                                  _ptr(z), _stream()))


def mesh_geometric_factors_hex3d(Nq, Nelements, x, y, z, D, gllw, ggeo, wJ, vgeo=None):
    check(L.load().libp_mesh_geometric_factors_hex3d(Nq, Nelements, _ptr(x), _ptr(y), _ptr(z), _ptr(D), _ptr(gllw),
                                                     _ptr(ggeo), _ptr(wJ), _ptr(vgeo) if vgeo is not None else None,
                                                     _stream()))


def mesh_surface_geometric_factors_hex3d(Nq, Nelements, x, y, z, D, gllw, sgeo, h):
    """step 1 of mesh_t::SurfaceGeometricFactorsHex3D: sgeo (all but IHID) and h = sJ/J per face node"""
    check(L.load().libp_mesh_surface_geometric_factors_hex3d(Nq, Nelements, _ptr(x), _ptr(y), _ptr(z), _ptr(D), _ptr(gllw),
                                                             _ptr(sgeo), _ptr(h), _stream()))


def mesh_surface_hinv_hex3d(Nq, Nelements, mapP, h, sgeo):
    """step 2: IHID = max(h-, h+) through mapP (h includes the exchanged halo part on several ranks)"""
    check(L.load().libp_mesh_surface_hinv_hex3d(Nq, Nelements, _ptr(mapP), _ptr(h), _ptr(sgeo), _stream()))


def rhs_bc_ipdg_hex3d(Nq, Nelements, tau, vgeo, sgeo, EToB, D, uD, gN, rhs):
    check(L.load().libp_elliptic_rhs_bc_ipdg_hex3d(Nq, Nelements, float(tau), _ptr(vgeo), _ptr(sgeo), _ptr(EToB), _ptr(D),
                                                   _ptr(uD) if uD is not None else None,
                                                   _ptr(gN) if gN is not None else None, _ptr(rhs), _stream()))


def elliptic_build_diagonal_ipdg_hex3d(Nq, Nelements, vgeo, sgeo, EToB, D, lam, tau, A):
    check(L.load().libp_elliptic_build_diagonal_ipdg_hex3d(Nq, Nelements, _ptr(vgeo), _ptr(sgeo), _ptr(EToB), _ptr(D),
                                                           float(lam), float(tau), _ptr(A), _stream()))


def elliptic_build_diagonal_hex3d(Nq, Nelements, ggeo, wJ, D, mapB, lam, boost, diagL):
    check(L.load().libp_elliptic_build_diagonal_hex3d(Nq, Nelements, _ptr(ggeo), _ptr(wJ), _ptr(D), _ptr(mapB), float(lam),
                                                      float(boost), _ptr(diagL), _stream()))


def ax_trilinear_hex3d(Nq, Nelements, elementList, GlobalToLocal, EXYZ, gllzw, D, lam, q, AqL):
    check(L.load().libp_ax_trilinear_hex3d(Nq, Nelements, _ptr(elementList) if elementList is not None else None,
                                           _ptr(GlobalToLocal) if GlobalToLocal is not None else None, _ptr(EXYZ),
                                           _ptr(gllzw), _ptr(D), float(lam), _ptr(q), _ptr(AqL), _stream()))


class Elliptic:
    """The C0 branch of elliptic_t::Operator as an operator_t."""

    def __init__(self, Nq, localList, globalList, GlobalToLocal, wJ, ggeo, D, lam, ogsMasked, mode=1):
        self.keep = (localList, globalList, GlobalToLocal, wJ, ggeo, D, ogsMasked)
        d = L.EllipticDesc()
        d.Nq = Nq
        d.NlocalGatherElements = 0 if localList is None else localList.numel()
        d.NglobalGatherElements = 0 if globalList is None else globalList.numel()
        d.Nelements = d.NlocalGatherElements + d.NglobalGatherElements
        d.localGatherElementList = _ptr(localList) if d.NlocalGatherElements else None
        d.globalGatherElementList = _ptr(globalList) if d.NglobalGatherElements else None
        d.GlobalToLocal, d.wJ, d.ggeo, d.D = _ptr(GlobalToLocal), _ptr(wJ), _ptr(ggeo), _ptr(D)
        d.lambda_ = float(lam)
        d.ogsMasked = ogsMasked.handle
        d.mode = mode
        self._h = C.c_void_p()
        check(L.load().libp_elliptic_create(C.byref(d), C.byref(self._h)))
        self.Ndofs = ogsMasked.Ngather
        self.Nhalo = ogsMasked.Nhalo

    @property
    def handle(self):
        return self._h

    def set_zero_ahead(self, on):
        """in-kernel zero-fill of the fused accumulator (libp_elliptic_set_zero_ahead)"""
        check(L.load().libp_elliptic_set_zero_ahead(self._h, int(bool(on))))

    def zero_ahead_errors(self):
        e = C.c_int(0)
        check(L.load().libp_elliptic_zero_ahead_errors(self._h, C.byref(e)))
        return e.value

    def set_chain(self, chain_elements, stages=1):
        """element-chain kernel: elements per chain (0 = off), TMA stages; see libp_elliptic_set_chain"""
        check(L.load().libp_elliptic_set_chain(self._h, int(chain_elements), int(stages)))

    def chain_stats(self, o_Aq):
        st = (C.c_longlong * 6)()
        check(L.load().libp_elliptic_chain_stats(self._h, _ptr(o_Aq), st, _stream()))
        return dict(chain=st[0], sectors=st[1], zero_sectors=st[2], positions=st[3], raw_elements=st[4], stages=st[5])

    def set_trilinear(self, EXYZ, gllz=None, gllw=None):
        """ELEMENT MAP = TRILINEAR: geometry recomputed from the element vertices EXYZ (device [E][3][8]);
        EXYZ=None switches back to the stored factors (libp_elliptic_set_trilinear)"""
        if EXYZ is None:
            check(L.load().libp_elliptic_set_trilinear(self._h, None, None, None))
            return
        self.keep = self.keep + (EXYZ,)
        z = np.ascontiguousarray(gllz, dtype=np.float64)
        w = np.ascontiguousarray(gllw, dtype=np.float64)
        check(L.load().libp_elliptic_set_trilinear(self._h, _ptr(EXYZ), _ptr(z), _ptr(w)))

    def set_chunk(self, chunk_elements):
        """elements per zero-fill piece of the fused operator (0 = off); see libp_elliptic_set_chunk"""
        check(L.load().libp_elliptic_set_chunk(self._h, int(chunk_elements)))

    def Operator(self, o_q, o_Aq):
        check(L.load().libp_elliptic_operator(self._h, _ptr(o_q), _ptr(o_Aq), _stream()))

    def OperatorTimed(self, o_q, o_Aq):
        """one apply with device timing of its parts: (ms zero-fill, ms exchange + Ax launches + combine)"""
        ms = (C.c_double * 2)()
        check(L.load().libp_elliptic_operator_timed(self._h, _ptr(o_q), _ptr(o_Aq), _stream(), ms))
        return ms[0], ms[1]

    def Free(self):
        if self._h:
            L.load().libp_elliptic_free(self._h)
            self._h = C.c_void_p()


# --------------------------------------------------------------------------- linAlg_t
class EllipticIpdg(Elliptic):
    """The IPDG branch of elliptic_t::Operator (ellipticOperator.cpp:108-160) as an operator_t; every array is the
    reference's (mesh_t::vgeo, sgeo, vmapM, vmapP, elliptic_t::EToB), traceHalo an Ogs of kind HALO set up from
    mesh_t::HaloTraceSetup's ids (None on one rank)."""

    def __init__(self, Nq, Nelements, vmapM, vmapP, vgeo, sgeo, EToB, D, lam, tau, NhaloElementsTotal=0, traceHalo=None,
                 internalElementIds=None, haloElementIds=None):
        self.keep = (vmapM, vmapP, vgeo, sgeo, EToB, D, traceHalo, internalElementIds, haloElementIds)
        d = L.IpdgDesc()
        d.Nq, d.Nelements, d.NhaloElementsTotal = Nq, int(Nelements), int(NhaloElementsTotal)
        d.NinternalElements = 0 if internalElementIds is None else internalElementIds.numel()
        d.NhaloElements = 0 if haloElementIds is None else haloElementIds.numel()
        d.internalElementIds = _ptr(internalElementIds) if d.NinternalElements else None
        d.haloElementIds = _ptr(haloElementIds) if d.NhaloElements else None
        d.vmapM, d.vmapP, d.vgeo, d.sgeo, d.EToB, d.D = (_ptr(a) for a in (vmapM, vmapP, vgeo, sgeo, EToB, D))
        d.lambda_, d.tau = float(lam), float(tau)
        d.traceHalo = traceHalo.handle if traceHalo is not None else None
        self._h = C.c_void_p()
        check(L.load().libp_elliptic_create_ipdg(C.byref(d), C.byref(self._h)))
        Np = Nq ** 3
        self.Ndofs, self.Nhalo = int(Nelements) * Np, int(NhaloElementsTotal) * Np

    def gradient(self):
        """elliptic_t::o_grad after the last apply: tensor [(Nelements + halo elements)*Np, 4] (a view, not a copy)"""
        p = C.c_void_p()
        check(L.load().libp_elliptic_ipdg_gradient(self._h, C.byref(p)))
        n = (self.Ndofs + self.Nhalo) * 4

        class _P:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(p.value), False), "version": 2}
        return torch.as_tensor(_P(), device="cuda").reshape(-1, 4)


class LinAlg:
    """linAlg_t (include/linAlg.hpp:52-120) on torch float64 tensors."""

    def __init__(self, comm=None):
        self.comm = comm

    def _c(self):
        return self.comm.handle if self.comm is not None else None

    def set(self, N, alpha, a): check(L.load().libp_linalg_set(N, alpha, _ptr(a), _stream()))
    def add(self, N, alpha, a): check(L.load().libp_linalg_add(N, alpha, _ptr(a), _stream()))
    def scale(self, N, alpha, a): check(L.load().libp_linalg_scale(N, alpha, _ptr(a), _stream()))
    def axpy(self, N, alpha, x, beta, y): check(L.load().libp_linalg_axpy(N, alpha, _ptr(x), beta, _ptr(y), _stream()))
    def zaxpy(self, N, alpha, x, beta, y, z): check(L.load().libp_linalg_zaxpy(N, alpha, _ptr(x), beta, _ptr(y), _ptr(z), _stream()))
    def amx(self, N, alpha, a, x): check(L.load().libp_linalg_amx(N, alpha, _ptr(a), _ptr(x), _stream()))
    def amxpy(self, N, alpha, a, x, beta, y): check(L.load().libp_linalg_amxpy(N, alpha, _ptr(a), _ptr(x), beta, _ptr(y), _stream()))
    def zamxpy(self, N, alpha, a, x, beta, y, z): check(L.load().libp_linalg_zamxpy(N, alpha, _ptr(a), _ptr(x), beta, _ptr(y), _ptr(z), _stream()))
    def adx(self, N, alpha, a, x): check(L.load().libp_linalg_adx(N, alpha, _ptr(a), _ptr(x), _stream()))
    def adxpy(self, N, alpha, a, x, beta, y): check(L.load().libp_linalg_adxpy(N, alpha, _ptr(a), _ptr(x), beta, _ptr(y), _stream()))
    def zadxpy(self, N, alpha, a, x, beta, y, z): check(L.load().libp_linalg_zadxpy(N, alpha, _ptr(a), _ptr(x), beta, _ptr(y), _ptr(z), _stream()))

    def _red(self, fn, N, *ptrs):
        out = C.c_double()
        check(fn(N, *[_ptr(p) for p in ptrs], self._c(), _stream(), C.byref(out)))
        return out.value

    def min(self, N, a): return self._red(L.load().libp_linalg_min, N, a)
    def max(self, N, a): return self._red(L.load().libp_linalg_max, N, a)
    def sum(self, N, a): return self._red(L.load().libp_linalg_sum, N, a)
    def norm2(self, N, a): return self._red(L.load().libp_linalg_norm2, N, a)
    def innerProd(self, N, x, y): return self._red(L.load().libp_linalg_inner_prod, N, x, y)
    def weightedNorm2(self, N, w, a): return self._red(L.load().libp_linalg_weighted_norm2, N, w, a)
    def weightedInnerProd(self, N, w, x, y): return self._red(L.load().libp_linalg_weighted_inner_prod, N, w, x, y)


# --------------------------------------------------------------------------- precon_t / pcg
class Precon:
    def __init__(self, handle, keep=None):
        self._h, self.keep = handle, keep

    @classmethod
    def Identity(cls, N):
        h = C.c_void_p()
        check(L.load().libp_precon_identity_create(int(N), C.byref(h)))
        return cls(h)

    @classmethod
    def Jacobi(cls, Ndofs, invDiagA, allNeumann=False, NglobalDofs=0, comm=None):
        h = C.c_void_p()
        check(L.load().libp_precon_jacobi_create(int(Ndofs), _ptr(invDiagA), int(allNeumann), int(NglobalDofs),
                                                 comm.handle if comm is not None else None, C.byref(h)))
        return cls(h)

    @classmethod
    def MultiGrid(cls, mg, allNeumann=False, NglobalDofs=0, comm=None):
        """MultiGridPrecon (solvers/elliptic/src/ellipticPreconMultiGrid.cpp:29-37)"""
        h = C.c_void_p()
        check(L.load().libp_precon_multigrid_create(mg.handle, int(allNeumann), int(NglobalDofs),
                                                    comm.handle if comm is not None else None, C.byref(h)))
        return cls(h, keep=mg)

    @property
    def handle(self):
        return self._h

    def Operator(self, o_r, o_Mr):
        check(L.load().libp_precon_apply(self._h, _ptr(o_r), _ptr(o_Mr), _stream()))

    def Free(self):
        if self._h:
            L.load().libp_precon_free(self._h)
            self._h = C.c_void_p()


# --------------------------------------------------------------------------- multigrid (apply path)
class MGLevel:
    """MGLevel (solvers/elliptic/ellipticPrecon.hpp:66-116): matrix-free p-multigrid level."""
    JACOBI, CHEBYSHEV = 1, 2

    def __init__(self, fine: Elliptic, coarse: Elliptic, NqF, NqC, P, invDiagA, weightG, smoother, lambda0, lambda1,
                 ChebyshevIterations=2):
        self.keep = (fine, coarse, P, invDiagA, weightG)
        d = L.MGLevelDesc()
        d.fine, d.coarse, d.NqF, d.NqC = fine.handle, coarse.handle, NqF, NqC
        d.P, d.invDiagA, d.weightG = _ptr(P), _ptr(invDiagA), _ptr(weightG)
        d.smoother, d.lambda0, d.lambda1, d.ChebyshevIterations = smoother, float(lambda0), float(lambda1), ChebyshevIterations
        self._h = C.c_void_p()
        check(L.load().libp_mglevel_create(C.byref(d), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def Operator(self, o_x, o_Ax): check(L.load().libp_mglevel_operator(self._h, _ptr(o_x), _ptr(o_Ax), _stream()))
    def smooth(self, o_rhs, o_x, x_is_zero): check(L.load().libp_mglevel_smooth(self._h, _ptr(o_rhs), _ptr(o_x), int(x_is_zero), _stream()))
    def residual(self, o_rhs, o_x, o_res): check(L.load().libp_mglevel_residual(self._h, _ptr(o_rhs), _ptr(o_x), _ptr(o_res), _stream()))
    def coarsen(self, o_x, o_Rx): check(L.load().libp_mglevel_coarsen(self._h, _ptr(o_x), _ptr(o_Rx), _stream()))
    def prolongate(self, o_xC, o_x): check(L.load().libp_mglevel_prolongate(self._h, _ptr(o_xC), _ptr(o_x), _stream()))


class Csr:
    """parCSR local block (include/parAlmond/parAlmondparCSR.hpp): host numpy arrays in, device copy inside."""

    def __init__(self, Nrows, Ncols, rowStarts, cols, vals, offd_nnz=0):
        rowStarts = np.ascontiguousarray(rowStarts, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.Nrows, self.Ncols = int(Nrows), int(Ncols)
        self._h = C.c_void_p()
        check(L.load().libp_csr_create(int(Nrows), int(Ncols), int(vals.size), _ptr(rowStarts), _ptr(cols), _ptr(vals),
                                       int(offd_nnz), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def SpMV(self, alpha, o_x, beta, o_y, o_z=None):
        check(L.load().libp_csr_spmv(self._h, float(alpha), _ptr(o_x), float(beta), _ptr(o_y),
                                     _ptr(o_y if o_z is None else o_z), _stream()))


class ParCsr(Csr):
    """Distributed parCSR row block (libp_parcsr_create): `d` is the dict amg_setup.split_rows returns."""

    def __init__(self, comm, d):
        self.keep = d
        self.Nrows, self.Ncols = int(d["Nrows"]), int(d["NlocalCols"]) + int(d["offd_colIds"].size)
        desc = L.ParCsrDesc(int(d["Nrows"]), int(d["NlocalCols"]), int(d["diag_vals"].size), _ptr(d["diag_rowStarts"]),
                            _ptr(d["diag_cols"]), _ptr(d["diag_vals"]), int(d["offd_vals"].size), int(d["offd_rows"].size),
                            _ptr(d["offd_rows"]), _ptr(d["offd_mRowStarts"]), _ptr(d["offd_cols"]), _ptr(d["offd_vals"]),
                            int(d["offd_colIds"].size), _ptr(d["offd_colIds"]), _ptr(d["globalColStarts"]))
        self._h = C.c_void_p()
        check(L.load().libp_parcsr_create(comm.handle, C.byref(desc), C.byref(self._h)))


class AmgLevel:
    """parAlmond::amgLevel apply (libs/parAlmond/parAlmondAMGLevel.cpp:48-84)."""
    DAMPED_JACOBI, CHEBYSHEV = 0, 1

    def __init__(self, A: Csr, P: Csr | None, R: Csr | None, diagInv, smoother, lam, lambda0, lambda1, ChebyshevIterations=2):
        self.keep = (A, P, R)
        diagInv = np.ascontiguousarray(diagInv, dtype=np.float64)
        self._h = C.c_void_p()
        check(L.load().libp_amglevel_create(A.handle, P.handle if P else None, R.handle if R else None, _ptr(diagInv),
                                            int(smoother), float(lam), float(lambda0), float(lambda1),
                                            int(ChebyshevIterations), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def smooth(self, o_rhs, o_x, x_is_zero): check(L.load().libp_amglevel_smooth(self._h, _ptr(o_rhs), _ptr(o_x), int(x_is_zero), _stream()))
    def residual(self, o_rhs, o_x, o_res): check(L.load().libp_amglevel_residual(self._h, _ptr(o_rhs), _ptr(o_x), _ptr(o_res), _stream()))
    def coarsen(self, o_x, o_Rx): check(L.load().libp_amglevel_coarsen(self._h, _ptr(o_x), _ptr(o_Rx), _stream()))
    def prolongate(self, o_xC, o_x): check(L.load().libp_amglevel_prolongate(self._h, _ptr(o_xC), _ptr(o_x), _stream()))


class CoarseExact:
    """parAlmond::exactSolver_t::solve (libs/parAlmond/parAlmondCoarseExact.cpp:35-73), single rank."""

    def __init__(self, N, diagInvAT):
        a = np.ascontiguousarray(diagInvAT, dtype=np.float64)
        assert a.size == N * N
        self._h = C.c_void_p()
        check(L.load().libp_coarse_exact_create(int(N), _ptr(a), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def solve(self, o_rhs, o_x): check(L.load().libp_coarse_solve(self._h, _ptr(o_rhs), _ptr(o_x), _stream()))


class CoarseExactPar(CoarseExact):
    """exactSolver_t on P ranks (parAlmondCoarseExact.cpp:80-196): rank r owns rows [offsets[r], offsets[r+1])."""

    def __init__(self, comm, N, coarseOffsets, diagInvAT, offdInvAT):
        off = np.ascontiguousarray(coarseOffsets, dtype=np.int64)
        a = np.ascontiguousarray(diagInvAT, dtype=np.float64)
        b = np.ascontiguousarray(offdInvAT, dtype=np.float64)
        assert a.size == N * N and b.size == N * (int(off[-1]) - N)
        self._h = C.c_void_p()
        check(L.load().libp_coarse_exact_create_par(comm.handle, int(N), _ptr(off), _ptr(a), _ptr(b), C.byref(self._h)))


class Multigrid:
    """parAlmond::multigrid_t V-cycle (libs/parAlmond/parAlmondVcycle.cpp:34-60)."""

    def __init__(self, comm=None):
        self._h = C.c_void_p()
        self.keep = []
        check(L.load().libp_multigrid_create(comm.handle if comm is not None else None, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def AddLevel(self, level):
        self.keep.append(level)
        if isinstance(level, MGLevel):
            check(L.load().libp_multigrid_add_mglevel(self._h, level.handle))
        else:
            check(L.load().libp_multigrid_add_amglevel(self._h, level.handle))

    def SetCoarse(self, coarse):
        self.keep.append(coarse)
        check(L.load().libp_multigrid_set_coarse(self._h, coarse.handle))

    def SetCycle(self, cycle="VCYCLE"):
        """PARALMOND CYCLE: VCYCLE | KCYCLE | NONSYM (K-cycle with the GMRES-type inner products)"""
        check(L.load().libp_multigrid_set_cycle(self._h, 0 if cycle == "VCYCLE" else 1, 1 if cycle == "NONSYM" else 0))

    def Operator(self, o_rhs, o_x):
        check(L.load().libp_multigrid_cycle(self._h, _ptr(o_rhs), _ptr(o_x), _stream()))

    def Vcycle(self, o_rhs, o_x):
        check(L.load().libp_multigrid_vcycle(self._h, _ptr(o_rhs), _ptr(o_x), _stream()))


class Pcg:
    """LinearSolver::pcg.  Solve() takes native handles; SolveCallbacks() takes python callables
    (in, out) -> None working on raw device pointers wrapped as tensors by the caller."""

    def __init__(self, N, Nhalo, comm=None, flexible=False, stopping="ABS/REL-INITRESID"):
        self._h = C.c_void_p()
        self.N, self.Nhalo = N, Nhalo
        st = 0 if stopping == "ABS/REL-INITRESID" else 1
        check(L.load().libp_pcg_create(int(N), int(Nhalo), int(flexible), st, comm.handle if comm is not None else None,
                                       C.byref(self._h)))

    def Solve(self, A: Elliptic, M: Precon, o_x, o_r, tol=1e-8, maxit=5000, verbose=0):
        it = C.c_int()
        check(L.load().libp_pcg_solve(self._h, A.handle, M.handle, _ptr(o_x), _ptr(o_r), float(tol), int(maxit),
                                      int(verbose), _stream(), C.byref(it)))
        return it.value

    def SolveCallbacks(self, A_fn, M_fn, o_x, o_r, tol=1e-8, maxit=5000, verbose=0):
        def wrap(fn):
            def cb(ctx, pin, pout, stream):
                try:
                    fn(pin, pout)
                    return 0
                except Exception as e:  # pragma: no cover
                    print("operator callback failed:", e)
                    return -1
            return L.OPERATOR_FN(cb)
        a, m = wrap(A_fn), wrap(M_fn)
        it = C.c_int()
        check(L.load().libp_pcg_solve_cb(self._h, a, None, m, None, _ptr(o_x), _ptr(o_r), float(tol), int(maxit),
                                         int(verbose), _stream(), C.byref(it)))
        return it.value

    def residual_history(self):
        p, n = C.c_void_p(), C.c_int()
        check(L.load().libp_pcg_residual_history(self._h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), (n.value,)).copy()

    def Free(self):
        if self._h:
            L.load().libp_pcg_free(self._h)
            self._h = C.c_void_p()


class NbPcg:
    """LinearSolver::nbpcg (libs/linearSolver/linearSolverNBPCG.cpp): non-blocking PCG, native handles."""

    def __init__(self, N, Nhalo, comm=None):
        self._h = C.c_void_p()
        check(L.load().libp_nbpcg_create(int(N), int(Nhalo), comm.handle if comm is not None else None, C.byref(self._h)))

    def Solve(self, A: Elliptic, M: Precon, o_x, o_r, tol=1e-8, maxit=5000, verbose=0):
        it = C.c_int()
        check(L.load().libp_nbpcg_solve(self._h, A.handle, M.handle, _ptr(o_x), _ptr(o_r), float(tol), int(maxit),
                                        int(verbose), _stream(), C.byref(it)))
        return it.value

    def residual_history(self):
        p, n = C.c_void_p(), C.c_int()
        check(L.load().libp_nbpcg_residual_history(self._h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), (n.value,)).copy()

    def Free(self):
        if self._h:
            L.load().libp_nbpcg_free(self._h)
            self._h = C.c_void_p()


class NbFPcg:
    """LinearSolver::nbfpcg (libs/linearSolver/linearSolverNBFPCG.cpp): non-blocking flexible PCG, native handles."""

    def __init__(self, N, Nhalo, comm=None):
        self._h = C.c_void_p()
        check(L.load().libp_nbfpcg_create(int(N), int(Nhalo), comm.handle if comm is not None else None, C.byref(self._h)))

    def Solve(self, A: Elliptic, M: Precon, o_x, o_r, tol=1e-8, maxit=5000, verbose=0):
        it = C.c_int()
        check(L.load().libp_nbfpcg_solve(self._h, A.handle, M.handle, _ptr(o_x), _ptr(o_r), float(tol), int(maxit),
                                         int(verbose), _stream(), C.byref(it)))
        return it.value

    def residual_history(self):
        p, n = C.c_void_p(), C.c_int()
        check(L.load().libp_nbfpcg_residual_history(self._h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), (n.value,)).copy()

    def Free(self):
        if self._h:
            L.load().libp_nbfpcg_free(self._h)
            self._h = C.c_void_p()


class InitialGuess:
    """InitialGuess strategies of linearSolver_t (libs/linearSolver/initialGuess.cpp): NONE, ZERO, CLASSIC, QR, EXTRAP.
    linearSolver_t::Solve calls FormInitialGuess before and Update after every solve (linearSolver.cpp:31-44)."""

    KINDS = {"NONE": 0, "ZERO": 1, "CLASSIC": 2, "QR": 3, "EXTRAP": 4}

    def __init__(self, strategy, N, Nhalo=0, history=0, extrap_degree=0, coeffs_method="MINNORM", comm=None):
        self._h = C.c_void_p()
        self._cb = None
        check(L.load().libp_ig_create(self.KINDS[strategy], int(N), int(Nhalo), int(history), int(extrap_degree),
                                      1 if coeffs_method == "CPQR" else 0, comm.handle if comm is not None else None,
                                      C.byref(self._h)))

    def FormInitialGuess(self, o_x, o_rhs):
        check(L.load().libp_ig_form_initial_guess(self._h, _ptr(o_x), _ptr(o_rhs), _stream()))

    def Update(self, A, o_x, o_rhs):
        """A: Elliptic handle, or a python callable (pin, pout) working on raw device pointers"""
        if isinstance(A, Elliptic):
            def fn(pin, pout, h=A.handle):
                check(L.load().libp_elliptic_operator(h, pin, pout, _stream()))
        else:
            fn = A

        def cb(ctx, pin, pout, stream):
            try:
                fn(pin, pout)
                return 0
            except Exception as e:  # pragma: no cover
                print("operator callback failed:", e)
                return -1
        self._cb = L.OPERATOR_FN(cb)
        check(L.load().libp_ig_update(self._h, self._cb, None, _ptr(o_x), _ptr(o_rhs), _stream()))

    @property
    def dimension(self):
        d = C.c_int()
        check(L.load().libp_ig_dimension(self._h, C.byref(d)))
        return d.value

    def Free(self):
        if self._h:
            L.load().libp_ig_free(self._h)
            self._h = C.c_void_p()


def extrap_coeffs(m, M, coeffs_method="MINNORM"):
    """Extrap::extrapCoeffs (initialGuess.cpp:440-466)"""
    c = (C.c_double * M)()
    check(L.load().libp_ig_extrap_coeffs(int(m), int(M), 1 if coeffs_method == "CPQR" else 0, c))
    return np.array(c[:])
