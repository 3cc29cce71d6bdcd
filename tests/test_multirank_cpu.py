"""world_size>1 checks on CPU (gloo): the C-ABI ogs setup driven by torch.distributed host collectives
must reproduce the oracle's simulated multi-rank setup (same algorithm as the reference's MPI path)
map for map: counters, signed ids, local/halo CSR maps, pairwise send lists and post-exchange combine."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.mesh_box import build_box_hex_mesh, masked_global_ids
from oracle.ogs_ref import SIGNED, ogs_setup_all


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, N, n, flag, outq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from libparanumal_b200 import _lib as L
        from libparanumal_b200.api import Comm, Ogs
        from libparanumal_b200.box_mesh import BoxMesh
        ctypes.CDLL("libc.so.6").srand(1)  # every MPI rank of the reference is a fresh process
        box = (n, n, n) if np.isscalar(n) else tuple(n)
        mesh = BoxMesh(N, box[0], box[1], box[2], rank, size, flag, geometry=False)
        _, ids = mesh.masked_global_ids()
        ids = ids.numpy().copy()
        comm = Comm(rank, size)
        ogs = Ogs().Setup(ids.size, ids, comm, kind=L.SIGNED, unique=True)
        res = dict(rank=rank, ids=ids, counts=[ogs.N, ogs.Ngather, ogs.NlocalT, ogs.NlocalP, ogs.NhaloT, ogs.NhaloP,
                                                ogs.NgatherGlobal],
                   local=ogs.maps("local"), halo=ogs.maps("halo"), postmpi=ogs.maps("postmpi"),
                   exN=ogs.exchange_lists(L.NOTRANS), exT=ogs.exchange_lists(L.TRANS),
                   g2l=ogs.SetupGlobalToLocalMapping(),
                   lists=(mesh.localGatherElementList.numpy(), mesh.globalGatherElementList.numpy()))
        outq.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,N,n,flag", [(2, 3, 4, 1), (4, 2, 4, 1), (8, 2, 4, -1), (3, 2, 5, 1)])
def test_multirank_ogs_setup_matches_oracle(size, N, n, flag):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, N, n, flag, q)) for r in range(size)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(size):
        r = q.get(timeout=120)
        results[r["rank"]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    ids = []
    for r in range(size):
        m = build_box_hex_mesh(N, n, n, n, r, size, flag, geometry=False)
        ids.append(masked_global_ids(m)[1])
    ref = ogs_setup_all(ids, SIGNED, True)
    total_halo = 0
    for r in range(size):
        got, o = results[r], ref[r]
        assert got["counts"] == [o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.NgatherGlobal]
        assert np.array_equal(got["ids"], o.ids)
        for key, op in (("local", o.gatherLocal), ("halo", o.gatherHalo), ("postmpi", o.exchange.postmpi)):
            for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
                assert np.array_equal(got[key][nm], getattr(op, nm)), (r, key, nm)
        assert np.array_equal(got["g2l"], o.global_to_local())
        ex = o.exchange
        assert np.array_equal(got["exN"]["sendIds"], ex.sendIdsN)
        assert np.array_equal(got["exT"]["sendIds"], ex.sendIdsT)
        for key, sc, rc in (("exN", ex.mpiSendCountsN, ex.mpiRecvCountsN), ("exT", ex.mpiSendCountsT, ex.mpiRecvCountsT)):
            assert np.array_equal(got[key]["sendRanks"], np.nonzero(sc)[0])
            assert np.array_equal(got[key]["sendCounts"], sc[sc > 0])
            assert np.array_equal(got[key]["recvRanks"], np.nonzero(rc)[0])
            assert np.array_equal(got[key]["recvCounts"], rc[rc > 0])
        total_halo += o.NhaloT
    assert total_halo > 0  # the partition really shares nodes
    # global invariants: every unmasked global node owned exactly once
    nglobal = len(np.unique(np.abs(np.concatenate(ids))[np.concatenate(ids) != 0]))
    assert ref[0].NgatherGlobal == nglobal


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MR_NAMES = ["mr_n3_e4x4x4_p2", "mr_n3_e4x4x4_p4", "mr_n3_e4x4x4_p8", "mr_n2_e5x4x3_p2", "mr_n2_e5x4x3_p4", "mr_n2_e5x4x3_p8",
            "mr_n7_e2x2x2_p8", "mr_n7_e4x2x2_p2", "mr_n2_e4x4x4_periodic_p4"]


@pytest.mark.parametrize("name", MR_NAMES)
def test_multirank_ogs_setup_matches_multirank_reference(name):
    """The C-ABI host setup on P processes (gloo host collectives standing in for MPI) against dumps of the UNMODIFIED
    reference run on P ranks (tests/golden/mr_*.npz, oracle/refbuild/make_golden_mr.py): box decomposition, owner
    choice, counters, gatherLocal / gatherHalo / postmpi maps, pairwise send lists, neighbour ranks, counts and
    offsets, GlobalToLocal, element lists - all bit-exact (libs/ogs/ogsSetup.cpp:69-886, ogsPairwise.cpp:194-415)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    N, box, flag, size = int(g["config_N"]), [int(x) for x in g["config_box"]], int(g["config_flag"]), int(g["config_P"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, N, box, flag, q)) for r in range(size)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(size):
        r = q.get(timeout=180)
        results[r["rank"]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(size):
        got, k = results[r], f"r{r}_"
        assert got["counts"] == [int(x) for x in g[k + "ogs_counts"][:7]]
        assert np.array_equal(got["ids"], g[k + "maskedGlobalIds"])
        assert np.array_equal(got["g2l"], g[k + "GlobalToLocal"])
        for key, ref in (("local", "gatherLocal"), ("halo", "gatherHalo"), ("postmpi", "pw_postmpi")):
            for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
                assert np.array_equal(got[key][nm], g[k + ref + "_" + nm]), (r, key, nm)
        for key, f in (("exN", "N"), ("exT", "T")):
            assert np.array_equal(got[key]["sendIds"], g[k + "pw_sendIds" + f])
            for nm in ("sendRanks", "sendCounts", "recvRanks", "recvCounts", "sendOffsets", "recvOffsets"):
                assert np.array_equal(got[key][nm], g[k + "pw_" + nm + f]), (r, key, nm)
        assert np.array_equal(got["lists"][0], g[k + "localGatherElementList"])
        assert np.array_equal(got["lists"][1], g[k + "globalGatherElementList"])


def _csr_worker(rank, size, port, outq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import ctypes as C

        import scipy.sparse as sp

        from libparanumal_b200 import _lib as L
        from libparanumal_b200 import amg_setup as am
        from libparanumal_b200.api import Comm, ParCsr
        comm = Comm(rank, size)
        # setup-time host collectives of the harness
        objs = comm.allgather_object({"rank": rank, "v": np.arange(rank + 1)})
        assert [o["rank"] for o in objs] == list(range(size)) and objs[-1]["v"].size == size
        assert comm.allreduce_sum(float(rank + 1)) == size * (size + 1) / 2
        # the same global matrix on every rank (as the replicated AMG setup produces it), split into row blocks
        n = 60
        M = sp.random(n, n, density=0.2, random_state=3, format="csr") + sp.identity(n, format="csr")
        starts = np.linspace(0, n, size + 1).astype(np.int64)
        d = am.split_rows(M.tocsr(), starts, starts, rank)
        A = ParCsr(comm, d)   # host part of libp_parcsr_create: who owns my non-local columns, what do I send
        nrows, nloc, ncols, nsend = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ids, ns, nr = C.c_void_p(), C.c_int(), C.c_int()
        L.check(L.load().libp_csr_info(A.handle, C.byref(nrows), C.byref(nloc), C.byref(ncols), C.byref(nsend), C.byref(ids),
                                       C.byref(ns), C.byref(nr)))
        send = np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_int32)), (nsend.value,)).copy() if nsend.value else np.zeros(0, np.int32)
        outq.put(dict(rank=rank, want=d["offd_colIds"], send=send + int(starts[rank]), shape=(nrows.value, nloc.value, ncols.value)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_multirank_parcsr_halo_plan(size):
    """Distributed CSR levels: the column-halo plan built by libp_parcsr_create through the host collectives sends
    exactly the entries the other ranks' off-rank blocks reference."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_csr_worker, args=(r, size, port, q)) for r in range(size)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(size):
        r = q.get(timeout=120)
        res[r["rank"]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 60
    starts = np.linspace(0, n, size + 1).astype(np.int64)
    for r in range(size):
        nl = int(starts[r + 1] - starts[r])
        assert res[r]["shape"] == (nl, nl, nl + res[r]["want"].size)
        wanted_from_r = np.sort(np.concatenate([w[(w >= starts[r]) & (w < starts[r + 1])]
                                                for k, w in ((k, res[k]["want"]) for k in range(size)) if k != r]))
        assert np.array_equal(np.sort(res[r]["send"]), wanted_from_r)


def _halo_worker(rank, size, port, name, outq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from libparanumal_b200 import _lib as L
        from libparanumal_b200.api import Comm, Ogs
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
        ids = g[f"r{rank}_traceGlobalIds"].astype(np.int64).copy()
        halo = Ogs().Setup(ids.size, ids, Comm(rank, size), kind=L.HALO, unique=False)
        outq.put(dict(rank=rank, ids=ids, counts=[halo.N, halo.NlocalT, halo.NhaloT, halo.NhaloP],
                      halo=halo.maps("halo"), exN=halo.exchange_lists(L.NOTRANS)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ipdg_n3_e4x4x4_p2", "ipdg_n2_e5x4x3_p4"])
def test_trace_halo_setup_from_reference_ids(name):
    """kind = Halo setup (mesh_t::HaloTraceSetup's ids of the P-rank reference run, libs/mesh/meshHaloTraceSetup.cpp):
    every flagged trace node becomes one halo row scattered to exactly that node, every owned node some other rank
    flagged becomes an owned halo row, and the rows the ranks send add up to the rows the others receive"""
    size = int(name.rsplit("_p", 1)[1])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, size, port, name, q)) for r in range(size)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(size)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    needed = {}
    for d in res:
        for v in -d["ids"][d["ids"] < 0]:
            needed[int(v)] = needed.get(int(v), 0) + 1
    for d in res:
        ids = d["ids"]
        N, NlocalT, NhaloT, NhaloP = d["counts"]
        neg = np.nonzero(ids < 0)[0]
        owned_needed = np.nonzero((ids > 0) & np.isin(ids, list(needed)))[0]
        assert NhaloT - NhaloP == neg.size and NhaloP == owned_needed.size
        h = d["halo"]
        assert h["NrowsN"] == NhaloP and h["NrowsT"] == NhaloT
        # owned rows gather from / scatter to the owned node, received rows scatter to the flagged node
        assert sorted(h["colIdsN"].tolist()) == sorted(owned_needed.tolist())
        assert sorted(h["colIdsT"][h["rowStartsT"][NhaloP]:].tolist()) == sorted(neg.tolist())
    sent = sum(int(d["exN"]["sendCounts"].sum()) for d in res)
    recvd = sum(int(d["exN"]["recvCounts"].sum()) for d in res)
    assert sent == recvd == sum(needed.values())
