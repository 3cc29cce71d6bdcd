"""ORACLE (test infrastructure only - never imported by the product path).

numpy restatement of libParanumal's gather-scatter *setup* and host apply, simulating all MPI
ranks inside one process (every collective becomes an explicit concatenation), so multi-rank
maps can be produced without MPI.

Follows (libParanumal 0.5.0):
  ogsBase_t::Setup            libs/ogs/ogsSetup.cpp:69-190
  FindSharedNodes             libs/ogs/ogsSetup.cpp:192-331   (unstable std::sort + rand() owner pick)
  ConstructSharedNodes        libs/ogs/ogsSetup.cpp:333-566
  LocalSignedSetup            libs/ogs/ogsSetup.cpp:569-683
  LocalHaloSetup              libs/ogs/ogsSetup.cpp:784-860
  SetupGlobalToLocalMapping   libs/ogs/ogsSetup.cpp:862-886
  ogsPairwise_t ctor          libs/ogs/ogsPairwise.cpp:194-415
  ogsOperator_t::Gather/Scatter (host)  libs/ogs/ogsOperator.cpp:64-113, 219-262

Third-party behaviour restated because it decides the maps bit-for-bit:
  * glibc rand() (TYPE_3 additive feedback generator, default seed 1) -> class GlibcRand
  * libstdc++ std::sort tie order -> oracle/csrc/oracle.c:oracle_libstdcxx_sort_perm
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

SIGNED, UNSIGNED, HALO = "Signed", "Unsigned", "Halo"

_lib = None


def _oracle_lib():
    global _lib
    if _lib is None:
        from . import build as _b
        _lib = ctypes.CDLL(_b.build())
        _lib.oracle_libstdcxx_sort_perm.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.oracle_libstdcxx_sort_perm.restype = None
    return _lib


def libstdcxx_sort_perm(k1: np.ndarray, k2: np.ndarray | None = None) -> np.ndarray:
    """Permutation produced by libstdc++ std::sort on records arriving in order 0..n-1 with
    strict weak order (k1, then k2 if given)."""
    k1 = np.ascontiguousarray(k1, dtype=np.int64)
    n = k1.shape[0]
    perm = np.arange(n, dtype=np.int64)
    k2p = None
    if k2 is not None:
        k2 = np.ascontiguousarray(k2, dtype=np.int64)
        k2p = k2.ctypes.data
    _oracle_lib().oracle_libstdcxx_sort_perm(n, k1.ctypes.data, k2p, perm.ctypes.data)
    return perm


class GlibcRand:
    """glibc rand() == random() with the default TYPE_3 state (r[i] = r[i-3] + r[i-31])."""

    def __init__(self, seed: int = 1):
        r = [0] * 34
        r[0] = seed if seed != 0 else 1
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            if w < 0:
                w += 2147483647
            r[i] = w
        for i in range(31, 34):
            r[i] = r[i - 31]
        self._r = r
        for _ in range(310):
            self._next_raw()

    def _next_raw(self):
        r = self._r
        v = (r[-31] + r[-3]) & 0xFFFFFFFF
        r.append(v)
        if len(r) > 64:
            del r[: len(r) - 34]
        return v

    def rand(self) -> int:
        return self._next_raw() >> 1

    def rand_many(self, n: int) -> np.ndarray:
        return np.fromiter((self.rand() for _ in range(n)), dtype=np.int64, count=n)


NODE_DT = np.dtype([("localId", np.int32), ("baseId", np.int64), ("newId", np.int32),
                    ("sign", np.int32), ("rank", np.int32), ("destRank", np.int32)])


@dataclass
class OgsOperator:
    Ncols: int = 0
    NrowsN: int = 0
    NrowsT: int = 0
    rowStartsN: np.ndarray = None
    rowStartsT: np.ndarray = None
    colIdsN: np.ndarray = None
    colIdsT: np.ndarray = None

    @property
    def nnzN(self):
        return int(self.rowStartsN[-1])

    @property
    def nnzT(self):
        return int(self.rowStartsT[-1])


@dataclass
class PairwiseExchange:
    Nhalo: int = 0
    NhaloP: int = 0
    sendIdsN: np.ndarray = None
    sendIdsT: np.ndarray = None
    # dense per-rank counts/offsets (length size / size+1)
    mpiSendCountsN: np.ndarray = None
    mpiSendCountsT: np.ndarray = None
    mpiRecvCountsN: np.ndarray = None
    mpiRecvCountsT: np.ndarray = None
    postmpi: OgsOperator = None


@dataclass
class OgsRank:
    N: int = 0
    kind: str = SIGNED
    unique: bool = False
    Ngather: int = 0
    NlocalT: int = 0
    NlocalP: int = 0
    NhaloT: int = 0
    NhaloP: int = 0
    NgatherGlobal: int = 0
    gather_defined: bool = True
    ids: np.ndarray = None  # (possibly sign-rewritten) ids
    gatherLocal: OgsOperator = None
    gatherHalo: OgsOperator = None
    exchange: PairwiseExchange = None
    extra: dict = field(default_factory=dict)

    @property
    def Nhalo(self):
        return self.NhaloT - self.NhaloP

    def global_to_local(self) -> np.ndarray:
        """ogs_t::SetupGlobalToLocalMapping."""
        g2l = np.full(self.N, -1, dtype=np.int32)
        for op, off in ((self.gatherLocal, 0), (self.gatherHalo, self.NlocalT)):
            if op is None or op.NrowsT == 0:
                continue
            rows = np.repeat(np.arange(op.NrowsT, dtype=np.int32), np.diff(op.rowStartsT))
            g2l[op.colIdsT] = rows + off
        return g2l


def _csr_from(gid, mask, lid, nrows):
    """rowStarts/colIds with columns in ascending traversal order (the fill loops of
    LocalSignedSetup walk nodes in local order)."""
    sel = np.nonzero(mask)[0]
    g = gid[sel]
    counts = np.bincount(g, minlength=nrows).astype(np.int32)
    rs = np.zeros(nrows + 1, dtype=np.int32)
    np.cumsum(counts, out=rs[1:])
    order = np.argsort(g, kind="stable")
    return rs, lid[sel][order].astype(np.int32)


def ogs_setup_all(ids_per_rank, kind=SIGNED, unique=False, rands=None):
    """Run ogsBase_t::Setup for every rank of a simulated communicator.

    ids_per_rank: list of int64 arrays (one per rank).  Returns list[OgsRank].  `rands` is a list
    of per-rank GlibcRand (each MPI rank is its own process with its own rand() state).
    """
    size = len(ids_per_rank)
    if rands is None:
        rands = [GlibcRand() for _ in range(size)]
    assert not ((kind == UNSIGNED and unique) or (kind == HALO and unique))
    out = [OgsRank(N=len(ids), kind=kind, unique=unique, ids=np.array(ids, dtype=np.int64)) for ids in ids_per_rank]

    # ---- node lists (zero ids squeezed out), in send order (grouped by destRank, local order inside)
    nodes = []
    nzmap = []
    for r in range(size):
        ids = out[r].ids
        nz = np.nonzero(ids)[0]
        nzmap.append(nz)
        nd = np.zeros(len(nz), dtype=NODE_DT)
        nd["localId"] = np.arange(len(nz))
        nd["baseId"] = np.abs(ids[nz]) if kind == UNSIGNED else ids[nz]
        nd["rank"] = r
        nd["destRank"] = np.abs(ids[nz]) % size
        nd = nd[np.argsort(nd["destRank"], kind="stable")]
        nodes.append(nd)

    # ---- FindSharedNodes
    send_slices = [[nodes[s][nodes[s]["destRank"] == d] for d in range(size)] for s in range(size)]
    returned = [[None] * size for _ in range(size)]  # returned[s][d]
    for d in range(size):
        recv = np.concatenate([send_slices[s][d] for s in range(size)]) if size else np.zeros(0, NODE_DT)
        counts = [len(send_slices[s][d]) for s in range(size)]
        n = len(recv)
        absid = np.abs(recv["baseId"])
        perm = libstdcxx_sort_perm(absid)
        srt = recv[perm]
        sabs = absid[perm]
        if n:
            brk = np.nonzero(np.diff(sabs))[0] + 1
            starts = np.concatenate([[0], brk])
            ends = np.concatenate([brk, [n]])
        else:
            starts = ends = np.zeros(0, dtype=np.int64)
        gsz = ends - starts
        if unique:
            m = rands[d].rand_many(len(starts)) % gsz
            srt["baseId"] = -sabs
            pick = starts + m
            srt["baseId"][pick] = sabs[pick]
        else:
            pos = np.add.reduceat((srt["baseId"] > 0).astype(np.int64), starts) if n else np.zeros(0)
            if np.any(pos != 1):
                for o in out:
                    o.gather_defined = False
            assert not (kind == HALO and np.any(pos != 1)), "halo needs exactly one positive id per group"
        if n:
            rmin = np.minimum.reduceat(srt["rank"], starts)
            rmax = np.maximum.reduceat(srt["rank"], starts)
            shared = np.where(rmin != rmax, 2, 1).astype(np.int32)
            srt["sign"] = np.repeat(shared, gsz)
        back = np.empty_like(srt)
        back[perm] = srt  # undo the sort: back to recv'd ordering
        off = 0
        for s in range(size):
            returned[s][d] = back[off:off + counts[s]]
            off += counts[s]
    for s in range(size):
        nodes[s] = np.concatenate(returned[s]) if size else nodes[s]

    # ---- ConstructSharedNodes (per rank part)
    send_shared = []
    for r in range(size):
        nd = nodes[r]
        n = len(nd)
        absid = np.abs(nd["baseId"])
        order = np.lexsort((-nd["baseId"], absid))  # group by |id|, positive first (ties immaterial)
        nd = nd[order]
        absid = absid[order]
        if n:
            brk = np.nonzero(np.diff(absid))[0] + 1
            starts = np.concatenate([[0], brk])
            gsz = np.diff(np.concatenate([starts, [n]]))
        else:
            starts = np.zeros(0, dtype=np.int64)
            gsz = starts
        lead = nd[starts] if n else nd[:0]
        sign = np.abs(lead["sign"])
        sign = np.where(lead["baseId"] < 0, -sign, sign).astype(np.int32)
        nd["sign"] = np.repeat(sign, gsz)
        o = out[r]
        o.NlocalT = int(np.sum(np.abs(sign) == 1))
        o.NlocalP = int(np.sum(sign == 1))
        o.NhaloT = int(np.sum(np.abs(sign) == 2))
        o.NhaloP = int(np.sum(sign == 2))
        o.Ngather = o.NlocalP + o.NhaloP
        nd["newId"] = np.repeat(np.arange(len(starts), dtype=np.int32), gsz)
        lead = nd[starts] if n else nd[:0]
        ss = lead[np.abs(lead["sign"]) == 2].copy()
        # back to (compressed) local order, renumber groups by first appearance
        nd = nd[np.argsort(nd["localId"], kind="stable")]
        newId = nd["newId"]
        first_idx = np.full(len(starts), n, dtype=np.int64)
        np.minimum.at(first_idx, newId, np.arange(n))
        gsign = np.zeros(len(starts), dtype=np.int32)
        gsign[newId] = nd["sign"]
        indexMap = np.full(len(starts), -1, dtype=np.int32)
        for sg, base in ((1, 0), (-1, o.NlocalP), (2, 0), (-2, o.NhaloP)):
            grp = np.nonzero(gsign == sg)[0]
            grp = grp[np.argsort(first_idx[grp], kind="stable")]
            indexMap[grp] = base + np.arange(len(grp), dtype=np.int32)
        nd["newId"] = indexMap[newId]
        ss["localId"] = indexMap[ss["newId"]]
        nodes[r] = nd
        send_shared.append(ss)
    NgG = sum(o.Ngather for o in out)
    for o in out:
        o.NgatherGlobal = NgG

    # shared-node rendezvous: leading node of each shared group goes to destRank, which sends the
    # full participant list back to every participant
    shared_nodes = [[] for _ in range(size)]
    for d in range(size):
        recv = np.concatenate([send_shared[s][send_shared[s]["destRank"] == d] for s in range(size)])
        if len(recv) == 0:
            continue
        absid = np.abs(recv["baseId"])
        order = np.argsort(absid, kind="stable")  # tie order immaterial (see module docstring of tests)
        recv = recv[order]
        absid = absid[order]
        brk = np.nonzero(np.diff(absid))[0] + 1
        starts = np.concatenate([[0], brk])
        ends = np.concatenate([brk, [len(recv)]])
        for s0, e0 in zip(starts, ends):
            grp = recv[s0:e0]
            for i in range(len(grp)):
                others = np.delete(grp, i).copy()
                others["newId"] = grp["localId"][i]
                others["sign"] = grp["sign"][i]
                shared_nodes[int(grp["rank"][i])].append(others)
    shared_nodes = [np.concatenate(s) if s else np.zeros(0, dtype=NODE_DT) for s in shared_nodes]

    # ---- write real local ids + signed ids back, local operators
    for r in range(size):
        o = out[r]
        nd = nodes[r]
        nd["localId"] = nzmap[r]
        if unique:
            o.ids[nzmap[r]] = nd["baseId"]
        gid = nd["newId"].astype(np.int64)
        lid = nd["localId"]
        is_local = np.abs(nd["sign"]) == 1
        pos = nd["baseId"] > 0
        gl = OgsOperator(Ncols=o.N, NrowsN=o.NlocalP, NrowsT=o.NlocalT)
        gh = OgsOperator(Ncols=o.N, NrowsN=o.NhaloP, NrowsT=o.NhaloT)
        if kind == SIGNED:
            gl.rowStartsN, gl.colIdsN = _csr_from(gid, is_local & pos, lid, o.NlocalT)
            gl.rowStartsT, gl.colIdsT = _csr_from(gid, is_local, lid, o.NlocalT)
            gh.rowStartsN, gh.colIdsN = _csr_from(gid, ~is_local & pos, lid, o.NhaloT)
            gh.rowStartsT, gh.colIdsT = _csr_from(gid, ~is_local, lid, o.NhaloT)
        elif kind == UNSIGNED:
            gl.rowStartsT, gl.colIdsT = _csr_from(gid, is_local, lid, o.NlocalT)
            gl.rowStartsN, gl.colIdsN = gl.rowStartsT, gl.colIdsT
            gh.rowStartsT, gh.colIdsT = _csr_from(gid, ~is_local, lid, o.NhaloT)
            gh.rowStartsN, gh.colIdsN = gh.rowStartsT, gh.colIdsT
        else:  # HALO: only shared nodes; N-map = nodes whose group sign is +2
            gl = None
            gh.rowStartsN, gh.colIdsN = _csr_from(gid, nd["sign"] == 2, lid, o.NhaloT)
            gh.rowStartsT, gh.colIdsT = _csr_from(gid, ~is_local, lid, o.NhaloT)
        o.gatherLocal, o.gatherHalo = gl, gh

    # ---- pairwise exchange
    sorted_shared = []
    for r in range(size):
        sn = shared_nodes[r]
        sn = sn[np.lexsort((sn["newId"], sn["rank"]))]
        sorted_shared.append(sn)
    for r in range(size):
        o = out[r]
        sn = sorted_shared[r]
        ex = PairwiseExchange(Nhalo=o.NhaloT, NhaloP=o.NhaloP)
        ex.mpiSendCountsT = np.bincount(sn["rank"], minlength=size).astype(np.int32)
        ex.mpiSendCountsN = np.bincount(sn["rank"][sn["sign"] > 0], minlength=size).astype(np.int32)
        ex.sendIdsN = sn["newId"][sn["sign"] == 2].astype(np.int32)
        ex.sendIdsT = sn["newId"].astype(np.int32)
        o.exchange = ex
    for r in range(size):
        o = out[r]
        ex = o.exchange
        # what rank r receives: from each source s (ascending), s's list for rank r in s's order
        recv = np.concatenate([sorted_shared[s][sorted_shared[s]["rank"] == r] for s in range(size)])
        src = np.concatenate([np.full(int(np.sum(sorted_shared[s]["rank"] == r)), s, dtype=np.int32) for s in range(size)])
        ex.mpiRecvCountsT = np.bincount(src, minlength=size).astype(np.int32)
        ex.mpiRecvCountsN = np.array([int(np.sum((sorted_shared[s]["rank"] == r) & (sorted_shared[s]["sign"] > 0)))
                                      for s in range(size)], dtype=np.int32)
        Nhalo, NhaloP = ex.Nhalo, ex.NhaloP
        rid = recv["localId"].astype(np.int64)
        rpos = recv["sign"] == 2
        pm = OgsOperator(Ncols=Nhalo + len(recv), NrowsN=Nhalo, NrowsT=Nhalo)
        # own value first (column n), then received copies in arrival order
        colsT_own = np.arange(Nhalo)
        colsT_rcv = np.arange(len(recv)) + Nhalo
        gidT = np.concatenate([colsT_own, rid])
        colT = np.concatenate([colsT_own, colsT_rcv])
        pm.rowStartsT, pm.colIdsT = _csr_from(gidT, np.ones(len(gidT), bool), colT, Nhalo)
        colsN_own = np.arange(NhaloP)
        colsN_rcv = Nhalo + np.arange(int(np.sum(rpos)))
        gidN = np.concatenate([colsN_own, rid[rpos]])
        colN = np.concatenate([colsN_own, colsN_rcv])
        pm.rowStartsN, pm.colIdsN = _csr_from(gidN, np.ones(len(gidN), bool), colN, Nhalo)
        ex.postmpi = pm
    return out


# ----------------------------------------------------------------------------- host apply
_OPS = {"Add": (np.add, 0), "Mul": (np.multiply, 1), "Max": (np.maximum, None), "Min": (np.minimum, None)}


def op_gather(op: OgsOperator, v: np.ndarray, trans: str, opname="Add", K=1) -> np.ndarray:
    """ogsOperator_t::Gather host path: sequential left-to-right reduction per row."""
    if trans == "NoTrans":
        nrows, rs, ci = op.NrowsN, op.rowStartsN, op.colIdsN
    else:
        nrows, rs, ci = op.NrowsT, op.rowStartsT, op.colIdsT
    v = np.asarray(v).reshape(-1, K)
    gv = np.zeros((nrows, K), dtype=v.dtype)
    for n in range(nrows):
        s, e = rs[n], rs[n + 1]
        if opname == "Add":
            val = np.zeros(K, dtype=v.dtype)
            for g in range(s, e):
                val = val + v[ci[g]]
        elif opname == "Mul":
            val = np.ones(K, dtype=v.dtype)
            for g in range(s, e):
                val = val * v[ci[g]]
        elif opname == "Max":
            val = np.full(K, np.finfo(v.dtype).min if v.dtype.kind == "f" else np.iinfo(v.dtype).min, dtype=v.dtype)
            # reference uses -max() for Max init (numeric_limits<T>::max() negated); equivalent for non-empty rows
            for g in range(s, e):
                val = np.maximum(val, v[ci[g]])
        else:
            val = np.full(K, np.finfo(v.dtype).max if v.dtype.kind == "f" else np.iinfo(v.dtype).max, dtype=v.dtype)
            for g in range(s, e):
                val = np.minimum(val, v[ci[g]])
        gv[n] = val
    return gv.reshape(-1) if K == 1 else gv.reshape(-1)


def op_gather_add_fast(op: OgsOperator, v: np.ndarray, trans: str) -> np.ndarray:
    """Vectorised Add gather (same left-to-right order per row: np.add.reduceat sums in order)."""
    if trans == "NoTrans":
        nrows, rs, ci = op.NrowsN, op.rowStartsN, op.colIdsN
    else:
        nrows, rs, ci = op.NrowsT, op.rowStartsT, op.colIdsT
    gv = np.zeros(nrows, dtype=v.dtype)
    cnt = np.diff(rs[: nrows + 1])
    maxc = int(cnt.max()) if nrows else 0
    for c in range(maxc):
        sel = np.nonzero(cnt > c)[0]
        gv[sel] = gv[sel] + v[ci[rs[sel] + c]]
    return gv


def op_scatter(op: OgsOperator, gv: np.ndarray, v: np.ndarray, trans: str, K=1):
    """ogsOperator_t::Scatter host path (Trans -> N maps, else T maps)."""
    if trans == "Trans":
        nrows, rs, ci = op.NrowsN, op.rowStartsN, op.colIdsN
    else:
        nrows, rs, ci = op.NrowsT, op.rowStartsT, op.colIdsT
    rows = np.repeat(np.arange(nrows), np.diff(rs[: nrows + 1]))
    if K == 1:
        v[ci[: len(rows)]] = gv[rows]
    else:
        v.reshape(-1, K)[ci[: len(rows)]] = gv.reshape(-1, K)[rows]
    return v
