"""ORACLE (test infrastructure): the reference's OWN CPU kernels for the hot path, as compiled by the reference's own
toolchain (OCCA OpenMP mode -> g++), loaded with ctypes.

oracle/refbuild/build_ref_kernels.sh runs the unmodified reference once per degree and keeps the JIT-compiled shared
objects under oracle/_ref/kernels/ (binaries only; git-ignored, they travel to the GPU box with the snapshot).  They
depend on libgomp / libstdc++ only, so they run without OCCA, the OKL sources or /root/reference.  Used by
tests/test_ref_kernels_cpu.py (oracle C port == reference kernels) and by bench.py's cpu_baseline / --impl reference
legs (`kind: "reference"`).  Never imported by the product.

Entry points (OCCA passes scalars by const reference):
  ellipticPartialAxHex3D(Nelements&, elementList, GlobalToLocal, wJ, ggeo, DT, S, MM, lambda&, q, Aq)
        solvers/elliptic/okl/ellipticAxHex3D.okl:156-295, called as in solvers/elliptic/src/ellipticOperator.cpp:43-48
  gather(Nblocks&, K&, blockStarts, gatherStarts, gatherIds, q, gatherq)
        libs/ogs/okl/ogsKernels.okl:86-122 (T = double, OGS_OP = +=), called as in libs/ogs/ogsOperator.cpp:175-203
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KDIR = os.path.join(HERE, "_ref", "kernels")
GATHER_NODES_PER_BLOCK = 512  # include/ogs/ogsOperator.hpp:123

_libs = {}


def available(N):
    return os.path.exists(os.path.join(KDIR, f"ellipticAxHex3D_N{N}.so")) and \
        os.path.exists(os.path.join(KDIR, "ogsKernels_double_add.so"))


def _load(name):
    if name not in _libs:
        _libs[name] = C.CDLL(os.path.join(KDIR, name))
    return _libs[name]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_num_threads(n):
    """the kernels are `#pragma omp parallel for` loops: libgomp's global thread count applies"""
    C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))


def max_threads():
    g = C.CDLL("libgomp.so.1")
    g.omp_get_max_threads.restype = C.c_int
    return int(g.omp_get_max_threads())


def row_blocks(rowStarts):
    """ogsOperator_t::setupRowBlocks (libs/ogs/ogsOperator.cpp:495-560): greedy blocks of rows holding at most
    gatherNodesPerBlock entries."""
    sizes = np.diff(rowStarts)
    assert sizes.size == 0 or sizes.max() <= GATHER_NODES_PER_BLOCK
    cs = np.asarray(rowStarts, dtype=np.int64)
    starts = [0]
    n = sizes.size
    i = 0
    while i < n:
        # furthest row j such that entries of rows [i, j) fit in one block
        j = int(np.searchsorted(cs, cs[i] + GATHER_NODES_PER_BLOCK, side="right")) - 1
        j = max(j, i + 1)
        if j >= n:
            break
        starts.append(j)
        i = j
    starts.append(n)
    return np.asarray(starts if n else [0], dtype=np.int32)


class RefOperator:
    """elliptic_t::Operator, continuous branch (ellipticOperator.cpp:31-106), single rank, on the reference's kernels."""

    def __init__(self, Nq, G2L, wJ, ggeo, D, lam, rowStartsT, colIdsT):
        N = Nq - 1
        self.ax = _load(f"ellipticAxHex3D_N{N}.so").ellipticPartialAxHex3D
        self.ax.restype = None
        self.gk = _load("ogsKernels_double_add.so").gather
        self.gk.restype = None
        self.Nq, self.Np = Nq, Nq ** 3
        self.G2L = np.ascontiguousarray(G2L, dtype=np.int32)
        self.wJ = np.ascontiguousarray(wJ, dtype=np.float64).reshape(-1)
        self.ggeo = np.ascontiguousarray(ggeo, dtype=np.float64).reshape(-1)
        self.D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1)
        self.lam = C.c_double(float(lam))
        self.E = self.wJ.size // self.Np
        self.rs = np.ascontiguousarray(rowStartsT, dtype=np.int32)
        self.ci = np.ascontiguousarray(colIdsT, dtype=np.int32)
        self.blocks = row_blocks(self.rs)
        self.Ng = self.rs.size - 1
        self.elems = np.arange(self.E, dtype=np.int32)   # localGatherElementList of a one-rank mesh
        self.AqL = np.zeros(self.E * self.Np)
        self.one = C.c_int(1)

    def _partial(self, start, count, q):
        if count <= 0:
            return
        n = C.c_int(count)
        el = self.elems[start:start + count]
        self.ax(C.byref(n), _p(el), _p(self.G2L), _p(self.wJ), _p(self.ggeo), _p(self.D), _p(self.D), _p(self.D),
                C.byref(self.lam), _p(q), _p(self.AqL))

    def __call__(self, q, out=None):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.empty(self.Ng) if out is None else out
        half = self.E // 2
        self._partial(0, half, q)                     # NlocalGatherElements/2 (no global elements on one rank)
        self._partial(half, self.E - half, q)         # (NlocalGatherElements+1)/2
        nb = C.c_int(self.blocks.size - 1)
        self.gk(C.byref(nb), C.byref(self.one), _p(self.blocks), _p(self.rs), _p(self.ci), _p(self.AqL), _p(out))
        return out


def transfer_available(NqF, NqC):
    return all(os.path.exists(os.path.join(KDIR, f"ellipticPrecon{k}Hex3D_F{NqF}_C{NqC}.so")) for k in ("Coarsen", "Prolongate"))


def coarsen(NqF, NqC, G2L_fine, P, qf_gathered):
    """ellipticPartialPreconCoarsenHex3D (okl/ellipticPreconCoarsenHex3D.okl:209-294): element-local coarse values
    qc[E*NqC^3] = (P^T x P^T x P^T) qf[GlobalToLocal] ; P is [NqF][NqC] row-major."""
    f = _load(f"ellipticPreconCoarsenHex3D_F{NqF}_C{NqC}.so").ellipticPartialPreconCoarsenHex3D
    f.restype = None
    G2L = np.ascontiguousarray(G2L_fine, dtype=np.int32)
    E = G2L.size // NqF ** 3
    el = np.arange(E, dtype=np.int32)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1)
    q = np.ascontiguousarray(qf_gathered, dtype=np.float64)
    out = np.zeros(E * NqC ** 3)
    n = C.c_int(E)
    f(C.byref(n), _p(el), _p(G2L), _p(P), _p(q), _p(out))
    return out


def prolongate(NqF, NqC, G2L_coarse, P, qc_gathered):
    """ellipticPartialPreconProlongateHex3D (okl/ellipticPreconProlongateHex3D.okl:216-301): element-local fine values
    qf[E*NqF^3] = (P x P x P) qc[GlobalToLocal]."""
    f = _load(f"ellipticPreconProlongateHex3D_F{NqF}_C{NqC}.so").ellipticPartialPreconProlongateHex3D
    f.restype = None
    G2L = np.ascontiguousarray(G2L_coarse, dtype=np.int32)
    E = G2L.size // NqC ** 3
    el = np.arange(E, dtype=np.int32)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1)
    q = np.ascontiguousarray(qc_gathered, dtype=np.float64)
    out = np.zeros(E * NqF ** 3)
    n = C.c_int(E)
    f(C.byref(n), _p(el), _p(G2L), _p(P), _p(q), _p(out))
    return out


def ogs_scatter(rowStarts, colIds, gv, nlocal):
    """scatter kernel (libs/ogs/okl/ogsKernels.okl:127-164, T = double): v[colIds[g]] = gv[row]"""
    f = _load("ogsKernels_double_add.so").scatter
    f.restype = None
    rs = np.ascontiguousarray(rowStarts, dtype=np.int32)
    ci = np.ascontiguousarray(colIds, dtype=np.int32)
    blocks = row_blocks(rs)
    gv = np.ascontiguousarray(gv, dtype=np.float64)
    out = np.zeros(nlocal)
    nb, one = C.c_int(blocks.size - 1), C.c_int(1)
    f(C.byref(nb), C.byref(one), _p(blocks), _p(rs), _p(ci), _p(gv), _p(out))
    return out


def ogs_gather_scatter(rowStarts, colIds, v):
    """gatherScatter kernel (libs/ogs/okl/ogsKernels.okl:32-81, T = double, Add), symmetric form: every copy of a
    node receives the sum over its copies, in place."""
    f = _load("ogsKernels_double_add.so").gatherScatter
    f.restype = None
    rs = np.ascontiguousarray(rowStarts, dtype=np.int32)
    ci = np.ascontiguousarray(colIds, dtype=np.int32)
    blocks = row_blocks(rs)
    v = np.array(v, dtype=np.float64)
    nb, one = C.c_int(blocks.size - 1), C.c_int(1)
    f(C.byref(nb), C.byref(one), _p(blocks), _p(rs), _p(ci), _p(rs), _p(ci), _p(v))
    return v
