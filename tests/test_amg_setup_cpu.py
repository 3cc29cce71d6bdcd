"""CPU tests of the host-side AMG setup (libparanumal_b200/amg_setup.py) against hierarchies dumped from the
unmodified reference (tests/golden/mg_n2_e12.npz, amg_n2_e24.npz; oracle/refbuild/make_golden_mg.py):
drand48, the degree-1 operator matrix, strength/aggregation/prolongator/Galerkin product/rho for every level,
and the row-block split used on more than one rank."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from golden_util import load

from libparanumal_b200 import amg_setup as am
from oracle.mesh_box import build_box_hex_mesh, masked_global_ids
from oracle.ogs_ref import SIGNED, ogs_setup_all


def csr(g, pre):
    m = g[pre + "_meta"]
    return sp.csr_matrix((g[pre + "_vals"], g[pre + "_cols"], g[pre + "_rowStarts"]), shape=(int(m[0]), int(m[1])))


def test_drand48_matches_glibc():
    libc = ctypes.CDLL("libc.so.6")
    libc.drand48.restype = ctypes.c_double
    for seed in (0, 3):
        libc.srand48(seed)
        ref = np.array([libc.drand48() for _ in range(3000)])
        r = am.Drand48(seed)
        got = np.concatenate([r.draw(1), r.draw(999), r.draw(0), r.draw(2000)])
        assert np.array_equal(ref, got)


@pytest.mark.parametrize("name", ["mg_n2_e12", "amg_n2_e24"])
def test_hierarchy_matches_reference(name):
    g = load(name)
    kinds = list(g["level_kinds"])
    first = kinds.index(1)
    A = csr(g, f"L{first}_A")
    rng = am.Drand48(0)
    for l in range(first):  # the matrix-free levels drew their Arnoldi start vectors first (Nrows each)
        rng.draw(int(g[f"L{l}_meta"][1]))
    n = A.shape[0]
    levels, Ac, rho_c, _ = am.setup_hierarchy(A, np.full(n, 1.0 / np.sqrt(n)), rng)
    assert len(levels) == kinds.count(1)
    for k, lv in enumerate(levels):
        pre = f"L{first + k}"
        lam = g[pre + "_lambda"]
        assert abs(lv["rho"] - lam[1]) <= 1e-10 * lam[1] and abs(lv["rho"] / 10 - lam[0]) <= 1e-10 * lam[0]
        for nm in ("A", "P", "R"):
            ref = csr(g, f"{pre}_{nm}")
            assert lv[nm].shape == ref.shape and lv[nm].nnz == ref.nnz, (pre, nm)
            assert abs(lv[nm] - ref).max() <= 1e-12 * abs(ref).max(), (pre, nm)
    ref = csr(g, "coarse_A")
    assert Ac.shape == ref.shape and abs(Ac - ref).max() <= 1e-12 * abs(ref).max()


def test_degree1_operator_matrix_matches_reference():
    g = load("amg_n2_e24")
    m = build_box_hex_mesh(1, 24, 24, 24)
    _, ids = masked_global_ids(m)
    # the degree-1 setup is the second `unique` ogs setup of the run (after degree 2): owner choices (N-maps)
    # differ with the rand() position, the gathered numbering (T-maps / GlobalToLocal) does not
    o = ogs_setup_all([ids], SIGNED, True)[0]
    gid = torch.from_numpy(o.global_to_local().astype(np.int64))
    r, c, v = am.element_matrix_triplets(2, torch.from_numpy(m.ggeo), torch.from_numpy(m.wJ), torch.from_numpy(m.D),
                                         1.0, gid)
    A = sp.coo_matrix((v, (r, c)), shape=(o.Ngather, o.Ngather)).tocsr()
    ref = csr(g, "L1_A")
    assert A.shape == ref.shape and A.nnz == ref.nnz
    assert abs(A - ref).max() <= 1e-13 * abs(ref).max()


def test_row_block_split_reassembles():
    rng = np.random.default_rng(5)
    n, nc = 97, 41
    M = sp.random(n, nc, density=0.15, random_state=7, format="csr")
    rs = np.array([0, 30, 30, 71, n])   # one empty rank
    cs = np.array([0, 10, 25, 25, nc])
    x = rng.standard_normal(nc)
    y = np.zeros(n)
    for rank in range(4):
        d = am.split_rows(M, rs, cs, rank)
        xl = np.concatenate([x[cs[rank]:cs[rank + 1]], x[d["offd_colIds"]]])
        assert np.all(np.diff(d["offd_colIds"]) > 0)
        assert not np.any((d["offd_colIds"] >= cs[rank]) & (d["offd_colIds"] < cs[rank + 1]))
        yl = sp.csr_matrix((d["diag_vals"], d["diag_cols"], d["diag_rowStarts"]), shape=(d["Nrows"], d["NlocalCols"])) @ \
            xl[: d["NlocalCols"]]
        for k, row in enumerate(d["offd_rows"]):
            a, b = d["offd_mRowStarts"][k], d["offd_mRowStarts"][k + 1]
            assert b > a
            yl[row] += np.dot(d["offd_vals"][a:b], xl[d["offd_cols"][a:b]])
        y[rs[rank]:rs[rank + 1]] = yl
    assert np.allclose(y, M @ x, rtol=1e-13, atol=1e-13)


def test_coarse_partition_follows_roots():
    roots = np.array([0, 4, 5, 9, 30, 31, 70])
    part = am.coarse_partition(np.array([0, 5, 5, 31, 80]), roots)
    assert list(part) == [0, 2, 2, 5, 7]
