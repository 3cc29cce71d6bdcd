"""Pins the oracle (oracle/*.py, oracle/csrc/oracle.c) against fixtures dumped from the UNMODIFIED
reference (libParanumal built in-container, see oracle/refbuild/).  CPU only."""
import numpy as np
import pytest

from golden_util import DIGEST, EDGE, FULL, Problem, load, relerr, sha
from oracle import elliptic_ref as er
from oracle.ogs_ref import GlibcRand, libstdcxx_sort_perm


@pytest.mark.parametrize("name", FULL)
def test_mesh_restatement_full(name):
    g = load(name)
    p = Problem(g)
    m = p.mesh
    # D = Vr/V comes out of LAPACK in the reference (meshBasis1D.cpp:114-136): agreement to rounding of its largest
    # entry (|D| reaches 24 at N=8)
    assert np.max(np.abs(m.D.ravel() - g["D"])) < 1e-14 * np.max(np.abs(g["D"]))
    assert np.max(np.abs(m.gllz - g["gllz"])) < 1e-14 and np.max(np.abs(m.gllw - g["gllw"])) < 1e-14
    assert relerr(m.x.ravel(), g["x"]) < 1e-14
    assert relerr(m.ggeo.ravel(), g["ggeo"]) < 1e-12 and relerr(m.wJ.ravel(), g["wJ"]) < 1e-12
    assert np.array_equal(m.globalIds, g["globalIds"])
    assert np.array_equal(m.mapB, g["meshMapB"])
    assert np.array_equal(p.mapB, g["mapB"])
    assert np.array_equal(m.localGatherElementList, g["localGatherElementList"])


@pytest.mark.parametrize("name", FULL)
def test_ogs_maps_bit_exact_full(name):
    g = load(name)
    p = Problem(g, geometry=False)
    o = p.ogs
    cnt = g["ogs_counts"]
    assert [o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.NgatherGlobal] == list(cnt[:7])
    assert np.array_equal(o.ids, g["maskedGlobalIds"])  # signs rewritten by the unique-owner pick
    for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
        assert np.array_equal(getattr(o.gatherLocal, nm), g["gatherLocal_" + nm]), nm
    assert np.array_equal(p.G2L, g["GlobalToLocal"])


@pytest.mark.parametrize("name", DIGEST)
def test_ogs_maps_bit_exact_digest(name):
    g = load(name)
    p = Problem(g, geometry=False)
    o = p.ogs
    assert np.array_equal(sha(p.mesh.globalIds), g["globalIds_sha256"])
    assert np.array_equal(sha(o.ids), g["maskedGlobalIds_sha256"])
    assert np.array_equal(sha(p.G2L), g["GlobalToLocal_sha256"])
    for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
        assert np.array_equal(sha(getattr(o.gatherLocal, nm)), g["gatherLocal_" + nm + "_sha256"]), nm
    assert list(g["ogs_counts"][:7]) == [o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.NgatherGlobal]


@pytest.mark.parametrize("name", FULL)
def test_operator_diag_rhs_pcg_full(name):
    g = load(name)
    p = Problem(g, geometry=False)
    o, Nq = p.ogs, p.Nq
    rsT, ciT = o.gatherLocal.rowStartsT, o.gatherLocal.colIdsT
    q = er.splitmix_uniform(1234, o.Ngather)
    assert np.array_equal(q, g["q"])
    Aq = er.operator(Nq, p.G2L, g["wJ"], g["ggeo"], g["D"], p.lam, rsT, ciT, q)
    assert relerr(Aq, g["Aq"]) < 1e-13
    dg = er.gather_add(rsT, ciT, er.build_diagonal_local(Nq, g["ggeo"], g["wJ"], g["D"], p.lam, p.mapB))
    assert relerr(dg, g["diagA"]) < 1e-13
    rL = er.rhs_sine3d(Nq, g["x"], g["y"], g["z"], g["wJ"], g["ggeo"], g["D"], p.lam, p.mapB)
    assert relerr(rL, g["rL"]) < 1e-12
    r = er.gather_add(rsT, ciT, g["rL"])
    assert relerr(r, g["r"]) < 1e-14
    inv = None if name.endswith("none") else 1.0 / g["diagA"]
    it, x, hist = er.pcg(Nq, p.G2L, g["wJ"], g["ggeo"], g["D"], p.lam, rsT, ciT, inv, np.zeros_like(r), g["r"])
    assert abs(it - int(g["iterations"][0])) <= 1
    assert relerr(x, g["xsol"]) < 1e-7
    n = min(len(hist) - 1, len(g["res_history"]))
    assert np.allclose(hist[1:n + 1], g["res_history"][:n], rtol=1e-5)


@pytest.mark.parametrize("name", DIGEST)
def test_operator_pcg_digest(name):
    """BASELINE config 1 (N=4, 10^3) and an N=7 BP5 (lambda=0) case: whole pipeline from the oracle's own
    mesh restatement, compared with strided samples of the reference's arrays."""
    g = load(name)
    p = Problem(g)
    o, Nq, m = p.ogs, p.Nq, p.mesh
    st = int(g["sample_stride"][0])
    rsT, ciT = o.gatherLocal.rowStartsT, o.gatherLocal.colIdsT
    assert relerr(m.ggeo.ravel()[::st], g["ggeo_sample"]) < 1e-12
    q = er.splitmix_uniform(1234, o.Ngather)
    Aq = er.operator(Nq, p.G2L, m.wJ, m.ggeo, m.D, p.lam, rsT, ciT, q)
    assert relerr(Aq[::st], g["Aq_sample"]) < 1e-12
    assert abs(np.linalg.norm(Aq) / g["Aq_norm2"][0] - 1) < 1e-12
    dg = er.gather_add(rsT, ciT, er.build_diagonal_local(Nq, m.ggeo, m.wJ, m.D, p.lam, p.mapB))
    assert relerr(dg[::st], g["diagA_sample"]) < 1e-12
    r = er.gather_add(rsT, ciT, er.rhs_sine3d(Nq, m.x, m.y, m.z, m.wJ, m.ggeo, m.D, p.lam, p.mapB))
    assert relerr(r[::st], g["r_sample"]) < 1e-11
    inv = None if name.endswith("none") else 1.0 / dg
    it, x, hist = er.pcg(Nq, p.G2L, m.wJ, m.ggeo, m.D, p.lam, rsT, ciT, inv, np.zeros_like(r), r)
    assert abs(it - int(g["iterations"][0])) <= 1
    assert relerr(x[::st], g["xsol_sample"]) < 1e-7


def test_glibc_rand_restatement():
    # first outputs of glibc rand() with the default seed (srand(1)): public known-answer values
    r = GlibcRand(1)
    assert [r.rand() for _ in range(5)] == [1804289383, 846930886, 1681692777, 1714636915, 1957747793]


def test_libstdcxx_sort_restatement_is_a_sort():
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 16, 17, 100, 5000):
        k = rng.integers(0, max(n // 4, 1), size=n)
        perm = libstdcxx_sort_perm(k)
        assert sorted(perm.tolist()) == list(range(n))
        assert np.all(np.diff(k[perm]) >= 0)


@pytest.mark.parametrize("name", EDGE)
def test_ogs_setup_edge_cases_from_reference_ids(name):
    """ids with many collisions inside one element (degenerate periodic boxes): the oracle's ogs setup on the
    reference's own ids reproduces the reference's signs, counters and all four local maps bit for bit."""
    from oracle.ogs_ref import SIGNED, ogs_setup_all
    g = load(name)
    ids_in = np.abs(g["maskedGlobalIds"]).astype(np.int64)
    assert np.bincount(np.unique(ids_in, return_inverse=True)[1]).max() >= 8
    o = ogs_setup_all([ids_in], SIGNED, True)[0]
    cnt = g["ogs_counts"]
    assert [o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.NgatherGlobal] == list(cnt[:7])
    assert np.array_equal(o.ids, g["maskedGlobalIds"])
    for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
        assert np.array_equal(getattr(o.gatherLocal, nm), g["gatherLocal_" + nm]), nm
    assert np.array_equal(o.global_to_local(), g["GlobalToLocal"])


@pytest.mark.parametrize("name", EDGE)
def test_operator_pcg_edge_cases_on_reference_arrays(name):
    """Operator / diagonal / PCG on the degenerate periodic boxes, everything (maps, geometry) taken from the
    reference dump: rows with many in-element copies, no masked node."""
    g = load(name)
    N, n, flag = (int(v) for v in g["config"])
    Nq, lam = N + 1, float(g["lambda"][0])
    G2L, rsT, ciT = g["GlobalToLocal"], g["gatherLocal_rowStartsT"], g["gatherLocal_colIdsT"]
    Ng = len(rsT) - 1
    q = er.splitmix_uniform(1234, Ng)
    assert np.array_equal(q, g["q"])
    Aq = er.operator(Nq, G2L, g["wJ"], g["ggeo"], g["D"], lam, rsT, ciT, q)
    assert relerr(Aq, g["Aq"]) < 1e-13
    dg = er.gather_add(rsT, ciT, er.build_diagonal_local(Nq, g["ggeo"], g["wJ"], g["D"], lam, g["mapB"]))
    assert relerr(dg, g["diagA"]) < 1e-13
    r = er.gather_add(rsT, ciT, g["rL"])
    assert relerr(r, g["r"]) < 1e-14
    inv = None if name == "hex_n2_e2_periodic" else 1.0 / g["diagA"]   # generated with PRECONDITIONER = NONE / JACOBI
    it, x, hist = er.pcg(Nq, G2L, g["wJ"], g["ggeo"], g["D"], lam, rsT, ciT, inv, np.zeros_like(r), g["r"])
    assert abs(it - int(g["iterations"][0])) <= 1
    assert relerr(x, g["xsol"]) < 1e-7
    # 30 unknowns, 32 iterations: past the exact-arithmetic termination CG is rounding-driven, so only the first
    # iterations are compared entry by entry
    k = min(len(hist) - 1, len(g["res_history"]), 10)
    assert np.allclose(hist[1:k + 1], g["res_history"][:k], rtol=1e-6)
