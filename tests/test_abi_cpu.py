"""CPU-side checks of the C ABI: the library loads without a GPU, exports every symbol the header
declares, and its HOST ogs setup reproduces the reference maps bit-exactly (golden fixtures)."""
import ctypes
import os
import re

import numpy as np
import pytest

from golden_util import DIGEST, EDGE, FULL, Problem, load, sha

import libparanumal_b200 as lp
from libparanumal_b200 import _lib as L
from libparanumal_b200.api import Comm, Ogs
from libparanumal_b200.box_mesh import BoxMesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libc = ctypes.CDLL("libc.so.6")


def fresh_rand():
    """The reference process never calls srand(): its first ogs setup sees glibc's default seed."""
    libc.srand(1)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "libp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(libp_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"libp_operator_fn"}
    assert len(declared) > 50
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/libp_b200.h but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature"
    assert b"sm_100a" in lib.libp_b200_version()


def test_errors_are_reported_not_thrown():
    lib = L.load()
    h = ctypes.c_void_p()
    ids = np.array([1, 2, 2], dtype=np.int64)
    comm = Comm()
    # Unsigned + unique is an invalid combination in the reference (ogsSetup.cpp:97-100)
    rc = lib.libp_ogs_setup(3, ids.ctypes.data, comm.handle, L.UNSIGNED, 1, 0, ctypes.byref(h))
    assert rc == -1 and b"Invalid ogs setup" in lib.libp_last_error()
    with pytest.raises(lp.LibpError):
        L.check(rc)


@pytest.mark.parametrize("name", FULL + DIGEST)
def test_ogs_setup_bit_exact_vs_reference(name):
    g = load(name)
    N, n, flag = (int(v) for v in g["config"])
    mesh = BoxMesh(N, n, n, n, boundary_flag=flag, geometry=False)
    _, ids = mesh.masked_global_ids()
    ids = ids.numpy().copy()
    fresh_rand()
    ogs = Ogs().Setup(ids.size, ids, Comm(), kind=L.SIGNED, unique=True)
    cnt = g["ogs_counts"]
    assert [ogs.N, ogs.Ngather, ogs.NlocalT, ogs.NlocalP, ogs.NhaloT, ogs.NhaloP, ogs.NgatherGlobal] == list(cnt[:7])
    m = ogs.maps("local")
    g2l = ogs.SetupGlobalToLocalMapping()
    if name in FULL:
        assert np.array_equal(ids, g["maskedGlobalIds"])
        for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
            assert np.array_equal(m[nm], g["gatherLocal_" + nm]), nm
        assert np.array_equal(g2l, g["GlobalToLocal"])
    else:
        assert np.array_equal(sha(ids), g["maskedGlobalIds_sha256"])
        for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
            assert np.array_equal(sha(m[nm]), g["gatherLocal_" + nm + "_sha256"]), nm
        assert np.array_equal(sha(g2l), g["GlobalToLocal_sha256"])
    ogs.Free()


@pytest.mark.parametrize("name", EDGE)
def test_ogs_setup_edge_cases_from_reference_ids(name):
    """the C-ABI host setup on the reference's own ids of degenerate periodic boxes (many copies of an id inside one
    element): signs, counters and maps bit-exact against the reference dump"""
    g = load(name)
    ids = np.abs(g["maskedGlobalIds"]).astype(np.int64)
    fresh_rand()
    ogs = Ogs().Setup(ids.size, ids, Comm(), kind=L.SIGNED, unique=True)
    cnt = g["ogs_counts"]
    assert [ogs.N, ogs.Ngather, ogs.NlocalT, ogs.NlocalP, ogs.NhaloT, ogs.NhaloP, ogs.NgatherGlobal] == list(cnt[:7])
    assert np.array_equal(ids, g["maskedGlobalIds"])
    m = ogs.maps("local")
    for nm in ("rowStartsN", "rowStartsT", "colIdsN", "colIdsT"):
        assert np.array_equal(m[nm], g["gatherLocal_" + nm]), nm
    assert np.array_equal(ogs.SetupGlobalToLocalMapping(), g["GlobalToLocal"])
    ogs.Free()


def test_ogs_kinds_and_empty_input():
    comm = Comm()
    # empty
    ids = np.zeros(0, dtype=np.int64)
    o = Ogs().Setup(0, ids, comm, kind=L.SIGNED, unique=False)
    assert o.Ngather == 0 and o.NlocalT == 0
    # all ids zero (everything masked)
    ids = np.zeros(7, dtype=np.int64)
    o = Ogs().Setup(7, ids, comm, kind=L.SIGNED, unique=True)
    assert o.Ngather == 0
    # Unsigned: N and T maps coincide; ragged multiplicities
    ids = np.array([5, 3, 5, 0, 9, 3, 5], dtype=np.int64)
    o = Ogs().Setup(7, ids, comm, kind=L.UNSIGNED, unique=False)
    m = o.maps("local")
    assert o.NlocalT == 3 and list(m["rowStartsT"]) == [0, 3, 5, 6]
    assert list(m["colIdsT"]) == [0, 2, 6, 1, 5, 4] and list(m["colIdsN"]) == list(m["colIdsT"])
    # Signed, not unique: one positive per group is respected; two positives -> gather undefined
    ids = np.array([-5, 3, 5, 0, 9, -3, -5], dtype=np.int64)
    o = Ogs().Setup(7, ids, comm, kind=L.SIGNED, unique=False)
    assert o.gather_defined == 1 and o.NlocalP == 3
    m = o.maps("local")
    assert list(m["rowStartsN"]) == [0, 1, 2, 3] and list(m["colIdsN"]) == [2, 1, 4]
    ids = np.array([5, 5], dtype=np.int64)
    o = Ogs().Setup(2, ids, comm, kind=L.SIGNED, unique=False)
    assert o.gather_defined == 0


def test_parallel_exact_sort_equals_std_sort():
    """ogsBase_t::Setup depends on the tie order of libstdc++'s std::sort (libs/ogs/ogsSetup.cpp:245-275): the library's
    task-parallel introsort must give the identical permutation (many ties, few ties, sizes around the thresholds)"""
    import ctypes as C
    lib = L.load()
    for n, nkeys, seed in [(0, 1, 1), (1, 1, 1), (16, 2, 2), (17, 3, 3), (33, 1, 4), (1000, 7, 5), (65536, 65536, 6),
                           (200000, 100, 7), (200000, 199999, 8), (1500000, 2, 9), (1500000, 300000, 10)]:
        same = C.c_int(0)
        L.check(lib.libp_ogs_sort_selftest(n, nkeys, seed, C.byref(same)))
        assert same.value == 1, (n, nkeys, seed)


def test_bulk_rand_draws_equal_glibc_rand():
    """ogsBase_t::FindSharedNodes draws one rand() per id group; the setup takes them in bulk from glibc's generator
    state (csrc/ogs_setup.cpp GlibcRandBulk) - same values, and the process-wide stream continues where n rand() calls
    would have left it"""
    import ctypes as C
    lib = L.load()
    for seed, n in [(1, 0), (1, 1), (1, 30), (1, 31), (7, 34), (123, 1000), (1, 1000003)]:
        same = C.c_int(0)
        L.check(lib.libp_ogs_rand_selftest(seed, n, C.byref(same)))
        assert same.value == 1, (seed, n)


def test_chain_kernel_shared_memory_layouts_are_consistent_and_conflict_free():
    """Audit of the shared-memory geometry compiled into the element-chain kernel (csrc/ax_chain.cu: ChT / ChainPerm /
    ChIdx), on the host: every column (layout C) and pencil (layouts A, B) has exactly one owner lane, offsets are in
    bounds, and under the half-warp / quarter-warp bank model of DESIGN.md 4.1c the orders of the degree sweep cost
    the ideal number of shared wavefronts (Nq = 5 keeps one two-way conflict on the s_u read of layout B; Nq = 2, 3 -
    p-multigrid levels only - still use padded rows)"""
    import ctypes as C
    lib = L.load()
    for Nq in range(2, 10):
        ok, wf = C.c_int(0), (C.c_int * 6)()
        L.check(lib.libp_ax_chain_layout_selftest(Nq, C.byref(ok), wf))
        assert ok.value == 1, Nq
        cA, cI, aA, aI, bA, bI = list(wf)
        assert cA >= cI and aA >= aI and bA >= bI
        if Nq in (4, 6, 7, 8, 9):
            assert (cA, aA, bA) == (cI, aI, bI), (Nq, list(wf))
        if Nq == 5:
            assert cA == cI and aA == aI and bA <= 1.25 * bI, list(wf)
