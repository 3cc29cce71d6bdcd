"""The reference's own regression tests for this path (test/testElliptic.py:95-98,199-213,244-248: Hex3D N=4 on a
10^3 box, `Solution norm` within TOL = 1e-5 of referenceNorm, test/test.py:63,131) run end to end through the
CUDA path: forcing + boundary lift, gather, PCG with each preconditioner, scatter, boundary data, mass-matrix norm
(EllipticProblem.run mirrors elliptic_t::Run).  Iteration counts are those of the reference build measured in this
container."""
import ctypes

import pytest

from libparanumal_b200 import api
from libparanumal_b200.problem import EllipticProblem

pytestmark = pytest.mark.gpu
libc = ctypes.CDLL("libc.so.6")
TOL = 1.0e-5  # test/test.py:63

# name, preconditioner, boundary flag, lambda, referenceNorm, multigrid arguments, reference iterations (or None)
CASES = [
    ("testEllipticHex_C0", "NONE", 1, 1.0, 0.353553390458384, {}, 113),
    ("testEllipticHex_C0_Jacobi", "JACOBI", 1, 1.0, 0.353553400508458, {}, 97),
    # the test-suite settings override PARALMOND AGGREGATION to UNSMOOTHED (test/testElliptic.py:41-45)
    ("testEllipticHex_C0_ParAlmond", "PARALMOND", 1, 1.0, 0.353553400508458, dict(aggregation="UNSMOOTHED"), 22),
    ("ParAlmond_smoothed", "PARALMOND", 1, 1.0, 0.353553400508458, {}, 11),
    ("testEllipticHex_C0_Multigrid", "MULTIGRID", 1, 1.0, 0.353553400508458, dict(aggregation="UNSMOOTHED"), 6),
    ("Multigrid_defaults", "MULTIGRID", 1, 1.0, 0.353553400508458, {}, 6),
    # test/testElliptic.py:244-248: periodic box (flag -1).  ellipticSettings() accepts Lambda=0.0 but never emits a
    # LAMBDA key (test/testElliptic.py:33-73), so the reference test really runs lambda = 1 (the default,
    # ellipticSettings.cpp): a non-singular periodic screened-Poisson problem.  Listed norm 0.059540839002614; the
    # unmodified reference built in this container prints 0.0595408388714726 in 5 iterations.
    ("testEllipticHex_C0_AllNeumann", "MULTIGRID", -1, 1.0, 0.059540839002614, dict(aggregation="UNSMOOTHED"), 5),
    # extra (not a reference test): the same rc file with LAMBDA = 0.0 added, i.e. the truly singular all-Neumann
    # operator: ZeroMean + rank-one coarse boost.  The reference prints 0.058046029189514 (5 iterations) for it.
    ("AllNeumann_lambda0", "MULTIGRID", -1, 0.0, 0.058046029189514, dict(aggregation="UNSMOOTHED"), 5),
]


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


@pytest.mark.parametrize("name,precon,flag,lam,ref_norm,mg,ref_it", CASES, ids=[c[0] for c in CASES])
def test_reference_suite_solution_norm(name, precon, flag, lam, ref_norm, mg, ref_it):
    libc.srand(1)
    p = EllipticProblem(4, 10, lam=lam, boundary_flag=flag, coords=True)
    it, norm, _ = p.run(precon, **mg)
    assert abs(norm - ref_norm) < TOL, (name, norm, ref_norm)
    if ref_it is not None:
        assert abs(it - ref_it) <= 1, (name, it, ref_it)
    assert it < 400


# LINEAR SOLVER = NBPCG on the same problem: iteration counts and the first residual norms printed by the
# unmodified reference built in this container (ellipticMain, Serial mode)
NBPCG = [("NONE", 113, [2.960718797524, 1.742998255149, 1.089704705958]),
         ("JACOBI", 97, [2.960718797524, 1.583783440215, 1.049754120587]),
         ("MULTIGRID", 6, [2.960718797524, 8.373782734443e-02, 2.514007040123e-03])]


@pytest.mark.parametrize("precon,ref_it,ref_hist", NBPCG, ids=[c[0] for c in NBPCG])
def test_nbpcg_matches_reference(precon, ref_it, ref_hist):
    import numpy as np

    from libparanumal_b200.api import NbPcg, Precon
    from libparanumal_b200.problem import MultigridHierarchy
    libc.srand(1)
    p = EllipticProblem(4, 10, lam=1.0, boundary_flag=1, coords=True)
    if precon == "NONE":
        M = Precon.Identity(p.Ndofs)
    elif precon == "JACOBI":
        M = p.jacobi()
    else:
        H = MultigridHierarchy.build(p)
        M = H.precon()
    r = p.rhs_sine3d()
    x = p.vec()
    solver = NbPcg(p.Ndofs, p.Nhalo, p.comm)
    it = solver.Solve(p.op, M, x, r, tol=1e-8, maxit=5000)
    assert abs(it - ref_it) <= 1, (it, ref_it)
    h = solver.residual_history()
    assert np.allclose(h[:3], ref_hist, rtol=1e-6), h[:3]
    # same solution as classic PCG
    x2, r2 = p.vec(), p.rhs_sine3d()
    p.pcg().Solve(p.op, M, x2, r2, tol=1e-8, maxit=5000)
    assert float((x[: p.Ndofs] - x2[: p.Ndofs]).abs().max() / x2[: p.Ndofs].abs().max()) < 1e-6


# PARALMOND CYCLE = KCYCLE (libs/parAlmond/parAlmondKcycle.cpp): iteration counts and residual norms printed by the
# unmodified reference built in this container (ellipticMain, Serial, Hex N=4 10^3, lambda=1).  The K-cycle differs
# from the V-cycle from the first iteration on (V-cycle: 8.3738e-02 after one MULTIGRID iteration).
KCYCLE = [("MULTIGRID", "SMOOTHED", 6, [2.960718797524, 8.704311358007e-02, 2.560919450235e-03, 1.450746244500e-04]),
          ("PARALMOND", "SMOOTHED", 10, [2.960718797524, 2.546306750764e-01, 5.399966726670e-02, 6.132016044317e-03]),
          ("PARALMOND", "UNSMOOTHED", 19, [2.960718797524, 3.554960101002e-01, 1.721931414151e-01, 5.092862981391e-02])]


@pytest.mark.parametrize("precon,agg,ref_it,ref_hist", KCYCLE, ids=[c[0] + "_" + c[1] for c in KCYCLE])
def test_kcycle_matches_reference(precon, agg, ref_it, ref_hist):
    import numpy as np

    from libparanumal_b200.problem import MultigridHierarchy
    libc.srand(1)
    p = EllipticProblem(4, 10, lam=1.0, boundary_flag=1, coords=True)
    H = MultigridHierarchy.build(p, aggregation=agg, algebraic_only=(precon == "PARALMOND"), cycle="KCYCLE")
    r = p.rhs_sine3d()
    x = p.vec()
    solver = p.pcg()
    it = solver.Solve(p.op, H.precon(), x, r, tol=1e-8, maxit=200)
    assert abs(it - ref_it) <= 1, (it, ref_it)
    h = solver.residual_history()
    assert np.allclose(h[:4], ref_hist, rtol=2e-4), h[:4]


# LINEAR SOLVER = NBFPCG on the same problem (libs/linearSolver/linearSolverNBFPCG.cpp): iteration counts and the
# first residual norms printed by the unmodified reference built in this container (ellipticMain, Serial mode;
# MULTIGRID with the parAlmond defaults).  The unpreconditioned variant takes 146 iterations in the reference too
# (its recurrences drift more than classic PCG's 113): the loop-exit rule and the recurrences are what is pinned.
NBFPCG = [("NONE", 146, [2.960718797524, 1.742998255149, 1.089704705958, 9.633990087332e-01]),
          ("JACOBI", 97, [2.960718797524, 1.583783440215, 1.049754120587, 9.100546237471e-01]),
          ("MULTIGRID", 6, [2.960718797524, 8.373782734443e-02, 2.514007040123e-03, 1.267123864858e-04])]


@pytest.mark.parametrize("precon,ref_it,ref_hist", NBFPCG, ids=[c[0] for c in NBFPCG])
def test_nbfpcg_matches_reference(precon, ref_it, ref_hist):
    import numpy as np

    from libparanumal_b200.api import NbFPcg, Precon
    from libparanumal_b200.problem import MultigridHierarchy
    libc.srand(1)
    p = EllipticProblem(4, 10, lam=1.0, boundary_flag=1, coords=True)
    if precon == "NONE":
        M = Precon.Identity(p.Ndofs)
    elif precon == "JACOBI":
        M = p.jacobi()
    else:
        M = MultigridHierarchy.build(p).precon()
    r = p.rhs_sine3d()
    x = p.vec()
    solver = NbFPcg(p.Ndofs, p.Nhalo, p.comm)
    it = solver.Solve(p.op, M, x, r, tol=1e-8, maxit=5000)
    # unpreconditioned NBFPCG is rounding-sensitive (the reference itself needs 146 iterations where PCG needs 113:
    # its recurrences drift): the count is pinned loosely there, the residual history (below) tightly
    assert abs(it - ref_it) <= (15 if precon == "NONE" else 1), (it, ref_it)
    h = solver.residual_history()
    assert np.allclose(h[:4], ref_hist, rtol=1e-6), h[:4]
    x2, r2 = p.vec(), p.rhs_sine3d()
    p.pcg().Solve(p.op, M, x2, r2, tol=1e-8, maxit=5000)
    assert float((x[: p.Ndofs] - x2[: p.Ndofs]).abs().max() / x2[: p.Ndofs].abs().max()) < 1e-6


# Degenerate periodic boxes (one / two elements per direction): the reference's ids repeat up to 27 times inside one
# element.  Everything is taken from the reference dump (ids, geometry, D, q): ogs setup through the C ABI, then the
# operator in both modes (many reductions of one block into the same row) and Jacobi / plain PCG.
@pytest.mark.parametrize("name,precon", [("hex_n3_e1_periodic", "JACOBI"), ("hex_n2_e2_periodic", "NONE")])
def test_edge_cases_on_reference_arrays(name, precon):
    import numpy as np
    import torch

    from golden_util import load, relerr
    from libparanumal_b200 import _lib as L
    from libparanumal_b200.api import Comm, Elliptic, Ogs, Pcg, Precon
    g = load(name)
    N = int(g["config"][0])
    Nq, lam = N + 1, float(g["lambda"][0])
    ids = np.abs(g["maskedGlobalIds"]).astype(np.int64)
    libc.srand(1)
    comm = Comm()
    ogs = Ogs().Setup(ids.size, ids, comm, kind=L.SIGNED, unique=True)
    assert np.array_equal(ids, g["maskedGlobalIds"])
    G2L = ogs.SetupGlobalToLocalMapping()
    assert np.array_equal(G2L, g["GlobalToLocal"])
    dev = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).cuda()
    E = ids.size // Nq ** 3
    elems = torch.arange(E, dtype=torch.int32, device="cuda")
    wJ, ggeo, D, dG2L = dev(g["wJ"]), dev(g["ggeo"]), dev(g["D"]), dev(G2L, np.int32)
    Ng = ogs.Ngather
    q = dev(g["q"])
    for mode in (0, 1):
        op = Elliptic(Nq, elems, None, dG2L, wJ, ggeo, D, lam, ogs, mode=mode)
        Aq = torch.zeros(Ng + ogs.Nhalo, dtype=torch.float64, device="cuda")
        op.Operator(q.clone(), Aq)
        assert relerr(Aq[:Ng].cpu().numpy(), g["Aq"]) < 1e-12, mode
    M = Precon.Jacobi(Ng, dev(1.0 / g["diagA"])) if precon == "JACOBI" else Precon.Identity(Ng)
    x = torch.zeros(Ng, dtype=torch.float64, device="cuda")
    it = Pcg(Ng, 0, comm).Solve(op, M, x, dev(g["r"]), tol=1e-8, maxit=200)
    assert abs(it - int(g["iterations"][0])) <= 1, (it, g["iterations"])
    assert relerr(x.cpu().numpy(), g["xsol"]) < 1e-6
