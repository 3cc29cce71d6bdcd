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
    # periodic box, lambda = 0: singular operator, ZeroMean + rank-one coarse boost (default precon MULTIGRID).
    # test/testElliptic.py:244-248 lists 0.059540839002614; the unmodified reference built in this container
    # prints 0.058046029189514 (5 iterations) for these settings, which is the value pinned here.
    ("testEllipticHex_C0_AllNeumann", "MULTIGRID", -1, 0.0, 0.058046029189514, dict(aggregation="UNSMOOTHED"), 5),
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
