"""GPU parity tests: the sm_100a CUDA path (through the C ABI) against the oracle and against the
reference's golden fixtures.  Bars: bit-exact for integer/index work, <= 1e-12 relative for the
FP64 operator (BASELINE.json north_star), PCG iteration count within +-1 of the reference."""
import ctypes

import numpy as np
import pytest
import torch

from golden_util import DIGEST, FULL, Problem, load, relerr

import libparanumal_b200 as lp
from libparanumal_b200 import _lib as L
from libparanumal_b200 import api
from libparanumal_b200.api import Comm, Elliptic, LinAlg, Ogs, Pcg, Precon
from libparanumal_b200.box_mesh import BoxMesh
from libparanumal_b200.problem import EllipticProblem
from oracle import elliptic_ref as er
from oracle import ogs_ref

pytestmark = pytest.mark.gpu
libc = ctypes.CDLL("libc.so.6")
AX_TOL = 1e-12


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ----------------------------------------------------------------------------- Ax kernel
@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("lam", [0.0, 1.3])
@pytest.mark.parametrize("path", ["general", "gll-registered", "pencil"])
def test_ax_hex3d_element_local_and_indexed(N, lam, path, request):
    """general: un-registered D (full Nq^2 contractions); gll-registered: D promised immutable, classified
    centro-antisymmetric -> even-odd kernels; pencil: the one-thread-per-column variant."""
    Nq, n = N + 1, 3
    lib = L.load()
    L.check(lib.libp_ax_hex3d_set_variant(0 if path == "pencil" else 1))
    request.addfinalizer(lambda: lib.libp_ax_hex3d_set_variant(1))
    _dev = dev
    if path == "gll-registered":
        Dreg = torch.from_numpy(BoxMesh(N, 1, 1, 1, device="cpu", geometry=False).D_host.reshape(-1).copy()).cuda()
        api.register_D(Nq, Dreg)
        request.addfinalizer(lambda: api.unregister_D(Dreg))
    mesh = BoxMesh(N, n, n + 1, n, device="cuda")
    E, Np = mesh.Nelements, mesh.Np
    rng = np.random.default_rng(N)
    # a deformed-geometry stand-in: random SPD-ish factors so all six G components are exercised
    ggeo = mesh.ggeo.cpu().numpy() * (1.0 + 0.3 * rng.standard_normal((E, 6, Np)))
    ggeo[:, [1, 2, 4]] += 0.1 * rng.standard_normal((E, 3, Np)) * np.abs(ggeo[:, [0]])
    wJ = mesh.wJ.cpu().numpy()
    D = mesh.D_host
    q = rng.uniform(-1, 1, E * Np)
    ref = er.ax_hex3d(Nq, wJ, ggeo, D, lam, q)
    out = torch.zeros(E * Np, dtype=torch.float64, device="cuda")
    dD = Dreg if path == "gll-registered" else dev(D)
    api.ax_hex3d(Nq, E, None, None, dev(wJ), dev(ggeo), dD, lam, dev(q), out)
    assert relerr(out.cpu().numpy(), ref) < AX_TOL
    # indexed (GlobalToLocal with -1 entries) on an element sub-list
    Ng = E * Np // 2
    G2L = rng.integers(-1, Ng, E * Np).astype(np.int32)
    qg = rng.uniform(-1, 1, Ng)
    elist = np.sort(rng.choice(E, size=max(E // 2, 1), replace=False)).astype(np.int32)
    ref = np.full(E * Np, 7.0)
    er.ax_hex3d(Nq, wJ, ggeo, D, lam, qg, G2L=G2L, element_list=elist, out=ref)
    out = torch.full((E * Np,), 7.0, dtype=torch.float64, device="cuda")
    api.ax_hex3d(Nq, len(elist), dev(elist), dev(G2L), dev(wJ), dev(ggeo), dD, lam, dev(qg), out)
    assert relerr(out.cpu().numpy(), ref) < AX_TOL  # untouched elements keep the sentinel
    # fused scatter-add epilogue: Aq[G2L] += A_e q over the sub-list
    refg = np.zeros(Ng)
    loc = ref.reshape(E, Np)
    g2 = G2L.reshape(E, Np)
    for e in elist:
        m = g2[e] >= 0
        np.add.at(refg, g2[e][m], loc[e][m])
    outg = torch.zeros(Ng, dtype=torch.float64, device="cuda")
    api.ax_hex3d_gather(Nq, len(elist), dev(elist), dev(G2L), dev(wJ), dev(ggeo), dD, lam, dev(qg), outg)
    assert relerr(outg.cpu().numpy(), refg) < AX_TOL


def test_ax_hex3d_non_gll_D_registered_takes_general_path():
    """A registered D that is NOT centro-antisymmetric must still give the right answer."""
    N, Nq, n = 3, 4, 2
    mesh = BoxMesh(N, n, n, n, device="cuda")
    E, Np = mesh.Nelements, mesh.Np
    rng = np.random.default_rng(5)
    D = rng.standard_normal((Nq, Nq))
    q = rng.uniform(-1, 1, E * Np)
    ref = er.ax_hex3d(Nq, mesh.wJ.cpu().numpy(), mesh.ggeo.cpu().numpy(), D, 0.7, q)
    dD = dev(D.reshape(-1))
    api.register_D(Nq, dD)
    try:
        out = torch.zeros(E * Np, dtype=torch.float64, device="cuda")
        api.ax_hex3d(Nq, E, None, None, mesh.wJ, mesh.ggeo, dD, 0.7, dev(q), out)
        assert relerr(out.cpu().numpy(), ref) < AX_TOL
    finally:
        api.unregister_D(dD)


@pytest.mark.parametrize("name", FULL)
@pytest.mark.parametrize("mode", [0, 1])
def test_operator_vs_reference_golden(name, mode):
    """elliptic_t::Operator on the reference's own D / ggeo / wJ / q: <= 1e-12 relative."""
    g = load(name)
    N, n, flag = (int(v) for v in g["config"])
    lam = float(g["lambda"][0])
    mesh = BoxMesh(N, n, n, n, boundary_flag=flag, device="cuda", geometry=False)
    E, Np = mesh.Nelements, mesh.Np
    mesh.ggeo, mesh.wJ, mesh.D = dev(g["ggeo"]).reshape(E, 6, Np), dev(g["wJ"]).reshape(E, Np), dev(g["D"])
    libc.srand(1)
    p = EllipticProblem(N, n, lam=lam, boundary_flag=flag, mesh=mesh, mode=mode)
    assert np.array_equal(p.G2L_host, g["GlobalToLocal"])
    q = p.vec()
    q[: p.Ndofs] = dev(g["q"])
    Aq = p.operator(q)
    assert relerr(Aq[: p.Ndofs].cpu().numpy(), g["Aq"]) < AX_TOL
    # Jacobi diagonal, rhs and the full solve
    inv = p.inv_diagonal()
    assert relerr(1.0 / inv[: p.Ndofs].cpu().numpy(), g["diagA"]) < 1e-12
    r = p.vec()
    p.ogs.Gather(r, dev(g["rL"]), 1, L.ADD, L.TRANS)
    assert relerr(r[: p.Ndofs].cpu().numpy(), g["r"]) < 1e-13
    M = Precon.Identity(p.Ndofs) if name.endswith("none") else Precon.Jacobi(p.Ndofs, inv)
    for native in (True, False):
        x, rr = p.vec(), r.clone()
        solver = p.pcg()
        if native:
            it = solver.Solve(p.op, M, x, rr)
        else:
            A_fn = lambda pin, pout: L.check(L.load().libp_elliptic_operator(p.op.handle, pin, pout, api._stream()))
            M_fn = lambda pin, pout: L.check(L.load().libp_precon_apply(M.handle, pin, pout, api._stream()))
            it = solver.SolveCallbacks(A_fn, M_fn, x, rr)
        assert abs(it - int(g["iterations"][0])) <= 1, (native, it)
        assert relerr(x[: p.Ndofs].cpu().numpy(), g["xsol"]) < 1e-7
        h = solver.residual_history()
        k = min(len(h) - 1, len(g["res_history"]))
        assert np.allclose(h[1:k + 1], g["res_history"][:k], rtol=1e-5)


@pytest.mark.parametrize("name", DIGEST)
def test_pipeline_vs_reference_digest(name):
    """BASELINE config 1 (N=4, 10^3 box, lambda=1) and an N=7 BP5 box, everything generated by the
    product-side harness: operator samples <= 1e-12, Jacobi/None PCG iteration counts +-1."""
    g = load(name)
    N, n, flag = (int(v) for v in g["config"])
    lam = float(g["lambda"][0])
    st = int(g["sample_stride"][0])
    libc.srand(1)
    p = EllipticProblem(N, n, lam=lam, boundary_flag=flag, coords=True)
    q = p.vec()
    q[: p.Ndofs] = dev(er.splitmix_uniform(1234, p.Ndofs))
    Aq = p.operator(q)[: p.Ndofs].cpu().numpy()
    assert relerr(Aq[::st], g["Aq_sample"]) < AX_TOL
    assert abs(np.linalg.norm(Aq) / g["Aq_norm2"][0] - 1) < AX_TOL
    r = p.rhs_sine3d()
    assert relerr(r[: p.Ndofs].cpu().numpy()[::st], g["r_sample"]) < 1e-11
    M = Precon.Identity(p.Ndofs) if name.endswith("none") else p.jacobi()
    x = p.vec()
    it = p.pcg().Solve(p.op, M, x, r)
    assert abs(it - int(g["iterations"][0])) <= 1
    assert relerr(x[: p.Ndofs].cpu().numpy()[::st], g["xsol_sample"]) < 1e-7


def test_operator_properties_large():
    """Size-independent properties on a mesh much larger than the oracle handles quickly
    (N=7, 24^3 elements, 7.1 M nodes): symmetry, linearity, fused == unfused."""
    N, n = 7, 24
    p1 = EllipticProblem(N, n, lam=1.0, mode=1)
    p0 = Elliptic(p1.Nq, p1.mesh.localGatherElementList, None, p1.GlobalToLocal, p1.mesh.wJ, p1.mesh.ggeo,
                  p1.mesh.D, 1.0, p1.ogs, mode=0)
    torch.manual_seed(0)
    x, y = p1.vec(), p1.vec()
    x[: p1.Ndofs] = torch.rand(p1.Ndofs, dtype=torch.float64, device="cuda") - 0.5
    y[: p1.Ndofs] = torch.rand(p1.Ndofs, dtype=torch.float64, device="cuda") - 0.5
    Ax, Ay = p1.operator(x), p1.operator(y)
    Ax0 = p1.vec()
    p0.Operator(x, Ax0)
    scale = float(Ax.abs().max())
    assert float((Ax - Ax0).abs().max()) / scale < 1e-13       # fused atomics vs ordered gather
    assert abs(float(y @ Ax - x @ Ay)) / abs(float(y @ Ax)) < 1e-12  # symmetry
    z = 2.0 * x - 3.0 * y
    Az = p1.operator(z)
    assert float((Az - (2.0 * Ax - 3.0 * Ay)).abs().max()) / scale < 1e-12  # linearity
    assert float(x @ Ax) > 0  # positive definite (lambda > 0, Dirichlet)


# ----------------------------------------------------------------------------- ogs device ops
OGS_IDS = np.array([5, 0, 3, 5, 9, -3, 3, 0, 5, 12, 9, 7, 7, 7, 7], dtype=np.int64)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64])
@pytest.mark.parametrize("op", ["Add", "Mul", "Max", "Min"])
@pytest.mark.parametrize("k", [1, 3])
def test_ogs_gather_scatter_ops(dtype, op, k):
    ids = OGS_IDS.copy()
    libc.srand(1)
    ogs = Ogs().Setup(ids.size, ids, Comm(), kind=L.SIGNED, unique=True)
    ref = ogs_ref.ogs_setup_all([OGS_IDS.copy()], ogs_ref.SIGNED, True)[0]
    assert np.array_equal(ids, ref.ids)
    rng = np.random.default_rng(3)
    v = (rng.integers(1, 4, ids.size * k)).astype(dtype) if np.issubdtype(dtype, np.integer) else \
        rng.uniform(0.5, 2.0, ids.size * k).astype(dtype)
    opk = {"Add": L.ADD, "Mul": L.MUL, "Max": L.MAX, "Min": L.MIN}[op]
    for trans, tname in ((L.TRANS, "Trans"), (L.NOTRANS, "NoTrans")):
        gv = torch.zeros(ref.NlocalT * k, dtype=torch.from_numpy(v).dtype, device="cuda")
        ogs.Gather(gv, dev(v), k, opk, trans)
        exp = ogs_ref.op_gather(ref.gatherLocal, v, tname, op, K=k)
        nrow = ref.NlocalT if trans == L.TRANS else ref.NlocalP
        got = gv.cpu().numpy()[: nrow * k]
        if np.issubdtype(dtype, np.integer) or op in ("Max", "Min"):
            assert np.array_equal(got, exp)
        else:
            assert np.allclose(got, exp, rtol=1e-6 if dtype == np.float32 else 1e-15)
    # scatter NoTrans writes every copy, masked nodes untouched
    gvals = (np.arange(ref.NlocalT * k) + 1).astype(dtype)
    out = torch.full((ids.size * k,), -7, dtype=torch.from_numpy(v).dtype, device="cuda")
    ogs.Scatter(out, dev(gvals), k, L.NOTRANS)
    exp = np.full(ids.size * k, -7, dtype=dtype)
    ogs_ref.op_scatter(ref.gatherLocal, gvals, exp, "NoTrans", K=k)
    assert np.array_equal(out.cpu().numpy(), exp)
    # gatherScatter Sym == scatter(gather_T)
    vv = dev(v).clone()
    ogs.GatherScatter(vv, k, opk, L.SYM)
    g = ogs_ref.op_gather(ref.gatherLocal, v, "Trans", op, K=k)
    exp = v.copy()
    ogs_ref.op_scatter(ref.gatherLocal, g, exp, "NoTrans", K=k)
    if np.issubdtype(dtype, np.integer) or op in ("Max", "Min"):
        assert np.array_equal(vv.cpu().numpy(), exp)
    else:
        assert np.allclose(vv.cpu().numpy(), exp, rtol=1e-6 if dtype == np.float32 else 1e-15)


def test_ogs_gather_add_is_bit_exact_with_reference_order():
    """Add gather sums left-to-right in colIds order like ogsOperator_t::Gather -> bitwise equal."""
    g = load("hex_n3_e3_jacobi")
    p = Problem(g, geometry=False)
    ids = p.ids.copy()
    libc.srand(1)
    ogs = Ogs().Setup(ids.size, ids, Comm(), kind=L.SIGNED, unique=True)
    v = np.random.default_rng(0).standard_normal(ids.size)
    gv = torch.zeros(ogs.NlocalT, dtype=torch.float64, device="cuda")
    ogs.Gather(gv, dev(v), 1, L.ADD, L.TRANS)
    exp = er.gather_add(p.ogs.gatherLocal.rowStartsT, p.ogs.gatherLocal.colIdsT, v)
    assert np.array_equal(gv.cpu().numpy(), exp)


# ----------------------------------------------------------------------------- linAlg
@pytest.mark.parametrize("n", [1, 2, 255, 1000, 100003])
def test_linalg_vector_ops_and_reductions(n):
    la = LinAlg()
    rng = np.random.default_rng(n)
    a, x, y = (rng.uniform(0.5, 2.0, n) for _ in range(3))
    da, dx = dev(a), dev(x)

    def run(fn, *args, out_idx=-1):
        targs = [dev(t) if isinstance(t, np.ndarray) else t for t in args]
        fn(n, *targs)
        return targs[out_idx].cpu().numpy()

    assert np.array_equal(run(la.set, 3.5, y.copy()), np.full(n, 3.5))
    assert np.allclose(run(la.add, 1.25, y.copy()), y + 1.25, rtol=1e-15)
    assert np.allclose(run(la.scale, 1.25, y.copy()), y * 1.25, rtol=1e-15)
    assert np.allclose(run(la.axpy, 2.0, x, -0.5, y.copy()), 2.0 * x - 0.5 * y, rtol=1e-15)
    nan = np.full(n, np.nan)
    assert np.allclose(run(la.axpy, 2.0, x, 0.0, nan.copy()), 2.0 * x, rtol=1e-15)  # beta==0 never reads y
    assert np.allclose(run(la.zaxpy, 2.0, x, -0.5, y, nan.copy()), 2.0 * x - 0.5 * y, rtol=1e-15)
    assert np.allclose(run(la.amx, 2.0, a, x.copy()), 2.0 * a * x, rtol=1e-15)
    assert np.allclose(run(la.amxpy, 2.0, a, x, 0.5, y.copy()), 2.0 * a * x + 0.5 * y, rtol=1e-15)
    assert np.allclose(run(la.amxpy, 2.0, a, x, 0.0, nan.copy()), 2.0 * a * x, rtol=1e-15)
    assert np.allclose(run(la.zamxpy, 2.0, a, x, 0.5, y, nan.copy()), 2.0 * a * x + 0.5 * y, rtol=1e-15)
    assert np.allclose(run(la.adx, 2.0, a, x.copy()), 2.0 * x / a, rtol=1e-15)
    assert np.allclose(run(la.adxpy, 2.0, a, x, 0.5, y.copy()), 2.0 * x / a + 0.5 * y, rtol=1e-15)
    assert np.allclose(run(la.adxpy, 2.0, a, x, 0.0, nan.copy()), 2.0 * x / a, rtol=1e-15)
    assert np.allclose(run(la.zadxpy, 2.0, a, x, 0.5, y, nan.copy()), 2.0 * x / a + 0.5 * y, rtol=1e-15)
    # unaligned views take the scalar path
    if n > 3:
        big = dev(np.concatenate([[0.0], y]))
        la.scale(n, 2.0, big[1:])
        assert np.allclose(big[1:].cpu().numpy(), 2.0 * y, rtol=1e-15)
    assert la.min(n, da) == a.min() and la.max(n, da) == a.max()
    assert abs(la.sum(n, da) / a.sum() - 1) < 1e-13
    assert abs(la.norm2(n, da) / np.linalg.norm(a) - 1) < 1e-13
    assert abs(la.innerProd(n, da, dx) / (a @ x) - 1) < 1e-13
    dy = dev(y)
    assert abs(la.weightedNorm2(n, dy, da) / np.sqrt(np.sum(y * a * a)) - 1) < 1e-13
    assert abs(la.weightedInnerProd(n, dy, da, dx) / np.sum(y * a * x) - 1) < 1e-13
    # deterministic: same bits on repeat
    assert la.innerProd(n, da, dx) == la.innerProd(n, da, dx)


def test_pcg_edge_cases():
    p = EllipticProblem(2, 3, lam=1.0, coords=True)
    M = p.jacobi()
    # zero rhs: zero iterations, like the reference's (iter==0 && rdotr0==0) exit
    x, r = p.vec(), p.vec()
    assert p.pcg().Solve(p.op, M, x, r) == 0
    # maxit cap
    r = p.rhs_sine3d()
    x = p.vec()
    assert p.pcg().Solve(p.op, M, x, r, maxit=3) == 3
    # flexible PCG converges to the same solution
    r1, r2, x1, x2 = p.rhs_sine3d(), p.rhs_sine3d(), p.vec(), p.vec()
    it1 = p.pcg().Solve(p.op, M, x1, r1)
    it2 = p.pcg(flexible=True).Solve(p.op, M, x2, r2)
    assert abs(it1 - it2) <= 2
    assert float((x1 - x2).abs().max()) < 1e-7


def test_all_neumann_periodic_poisson():
    """lambda=0 on a periodic box: allNeumann path (ZeroMean on rhs and on every precon output)."""
    N, n = 3, 4
    libc.srand(1)
    p = EllipticProblem(N, n, lam=0.0, boundary_flag=-1, coords=True)
    assert p.allNeumann
    m = p.mesh
    f = (torch.sin(2 * np.pi * m.x) * torch.cos(2 * np.pi * m.y)).reshape(-1)  # zero-mean forcing
    rL = (m.wJ.reshape(-1) * f).contiguous()
    r = p.vec()
    p.ogs.Gather(r, rL, 1, L.ADD, L.TRANS)
    la = LinAlg()
    mean = la.sum(p.Ndofs, r) / p.NglobalDofs
    la.add(p.Ndofs, -mean, r)
    x = p.vec()
    it = p.pcg().Solve(p.op, p.jacobi(), x, r.clone())
    assert 0 < it < 200
    Ax = p.operator(x)
    assert float((Ax[: p.Ndofs] - r[: p.Ndofs]).norm() / r[: p.Ndofs].norm()) < 1e-6


@pytest.mark.parametrize("N,n,lam", [(7, 16, 1.0), (7, 13, 0.0), (3, 26, 1.0), (1, 34, 0.5), (4, 20, 0.0), (5, 18, 1.0)])
def test_zero_ahead_operator_and_pcg_match_memset_path(N, n, lam):
    """In-kernel zero-fill of the fused accumulator (ticket + per-group counters, csrc/ax_hex3d.cu kZA): same
    operator to rounding (the reductions have no fixed order in either path), no protocol time-outs, same PCG
    iteration count.  Sizes exceed the zero-ahead distance (2048 blocks) so producers and consumers both run."""
    libc.srand(1)
    p = EllipticProblem(N, n, lam=lam, coords=True)
    q = p.vec()
    q[: p.Ndofs] = dev(er.splitmix_uniform(99, p.Ndofs))
    ref = p.operator(q).clone()
    p.op.set_zero_ahead(True)
    junk = p.vec(fill=123.0)          # the accumulator must not need to be clean on entry
    for _ in range(3):
        out = p.operator(q, junk)
    assert p.op.zero_ahead_errors() == 0
    scale = float(ref[: p.Ndofs].abs().max())
    assert float((out[: p.Ndofs] - ref[: p.Ndofs]).abs().max()) / scale < 1e-13
    if lam:
        r0 = p.rhs_sine3d()
        its = []
        for za in (False, True):
            p.op.set_zero_ahead(za)
            x, r = p.vec(), r0.clone()
            its.append(p.pcg().Solve(p.op, p.jacobi(), x, r, tol=1e-8, maxit=60))
        assert abs(its[0] - its[1]) <= 1, its
        assert p.op.zero_ahead_errors() == 0
