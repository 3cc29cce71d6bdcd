"""Device-side setup kernels against dumps of the unmodified reference and the oracle:
  physical nodes / geometric factors   libs/mesh/meshPhysicalNodesHex3D.cpp, meshGeometricFactorsHex3D.cpp:94-174
  operator diagonal                    solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:998-1057
  trilinear on-the-fly Ax              solvers/elliptic/okl/ellipticAxHex3D.okl:440-627"""
import numpy as np
import pytest
import torch

from golden_util import FULL, load
from libparanumal_b200 import _lib as L
from libparanumal_b200 import api
from libparanumal_b200.box_mesh import BoxMesh
from oracle import elliptic_ref as er

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


def dev(a, dtype=None):
    return torch.from_numpy(np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))).cuda()


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


@pytest.mark.parametrize("name", FULL)
def test_physical_nodes_and_geometric_factors_vs_reference(name):
    g = load(name)
    N, n, flag = (int(v) for v in g["config"])
    Nq, Np = N + 1, (N + 1) ** 3
    mesh = BoxMesh(N, n, n, n, boundary_flag=flag, device="cuda", geometry=False)
    E = mesh.Nelements
    ex, ey, ez = mesh.element_vertices()
    x, y, z = (torch.empty(E * Np, dtype=torch.float64, device="cuda") for _ in range(3))
    api.mesh_physical_nodes_hex3d(Nq, E, ex, ey, ez, dev(g["gllz"]), x, y, z)
    for got, key in ((x, "x"), (y, "y"), (z, "z")):
        assert np.abs(got.cpu().numpy() - g[key]).max() < 2e-15, key
    # the factors from the reference's own coordinates, D and weights
    ggeo = torch.empty(E * 6 * Np, dtype=torch.float64, device="cuda")
    wJ = torch.empty(E * Np, dtype=torch.float64, device="cuda")
    vgeo = torch.empty(E * 12 * Np, dtype=torch.float64, device="cuda")
    api.mesh_geometric_factors_hex3d(Nq, E, dev(g["x"]), dev(g["y"]), dev(g["z"]), dev(g["D"]), dev(g["gllw"]), ggeo, wJ, vgeo)
    assert rel(ggeo.cpu().numpy(), g["ggeo"]) < 1e-13
    assert rel(wJ.cpu().numpy(), g["wJ"]) < 1e-13
    v = vgeo.cpu().numpy().reshape(E, 12, Np)
    assert rel(v[:, 10], g["wJ"].reshape(E, Np)) < 1e-13 and rel(v[:, 11] * v[:, 10], np.ones((E, Np))) < 1e-14


def test_geometric_factors_rejects_inverted_elements():
    N, Nq, Np = 2, 3, 27
    mesh = BoxMesh(N, 2, 2, 2, device="cuda", coords=True)
    x = mesh.x.reshape(-1).clone()
    ggeo = torch.empty(8 * 6 * Np, dtype=torch.float64, device="cuda")
    wJ = torch.empty(8 * Np, dtype=torch.float64, device="cuda")
    with pytest.raises(L.LibpError, match="Negative J"):
        api.mesh_geometric_factors_hex3d(Nq, 8, (-x).contiguous(), mesh.y.reshape(-1), mesh.z.reshape(-1), mesh.D,
                                         dev(mesh.gllw), ggeo, wJ)


@pytest.mark.parametrize("name", FULL)
def test_build_diagonal_vs_reference(name):
    """element-local diagonal from the reference's ggeo / wJ / D, gathered by the oracle: == reference diagA"""
    g = load(name)
    N, n, flag = (int(v) for v in g["config"])
    lam = float(g["lambda"][0])
    Nq, Np = N + 1, (N + 1) ** 3
    E = g["wJ"].size // Np
    A = torch.empty(E * Np, dtype=torch.float64, device="cuda")
    api.elliptic_build_diagonal_hex3d(Nq, E, dev(g["ggeo"]), dev(g["wJ"]), dev(g["D"]), dev(g["mapB"], np.int32), lam, 0.0, A)
    ref = er.build_diagonal_local(Nq, g["ggeo"], g["wJ"], g["D"], lam, g["mapB"])
    assert rel(A.cpu().numpy(), ref) < 1e-14
    dg = er.gather_add(g["gatherLocal_rowStartsT"], g["gatherLocal_colIdsT"], A.cpu().numpy())
    assert rel(dg, g["diagA"]) < 1e-13


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("lam", [0.0, 0.9])
def test_ax_trilinear_vs_oracle_and_stored_geometry(N, lam):
    """distorted (non-affine) trilinear elements: on-the-fly geometry == oracle restatement of the OKL kernel, and ==
    the stored-geometry kernel fed with libp_mesh_geometric_factors_hex3d of the same trilinear nodes"""
    Nq, Np, n = N + 1, (N + 1) ** 3, 3
    mesh = BoxMesh(N, n, n, n, device="cuda", geometry=False)
    E = mesh.Nelements
    ex, ey, ez = mesh.element_vertices()
    rng = np.random.default_rng(10 + N)
    # consistent distortion: move the lattice vertices, not the element copies
    lat = {}
    EX = np.stack([ex.cpu().numpy(), ey.cpu().numpy(), ez.cpu().numpy()], axis=1)  # [E,3,8]
    for e in range(E):
        for v in range(8):
            key = tuple(np.round(EX[e, :, v] * 3 * 64).astype(int))
            if key not in lat:
                lat[key] = rng.uniform(-0.04, 0.04, 3)
            EX[e, :, v] += lat[key]
    gllzw = np.concatenate([mesh.gllz, mesh.gllw])
    q = rng.uniform(-1, 1, E * Np)
    ref = er.ax_trilinear_hex3d(Nq, EX, mesh.gllz, mesh.gllw, mesh.D_host, lam, q)
    out = torch.empty(E * Np, dtype=torch.float64, device="cuda")
    api.ax_trilinear_hex3d(Nq, E, None, None, dev(EX.reshape(-1)), dev(gllzw), mesh.D, lam, dev(q), out)
    assert rel(out.cpu().numpy(), ref) < 1e-12
    # stored-geometry path on the same elements
    x, y, z = (torch.empty(E * Np, dtype=torch.float64, device="cuda") for _ in range(3))
    api.mesh_physical_nodes_hex3d(Nq, E, dev(EX[:, 0].copy()), dev(EX[:, 1].copy()), dev(EX[:, 2].copy()), dev(mesh.gllz), x, y, z)
    ggeo = torch.empty(E * 6 * Np, dtype=torch.float64, device="cuda")
    wJ = torch.empty(E * Np, dtype=torch.float64, device="cuda")
    api.mesh_geometric_factors_hex3d(Nq, E, x, y, z, mesh.D, dev(mesh.gllw), ggeo, wJ)
    out2 = torch.empty(E * Np, dtype=torch.float64, device="cuda")
    api.ax_hex3d(Nq, E, None, None, wJ, ggeo, mesh.D, lam, dev(q), out2)
    assert rel(out2.cpu().numpy(), ref) < 1e-11
    # gathered input through GlobalToLocal with an element list
    G2L = rng.integers(-1, 50, E * Np).astype(np.int32)
    qg = rng.uniform(-1, 1, 50)
    elist = np.array([2, 0, 5, 7], dtype=np.int32)
    ref3 = er.ax_trilinear_hex3d(Nq, EX, mesh.gllz, mesh.gllw, mesh.D_host, lam, qg, G2L=G2L, element_list=elist)
    out3 = torch.zeros(E * Np, dtype=torch.float64, device="cuda")
    api.ax_trilinear_hex3d(Nq, len(elist), dev(elist), dev(G2L), dev(EX.reshape(-1)), dev(gllzw), mesh.D, lam, dev(qg), out3)
    sel = np.zeros(E * Np, dtype=bool)
    for e in elist:
        sel[e * Np:(e + 1) * Np] = True
    assert rel(out3.cpu().numpy()[sel], ref3[sel]) < 1e-12


def _exyz(mesh, distort=0.0, seed=0):
    """[E,3,8] vertex array of a BoxMesh, optionally with a consistent random displacement of the lattice vertices"""
    ex, ey, ez = mesh.element_vertices()
    EX = np.stack([ex.cpu().numpy(), ey.cpu().numpy(), ez.cpu().numpy()], axis=1)
    if distort > 0:
        rng = np.random.default_rng(seed)
        lat = {}
        for e in range(EX.shape[0]):
            for v in range(8):
                key = tuple(np.round(EX[e, :, v] * 3 * 64).astype(int))
                if key not in lat:
                    lat[key] = rng.uniform(-distort, distort, 3)
                EX[e, :, v] += lat[key]
    return EX


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("distort", [0.0, 0.03])
def test_trilinear_operator_mode(N, distort):
    """libp_elliptic_set_trilinear: the operator with geometry on the fly (chain kernel, affine shortcut for the box,
    general trilinear otherwise) == the operator on stored factors of the same elements, and == the oracle"""
    import ctypes
    from libparanumal_b200.problem import EllipticProblem
    n, lam = 4, 0.7
    Nq, Np = N + 1, (N + 1) ** 3
    ctypes.CDLL("libc.so.6").srand(1)
    mesh = BoxMesh(N, n, n, n, device="cuda", geometry=False)
    E = mesh.Nelements
    EX = _exyz(mesh, distort, 20 + N)
    x, y, z = (torch.empty(E * Np, dtype=torch.float64, device="cuda") for _ in range(3))
    api.mesh_physical_nodes_hex3d(Nq, E, dev(EX[:, 0].copy()), dev(EX[:, 1].copy()), dev(EX[:, 2].copy()), dev(mesh.gllz), x, y, z)
    mesh.ggeo = torch.empty(E, 6, Np, dtype=torch.float64, device="cuda")
    mesh.wJ = torch.empty(E, Np, dtype=torch.float64, device="cuda")
    api.mesh_geometric_factors_hex3d(Nq, E, x, y, z, mesh.D, dev(mesh.gllw), mesh.ggeo, mesh.wJ)
    p = EllipticProblem(N, n, lam=lam, mesh=mesh)
    g = torch.Generator(device="cuda").manual_seed(5)
    q = p.vec()
    q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    ref = p.operator(q)[: p.Ndofs].cpu().numpy()
    AqL = er.ax_trilinear_hex3d(Nq, EX, mesh.gllz, mesh.gllw, mesh.D_host, lam, q[: p.Ndofs].cpu().numpy(), G2L=p.G2L_host)
    mp = p.ogs.maps("local")
    oracle = er.gather_add(mp["rowStartsT"], mp["colIdsT"], AqL)
    assert rel(ref, oracle) < 1e-11
    dEX = dev(EX.reshape(-1))
    p.op.set_trilinear(dEX, mesh.gllz, mesh.gllw)
    for L_ in (8, 1, 3):
        p.op.set_chain(L_, 1)
        Aq = p.vec(fill=float("nan"))
        p.op.Operator(q, Aq)
        assert rel(Aq[: p.Ndofs].cpu().numpy(), oracle) < 1e-12, (N, distort, L_)
    # Jacobi-PCG through the trilinear operator: same iteration count as on stored factors
    M = p.jacobi()
    r = p.vec()
    r[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda", generator=g)
    its = []
    for tri in (True, False):
        p.op.set_trilinear(dEX if tri else None, mesh.gllz, mesh.gllw)
        xs = p.vec()
        its.append(p.pcg().Solve(p.op, M, xs, r.clone(), tol=1e-8, maxit=500))
    assert abs(its[0] - its[1]) <= 1, its
