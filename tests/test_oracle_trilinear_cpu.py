"""Oracle restatement of the on-the-fly trilinear geometry (solvers/elliptic/okl/ellipticAxHex3D.okl:527-566) pinned
against the stored factors the unmodified reference dumps for box meshes (meshGeometricFactorsHex3D.cpp:94-174): for
trilinear elements both must agree."""
import numpy as np
import pytest

from golden_util import FULL, load, relerr
from oracle import elliptic_ref as er


def element_vertices(g):
    """EXYZ [E,3,8] from the dumped nodal coordinates: the element corners are nodes of the GLL lattice"""
    N = int(g["config"][0])
    Nq, Np = N + 1, (N + 1) ** 3
    E = g["x"].size // Np
    corner = [0, N, N + N * Nq, N * Nq]
    corner = corner + [c + N * Nq * Nq for c in corner]
    out = np.zeros((E, 3, 8))
    for d, key in enumerate(("x", "y", "z")):
        out[:, d, :] = g[key].reshape(E, Np)[:, corner]
    return out


@pytest.mark.parametrize("name", FULL)
def test_trilinear_factors_equal_reference_factors_on_boxes(name):
    g = load(name)
    N = int(g["config"][0])
    ggeo, wJ = er.trilinear_factors(N + 1, element_vertices(g), g["gllz"], g["gllw"])
    assert relerr(ggeo.reshape(-1), g["ggeo"]) < 1e-12
    assert relerr(wJ.reshape(-1), g["wJ"]) < 1e-12
    q = g["q"]
    G2L = g["GlobalToLocal"]
    a = er.ax_trilinear_hex3d(N + 1, element_vertices(g), g["gllz"], g["gllw"], g["D"], float(g["lambda"][0]), q, G2L=G2L)
    b = er.ax_hex3d(N + 1, g["wJ"], g["ggeo"], g["D"], float(g["lambda"][0]), q, G2L=G2L)
    assert relerr(a, b) < 1e-11
