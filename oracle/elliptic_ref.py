"""ORACLE (test infrastructure only - never imported by the product path).

ctypes/numpy front-end of oracle/csrc/oracle.c plus numpy restatements of the host-side setup
steps that feed the elliptic hot path (libParanumal 0.5.0):

  ax_hex3d / operator      solvers/elliptic/okl/ellipticAxHex3D.okl:28-295,
                           solvers/elliptic/src/ellipticOperator.cpp:31-106
  build_diagonal           solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:998-1057 (+ :76-77 gather)
  rhs_sine3d               solvers/elliptic/okl/ellipticRhsHex3D.okl:28-50,
                           solvers/elliptic/okl/ellipticRhsBCHex3D.okl (Dirichlet lift),
                           solvers/elliptic/data/ellipticSine3D.h
  pcg                      libs/linearSolver/linearSolverPCG.cpp:67-171
  coarsen / prolongate     solvers/elliptic/okl/ellipticPreconCoarsenHex3D.okl:209-294,
                           solvers/elliptic/okl/ellipticPreconProlongateHex3D.okl:216-301
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build

_lib = None
_vp = ctypes.c_void_p


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        L.oracle_ax_hex3d.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _vp, _vp]
        L.oracle_ax_hex3d.restype = None
        L.oracle_gather_add.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp]
        L.oracle_gather_add.restype = None
        L.oracle_scatter.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp]
        L.oracle_scatter.restype = None
        L.oracle_operator.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_double,
                                      ctypes.c_int, _vp, _vp, _vp, _vp, _vp]
        L.oracle_operator.restype = None
        L.oracle_pcg.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_double, ctypes.c_int,
                                 _vp, _vp, _vp, _vp, _vp, ctypes.c_double, ctypes.c_int, ctypes.c_int, _vp]
        L.oracle_pcg.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return lib().oracle_num_threads()


def ax_hex3d(Nq, wJ, ggeo, D, lam, q, G2L=None, element_list=None, out=None, Nelements=None):
    """AqL[E*Np] = A_L q (element-local when G2L is None, else q is the gathered vector)."""
    Np = Nq ** 3
    wJ = _c(wJ, np.float64)
    ggeo = _c(ggeo, np.float64)
    D = _c(D, np.float64)
    q = _c(q, np.float64)
    G2L = _c(G2L, np.int32)
    element_list = _c(element_list, np.int32)
    E = wJ.size // Np
    n = len(element_list) if element_list is not None else (E if Nelements is None else Nelements)
    if out is None:
        out = np.zeros(E * Np, dtype=np.float64)
    lib().oracle_ax_hex3d(Nq, n, _p(element_list), _p(G2L), _p(wJ), _p(ggeo), _p(D), float(lam), _p(q), _p(out))
    return out


def trilinear_factors(Nq, EXYZ, gllz, gllw):
    """Geometric factors of trilinear elements evaluated at the GLL nodes, as ellipticPartialAxTrilinearHex3D computes
    them on the fly (solvers/elliptic/okl/ellipticAxHex3D.okl:527-566; delayed J scaling: G = (W/J) * cofactor
    products, GwJ = W*J).  EXYZ [E][3][8] vertex coordinates.  Returns ggeo [E,6,Np], wJ [E,Np]."""
    E = np.asarray(EXYZ).size // 24
    v = np.asarray(EXYZ, dtype=np.float64).reshape(E, 3, 8)
    z, w = np.asarray(gllz, dtype=np.float64), np.asarray(gllw, dtype=np.float64)
    rn = z[None, None, None, :] * np.ones((1, Nq, Nq, 1))
    sn = z[None, None, :, None] * np.ones((1, Nq, 1, Nq))
    tn = z[None, :, None, None] * np.ones((1, 1, Nq, Nq))
    c = lambda d, a: v[:, d, a][:, None, None, None]

    def jac(d):
        fr = 0.125 * ((1 - tn) * (1 - sn) * (c(d, 1) - c(d, 0)) + (1 - tn) * (1 + sn) * (c(d, 2) - c(d, 3))
                      + (1 + tn) * (1 - sn) * (c(d, 5) - c(d, 4)) + (1 + tn) * (1 + sn) * (c(d, 6) - c(d, 7)))
        fs = 0.125 * ((1 - tn) * (1 - rn) * (c(d, 3) - c(d, 0)) + (1 - tn) * (1 + rn) * (c(d, 2) - c(d, 1))
                      + (1 + tn) * (1 - rn) * (c(d, 7) - c(d, 4)) + (1 + tn) * (1 + rn) * (c(d, 6) - c(d, 5)))
        ft = 0.125 * ((1 - rn) * (1 - sn) * (c(d, 4) - c(d, 0)) + (1 + rn) * (1 - sn) * (c(d, 5) - c(d, 1))
                      + (1 + rn) * (1 + sn) * (c(d, 6) - c(d, 2)) + (1 - rn) * (1 + sn) * (c(d, 7) - c(d, 3)))
        return fr, fs, ft
    xr, xs, xt = jac(0)
    yr, ys, yt = jac(1)
    zr, zs, zt = jac(2)
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    rx, ry, rz = (ys * zt - zs * yt), -(xs * zt - zs * xt), (xs * yt - ys * xt)
    sx, sy, sz = -(yr * zt - zr * yt), (xr * zt - zr * xt), -(xr * yt - yr * xt)
    tx, ty, tz = (yr * zs - zr * ys), -(xr * zs - zr * xs), (xr * ys - yr * xs)
    W = w[None, None, None, :] * w[None, None, :, None] * w[None, :, None, None]
    sc = W / J
    G = np.stack([sc * (rx * rx + ry * ry + rz * rz), sc * (rx * sx + ry * sy + rz * sz), sc * (rx * tx + ry * ty + rz * tz),
                  sc * (sx * sx + sy * sy + sz * sz), sc * (sx * tx + sy * ty + sz * tz), sc * (tx * tx + ty * ty + tz * tz)],
                 axis=1)
    return G.reshape(E, 6, Nq ** 3), (W * J).reshape(E, Nq ** 3)


def ax_trilinear_hex3d(Nq, EXYZ, gllz, gllw, D, lam, q, G2L=None, element_list=None):
    """ellipticPartialAxTrilinearHex3D (solvers/elliptic/okl/ellipticAxHex3D.okl:440-627): the same tensor-product
    apply as ellipticPartialAxHex3D with the factors of trilinear_factors()."""
    ggeo, wJ = trilinear_factors(Nq, EXYZ, gllz, gllw)
    return ax_hex3d(Nq, wJ, ggeo, D, lam, q, G2L=G2L, element_list=element_list)


def gather_add(rowStarts, colIds, v, nrows=None):
    rowStarts = _c(rowStarts, np.int32)
    colIds = _c(colIds, np.int32)
    v = _c(v, np.float64)
    nrows = len(rowStarts) - 1 if nrows is None else nrows
    gv = np.zeros(nrows, dtype=np.float64)
    lib().oracle_gather_add(nrows, _p(rowStarts), _p(colIds), _p(v), _p(gv))
    return gv


def operator(Nq, G2L, wJ, ggeo, D, lam, rowStartsT, colIdsT, q):
    """Single-rank elliptic_t::Operator (C0): returns Aq[Ngather]."""
    AqL = ax_hex3d(Nq, wJ, ggeo, D, lam, q, G2L=G2L)
    return gather_add(rowStartsT, colIdsT, AqL)


def pcg(Nq, G2L, wJ, ggeo, D, lam, rowStartsT, colIdsT, invDiag, x, r, tol=1e-8, maxit=5000, flexible=False):
    """Returns (iterations, x, residual history [iterations+1])."""
    Np = Nq ** 3
    G2L = _c(G2L, np.int32)
    wJ = _c(wJ, np.float64)
    ggeo = _c(ggeo, np.float64)
    D = _c(D, np.float64)
    rowStartsT = _c(rowStartsT, np.int32)
    colIdsT = _c(colIdsT, np.int32)
    invDiag = _c(invDiag, np.float64)
    x = np.array(x, dtype=np.float64)
    r = np.array(r, dtype=np.float64)
    N = len(rowStartsT) - 1
    E = wJ.size // Np
    hist = np.zeros(maxit + 1, dtype=np.float64)
    it = lib().oracle_pcg(Nq, E, _p(G2L), _p(wJ), _p(ggeo), _p(D), float(lam), N, _p(rowStartsT), _p(colIdsT),
                          _p(invDiag), _p(x), _p(r), float(tol), int(maxit), int(flexible), _p(hist))
    return it, x, hist[: it + 1]


def build_diagonal_local(Nq, ggeo, wJ, D, lam, mapB, all_neumann_boost=0.0):
    """BuildOperatorDiagonalContinuousHex3D: element-local diagonal diagAL[E*Np]."""
    Np = Nq ** 3
    E = wJ.size // Np
    G = np.asarray(ggeo).reshape(E, 6, Nq, Nq, Nq)  # [e, comp, k, j, i]
    D = np.asarray(D).reshape(Nq, Nq)
    dd = np.diag(D)
    A = np.zeros((E, Nq, Nq, Nq))
    di = dd[None, None, None, :]
    dj = dd[None, None, :, None]
    dk = dd[None, :, None, None]
    A += 2 * G[:, 1] * di * dj
    A += 2 * G[:, 2] * di * dk
    A += 2 * G[:, 4] * dj * dk
    D2 = D * D  # D2[k, n] = D[k*Nq+n]^2 ; reference uses D[nx + k*Nq]
    A += np.einsum("ezyk,kx->ezyx", G[:, 0], D2)
    A += np.einsum("ezkx,ky->ezyx", G[:, 3], D2)
    A += np.einsum("ekyx,kz->ezyx", G[:, 5], D2)
    A += np.asarray(wJ).reshape(E, Nq, Nq, Nq) * lam
    A = A.reshape(-1)
    masked = np.asarray(mapB).reshape(-1) == 1
    A[~masked] += all_neumann_boost
    A[masked] = 1.0
    return A


def rhs_sine3d(Nq, mesh_x, mesh_y, mesh_z, wJ, ggeo, D, lam, mapB):
    """Element-local rhs rL = wJ*f - A_L(u_D on Dirichlet nodes) for data/ellipticSine3D.h."""
    PI = 3.14159265358979323846
    x = np.asarray(mesh_x).reshape(-1)
    y = np.asarray(mesh_y).reshape(-1)
    z = np.asarray(mesh_z).reshape(-1)
    f = (3 * PI * PI + lam) * np.sin(PI * x) * np.sin(PI * y) * np.sin(PI * z)
    rL = np.asarray(wJ).reshape(-1) * f
    uD = np.where(np.asarray(mapB).reshape(-1) == 1, np.sin(PI * x) * np.sin(PI * y) * np.sin(PI * z), 0.0)
    if np.any(uD != 0.0):
        rL = rL - ax_hex3d(Nq, wJ, ggeo, D, lam, uD)
    return rL


def splitmix_uniform(seed, n):
    """Counter-based uniform(-1,1) stream shared with oracle/refbuild/dump_driver.cpp."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0


def coarsen_hex3d(NqF, NqC, P, qf_local):
    """qc = (P^T x P^T x P^T) qf per element; P is [NqF, NqC] (ellipticPreconCoarsenHex3D.okl)."""
    E = qf_local.size // NqF ** 3
    Q = qf_local.reshape(E, NqF, NqF, NqF)
    return np.einsum("ia,jb,kc,ekji->ecba", P, P, P, Q).reshape(-1)


def prolongate_hex3d(NqF, NqC, P, qc_local):
    """qf = (P x P x P) qc per element (ellipticPreconProlongateHex3D.okl)."""
    E = qc_local.size // NqC ** 3
    Q = qc_local.reshape(E, NqC, NqC, NqC)
    return np.einsum("ia,jb,kc,ecba->ekji", P, P, P, Q).reshape(-1)
