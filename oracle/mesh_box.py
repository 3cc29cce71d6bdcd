"""ORACLE (test infrastructure only - never imported by the product path).

numpy restatement of the reference's structured hexahedral BOX mesh producer, i.e. the
immutable inputs of the elliptic hot path.  Everything here follows libParanumal 0.5.0:

  rank grid            libs/core/rankDecomp.cpp:52-91 (Factor3), :166-214 (RankDecomp3)
  box elements/verts   libs/mesh/meshSetupBoxHex3D.cpp:31-157
  GLL nodes / weights  libs/mesh/meshBasis1D.cpp:260-292 (JacobiGLL), :294-346 (JacobiGQ)
  D matrix             libs/mesh/meshBasis1D.cpp:114-136 (Dmatrix1D): D[i*Nq+m] = phi'_m(r_i)
  node ordering        libs/mesh/meshBasisHex3D.cpp:34-58  (n = i + j*Nq + k*Nq^2)
  physical nodes       libs/mesh/meshPhysicalNodesHex3D.cpp:31-110 (trilinear map)
  geometric factors    libs/mesh/meshGeometricFactorsHex3D.cpp:94-174 (ggeo ids G00..G22 = 0..5)
  global node ids      libs/mesh/meshConnectNodes.cpp:32-117 (min of 1+local index over copies)
  node boundary flags  libs/mesh/meshConnectNodes.cpp:49-64,92-101
  gather element lists libs/mesh/meshGatherScatterSetup.cpp:32-130

The reference obtains GLL nodes from a LAPACK eigen-solve and D from a LAPACK solve; here they
come from Newton iteration on Legendre polynomials and the closed-form Lagrange derivative,
which agree to ~1e-15 (checked against the reference dump in tests/test_oracle_golden.py).
For 1e-12 operator parity D/ggeo/wJ always cross the boundary as *data* (SURVEY section 7).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------- rank grid
def factor3(n: int):
    """libs/core/rankDecomp.cpp:52-91."""
    nx = int(round(n ** (1.0 / 3.0))) if n > 0 else 1
    nx = max(nx, 1)
    ny = nz = 1
    while nx < n:
        if n % nx == 0:
            f = n // nx
            ny = int(round(math.sqrt(f)))
            while ny < f:
                if f % ny == 0:
                    nz = f // ny
                    if ny > nx:
                        nx, ny = ny, nx
                    if nz > ny:
                        ny, nz = nz, ny
                    if ny > nx:
                        nx, ny = ny, nx
                    return nx, ny, nz
                ny += 1
            ny, nz = f, 1
            if ny > nx:
                nx, ny = ny, nx
            return nx, ny, nz
        nx += 1
    return n, 1, 1


def _max_prime_factor(n: int) -> int:
    p = -1
    while n % 2 == 0:
        p = 2
        n >>= 1
    i = 3
    while i * i <= n:
        while n % i == 0:
            p = i
            n //= i
        i += 2
    if n > 2:
        p = n
    return p


def rank_decomp3(sx: int, sy: int, sz: int, rank: int):
    """libs/core/rankDecomp.cpp:166-214."""
    size = sx * sy * sz
    if size == 1:
        return 0, 0, 0
    if sz >= sx and sz >= sy:
        p = _max_prime_factor(sz)
        csize = size // p
        rx, ry, crz = rank_decomp3(sx, sy, sz // p, rank % csize)
        return rx, ry, crz + (rank // csize) * (sz // p)
    if sy >= sx and sy >= sz:
        p = _max_prime_factor(sy)
        csize = size // p
        rx, cry, rz = rank_decomp3(sx, sy // p, sz, rank % csize)
        return rx, cry + (rank // csize) * (sy // p), rz
    p = _max_prime_factor(sx)
    csize = size // p
    crx, ry, rz = rank_decomp3(sx // p, sy, sz, rank % csize)
    return crx + (rank // csize) * (sx // p), ry, rz


# ----------------------------------------------------------------------------- 1-D basis
def _legendre(n: int, x: np.ndarray):
    """P_n(x) and P_n'(x) by the three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = n * (x * p1 - p0) / (x * x - 1.0)
    return p1, dp


def gll_nodes_weights(N: int):
    """Gauss-Lobatto-Legendre nodes/weights on [-1,1] (JacobiGLL, meshBasis1D.cpp:260-292)."""
    if N == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    # Chebyshev-Lobatto initial guess, Newton on (1-x^2) P_N'(x)
    x = -np.cos(np.pi * np.arange(N + 1) / N)
    for _ in range(100):
        xi = x[1:-1]
        pn, dpn = _legendre(N, xi)
        # q(x) = P_N'(x);  q'(x) = (2x P_N' - N(N+1) P_N)/(1-x^2)
        d2 = (2.0 * xi * dpn - N * (N + 1) * pn) / (1.0 - xi * xi)
        dx = dpn / d2
        x[1:-1] = xi - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    pn, _ = _legendre(N, x)
    w = 2.0 / (N * (N + 1) * pn * pn)
    return x, w


def dmatrix1d(N: int, r: np.ndarray):
    """D[i, m] = l_m'(r_i) for the Lagrange basis on GLL nodes r (Dmatrix1D)."""
    Nq = N + 1
    pn, _ = _legendre(N, r)
    D = np.zeros((Nq, Nq))
    for i in range(Nq):
        for m in range(Nq):
            if i != m:
                D[i, m] = pn[i] / (pn[m] * (r[i] - r[m]))
    D[0, 0] = -N * (N + 1) / 4.0
    D[N, N] = N * (N + 1) / 4.0
    return D


def degree_raise_1d(Nc: int, Nf: int):
    """P[NqF, NqC]: interpolate the degree-Nc GLL Lagrange basis at the degree-Nf GLL nodes
    (mesh_t::DegreeRaiseMatrix1D, libs/mesh/meshBasis1D.cpp)."""
    rc, _ = gll_nodes_weights(Nc)
    rf, _ = gll_nodes_weights(Nf)
    P = np.ones((Nf + 1, Nc + 1))
    for m in range(Nc + 1):
        for l in range(Nc + 1):
            if l != m:
                P[:, m] *= (rf - rc[l]) / (rc[m] - rc[l])
    return P


# ----------------------------------------------------------------------------- the mesh
@dataclass
class BoxHexMesh:
    N: int
    NX: int
    NY: int
    NZ: int
    rank: int = 0
    size: int = 1
    boundary_flag: int = 1
    dims: tuple = (1.0, 1.0, 1.0)
    # filled by build()
    Nq: int = 0
    Np: int = 0
    Nelements: int = 0
    nx: int = 0
    ny: int = 0
    nz: int = 0
    offsets: tuple = (0, 0, 0)
    gllz: np.ndarray = None
    gllw: np.ndarray = None
    D: np.ndarray = None
    x: np.ndarray = None
    y: np.ndarray = None
    z: np.ndarray = None
    ggeo: np.ndarray = None  # [E, 6, Np]
    wJ: np.ndarray = None  # [E, Np]
    globalIds: np.ndarray = None  # [E*Np] int64
    mapB: np.ndarray = None  # [E*Np] int32 (mesh flag: boundary_flag on the boundary else -1)
    localGatherElementList: np.ndarray = None
    globalGatherElementList: np.ndarray = None
    extra: dict = field(default_factory=dict)


def _local_box(NX, NY, NZ, size, rank):
    sx, sy, sz = factor3(size)
    rx, ry, rz = rank_decomp3(sx, sy, sz, rank)
    nx = NX // sx + (1 if rx < NX % sx else 0)
    ny = NY // sy + (1 if ry < NY % sy else 0)
    nz = NZ // sz + (1 if rz < NZ % sz else 0)
    ox = rx * (NX // sx) + min(rx, NX % sx)
    oy = ry * (NY // sy) + min(ry, NY % sy)
    oz = rz * (NZ // sz) + min(rz, NZ % sz)
    return (nx, ny, nz), (ox, oy, oz)


def _lattice_index(N, NX, NY, NZ, nloc, off, periodic):
    """Global GLL-lattice index of every local node [E, Np] (element e = i + j*nx + k*nx*ny)."""
    nx, ny, nz = nloc
    ox, oy, oz = off
    Nq = N + 1
    LX = NX * N if periodic else NX * N + 1
    LY = NY * N if periodic else NY * N + 1
    LZ = NZ * N if periodic else NZ * N + 1
    ex = np.arange(nx)[None, None, :, None, None, None]
    ey = np.arange(ny)[None, :, None, None, None, None]
    ez = np.arange(nz)[:, None, None, None, None, None]
    i = np.arange(Nq)[None, None, None, None, None, :]
    j = np.arange(Nq)[None, None, None, None, :, None]
    k = np.arange(Nq)[None, None, None, :, None, None]
    gi = ((ex + ox) * N + i) % LX
    gj = ((ey + oy) * N + j) % LY
    gk = ((ez + oz) * N + k) % LZ
    lat = (gi + LX * (gj + LY * gk)).astype(np.int64)
    lat = np.broadcast_to(lat, (nz, ny, nx, Nq, Nq, Nq)).reshape(nx * ny * nz, Nq ** 3)
    on_bdry = np.zeros((nz, ny, nx, Nq, Nq, Nq), dtype=bool)
    if not periodic:
        on_bdry = ((gi == 0) | (gi == LX - 1)) | ((gj == 0) | (gj == LY - 1)) | ((gk == 0) | (gk == LZ - 1))
        on_bdry = np.broadcast_to(on_bdry, (nz, ny, nx, Nq, Nq, Nq))
    return lat, on_bdry.reshape(nx * ny * nz, Nq ** 3), LX * LY * LZ


def build_box_hex_mesh(N, NX, NY, NZ, rank=0, size=1, boundary_flag=1, dims=(1.0, 1.0, 1.0),
                       geometry=True) -> BoxHexMesh:
    """Restates mesh_t::Setup for MESH FILE=BOX, ELEMENT TYPE=12 with BOX GLOBAL NX/NY/NZ given."""
    m = BoxHexMesh(N=N, NX=NX, NY=NY, NZ=NZ, rank=rank, size=size, boundary_flag=boundary_flag, dims=dims)
    periodic = boundary_flag == -1
    Nq = N + 1
    Np = Nq ** 3
    m.Nq, m.Np = Nq, Np
    (nx, ny, nz), (ox, oy, oz) = _local_box(NX, NY, NZ, size, rank)
    m.nx, m.ny, m.nz, m.offsets = nx, ny, nz, (ox, oy, oz)
    E = nx * ny * nz
    m.Nelements = E
    m.gllz, m.gllw = gll_nodes_weights(N)
    m.D = dmatrix1d(N, m.gllz)

    # ---- global ids: min over all copies (all ranks) of 1 + local index + rank offset
    lat, on_bdry, nlat = _lattice_index(N, NX, NY, NZ, (nx, ny, nz), (ox, oy, oz), periodic)
    latmin = np.full(nlat, np.iinfo(np.int64).max, dtype=np.int64)
    start = 0
    my_start = 0
    for r in range(size):
        (rnx, rny, rnz), roff = _local_box(NX, NY, NZ, size, r)
        rE = rnx * rny * rnz
        if r == rank:
            my_start = start
            rlat = lat
        else:
            rlat, _, _ = _lattice_index(N, NX, NY, NZ, (rnx, rny, rnz), roff, periodic)
        ids = 1 + start + np.arange(rE * Np, dtype=np.int64)
        np.minimum.at(latmin, rlat.reshape(-1), ids)
        start += rE * Np
    m.globalIds = latmin[lat.reshape(-1)]
    m.extra["gatherNodeStart"] = my_start
    m.extra["lattice"] = lat
    m.mapB = np.where(on_bdry.reshape(-1), boundary_flag, -1).astype(np.int32)

    # ---- gather element lists: an element is "global" if any vertex is shared with another rank
    if size == 1:
        m.localGatherElementList = np.arange(E, dtype=np.int32)
        m.globalGatherElementList = np.zeros(0, dtype=np.int32)
    else:
        # vertex lattice (N=1 lattice) owner ranks: min/max rank touching each vertex
        vminr = {}
        NnX = NX if periodic else NX + 1
        NnY = NY if periodic else NY + 1
        NnZ = NZ if periodic else NZ + 1
        vmin = np.full(NnX * NnY * NnZ, size, dtype=np.int64)
        vmax = np.full(NnX * NnY * NnZ, -1, dtype=np.int64)
        myv = None
        for r in range(size):
            (rnx, rny, rnz), roff = _local_box(NX, NY, NZ, size, r)
            vl, _, _ = _lattice_index(1, NX, NY, NZ, (rnx, rny, rnz), roff, periodic)
            np.minimum.at(vmin, vl.reshape(-1), r)
            np.maximum.at(vmax, vl.reshape(-1), r)
            if r == rank:
                myv = vl
        is_halo = np.any((vmin[myv] != rank) | (vmax[myv] != rank), axis=1)
        m.localGatherElementList = np.nonzero(~is_halo)[0].astype(np.int32)
        m.globalGatherElementList = np.nonzero(is_halo)[0].astype(np.int32)

    if not geometry:
        return m

    # ---- physical nodes (trilinear map of the 8 vertices)
    DIMX, DIMY, DIMZ = dims
    dx, dy, dz = DIMX / NX, DIMY / NY, DIMZ / NZ
    X0 = -DIMX / 2.0 + ox * dx
    Y0 = -DIMY / 2.0 + oy * dy
    Z0 = -DIMZ / 2.0 + oz * dz
    e = np.arange(E)
    ei, ej, ek = e % nx, (e // nx) % ny, e // (nx * ny)
    x0 = X0 + dx * ei
    y0 = Y0 + dy * ej
    z0 = Z0 + dz * ek
    r = np.tile(m.gllz, Nq * Nq)
    s = np.tile(np.repeat(m.gllz, Nq), Nq)
    t = np.repeat(m.gllz, Nq * Nq)
    ex = np.stack([x0, x0 + dx, x0 + dx, x0, x0, x0 + dx, x0 + dx, x0], axis=1)
    ey = np.stack([y0, y0, y0 + dy, y0 + dy, y0, y0, y0 + dy, y0 + dy], axis=1)
    ez = np.stack([z0, z0, z0, z0, z0 + dz, z0 + dz, z0 + dz, z0 + dz], axis=1)
    shp = np.stack([
        0.125 * (1 - r) * (1 - s) * (1 - t), 0.125 * (1 + r) * (1 - s) * (1 - t),
        0.125 * (1 + r) * (1 + s) * (1 - t), 0.125 * (1 - r) * (1 + s) * (1 - t),
        0.125 * (1 - r) * (1 - s) * (1 + t), 0.125 * (1 + r) * (1 - s) * (1 + t),
        0.125 * (1 + r) * (1 + s) * (1 + t), 0.125 * (1 - r) * (1 + s) * (1 + t)], axis=0)  # [8, Np]
    m.x = ex @ shp
    m.y = ey @ shp
    m.z = ez @ shp
    m.ggeo, m.wJ = geometric_factors_hex3d(m.D, m.gllw, m.x, m.y, m.z)
    return m


def geometric_factors_hex3d(D, gllw, x, y, z):
    """ggeo[E,6,Np], wJ[E,Np] from nodal coordinates (meshGeometricFactorsHex3D.cpp:94-174)."""
    Nq = D.shape[0]
    E = x.shape[0]
    X = x.reshape(E, Nq, Nq, Nq)
    Y = y.reshape(E, Nq, Nq, Nq)
    Z = z.reshape(E, Nq, Nq, Nq)

    def ddr(F):
        return np.einsum("im,ekjm->ekji", D, F)

    def dds(F):
        return np.einsum("jm,ekmi->ekji", D, F)

    def ddt(F):
        return np.einsum("km,emji->ekji", D, F)

    xr, xs, xt = ddr(X), dds(X), ddt(X)
    yr, ys, yt = ddr(Y), dds(Y), ddt(Y)
    zr, zs, zt = ddr(Z), dds(Z), ddt(Z)
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    rx, ry, rz = (ys * zt - zs * yt) / J, -(xs * zt - zs * xt) / J, (xs * yt - ys * xt) / J
    sx, sy, sz = -(yr * zt - zr * yt) / J, (xr * zt - zr * xt) / J, -(xr * yt - yr * xt) / J
    tx, ty, tz = (yr * zs - zr * ys) / J, -(xr * zs - zr * xs) / J, (xr * ys - yr * xs) / J
    W = gllw[None, None, None, :] * gllw[None, None, :, None] * gllw[None, :, None, None]
    JW = J * W
    g = np.stack([
        JW * (rx * rx + ry * ry + rz * rz), JW * (rx * sx + ry * sy + rz * sz),
        JW * (rx * tx + ry * ty + rz * tz), JW * (sx * sx + sy * sy + sz * sz),
        JW * (sx * tx + sy * ty + sz * tz), JW * (tx * tx + ty * ty + tz * tz)], axis=1)
    return np.ascontiguousarray(g.reshape(E, 6, Nq ** 3)), np.ascontiguousarray(JW.reshape(E, Nq ** 3))


def masked_global_ids(mesh: BoxHexMesh, bc_type=(0, 1, 2)):
    """elliptic_t::BoundarySetup mask (solvers/elliptic/src/ellipticBoundarySetup.cpp:55-86):
    returns (mapB elliptic [0/BC], maskedGlobalIds with Dirichlet nodes zeroed)."""
    mapB = np.zeros_like(mesh.mapB)
    pos = mesh.mapB > 0
    mapB[pos] = np.asarray(bc_type, dtype=np.int32)[mesh.mapB[pos]]
    ids = mesh.globalIds.copy()
    ids[mapB == 1] = 0
    return mapB, ids
