"""Synthetic structured BOX hexahedral mesh producer for the harness (bench.py, tests, smoke).

The mesh is an *input producer* for the hot path (out of scope to
re-implement in general); this module only covers what the BASELINE configs need:
MESH FILE=BOX, ELEMENT TYPE=12, ELEMENT MAP=ISOPARAMETRIC on [-DIM/2, DIM/2]^3 with the
reference's rank decomposition, node numbering and per-node geometric factors, built with torch
so that the 64^3 x 512-node arrays are generated directly in HBM.

Reference behaviour mirrored (libParanumal 0.5.0):
  libs/core/rankDecomp.cpp:52-91,166-214        Factor3 / RankDecomp3
  libs/mesh/meshSetupBoxHex3D.cpp:31-157        local boxes, element order e = i + j*nx + k*nx*ny
  libs/mesh/meshBasis1D.cpp:114-136,260-292     GLL nodes / weights / D
  libs/mesh/meshPhysicalNodesHex3D.cpp          trilinear nodes
  libs/mesh/meshGeometricFactorsHex3D.cpp:94-174 ggeo (G00,G01,G02,G11,G12,G22) and wJ per node
  libs/mesh/meshConnectNodes.cpp:32-117         global ids = min over copies of 1+index+rank offset
  libs/mesh/meshGatherScatterSetup.cpp:32-130   local / global gather element lists
  solvers/elliptic/src/ellipticBoundarySetup.cpp:55-86  Dirichlet mask -> id 0
"""
from __future__ import annotations

import math

import numpy as np
import torch


def factor3(n):
    nx = max(int(round(n ** (1.0 / 3.0))), 1)
    ny = nz = 1
    while nx < n:
        if n % nx == 0:
            f = n // nx
            ny = int(round(math.sqrt(f)))
            while ny < f:
                if f % ny == 0:
                    nz = f // ny
                    if ny > nx: nx, ny = ny, nx
                    if nz > ny: ny, nz = nz, ny
                    if ny > nx: nx, ny = ny, nx
                    return nx, ny, nz
                ny += 1
            ny, nz = f, 1
            if ny > nx: nx, ny = ny, nx
            return nx, ny, nz
        nx += 1
    return n, 1, 1


def _maxprime(n):
    p = -1
    while n % 2 == 0:
        p, n = 2, n >> 1
    i = 3
    while i * i <= n:
        while n % i == 0:
            p, n = i, n // i
        i += 2
    return n if n > 2 else p


def rank_decomp3(sx, sy, sz, rank):
    size = sx * sy * sz
    if size == 1:
        return 0, 0, 0
    if sz >= sx and sz >= sy:
        p = _maxprime(sz); cs = size // p
        rx, ry, c = rank_decomp3(sx, sy, sz // p, rank % cs)
        return rx, ry, c + (rank // cs) * (sz // p)
    if sy >= sx and sy >= sz:
        p = _maxprime(sy); cs = size // p
        rx, c, rz = rank_decomp3(sx, sy // p, sz, rank % cs)
        return rx, c + (rank // cs) * (sy // p), rz
    p = _maxprime(sx); cs = size // p
    c, ry, rz = rank_decomp3(sx // p, sy, sz, rank % cs)
    return c + (rank // cs) * (sx // p), ry, rz


def gll(N):
    """GLL nodes, weights and D[i, m] = l_m'(r_i) (float64 numpy)."""
    def leg(n, x):
        p0, p1 = np.ones_like(x), x.copy()
        if n == 0:
            return p0
        for k in range(2, n + 1):
            p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        return p1
    if N == 1:
        r = np.array([-1.0, 1.0])
    else:
        r = -np.cos(np.pi * np.arange(N + 1) / N)
        for _ in range(100):
            xi = r[1:-1]
            pn, pm = leg(N, xi), leg(N - 1, xi)
            dpn = N * (pm - xi * pn) / (1 - xi * xi)
            d2 = (2 * xi * dpn - N * (N + 1) * pn) / (1 - xi * xi)
            dx = dpn / d2
            r[1:-1] = xi - dx
            if np.max(np.abs(dx)) < 1e-16:
                break
        r[0], r[-1] = -1.0, 1.0
        r = 0.5 * (r - r[::-1])
    pn = leg(N, r)
    w = 2.0 / (N * (N + 1) * pn * pn)
    Nq = N + 1
    D = np.zeros((Nq, Nq))
    for i in range(Nq):
        for m in range(Nq):
            if i != m:
                D[i, m] = pn[i] / (pn[m] * (r[i] - r[m]))
    D[0, 0] = -N * (N + 1) / 4.0
    D[N, N] = N * (N + 1) / 4.0
    return r, w, D


def local_box(NX, NY, NZ, size, rank):
    sx, sy, sz = factor3(size)
    rx, ry, rz = rank_decomp3(sx, sy, sz, rank)
    n = (NX // sx + (1 if rx < NX % sx else 0), NY // sy + (1 if ry < NY % sy else 0),
         NZ // sz + (1 if rz < NZ % sz else 0))
    o = (rx * (NX // sx) + min(rx, NX % sx), ry * (NY // sy) + min(ry, NY % sy), rz * (NZ // sz) + min(rz, NZ % sz))
    return n, o


def _lattice(N, NX, NY, NZ, nloc, off, periodic, device):
    """[E, Np] int64 lattice index + boundary mask of the local nodes (element e = i + j*nx + k*nx*ny)."""
    nx, ny, nz = nloc
    ox, oy, oz = off
    Nq = N + 1
    LX = NX * N + (0 if periodic else 1)
    LY = NY * N + (0 if periodic else 1)
    LZ = NZ * N + (0 if periodic else 1)
    ar = lambda n: torch.arange(n, device=device, dtype=torch.int64)
    gi = ((ar(nx) + ox)[:, None] * N + ar(Nq)[None, :]) % LX  # [nx, Nq]
    gj = ((ar(ny) + oy)[:, None] * N + ar(Nq)[None, :]) % LY
    gk = ((ar(nz) + oz)[:, None] * N + ar(Nq)[None, :]) % LZ
    # dims: [ez, ey, ex, k, j, i]
    GI = gi[None, None, :, None, None, :]
    GJ = gj[None, :, None, None, :, None]
    GK = gk[:, None, None, :, None, None]
    lat = (GI + LX * (GJ + LY * GK)).reshape(nx * ny * nz, Nq ** 3)
    if periodic:
        bd = torch.zeros_like(lat, dtype=torch.bool)
    else:
        bd = ((GI == 0) | (GI == LX - 1) | (GJ == 0) | (GJ == LY - 1) | (GK == 0) | (GK == LZ - 1))
        bd = bd.expand(nz, ny, nx, Nq, Nq, Nq).reshape(nx * ny * nz, Nq ** 3)
    return lat, bd, LX * LY * LZ


class BoxMesh:
    """Per-rank arrays of the box mesh (torch tensors on `device`)."""

    def __init__(self, N, NX, NY, NZ, rank=0, size=1, boundary_flag=1, dims=(1.0, 1.0, 1.0), device="cpu",
                 geometry=True, coords=False):
        self.N, self.Nq, self.Np = N, N + 1, (N + 1) ** 3
        self.NX, self.NY, self.NZ = NX, NY, NZ
        self.rank, self.size, self.boundary_flag, self.dims = rank, size, boundary_flag, dims
        self.device = torch.device(device)
        periodic = boundary_flag == -1
        (nx, ny, nz), off = local_box(NX, NY, NZ, size, rank)
        self.nloc, self.off = (nx, ny, nz), off
        E = nx * ny * nz
        self.Nelements = E
        Np = self.Np
        r, w, D = gll(N)
        self.gllz, self.gllw, self.D_host = r, w, D
        self.D = torch.from_numpy(D.reshape(-1).copy()).to(self.device)

        # ---- global ids (min over every copy on every rank) and boundary flags
        lat, bd, nlat = _lattice(N, NX, NY, NZ, (nx, ny, nz), off, periodic, self.device)
        latmin = torch.full((nlat,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=self.device)
        start = 0
        for rr in range(size):
            nl, of = local_box(NX, NY, NZ, size, rr)
            rE = nl[0] * nl[1] * nl[2]
            rlat = lat if rr == rank else _lattice(N, NX, NY, NZ, nl, of, periodic, self.device)[0]
            ids = torch.arange(1 + start, 1 + start + rE * Np, dtype=torch.int64, device=self.device)
            latmin.scatter_reduce_(0, rlat.reshape(-1), ids, reduce="amin", include_self=True)
            start += rE * Np
            del rlat, ids
        self.globalIds = latmin[lat.reshape(-1)]
        del latmin
        self.mapB = torch.where(bd.reshape(-1), torch.tensor(boundary_flag, dtype=torch.int32, device=self.device),
                                torch.tensor(-1, dtype=torch.int32, device=self.device))
        del lat, bd

        # ---- gather element lists
        if size == 1:
            self.localGatherElementList = torch.arange(E, dtype=torch.int32, device=self.device)
            self.globalGatherElementList = torch.zeros(0, dtype=torch.int32, device=self.device)
        else:
            nv = (NX + (0 if periodic else 1)) * (NY + (0 if periodic else 1)) * (NZ + (0 if periodic else 1))
            vmin = torch.full((nv,), size, dtype=torch.int64, device=self.device)
            vmax = torch.full((nv,), -1, dtype=torch.int64, device=self.device)
            myv = None
            for rr in range(size):
                nl, of = local_box(NX, NY, NZ, size, rr)
                vl = _lattice(1, NX, NY, NZ, nl, of, periodic, self.device)[0]
                rv = torch.full((vl.numel(),), rr, dtype=torch.int64, device=self.device)
                vmin.scatter_reduce_(0, vl.reshape(-1), rv, reduce="amin", include_self=True)
                vmax.scatter_reduce_(0, vl.reshape(-1), rv, reduce="amax", include_self=True)
                if rr == rank:
                    myv = vl
            is_halo = ((vmin[myv] != rank) | (vmax[myv] != rank)).any(dim=1)
            self.localGatherElementList = torch.nonzero(~is_halo).reshape(-1).to(torch.int32)
            self.globalGatherElementList = torch.nonzero(is_halo).reshape(-1).to(torch.int32)

        self.x = self.y = self.z = None
        self.ggeo = self.wJ = None
        if geometry:
            self._geometry(coords)

    def element_vertices(self, e0=0, e1=None):
        """EX, EY, EZ [n, 8] of elements e0..e1 (vertex order of the reference: libs/mesh/meshSetupBoxHex3D.cpp)"""
        nx, ny, nz = self.nloc
        ox, oy, oz = self.off
        DIMX, DIMY, DIMZ = self.dims
        dx, dy, dz = DIMX / self.NX, DIMY / self.NY, DIMZ / self.NZ
        X0, Y0, Z0 = -DIMX / 2.0 + ox * dx, -DIMY / 2.0 + oy * dy, -DIMZ / 2.0 + oz * dz
        e = torch.arange(e0, self.Nelements if e1 is None else e1, device=self.device)
        x0 = X0 + dx * (e % nx).double()
        y0 = Y0 + dy * ((e // nx) % ny).double()
        z0 = Z0 + dz * (e // (nx * ny)).double()
        ex = torch.stack([x0, x0 + dx, x0 + dx, x0, x0, x0 + dx, x0 + dx, x0], dim=1)
        ey = torch.stack([y0, y0, y0 + dy, y0 + dy, y0, y0, y0 + dy, y0 + dy], dim=1)
        ez = torch.stack([z0, z0, z0, z0, z0 + dz, z0 + dz, z0 + dz, z0 + dz], dim=1)
        return ex.contiguous(), ey.contiguous(), ez.contiguous()

    def _geometry_device(self, keep_coords, chunk=32768):
        """physical nodes + geometric factors by the library's device kernels (libp_mesh_physical_nodes_hex3d,
        libp_mesh_geometric_factors_hex3d), chunked over elements so that the temporaries x, y, z stay small"""
        from . import api
        Nq, Np, E, dev = self.Nq, self.Np, self.Nelements, self.device
        gz = torch.from_numpy(self.gllz).to(dev)
        gw = torch.from_numpy(self.gllw).to(dev)
        self.ggeo = torch.empty((E, 6, Np), dtype=torch.float64, device=dev)
        self.wJ = torch.empty((E, Np), dtype=torch.float64, device=dev)
        if keep_coords:
            self.x = torch.empty((E, Np), dtype=torch.float64, device=dev)
            self.y = torch.empty_like(self.x)
            self.z = torch.empty_like(self.x)
        else:
            n = min(chunk, max(E, 1))
            tx, ty, tz = (torch.empty((n, Np), dtype=torch.float64, device=dev) for _ in range(3))
        for e0 in range(0, E, chunk):
            e1 = min(E, e0 + chunk)
            ex, ey, ez = self.element_vertices(e0, e1)
            if keep_coords:
                x, y, z = self.x[e0:e1], self.y[e0:e1], self.z[e0:e1]
            else:
                x, y, z = tx[: e1 - e0], ty[: e1 - e0], tz[: e1 - e0]
            api.mesh_physical_nodes_hex3d(Nq, e1 - e0, ex, ey, ez, gz, x, y, z)
            api.mesh_geometric_factors_hex3d(Nq, e1 - e0, x, y, z, self.D, gw, self.ggeo[e0:e1], self.wJ[e0:e1])

    # physical nodes + geometric factors, chunked over elements to bound temporaries
    def _geometry(self, keep_coords, chunk=16384):
        if self.device.type == "cuda":
            return self._geometry_device(keep_coords)
        N, Nq, Np, E = self.N, self.Nq, self.Np, self.Nelements
        nx, ny, nz = self.nloc
        ox, oy, oz = self.off
        dev = self.device
        DIMX, DIMY, DIMZ = self.dims
        dx, dy, dz = DIMX / self.NX, DIMY / self.NY, DIMZ / self.NZ
        X0, Y0, Z0 = -DIMX / 2.0 + ox * dx, -DIMY / 2.0 + oy * dy, -DIMZ / 2.0 + oz * dz
        gz = torch.from_numpy(self.gllz).to(dev)
        gw = torch.from_numpy(self.gllw).to(dev)
        D = torch.from_numpy(self.D_host).to(dev)
        r = gz.repeat(Nq * Nq)
        s = gz.repeat_interleave(Nq).repeat(Nq)
        t = gz.repeat_interleave(Nq * Nq)
        shp = torch.stack([0.125 * (1 - r) * (1 - s) * (1 - t), 0.125 * (1 + r) * (1 - s) * (1 - t),
                           0.125 * (1 + r) * (1 + s) * (1 - t), 0.125 * (1 - r) * (1 + s) * (1 - t),
                           0.125 * (1 - r) * (1 - s) * (1 + t), 0.125 * (1 + r) * (1 - s) * (1 + t),
                           0.125 * (1 + r) * (1 + s) * (1 + t), 0.125 * (1 - r) * (1 + s) * (1 + t)], dim=0)
        W = (gw[None, None, :] * gw[None, :, None] * gw[:, None, None]).reshape(1, Nq, Nq, Nq)
        self.ggeo = torch.empty((E, 6, Np), dtype=torch.float64, device=dev)
        self.wJ = torch.empty((E, Np), dtype=torch.float64, device=dev)
        if keep_coords:
            self.x = torch.empty((E, Np), dtype=torch.float64, device=dev)
            self.y = torch.empty_like(self.x)
            self.z = torch.empty_like(self.x)
        for e0 in range(0, E, chunk):
            e = torch.arange(e0, min(E, e0 + chunk), device=dev)
            x0 = X0 + dx * (e % nx).double()
            y0 = Y0 + dy * ((e // nx) % ny).double()
            z0 = Z0 + dz * (e // (nx * ny)).double()
            ex = torch.stack([x0, x0 + dx, x0 + dx, x0, x0, x0 + dx, x0 + dx, x0], dim=1)
            ey = torch.stack([y0, y0, y0 + dy, y0 + dy, y0, y0, y0 + dy, y0 + dy], dim=1)
            ez = torch.stack([z0, z0, z0, z0, z0 + dz, z0 + dz, z0 + dz, z0 + dz], dim=1)
            xc, yc, zc = ex @ shp, ey @ shp, ez @ shp
            if keep_coords:
                self.x[e0:e0 + len(e)], self.y[e0:e0 + len(e)], self.z[e0:e0 + len(e)] = xc, yc, zc
            n = len(e)
            X, Y, Z = xc.reshape(n, Nq, Nq, Nq), yc.reshape(n, Nq, Nq, Nq), zc.reshape(n, Nq, Nq, Nq)
            dr = lambda F: torch.einsum("im,ekjm->ekji", D, F)
            ds = lambda F: torch.einsum("jm,ekmi->ekji", D, F)
            dt = lambda F: torch.einsum("km,emji->ekji", D, F)
            xr, xs, xt = dr(X), ds(X), dt(X)
            yr, ys, yt = dr(Y), ds(Y), dt(Y)
            zr, zs, zt = dr(Z), ds(Z), dt(Z)
            J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
            rx, ry, rz = (ys * zt - zs * yt) / J, -(xs * zt - zs * xt) / J, (xs * yt - ys * xt) / J
            sx, sy, sz = -(yr * zt - zr * yt) / J, (xr * zt - zr * xt) / J, -(xr * yt - yr * xt) / J
            tx, ty, tz = (yr * zs - zr * ys) / J, -(xr * zs - zr * xs) / J, (xr * ys - yr * xs) / J
            JW = J * W
            g = self.ggeo[e0:e0 + n]
            g[:, 0] = (JW * (rx * rx + ry * ry + rz * rz)).reshape(n, Np)
            g[:, 1] = (JW * (rx * sx + ry * sy + rz * sz)).reshape(n, Np)
            g[:, 2] = (JW * (rx * tx + ry * ty + rz * tz)).reshape(n, Np)
            g[:, 3] = (JW * (sx * sx + sy * sy + sz * sz)).reshape(n, Np)
            g[:, 4] = (JW * (sx * tx + sy * ty + sz * tz)).reshape(n, Np)
            g[:, 5] = (JW * (tx * tx + ty * ty + tz * tz)).reshape(n, Np)
            self.wJ[e0:e0 + n] = JW.reshape(n, Np)

    # ---- discontinuous-Galerkin connectivity of the box (one rank): mesh_t::ConnectFaceNodes for this mesh
    def face_nodes(self):
        """[6, Nq^2] volume node of every face node; faces 0..5 = t=-1, s=-1, r=+1, s=+1, r=-1, t=+1 (the reference's
        hexahedron: libs/mesh/meshReferenceNodesHex3D.cpp faceNodes)"""
        Nq, N = self.Nq, self.N
        n = torch.arange(Nq * Nq, device=self.device)
        a, b = n % Nq, n // Nq
        return torch.stack([a + b * Nq, a + b * Nq * Nq, N + a * Nq + b * Nq * Nq, a + N * Nq + b * Nq * Nq,
                            a * Nq + b * Nq * Nq, a + b * Nq + N * Nq * Nq]).to(torch.int64)

    def _dg_connectivity_one_rank(self):
        """dg_connectivity on one rank (no halo elements; wrap-around neighbours of a periodic box are local)"""
        nx, ny, nz = self.nloc
        E, Np, Nfp = self.Nelements, self.Np, self.Nq * self.Nq
        dev = self.device
        e = torch.arange(E, device=dev)
        ex, ey, ez = e % nx, (e // nx) % ny, e // (nx * ny)
        fn = self.face_nodes()
        periodic = self.boundary_flag == -1
        shifts = [(0, 0, -1, 5), (0, -1, 0, 3), (1, 0, 0, 4), (0, 1, 0, 1), (-1, 0, 0, 2), (0, 0, 1, 0)]
        n = torch.arange(Nfp, device=dev)
        vmapM = (e[:, None, None] * Np + fn[None]).reshape(E, 6 * Nfp)
        vmapP = torch.empty((E, 6, Nfp), dtype=torch.int64, device=dev)
        mapP = torch.empty((E, 6, Nfp), dtype=torch.int64, device=dev)
        EToB = torch.full((E, 6), -1, dtype=torch.int32, device=dev)
        for f, (sx, sy, sz, fP) in enumerate(shifts):
            px, py, pz = ex + sx, ey + sy, ez + sz
            outside = (px < 0) | (px >= nx) | (py < 0) | (py >= ny) | (pz < 0) | (pz >= nz)
            eP = (px % nx) + (py % ny) * nx + (pz % nz) * nx * ny
            bnd = outside & (not periodic)
            vP = eP[:, None] * Np + fn[fP][None]
            mP = eP[:, None] * 6 * Nfp + fP * Nfp + n[None]
            vM = e[:, None] * Np + fn[f][None]
            mM = e[:, None] * 6 * Nfp + f * Nfp + n[None]
            vmapP[:, f] = torch.where(bnd[:, None], vM, vP)
            mapP[:, f] = torch.where(bnd[:, None], mM, mP)
            EToB[:, f] = torch.where(bnd, torch.tensor(self.boundary_flag, dtype=torch.int32, device=dev),
                                     torch.tensor(-1, dtype=torch.int32, device=dev))
        return (vmapM.to(torch.int32).contiguous(), vmapP.reshape(E, 6 * Nfp).to(torch.int32).contiguous(),
                mapP.reshape(E, 6 * Nfp).to(torch.int32).contiguous(), EToB.contiguous())

    def dg_connectivity(self):
        """(vmapM, vmapP, mapP [E, 6*Nq^2] int32, EToB [E, 6] int32 mesh boundary flag, -1 = interior face).
        Neighbouring box elements see a shared face with the same (a, b) face-node numbering, so face node n of face f
        meets face node n of the opposite face of the neighbour; a boundary face maps to itself.  On several ranks a
        neighbour on another rank becomes a halo element: slots Nelements, Nelements+1, ... in (element, face) order,
        one per (element, face) pair, as mesh_t::HaloSetup numbers them (libs/mesh/meshHaloSetup.cpp:97-111); the
        details of that halo are kept in self.dg_halo (see dg_halo_info)."""
        if self.size == 1:
            out = self._dg_connectivity_one_rank()
            E = self.Nelements
            z = torch.zeros((E, 6), dtype=torch.int64, device=self.device)
            self.dg_halo = dict(totalHaloPairs=0, remote=torch.zeros((E, 6), dtype=torch.bool, device=self.device),
                                neighbourRank=z, neighbourLocal=z, slot=z - 1,
                                elementOffsets=torch.tensor([0, E], dtype=torch.int64, device=self.device),
                                internalElementIds=torch.arange(E, dtype=torch.int32, device=self.device),
                                haloElementIds=torch.zeros(0, dtype=torch.int32, device=self.device))
            return out
        nx, ny, nz = self.nloc
        ox, oy, oz = self.off
        NX, NY, NZ = self.NX, self.NY, self.NZ
        E, Np, Nfp = self.Nelements, self.Np, self.Nq * self.Nq
        dev = self.device
        e = torch.arange(E, device=dev)
        ex, ey, ez = e % nx, (e // nx) % ny, e // (nx * ny)
        fn = self.face_nodes()
        periodic = self.boundary_flag == -1
        shifts = [(0, 0, -1, 5), (0, -1, 0, 3), (1, 0, 0, 4), (0, 1, 0, 1), (-1, 0, 0, 2), (0, 0, 1, 0)]
        n = torch.arange(Nfp, device=dev)
        # owner rank, local element number and first global element of every element of the global box
        owner = torch.empty((NZ, NY, NX), dtype=torch.int64, device=dev)
        local = torch.empty((NZ, NY, NX), dtype=torch.int64, device=dev)
        goff = [0]
        for rr in range(self.size):
            (mx, my, mz), (px, py, pz) = local_box(NX, NY, NZ, self.size, rr)
            owner[pz:pz + mz, py:py + my, px:px + mx] = rr
            local[pz:pz + mz, py:py + my, px:px + mx] = torch.arange(mx * my * mz, device=dev).reshape(mz, my, mx)
            goff.append(goff[-1] + mx * my * mz)
        goff = torch.tensor(goff, dtype=torch.int64, device=dev)
        # neighbour of every (element, face): global coordinates, owner, boundary
        gP = torch.empty((E, 6, 3), dtype=torch.int64, device=dev)
        bnd = torch.empty((E, 6), dtype=torch.bool, device=dev)
        for f, (sx, sy, sz, _) in enumerate(shifts):
            qx, qy, qz = ex + ox + sx, ey + oy + sy, ez + oz + sz
            outside = (qx < 0) | (qx >= NX) | (qy < 0) | (qy >= NY) | (qz < 0) | (qz >= NZ)
            bnd[:, f] = outside & (not periodic)
            gP[:, f, 0], gP[:, f, 1], gP[:, f, 2] = qx % NX, qy % NY, qz % NZ
        rP = owner[gP[..., 2], gP[..., 1], gP[..., 0]]                  # [E, 6]
        lP = local[gP[..., 2], gP[..., 1], gP[..., 0]]
        remote = (rP != self.rank) & ~bnd
        slot = torch.cumsum(remote.reshape(-1).to(torch.int64), 0).reshape(E, 6) - 1   # (element, face) order
        eP = torch.where(remote, E + slot, lP)                           # element index in the (local + halo) numbering
        vmapM = (e[:, None, None] * Np + fn[None]).reshape(E, 6 * Nfp)
        vmapP = torch.empty((E, 6, Nfp), dtype=torch.int64, device=dev)
        mapP = torch.empty((E, 6, Nfp), dtype=torch.int64, device=dev)
        EToB = torch.full((E, 6), -1, dtype=torch.int32, device=dev)
        for f, (_, _, _, fP) in enumerate(shifts):
            vP = eP[:, f, None] * Np + fn[fP][None]
            mP = eP[:, f, None] * 6 * Nfp + fP * Nfp + n[None]
            vM = e[:, None] * Np + fn[f][None]
            mM = e[:, None] * 6 * Nfp + f * Nfp + n[None]
            b = bnd[:, f, None]
            vmapP[:, f] = torch.where(b, vM, vP)
            mapP[:, f] = torch.where(b, mM, mP)
            EToB[:, f] = torch.where(bnd[:, f], torch.tensor(self.boundary_flag, dtype=torch.int32, device=dev),
                                     torch.tensor(-1, dtype=torch.int32, device=dev))
        is_halo = remote.any(dim=1)
        self.dg_halo = dict(totalHaloPairs=int(remote.sum()), remote=remote, neighbourRank=rP, neighbourLocal=lP,
                            slot=slot, elementOffsets=goff,
                            internalElementIds=torch.nonzero(~is_halo).reshape(-1).to(torch.int32),
                            haloElementIds=torch.nonzero(is_halo).reshape(-1).to(torch.int32))
        return (vmapM.to(torch.int32).contiguous(), vmapP.reshape(E, 6 * Nfp).to(torch.int32).contiguous(),
                mapP.reshape(E, 6 * Nfp).to(torch.int32).contiguous(), EToB.contiguous())

    def dg_trace_halo_ids(self):
        """ids mesh_t::HaloTraceSetup(1) hands to halo_t::Setup (libs/mesh/meshHaloTraceSetup.cpp:38-83), int64
        [(Nelements + totalHaloPairs) * Np]: own nodes carry (global element * Np + node + 1), the face nodes of the
        halo elements this rank's faces look at carry the negative id of the remote node, everything else 0.
        Call after dg_connectivity()."""
        h = self.dg_halo
        E, Np, Nfp = self.Nelements, self.Np, self.Nq * self.Nq
        dev = self.device
        Eh = h["totalHaloPairs"]
        ids = torch.zeros((E + Eh) * Np, dtype=torch.int64, device=dev)
        ids[: E * Np] = int(h["elementOffsets"][self.rank]) * Np + torch.arange(E * Np, device=dev) + 1
        if Eh:
            fn = self.face_nodes()
            opposite = [5, 3, 4, 1, 2, 0]
            ef = torch.nonzero(h["remote"])                                # rows (element, face) in slot order
            el, fa = ef[:, 0], ef[:, 1]
            gl = h["neighbourLocal"][el, fa] + h["elementOffsets"][h["neighbourRank"][el, fa]]
            fP = torch.tensor(opposite, device=dev)[fa]
            nodes = fn[fP]                                                 # [Eh, Nfp] nodes of the neighbour's face
            slots = E + torch.arange(Eh, device=dev)
            ids[(slots[:, None] * Np + nodes).reshape(-1)] = -(gl[:, None] * Np + nodes + 1).reshape(-1)
        return ids

    def masked_global_ids(self, bc_type=(0, 1, 2)):
        """(elliptic mapB, maskedGlobalIds): Dirichlet nodes get id 0 (ellipticBoundarySetup.cpp:55-86)."""
        bt = torch.tensor(bc_type, dtype=torch.int32, device=self.device)
        mapB = torch.zeros_like(self.mapB)
        pos = self.mapB > 0
        mapB[pos] = bt[self.mapB[pos].long()]
        ids = self.globalIds.clone()
        ids[mapB == 1] = 0
        return mapB, ids
