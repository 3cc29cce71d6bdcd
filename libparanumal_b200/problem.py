"""Harness: assembles one rank's elliptic problem (mesh arrays in HBM, masked ogs, operator,
Jacobi diagonal, right-hand side) the way elliptic_t::Setup / Run do, using only the C ABI for the
hot path.  Setup-only host/torch code here mirrors:
  solvers/elliptic/src/ellipticBoundarySetup.cpp:30-140   masked ids -> ogsMasked -> GlobalToLocal
  solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:998-1057 (+ gather :76-77)  Jacobi diagonal
  solvers/elliptic/src/ellipticRun.cpp:139-185            forcing + Dirichlet lift + gather of the rhs
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L
from .api import Comm, Elliptic, Ogs, Pcg, Precon, ax_hex3d
from .box_mesh import BoxMesh


class EllipticProblem:
    def __init__(self, N, NX, NY=None, NZ=None, lam=1.0, boundary_flag=1, comm: Comm | None = None, device="cuda",
                 mode=1, coords=False, mesh: BoxMesh | None = None):
        NY = NX if NY is None else NY
        NZ = NX if NZ is None else NZ
        self.comm = comm if comm is not None else Comm()
        self.device = torch.device(device)
        self.N, self.Nq, self.lam, self.mode = N, N + 1, float(lam), mode
        self.mesh = mesh if mesh is not None else BoxMesh(N, NX, NY, NZ, self.comm.rank, self.comm.size,
                                                          boundary_flag, device=device, coords=coords)
        m = self.mesh
        self.mapB, ids = m.masked_global_ids()
        self.maskedGlobalIds = ids.cpu().numpy().copy()
        del ids
        self.ogs = Ogs().Setup(self.maskedGlobalIds.size, self.maskedGlobalIds, self.comm, kind=L.SIGNED, unique=True)
        self.G2L_host = self.ogs.SetupGlobalToLocalMapping()
        self.GlobalToLocal = torch.from_numpy(self.G2L_host).to(self.device)
        self.Ndofs, self.Nhalo = self.ogs.Ngather, self.ogs.Nhalo
        self.Nall = self.Ndofs + self.Nhalo
        self.NglobalDofs = self.ogs.NgatherGlobal
        # allNeumann: lambda==0 and no Dirichlet face anywhere (ellipticBoundarySetup.cpp:33-51)
        self.allNeumann = (self.lam == 0.0) and (boundary_flag == -1)
        self.op = Elliptic(self.Nq, m.localGatherElementList, m.globalGatherElementList, self.GlobalToLocal,
                           m.wJ, m.ggeo, m.D, self.lam, self.ogs, mode=mode)

    def set_lambda(self, lam):
        """Same mesh, maps and ogs with a different screening parameter: only the operator handle changes."""
        self.op.Free()
        self.lam = float(lam)
        m = self.mesh
        self.allNeumann = (self.lam == 0.0) and (m.boundary_flag == -1)
        self.op = Elliptic(self.Nq, m.localGatherElementList, m.globalGatherElementList, self.GlobalToLocal,
                           m.wJ, m.ggeo, m.D, self.lam, self.ogs, mode=self.mode)

    def vec(self, fill=0.0):
        return torch.full((self.Nall,), fill, dtype=torch.float64, device=self.device)

    def operator(self, q, Aq=None):
        Aq = self.vec() if Aq is None else Aq
        self.op.Operator(q, Aq)
        return Aq

    # ---- setup-time pieces (torch, not the hot path) ---------------------------------------------
    def diagonal_local(self):
        """BuildOperatorDiagonalContinuousHex3D: element-local diagonal [E*Np]."""
        m, Nq = self.mesh, self.Nq
        E = m.Nelements
        G = m.ggeo.reshape(E, 6, Nq, Nq, Nq)
        D = m.D.reshape(Nq, Nq)
        dd = torch.diagonal(D)
        di, dj, dk = dd[None, None, None, :], dd[None, None, :, None], dd[None, :, None, None]
        A = 2 * G[:, 1] * di * dj + 2 * G[:, 2] * di * dk + 2 * G[:, 4] * dj * dk
        D2 = D * D
        A = A + torch.einsum("ezyk,kx->ezyx", G[:, 0], D2)
        A = A + torch.einsum("ezkx,ky->ezyx", G[:, 3], D2)
        A = A + torch.einsum("ekyx,kz->ezyx", G[:, 5], D2)
        A = A + m.wJ.reshape(E, Nq, Nq, Nq) * self.lam
        A = A.reshape(-1).clone()
        masked = self.mapB == 1
        if self.allNeumann:
            scale = 1.0 / math.sqrt(float(self.NglobalDofs))
            A[~masked] += 1.0 * scale * scale
        A[masked] = 1.0
        return A

    def inv_diagonal(self):
        diagL = self.diagonal_local()
        diag = self.vec()
        self.ogs.Gather(diag, diagL, 1, L.ADD, L.TRANS)
        inv = self.vec()
        inv[: self.Ndofs] = 1.0 / diag[: self.Ndofs]
        return inv

    def jacobi(self):
        inv = self.inv_diagonal()
        return Precon.Jacobi(self.Ndofs, inv, self.allNeumann, self.NglobalDofs, self.comm)

    def rhs_sine3d(self):
        """Gathered right-hand side for data/ellipticSine3D.h (needs mesh coords)."""
        m = self.mesh
        PI = 3.14159265358979323846
        s = torch.sin(PI * m.x) * torch.sin(PI * m.y) * torch.sin(PI * m.z)
        rL = (m.wJ * ((3 * PI * PI + self.lam) * s)).reshape(-1)
        uD = torch.where(self.mapB.reshape(m.x.shape) == 1, s, torch.zeros_like(s)).reshape(-1).contiguous()
        if bool((uD != 0).any()):
            AuD = torch.empty_like(uD)
            ax_hex3d(self.Nq, m.Nelements, None, None, m.wJ, m.ggeo, m.D, self.lam, uD, AuD)
            rL = rL - AuD
        r = self.vec()
        self.ogs.Gather(r, rL.contiguous(), 1, L.ADD, L.TRANS)
        return r

    def pcg(self, flexible=False):
        return Pcg(self.Ndofs, self.Nhalo, self.comm, flexible=flexible)
