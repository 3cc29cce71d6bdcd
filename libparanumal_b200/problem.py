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

    def weightG(self):
        """elliptic_t::weightG (ellipticBoundarySetup.cpp:100-110): inverse multiplicity of every gathered DOF."""
        ones = torch.ones(self.mesh.Nelements * self.mesh.Np, dtype=torch.float64, device=self.device)
        w = self.vec()
        self.ogs.Gather(w, ones, 1, L.ADD, L.TRANS)
        w[: self.Ndofs] = torch.where(w[: self.Ndofs] > 0, 1.0 / w[: self.Ndofs], w[: self.Ndofs])
        w[self.Ndofs:] = 0
        return w

    def pcg(self, flexible=False):
        return Pcg(self.Ndofs, self.Nhalo, self.comm, flexible=flexible)


def degree_raise_1d(Nc, Nf):
    """mesh_t::DegreeRaiseMatrix1D (libs/mesh/meshBasis1D.cpp): P[NqF, NqC], degree-Nc GLL Lagrange basis
    evaluated at the degree-Nf GLL nodes."""
    from .box_mesh import gll
    rc, rf = gll(Nc)[0], gll(Nf)[0]
    P = np.ones((Nf + 1, Nc + 1))
    for m in range(Nc + 1):
        for l in range(Nc + 1):
            if l != m:
                P[:, m] *= (rf - rc[l]) / (rc[m] - rc[l])
    return P


def halfdofs_ladder(N):
    """Degree ladder of MULTIGRID COARSENING = HALFDOFS for hexes (ellipticPreconMultiGrid.cpp:69-84)."""
    lad, Nf = [N], N
    while Nf > 1:
        Nc, NpF = Nf, (Nf + 1) ** 3
        NpC = NpF
        while NpC > NpF // 2 and Nc > 1:
            Nc -= 1
            NpC = (Nc + 1) ** 3
        lad.append(Nc)
        Nf = Nc
    return lad


class MultigridHierarchy:
    """Harness-side assembly of MultiGridPrecon (solvers/elliptic/src/ellipticPreconMultiGrid.cpp:40-154) from
    setup products: one EllipticProblem per degree of the ladder (built in the reference's order, so every
    `unique` ogs setup consumes rand() in the same sequence), smoother parameters and the AMG levels / coarse
    inverse that the reference's setup produced (`levels` = list of dicts, see tests/test_gpu_multigrid.py)."""

    def __init__(self, fine: EllipticProblem, ladder, level_data, amg_data, coarse_invAT, coarse_N):
        from .api import AmgLevel, CoarseExact, Csr, MGLevel, Multigrid
        self.fine = fine
        self.problems = []
        m = fine.mesh
        for Nl in list(ladder) + ([1] if ladder[-1] != 1 else []):
            self.problems.append(EllipticProblem(Nl, m.NX, m.NY, m.NZ, lam=fine.lam, boundary_flag=m.boundary_flag,
                                                 comm=fine.comm, device=fine.device, mode=1))
        self.mg = Multigrid(fine.comm)
        self.keep = []
        for l, ld in enumerate(level_data):
            pF, pC = self.problems[l], self.problems[l + 1]
            P = torch.from_numpy(np.ascontiguousarray(ld["P"], dtype=np.float64)).to(fine.device)
            inv = ld.get("invDiagA")
            if inv is None:
                inv = pF.inv_diagonal()[: pF.Ndofs].clone()
                if ld["smoother"] == MGLevel.JACOBI:
                    inv *= ld["lambda0"]
            else:
                inv = torch.from_numpy(np.ascontiguousarray(inv)).to(fine.device)
            wG = pF.weightG()
            self.keep += [P, inv, wG]
            self.mg.AddLevel(MGLevel(pF.op, pC.op, pF.Nq, pC.Nq, P, inv, wG, ld["smoother"], ld["lambda0"], ld["lambda1"],
                                     ld.get("ChebyshevIterations", 2)))
        for ad in amg_data:
            A = Csr(*ad["A"])
            Pm = Csr(*ad["P"]) if ad.get("P") is not None else None
            R = Csr(*ad["R"]) if ad.get("R") is not None else None
            self.mg.AddLevel(AmgLevel(A, Pm, R, ad["diagInv"], ad["smoother"], ad["lambda"], ad["lambda0"], ad["lambda1"],
                                      ad.get("ChebyshevIterations", 2)))
        self.mg.SetCoarse(CoarseExact(coarse_N, coarse_invAT))

    def precon(self):
        from .api import Precon
        return Precon.MultiGrid(self.mg, self.fine.allNeumann, self.fine.NglobalDofs, self.fine.comm)
