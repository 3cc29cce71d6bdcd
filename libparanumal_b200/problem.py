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
from .api import Comm, Elliptic, Ogs, Pcg, Precon
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
        """BuildOperatorDiagonalContinuousHex3D: element-local diagonal [E*Np] (device kernel
        libp_elliptic_build_diagonal_hex3d)."""
        from .api import elliptic_build_diagonal_hex3d
        m = self.mesh
        boost = 0.0
        if self.allNeumann:
            scale = 1.0 / math.sqrt(float(self.NglobalDofs))
            boost = 1.0 * scale * scale  # allNeumannPenalty * allNeumannScale^2
        A = torch.empty(m.Nelements * m.Np, dtype=torch.float64, device=self.device)
        elliptic_build_diagonal_hex3d(self.Nq, m.Nelements, m.ggeo, m.wJ, m.D, self.mapB.contiguous(), self.lam, boost, A)
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

    def _sine3d(self):
        m = self.mesh
        PI = 3.14159265358979323846
        return (torch.sin(PI * m.x) * torch.sin(PI * m.y) * torch.sin(PI * m.z)).reshape(-1).contiguous()

    def rhs_sine3d(self):
        """Gathered right-hand side for data/ellipticSine3D.h (needs mesh coords): forcing kernel, boundary lift,
        ogsMasked.Gather(Add, Trans) - ellipticRun.cpp:139-183."""
        from .api import rhs_bc_hex3d, rhs_forcing_hex3d
        m = self.mesh
        PI = 3.14159265358979323846
        s = self._sine3d()
        f = (3 * PI * PI + self.lam) * s                       # ellipticForcing3D of the data file
        rL = torch.empty_like(f)
        rhs_forcing_hex3d(m.Nelements, m.Np, m.wJ, f, rL)
        uD = torch.where(self.mapB.reshape(-1) == 1, s, torch.zeros_like(s)).contiguous()  # ellipticBoundaryConditions3D
        if bool((uD != 0).any()):
            rhs_bc_hex3d(self.Nq, m.Nelements, m.wJ, m.ggeo, m.D, self.lam, uD, None, rL)
        r = self.vec()
        self.ogs.Gather(r, rL, 1, L.ADD, L.TRANS)
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

    def run(self, precon="JACOBI", tol=1e-8, maxit=5000, **mg_args):
        """elliptic_t::Run (solvers/elliptic/src/ellipticRun.cpp:139-246) for data/ellipticSine3D.h: forcing and
        boundary lift, gather, PCG solve, scatter, Dirichlet data on the masked nodes, and the mass-matrix norm the
        reference prints as "Solution norm".  precon: NONE | JACOBI | MULTIGRID | PARALMOND.
        Returns (iterations, solution norm, xL)."""
        m = self.mesh
        if precon == "NONE":
            M = Precon.Identity(self.Ndofs)
        elif precon == "JACOBI":
            M = self.jacobi()
        else:
            self._hier = MultigridHierarchy.build(self, algebraic_only=(precon == "PARALMOND"), **mg_args)
            M = self._hier.precon()
        r = self.rhs_sine3d()
        x = self.vec()
        if self.allNeumann:  # ellipticSolve.cpp:34 ZeroMean(o_r)
            r[: self.Ndofs] -= self.comm.allreduce_sum(float(r[: self.Ndofs].sum())) / float(self.NglobalDofs)
        it = self.pcg().Solve(self.op, M, x, r, tol=tol, maxit=maxit)
        xL = torch.zeros(m.Nelements * m.Np, dtype=torch.float64, device=self.device)
        self.ogs.Scatter(xL, x, 1, L.NOTRANS)
        from .api import add_bc_hex3d, mass_matrix_apply_hex3d
        add_bc_hex3d(m.Nelements, m.Np, self.mapB.reshape(-1).contiguous(), self._sine3d(), xL)   # addBCKernel
        MxL = torch.empty_like(xL)
        mass_matrix_apply_hex3d(m.Nelements, m.Np, m.wJ, xL, MxL)   # mesh_t::MassMatrixApply (collocated GLL)
        norm = float(np.sqrt(self.comm.allreduce_sum(float(torch.dot(xL, MxL)))))
        return it, norm, xL

    def global_numbering(self):
        """Global number of every gathered DOF this rank sees (owned, then halo) and the row partition
        (elliptic_t::maskedGlobalNumbering, ellipticSetup.cpp; A.globalRowStarts,
        ellipticBuildOperatorMatrixContinuous.cpp:699-710)."""
        counts = np.array([int(c) for c in self.comm.allgather_object(int(self.Ndofs))], dtype=np.int64)
        starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        num = torch.full((self.Nall,), -1, dtype=torch.int64, device=self.device)
        num[: self.Ndofs] = torch.arange(self.Ndofs, dtype=torch.int64, device=self.device) + int(starts[self.comm.rank])
        if self.comm.size > 1 and self.Nhalo > 0:
            self.ogs.Exchange(num)
        return num, starts

    def operator_matrix(self, threshold=1e-7):
        """BuildOperatorMatrixContinuousHex3D (ellipticBuildOperatorMatrixContinuous.cpp:697-870): unassembled
        element matrices A_e = sum_ab D_a^T diag(G_ab) D_b + lambda diag(wJ) with masked rows/columns removed and
        entries below `threshold` dropped, as (global row, global col, value) triplets + the row partition."""
        from .amg_setup import element_matrix_triplets
        m = self.mesh
        num, starts = self.global_numbering()
        gid = torch.where(self.GlobalToLocal >= 0, num[self.GlobalToLocal.clamp(min=0).long()],
                          torch.full_like(self.GlobalToLocal, -1, dtype=torch.int64))
        r, c, v = element_matrix_triplets(self.Nq, m.ggeo, m.wJ, m.D, self.lam, gid, threshold)
        return r, c, v, starts

    def max_eig_smooth_ax(self, invDiag, rng):
        """MGLevel::maxEigSmoothAx (ellipticPreconMultiGridLevel.cpp:393-473): Arnoldi estimate of
        rho(D^-1 A) from a drand48 start vector, with this problem's device operator."""
        from .amg_setup import arnoldi_rho
        n = self.Ndofs
        buf, out = self.vec(), self.vec()

        def apply(v):
            buf[:n] = v
            self.op.Operator(buf, out)
            return out[:n] * invDiag[:n]

        dot = lambda a, b: self.comm.allreduce_sum(float(torch.dot(a, b)))
        v0 = torch.from_numpy(rng.draw(n)).to(self.device)
        return arnoldi_rho(apply, v0, self.NglobalDofs, dot=dot)


class IpdgProblem:
    """DISCRETIZATION = IPDG on the box (one rank): the elliptic_t setup pieces the IPDG operator needs
    (ellipticSetup.cpp:55-80 tau and gradient buffer, mesh_t vgeo / sgeo / vmapM / vmapP, ellipticBoundarySetup.cpp:37-48
    EToB translation) and elliptic_t::Run for data/ellipticSine3D.h (ellipticRun.cpp:139-246, IPDG branch).
    Vectors are element-local [Nelements*Np]."""

    def __init__(self, N, NX, NY=None, NZ=None, lam=1.0, boundary_flag=1, device="cuda", bc_type=(0, 1, 2)):
        from . import api
        NY = NX if NY is None else NY
        NZ = NX if NZ is None else NZ
        self.comm = Comm()
        self.device = torch.device(device)
        self.N, self.Nq, self.lam = N, N + 1, float(lam)
        m = self.mesh = BoxMesh(N, NX, NY, NZ, 0, 1, boundary_flag, device=device, coords=True)
        E, Np, Nq, Nfp = m.Nelements, m.Np, m.Nq, m.Nq * m.Nq
        self.tau = 2.0 * (N + 1) * (N + 3)
        gw = torch.from_numpy(m.gllw).to(self.device)
        self.vgeo = torch.empty(E * 12 * Np, dtype=torch.float64, device=self.device)
        api.mesh_geometric_factors_hex3d(Nq, E, m.x, m.y, m.z, m.D, gw, m.ggeo, m.wJ, self.vgeo)
        self.vmapM, self.vmapP, self.mapP, meshEToB = m.dg_connectivity()
        bt = torch.tensor(bc_type, dtype=torch.int32, device=self.device)
        self.EToB = torch.where(meshEToB > 0, bt[meshEToB.clamp(min=0).long()], torch.zeros_like(meshEToB)).contiguous()
        self.sgeo = torch.empty(E * 6 * Nfp * 8, dtype=torch.float64, device=self.device)
        h = torch.empty(E * 6 * Nfp, dtype=torch.float64, device=self.device)
        api.mesh_surface_geometric_factors_hex3d(Nq, E, m.x, m.y, m.z, m.D, gw, self.sgeo, h)
        api.mesh_surface_hinv_hex3d(Nq, E, self.mapP, h, self.sgeo)
        self.op = api.EllipticIpdg(Nq, E, self.vmapM, self.vmapP, self.vgeo, self.sgeo, self.EToB, m.D, self.lam, self.tau)
        self.Ndofs, self.Nhalo = E * Np, 0
        self.NglobalDofs = self.Ndofs
        self.allNeumann = (self.lam == 0.0) and (boundary_flag == -1)

    def vec(self, fill=0.0):
        return torch.full((self.Ndofs,), fill, dtype=torch.float64, device=self.device)

    def operator(self, q, Aq=None):
        Aq = self.vec() if Aq is None else Aq
        self.op.Operator(q, Aq)
        return Aq

    def diagonal(self):
        from .api import elliptic_build_diagonal_ipdg_hex3d
        m = self.mesh
        A = self.vec()
        elliptic_build_diagonal_ipdg_hex3d(self.Nq, m.Nelements, self.vgeo, self.sgeo, self.EToB, m.D, self.lam, self.tau, A)
        return A

    def jacobi(self):
        return Precon.Jacobi(self.Ndofs, 1.0 / self.diagonal(), self.allNeumann, self.NglobalDofs, self.comm)

    def pcg(self, flexible=False):
        return Pcg(self.Ndofs, self.Nhalo, self.comm, flexible=flexible)

    def rhs_sine3d(self):
        """forcing + boundary data of data/ellipticSine3D.h: u = sin(pi x) sin(pi y) sin(pi z), Dirichlet faces carry u"""
        from .api import rhs_bc_ipdg_hex3d, rhs_forcing_hex3d
        m = self.mesh
        PI = 3.14159265358979323846
        s = (torch.sin(PI * m.x) * torch.sin(PI * m.y) * torch.sin(PI * m.z)).reshape(-1).contiguous()
        f = (3 * PI * PI + self.lam) * s
        r = torch.empty_like(f)
        rhs_forcing_hex3d(m.Nelements, m.Np, m.wJ, f, r)
        uD = s[self.vmapM.reshape(-1).long()].contiguous()  # boundary values at the face nodes
        rhs_bc_ipdg_hex3d(self.Nq, m.Nelements, self.tau, self.vgeo, self.sgeo, self.EToB, m.D, uD, None, r)
        return r

    def run(self, precon="NONE", tol=1e-8, maxit=5000):
        """Returns (iterations, solution norm as elliptic_t::Run prints it, x)"""
        M = Precon.Identity(self.Ndofs) if precon == "NONE" else self.jacobi()
        r = self.rhs_sine3d()
        x = self.vec()
        it = self.pcg().Solve(self.op, M, x, r, tol=tol, maxit=maxit)
        norm = float(torch.sqrt(torch.sum(x * self.mesh.wJ.reshape(-1) * x)))
        return it, norm, x


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
        # elliptic_t::SetupNewDegree returns the original solver for the original degree
        # (ellipticSetupNewDegree.cpp:32-33): level 0 works on the caller's vectors, so it must share the fine
        # problem's ogs (on several ranks a second `unique` setup picks other owners for the shared nodes)
        for Nl in list(ladder) + ([1] if ladder[-1] != 1 else []):
            if Nl == fine.N and not self.problems:
                self.problems.append(fine)
                continue
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

    @classmethod
    def build(cls, fine: EllipticProblem, smoother="CHEBYSHEV", chebyshev_degree=2, amg_smoother=None,
              strength="SYMMETRIC", aggregation="SMOOTHED", coarse_target=1000, level_rho=None, algebraic_only=False,
              cycle="VCYCLE"):
        """MultiGridPrecon::MultiGridPrecon (ellipticPreconMultiGrid.cpp:40-154) without the reference: the
        HALFDOFS degree ladder, one problem per degree (built in the reference's order: every `unique` ogs setup
        consumes rand() in sequence), Chebyshev/Jacobi bounds from the Arnoldi estimate on the device operator,
        the degree-1 matrix, and the algebraic levels of amg_setup.setup_hierarchy split into row blocks.
        Smoothers: "CHEBYSHEV" (degree `chebyshev_degree`) or "DAMPEDJACOBI" (MULTIGRID SMOOTHER /
        PARALMOND SMOOTHER, ellipticSettings / parAlmondSettings.cpp:35-53).  level_rho: optional list of
        rho(D^-1 A) for the matrix-free levels (skips their Arnoldi estimates, e.g. to run the very same
        preconditioner on a different number of ranks).  algebraic_only: PRECONDITIONER = PARALMOND
        (ellipticPreconParAlmond.cpp:40-68) - no matrix-free levels, AMG on the assembled degree-N matrix."""
        import scipy.sparse as sp

        from . import amg_setup as am
        from .api import AmgLevel, CoarseExact, CoarseExactPar, Csr, MGLevel, Multigrid, ParCsr
        self = cls.__new__(cls)
        comm, m = fine.comm, fine.mesh
        rank, size = comm.rank, comm.size
        amg_smoother = smoother if amg_smoother is None else amg_smoother
        ladder = [fine.N] if algebraic_only else halfdofs_ladder(fine.N)
        self.fine, self.ladder = fine, ladder
        self.problems = []
        for Nl in ([] if algebraic_only else ladder):  # SetupNewDegree(Nf) for every ladder degree (1 is the last)
            # the original degree re-uses the original solver (ellipticSetupNewDegree.cpp:32-33): level 0 works on
            # the caller's vectors and must share the fine problem's ogs / DOF layout
            if Nl == fine.N and not self.problems:
                self.problems.append(fine)
                continue
            self.problems.append(EllipticProblem(Nl, m.NX, m.NY, m.NZ, lam=fine.lam, boundary_flag=m.boundary_flag,
                                                 comm=comm, device=fine.device, mode=1))
        self.mg = Multigrid(comm)
        self.keep = []
        rng = am.Drand48(rank)      # srand48(rank), parAlmondKernels.cpp:56-58
        drawn_global = 0            # position of a one-rank stream after the p-MG levels (AMG setup stream)
        self.level_info = []
        for l in range(len(ladder) - 1):
            pF, pC = self.problems[l], self.problems[l + 1]
            P = torch.from_numpy(degree_raise_1d(pC.N, pF.N)).to(fine.device).contiguous()
            inv = pF.inv_diagonal()[: pF.Ndofs].clone()
            rho = float(level_rho[l]) if level_rho is not None else pF.max_eig_smooth_ax(inv, rng)
            drawn_global += int(pF.NglobalDofs)
            if smoother == "CHEBYSHEV":
                kind, l0, l1 = MGLevel.CHEBYSHEV, rho / 10.0, rho
            else:
                kind, l0, l1 = MGLevel.JACOBI, (4.0 / 3.0) / rho, 0.0
                inv *= l0
            wG = pF.weightG()
            self.keep += [P, inv, wG]
            self.mg.AddLevel(MGLevel(pF.op, pC.op, pF.Nq, pC.Nq, P, inv, wG, kind, l0, l1, chebyshev_degree))
            self.level_info.append(dict(kind="pMG", degree=pF.N, rows=int(pF.NglobalDofs), rho=rho))
        # ---- degree-1 matrix, replicated (BuildOperatorMatrixContinuous + parCSR(cooA))
        p1 = fine if algebraic_only else self.problems[-1]
        r_, c_, v_, starts = p1.operator_matrix()
        parts = comm.allgather_object((r_, c_, v_))
        ntot = int(starts[-1])
        A = sp.coo_matrix((np.concatenate([p[2] for p in parts]),
                           (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))),
                          shape=(ntot, ntot)).tocsr()
        del parts
        # ---- algebraic hierarchy: the one-rank algorithm on the global matrix, same on every rank
        arng = am.Drand48(0)
        arng.draw(drawn_global)
        null = np.full(ntot, 1.0 / np.sqrt(ntot))
        levels, Ac, rho_c, null_c = am.setup_hierarchy(A, null, arng, coarse_target, strength, aggregation)
        self.amg_levels, self.coarse_A = levels, Ac
        part = np.asarray(starts, dtype=np.int64)
        cheb = amg_smoother == "CHEBYSHEV"
        self.amg_handles, self.amg_parts = [], []
        for lv in levels:
            cpart = am.coarse_partition(part, lv["roots"])
            dInv = 1.0 / lv["A"].diagonal()[part[rank]: part[rank + 1]]
            if size == 1:
                mk = lambda M, rp, cp: Csr(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
            else:
                mk = lambda M, rp, cp: ParCsr(comm, am.split_rows(M, rp, cp, rank))
            cA, cP, cR = mk(lv["A"], part, part), mk(lv["P"], part, cpart), mk(lv["R"], cpart, part)
            rho = lv["rho"]
            self.amg_handles.append(AmgLevel(cA, cP, cR, dInv, AmgLevel.CHEBYSHEV if cheb else AmgLevel.DAMPED_JACOBI,
                                             (4.0 / 3.0) / rho, rho / 10.0, rho, chebyshev_degree))
            self.amg_parts.append((part, cpart))
            self.mg.AddLevel(self.amg_handles[-1])
            self.level_info.append(dict(kind="AMG", rows=int(lv["A"].shape[0]), nnz=int(lv["A"].nnz), rho=rho))
            part = cpart
        # ---- exact coarse solve (exactSolver_t::setup): dense inverse, stored transposed
        Acd = Ac.toarray()
        if fine.allNeumann:  # rank-one boost of the singular coarse operator (allNeumannPenalty = 1)
            Acd += np.outer(null_c, null_c)
        inv = np.linalg.inv(Acd)
        n0, n1 = int(part[rank]), int(part[rank + 1])
        N = n1 - n0
        diagT = np.ascontiguousarray(inv[n0:n1, n0:n1].T)          # diagInvAT[n + m*N] = inv[n0+n, n0+m]
        if size == 1:
            self.coarse_handle = CoarseExact(N, diagT.reshape(-1))
        else:
            others = np.concatenate([np.arange(0, n0), np.arange(n1, Ac.shape[0])])
            offdT = np.ascontiguousarray(inv[n0:n1][:, others].T)  # offdInvAT[n + m*N], other ranks' rows ascending
            self.coarse_handle = CoarseExactPar(comm, N, part, diagT.reshape(-1), offdT.reshape(-1))
        self.mg.SetCoarse(self.coarse_handle)
        self.coarse_part, self.coarse_dense = part, Acd
        self.level_info.append(dict(kind="exact", rows=int(Ac.shape[0]), nnz=int(Ac.nnz)))
        self.mg.SetCycle(cycle)  # PARALMOND CYCLE
        return self
