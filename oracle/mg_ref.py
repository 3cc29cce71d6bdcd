"""ORACLE (test infrastructure): CPU restatement of the multigrid preconditioner apply path, single rank.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned against the unmodified reference through tests/golden/mg_*.npz (tests/test_oracle_mg_golden.py).

  MGLevel::{residual, coarsen, prolongate, smoothJacobi, smoothChebyshev}
                                     solvers/elliptic/src/ellipticPreconMultiGridLevel.cpp:34-206
  amgLevel / parCSR smoothers        libs/parAlmond/parAlmondAMGLevel.cpp:48-74, parAlmondAMGSmoother.cpp:35-160
  exactSolver_t::solve               libs/parAlmond/parAlmondCoarseExact.cpp:35-73
  multigrid_t::vcycle                libs/parAlmond/parAlmondVcycle.cpp:34-60
  MultiGridPrecon::Operator          solvers/elliptic/src/ellipticPreconMultiGrid.cpp:29-37
"""
import numpy as np
import scipy.sparse as sp

from . import elliptic_ref as er
from .mesh_box import build_box_hex_mesh, masked_global_ids
from .ogs_ref import SIGNED, ogs_setup_all

JACOBI, CHEBYSHEV = 1, 2          # MGLevel::SmootherType
DAMPED_JACOBI, AMG_CHEBYSHEV = 0, 1  # parAlmond SmoothType


class DegreeProblem:
    """One degree of the p-multigrid ladder on a box: mesh, masked ogs, operator, diagonal, weights."""

    def __init__(self, N, n, lam, flag):
        self.N, self.Nq, self.lam = N, N + 1, lam
        self.mesh = m = build_box_hex_mesh(N, n, n, n, boundary_flag=flag)
        self.mapB, ids = masked_global_ids(m)
        self.ogs = o = ogs_setup_all([ids], SIGNED, True)[0]
        self.G2L = o.global_to_local()
        self.rs, self.ci = o.gatherLocal.rowStartsT, o.gatherLocal.colIdsT
        self.rsN, self.ciN = o.gatherLocal.rowStartsN, o.gatherLocal.colIdsN
        self.Ndofs = o.Ngather

    def operator(self, q):
        m = self.mesh
        return er.operator(self.Nq, self.G2L, m.wJ, m.ggeo, m.D, self.lam, self.rs, self.ci, q)

    def inv_diagonal(self):
        m = self.mesh
        dl = er.build_diagonal_local(self.Nq, m.ggeo, m.wJ, m.D, self.lam, self.mapB)
        return 1.0 / er.gather_add(self.rs, self.ci, dl)

    def weightG(self):
        """elliptic_t::weightG (ellipticBoundarySetup.cpp:100-110): inverse multiplicity of every gathered DOF"""
        cnt = er.gather_add(self.rs, self.ci, np.ones(self.mesh.Nelements * self.mesh.Np))
        return np.where(cnt > 0, 1.0 / np.where(cnt > 0, cnt, 1.0), cnt)

    def local(self, q):
        """scatter: element-local copy of a gathered vector (masked nodes -> 0)"""
        return np.where(self.G2L >= 0, q[np.maximum(self.G2L, 0)], 0.0)


class MGLevelRef:
    """MGLevel for the continuous hex discretisation."""

    def __init__(self, fine: DegreeProblem, coarse: DegreeProblem, P, invDiagA, smoother, lambda0, lambda1, cheb_iters=2):
        self.F, self.C = fine, coarse
        self.P = np.asarray(P, dtype=np.float64).reshape(fine.Nq, coarse.Nq)
        self.invDiagA = invDiagA          # already times lambda0 for the Jacobi smoother (:374-380)
        self.wG = fine.weightG()
        self.smoother, self.l0, self.l1, self.cheb_iters = smoother, lambda0, lambda1, cheb_iters

    def residual(self, rhs, x):
        return rhs - self.F.operator(x)

    def coarsen(self, x):
        F, C = self.F, self.C
        RxL = er.coarsen_hex3d(F.Nq, C.Nq, self.P, F.local(self.wG * x))
        RxL = np.where(C.G2L >= 0, RxL, 0.0)
        return er.gather_add(C.rs, C.ci, RxL)            # ogsMasked.Gather(Add, Trans)

    def prolongate(self, xC, x):
        F, C = self.F, self.C
        PxL = er.prolongate_hex3d(F.Nq, C.Nq, self.P, C.local(xC))
        return x + er.gather_add(F.rsN, F.ciN, PxL)      # Gather(Add, NoTrans): one copy per DOF

    def smooth(self, rhs, x, x_is_zero):
        S = self.invDiagA
        if self.smoother == JACOBI:                      # smoothJacobi (:137-154)
            if x_is_zero:
                return S * rhs
            return x + S * (rhs - self.F.operator(x))
        theta, delta = 0.5 * (self.l1 + self.l0), 0.5 * (self.l1 - self.l0)   # smoothChebyshev (:156-206)
        sigma = theta / delta
        rho_n = 1.0 / sigma
        res = S * rhs if x_is_zero else S * (rhs - self.F.operator(x))
        d = res / theta
        x = np.zeros_like(rhs) if x_is_zero else x.copy()
        for _ in range(self.cheb_iters):
            x = x + d
            res = res - S * self.F.operator(d)
            rho_np1 = 1.0 / (2.0 * sigma - rho_n)
            d = rho_np1 * rho_n * d + 2.0 * rho_np1 / delta * res
            rho_n = rho_np1
        return x + d


class AmgLevelRef:
    def __init__(self, A, P, R, diagInv, smoother, lam, lambda0, lambda1, cheb_iters=2):
        self.A, self.P, self.R = sp.csr_matrix(A), sp.csr_matrix(P), sp.csr_matrix(R)
        self.dInv, self.smoother, self.lam, self.l0, self.l1, self.cheb_iters = diagInv, smoother, lam, lambda0, lambda1, cheb_iters

    def residual(self, rhs, x):
        return rhs - self.A @ x

    def coarsen(self, x):
        return self.R @ x

    def prolongate(self, xC, x):
        return x + self.P @ xC

    def smooth(self, rhs, x, x_is_zero):
        A, dInv = self.A, self.dInv
        if self.smoother == DAMPED_JACOBI:               # parCSR::smoothDampedJacobi
            if x_is_zero:
                return self.lam * dInv * rhs
            return x + self.lam * dInv * (rhs - A @ x)
        theta, delta = 0.5 * (self.l1 + self.l0), 0.5 * (self.l1 - self.l0)   # parCSR::smoothChebyshev
        sigma = theta / delta
        rho_n = 1.0 / sigma
        r = dInv * rhs if x_is_zero else dInv * (rhs - A @ x)
        d = r / theta
        x = d.copy() if x_is_zero else x + d
        for _ in range(self.cheb_iters):
            r = r - dInv * (A @ d)
            rho_np1 = 1.0 / (2.0 * sigma - rho_n)
            d = rho_np1 * rho_n * d + 2.0 * rho_np1 / delta * r
            x = x + d
            rho_n = rho_np1
        return x


class MultigridRef:
    """multigrid_t V-cycle + MultiGridPrecon::Operator"""

    def __init__(self, levels, coarse_inv):
        self.levels, self.coarse_inv = levels, coarse_inv

    def vcycle(self, k, rhs):
        if k == len(self.levels):
            return self.coarse_inv @ rhs                 # exactSolver_t::solve
        L = self.levels[k]
        x = L.smooth(rhs, None, True)
        res = L.residual(rhs, x)
        xC = self.vcycle(k + 1, L.coarsen(res))
        x = L.prolongate(xC, x)
        return L.smooth(rhs, x, False)

    # ---- PARALMOND CYCLE = KCYCLE (libs/parAlmond/parAlmondKcycle.cpp:33-133): on the first NUMKCYCLES coarse levels the
    # coarse correction is improved by up to two steps of a flexible Krylov iteration preconditioned by the cycle itself
    NUMKCYCLES, KCYCLETOL = 3, 0.2   # include/parAlmond/parAlmondDefines.hpp:35-36

    def level_operator(self, k, x):
        L = self.levels[k]
        return L.F.operator(x) if isinstance(L, MGLevelRef) else L.A @ x

    def kcycle(self, k, rhs):
        if k == len(self.levels):
            return self.coarse_inv @ rhs
        L = self.levels[k]
        x = L.smooth(rhs, None, True)
        res = L.residual(rhs, x)
        rhsC = L.coarsen(res)
        if k + 1 > self.NUMKCYCLES:
            xC = self.vcycle(k + 1, rhsC)
        elif k + 1 == len(self.levels):
            # the exact coarse solve is next: its Krylov step is the identity (alpha1/rho1 = 1, zero residual)
            xC = self.kcycle(k + 1, rhsC)
        else:
            xC = self.kcycle(k + 1, rhsC)                      # first inner iteration
            ck = xC.copy()                                      # kcycleOp1 (:89-112)
            vk = self.level_operator(k + 1, ck)
            alpha1, rho1, norm_rhs = float(ck @ rhsC), float(ck @ vk), np.sqrt(float(rhsC @ rhsC))
            rhsC = rhsC - (alpha1 / rho1) * vk
            norm_rhstilde = np.sqrt(float(rhsC @ rhsC))
            if norm_rhstilde < self.KCYCLETOL * norm_rhs:
                xC = (alpha1 / rho1) * xC
            else:
                xC = self.kcycle(k + 1, rhsC)                  # second inner iteration
                if abs(rho1) > 1e-20:                           # kcycleOp2 (:114-141)
                    wk = self.level_operator(k + 1, xC)
                    gamma, beta, alpha2 = float(xC @ vk), float(xC @ wk), float(xC @ rhsC)
                    rho2 = beta - gamma * gamma / rho1
                    if abs(rho2) > 1e-20:
                        xC = (alpha1 / rho1 - gamma * alpha2 / (rho1 * rho2)) * ck + (alpha2 / rho2) * xC
        x = L.prolongate(xC, x)
        return L.smooth(rhs, x, False)

    def apply(self, r):
        return self.vcycle(0, r)

    def apply_kcycle(self, r):
        return self.kcycle(0, r)


def pcg(A, M, x, r, tol=1e-8, maxit=5000):
    """LinearSolver::pcg::Solve with callable operator / preconditioner (linearSolverPCG.cpp:67-151).
    Returns (iterations, x, residual norms [iterations+1])."""
    x, r = x.copy(), r - A(x)
    rdotr = float(r @ r)
    TOL = max(tol * tol * rdotr, tol * tol)
    hist = [np.sqrt(rdotr)]
    rdotz1, p = 0.0, np.zeros_like(r)
    it = 0
    while it < maxit:
        if (it == 0 and rdotr == 0.0) or (it > 0 and rdotr <= TOL):
            break
        z = M(r)
        rdotz2, rdotz1 = rdotz1, float(r @ z)
        beta = 0.0 if it == 0 else rdotz1 / rdotz2
        p = z + beta * p
        Ap = A(p)
        alpha = rdotz1 / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rdotr = float(r @ r)
        hist.append(np.sqrt(rdotr))
        it += 1
    return it, x, np.array(hist)


def nbpcg(A, M, x, r, tol=1e-8, maxit=5000):
    """LinearSolver::nbpcg::Solve (libs/linearSolver/linearSolverNBPCG.cpp:68-173), Gropp's non-blocking PCG with
    callable operator / preconditioner.  Returns (iterations, x, residual norms [iterations+1])."""
    x, r = x.copy(), r - A(x)
    z = M(r)
    gamma0, rdotr = float(r @ z), float(r @ r)       # Update2NBPCG with alpha = 0
    Z = A(z)
    TOL = max(tol * tol * rdotr, tol * tol)
    hist = [np.sqrt(rdotr)]
    p, s, beta = np.zeros_like(r), np.zeros_like(r), 0.0
    it = 0
    while it < maxit:
        if rdotr <= TOL:
            break
        p = z + beta * p                              # Update1NBPCG
        s = Z + beta * s
        delta = float(p @ s)
        S = M(s)
        alpha = gamma0 / delta
        r = r - alpha * s                             # Update2NBPCG
        z = z - alpha * S
        gamma1, gamma0, rdotr = gamma0, float(r @ z), float(r @ r)
        x = x + alpha * p
        Z = A(z)
        beta = gamma0 / gamma1
        hist.append(np.sqrt(rdotr))
        it += 1
    return it, x, np.array(hist)


def nbfpcg(A, M, x, r, tol=1e-8, maxit=5000):
    """LinearSolver::nbfpcg::Solve (libs/linearSolver/linearSolverNBFPCG.cpp:71-171), the non-blocking flexible PCG of
    Sanan et al. with callable operator / preconditioner; Update0NBFPCG / Update1NBFPCG (:174-246) are the fused
    vector-update + dot-product kernels.  Returns (iterations, x, residual norms [iterations+1])."""
    x, r = x.copy(), r - A(x)
    u = M(r)
    p = u.copy()
    w = A(p)
    gamma, delta, rdotr = float(u @ r), float(u @ w), float(r @ r)   # Update0NBFPCG
    m = M(w)
    n = A(m)
    s, q, z = w.copy(), m.copy(), n.copy()
    eta = delta
    alpha = gamma / eta
    TOL = max(tol * tol * rdotr, tol * tol)
    hist = [np.sqrt(rdotr)]
    it = 0
    while it < maxit:
        if rdotr <= TOL:
            break
        x = x + alpha * p                                             # Update1NBFPCG
        r = r - alpha * s
        u = u - alpha * q
        w = w - alpha * z
        gamma, uds, delta, rdotr = float(u @ r), float(u @ s), float(u @ w), float(r @ r)
        m = u + M(w - r)
        n = A(m)
        beta = -uds / eta
        p = u + beta * p
        s = w + beta * s
        q = m + beta * q
        z = n + beta * z
        eta = delta - beta * beta * eta
        alpha = gamma / eta
        hist.append(np.sqrt(rdotr))
        it += 1
    return it, x, np.array(hist)
