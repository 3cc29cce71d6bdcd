"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the initial-guess strategies of
libs/linearSolver/initialGuess.cpp (used by repeated solves, e.g. the flow solvers' pressure / velocity systems):

  Zero                   :78-85      x = 0
  ClassicProjection      :117-198    history of A-images orthonormalised by two Gram-Schmidt passes, restart when full
  RollingQRProjection    :200-329    same space kept as a QR factorisation; the oldest column is dropped by Givens
                                     rotations (okl/igDropQRFirstColumn.okl) instead of restarting
  Extrap                 :343-469    polynomial extrapolation of the solution history; coefficients from an
                                     underdetermined Vandermonde system (minimum norm or column-pivoted QR,
                                     libs/linAlg/linAlgMatrixRightSolve.cpp:269-318)

Pinned against a dump of the reference's own classes driven over a sequence of solves
(tests/golden/ig_tridiag_n400.npz, oracle/refbuild/dump_ig_driver.cpp) by tests/test_oracle_initial_guess_cpu.py.
Only tests/ may import this."""
import numpy as np
import scipy.linalg


class Zero:
    def __init__(self, N, **_):
        self.N = N

    def form(self, x, rhs):
        return np.zeros(self.N)

    def update(self, A, x, rhs):
        pass


class ClassicProjection:
    def __init__(self, N, history=4, **_):
        self.N, self.maxDim, self.curDim = N, history, 0
        self.B = np.zeros((history, N))   # orthonormal images A x_i
        self.X = np.zeros((history, N))   # the matching combinations of solutions

    def form(self, x, rhs):
        if self.curDim == 0:
            return x                        # FormInitialGuess leaves x alone until there is a history
        k = self.curDim
        alphas = self.B[:k] @ rhs
        return alphas @ self.X[:k]

    def update(self, A, x, rhs):
        bt = A(x)
        if self.curDim >= self.maxDim or self.curDim == 0:
            nrm = np.linalg.norm(bt)
            if nrm > 0:
                self.B[0], self.X[0] = bt / nrm, x / nrm
                self.curDim = 1
            return
        k = self.curDim
        xt = x.copy()
        for _ in range(2):                  # Nreorth
            alphas = self.B[:k] @ bt
            bt = bt - alphas @ self.B[:k]
            xt = xt - alphas @ self.X[:k]
        inv = 1.0 / np.linalg.norm(bt)
        self.B[k], self.X[k] = inv * bt, inv * xt
        self.curDim += 1


def _givens(a, b):
    if b != 0:
        d = 1.0 / np.hypot(a, b)
        return abs(a) * d, np.copysign(d, a) * b
    return 1.0, 0.0


class RollingQRProjection(ClassicProjection):
    def __init__(self, N, history=4, **_):
        super().__init__(N, history)
        self.R = np.zeros((history, history))

    def update(self, A, x, rhs):
        bt = A(x)
        M = self.maxDim
        if self.curDim == M:
            R = self.R
            R[:, :-1] = R[:, 1:].copy()     # drop the first column of R
            R[:, -1] = 0.0
            for i in range(M - 1):          # restore the triangle; the same rotations act on the columns of B and X
                c, s = _givens(R[i, i], R[i + 1, i])
                Ri, Rp = R[i].copy(), R[i + 1].copy()
                R[i], R[i + 1] = c * Ri + s * Rp, -s * Ri + c * Rp
                for Q in (self.B, self.X):
                    qi, qp = Q[i].copy(), Q[i + 1].copy()
                    Q[i], Q[i + 1] = c * qi + s * qp, -s * qi + c * qp
            self.B[M - 1] = 0.0
            self.X[M - 1] = 0.0
            self.curDim -= 1
        if self.curDim == 0:
            nrm = np.linalg.norm(bt)
            if nrm > 0:
                self.B[0], self.X[0] = bt / nrm, x / nrm
                self.R[0, 0] = nrm
                self.curDim = 1
            return
        k = self.curDim
        xt = x.copy()
        nrm = np.linalg.norm(bt)
        self.R[:k, k] = 0.0
        for _ in range(2):
            alphas = self.B[:k] @ bt
            bt = bt - alphas @ self.B[:k]
            xt = xt - alphas @ self.X[:k]
            self.R[:k, k] += alphas
        nproj = np.linalg.norm(bt)
        if nproj / nrm > 1.0e-10:
            self.B[k], self.X[k] = bt / nproj, xt / nproj
            self.R[k, k] = nproj
            self.curDim += 1


def vandermonde1d(m, r):
    """mesh_t::Vandermonde1D: orthonormal Legendre polynomials P_0..P_m at the points r"""
    r = np.asarray(r, dtype=np.float64)
    V = np.empty((r.size, m + 1))
    for j in range(m + 1):
        c = np.zeros(j + 1)
        c[j] = 1.0
        V[:, j] = np.polynomial.legendre.legval(r, c) * np.sqrt((2 * j + 1) / 2.0)
    return V


def extrap_coeffs(m, M, method="MINNORM"):
    """Extrap::extrapCoeffs: c with sum_i c_i p(r_i) = p(1 + h) for polynomials of degree <= m, r_i = -1 + i h"""
    if M < m + 1:
        raise ValueError(f"Extrapolation space dimension ({M}) too low for degree ({m}).")
    h = 2.0 / (M - 1)
    r = -1.0 + h * np.arange(M)
    V = vandermonde1d(m, r)                 # M x (m+1): solve c V = b
    b = vandermonde1d(m, [1.0 + h])[0]
    if method == "MINNORM":
        return V @ np.linalg.solve(V.T @ V, b)
    Q, Rm, piv = scipy.linalg.qr(V.T, pivoting=True)   # V^T P = Q R (dgeqp3)
    y = scipy.linalg.solve_triangular(Rm[:, : m + 1], Q.T @ b)
    c = np.zeros(M)
    c[piv[: m + 1]] = y
    return c


class Extrap:
    def __init__(self, N, history=4, extrap_degree=2, coeffs_method="MINNORM", **_):
        self.N, self.Nh, self.m, self.method = N, history, extrap_degree, coeffs_method
        self.entry, self.shift = 0, 0
        self.xh = np.zeros((history, N))
        self.d = np.zeros(history)

    def form(self, x, rhs):
        Nh = self.Nh
        if self.entry < Nh:                 # the first solves use the history that exists so far
            if self.entry == Nh - 1:
                M, m = Nh, self.m
            else:
                M = max(1, self.entry + 1)
                m = int(np.sqrt(float(M)))
            d = np.zeros(Nh)
            if M == 1:
                d[Nh - 1] = 1.0
            else:
                d[Nh - M:] = extrap_coeffs(m, M, self.method)
            self.d = d
            self.entry += 1
        out = np.zeros(self.N)
        for i in range(Nh):
            ci = self.d[i]
            if (self.method == "MINNORM" and ci != 0.0) or (self.method != "MINNORM" and abs(ci) > 1e-14):
                out += ci * self.xh[(i + self.shift) % Nh]
        return out

    def update(self, A, x, rhs):
        self.xh[self.shift] = x
        self.shift = (self.shift + 1) % self.Nh


KINDS = {"ZERO": Zero, "CLASSIC": ClassicProjection, "QR": RollingQRProjection, "EXTRAP": Extrap}
