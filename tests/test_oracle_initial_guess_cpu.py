"""The initial-guess oracle (oracle/initial_guess_ref.py) against the guesses the UNMODIFIED reference's InitialGuess
classes form over the same sequence of solves (tests/golden/ig_tridiag_n400.npz, oracle/refbuild/dump_ig_driver.cpp)."""
import os

import numpy as np
import pytest

from oracle import initial_guess_ref as ig

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ig_tridiag_n400.npz")))
N, K = int(G["N"]), int(G["K"])
CASES = [("zero", "ZERO", 0, 0, "MINNORM"), ("classic4", "CLASSIC", 4, 0, "MINNORM"), ("qr5", "QR", 5, 0, "MINNORM"),
         ("qr3", "QR", 3, 0, "MINNORM"), ("extrap_m2_M4_minnorm", "EXTRAP", 4, 2, "MINNORM"),
         ("extrap_m3_M6_minnorm", "EXTRAP", 6, 3, "MINNORM"), ("extrap_m2_M5_cpqr", "EXTRAP", 5, 2, "CPQR")]


def A(q):  # the dump driver's tridiagonal operator
    out = G["diag"] * q
    out[1:] -= q[:-1]
    out[:-1] -= q[1:]
    return out


@pytest.mark.parametrize("key,kind,hist,deg,method", CASES, ids=[c[0] for c in CASES])
def test_initial_guess_oracle_matches_reference(key, kind, hist, deg, method):
    rhs, sol, ref = G["rhs"].reshape(K, N), G["sol"].reshape(K, N), G["guess_" + key].reshape(K, N)
    s = ig.KINDS[kind](N, history=hist, extrap_degree=deg, coeffs_method=method)
    x = np.zeros(N)
    for k in range(K):
        x = s.form(x, rhs[k])
        scale = max(np.abs(ref[k]).max(), np.abs(sol[k]).max())
        assert np.abs(x - ref[k]).max() / scale < 1e-9, (key, k, np.abs(x - ref[k]).max() / scale)
        x = sol[k].copy()
        s.update(A, x, rhs[k])


def test_extrap_coefficients_reproduce_polynomials():
    for m, M in [(1, 2), (2, 4), (3, 6), (2, 8), (4, 5)]:
        h = 2.0 / (M - 1)
        r = -1.0 + h * np.arange(M)
        cm, cq = ig.extrap_coeffs(m, M, "MINNORM"), ig.extrap_coeffs(m, M, "CPQR")
        for deg in range(m + 1):
            for c in (cm, cq):
                assert abs(np.dot(c, r ** deg) - (1.0 + h) ** deg) < 1e-10 * max(1.0, (1 + h) ** deg)
        assert np.linalg.norm(cm) <= np.linalg.norm(cq) * (1 + 1e-12)
        assert np.count_nonzero(np.abs(cq) > 1e-14) <= m + 1


def test_library_extrap_coefficients_equal_the_oracle():
    """Extrap::extrapCoeffs is host code on both sides: libp_ig_extrap_coeffs (no GPU needed) == the oracle"""
    from libparanumal_b200 import api
    for m, M in [(1, 2), (2, 4), (3, 6), (2, 8), (4, 5), (2, 5)]:
        for method in ("MINNORM", "CPQR"):
            got, ref = api.extrap_coeffs(m, M, method), ig.extrap_coeffs(m, M, method)
            assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()), (m, M, method, got, ref)
