"""Initial-guess strategies (csrc/initial_guess.cu) against the guesses the UNMODIFIED reference's InitialGuess classes
form over the same sequence of solves (tests/golden/ig_tridiag_n400.npz from oracle/refbuild/dump_ig_driver.cpp;
libs/linearSolver/initialGuess.cpp: ZERO, CLASSIC, QR, EXTRAP with MINNORM and CPQR coefficients)."""
import os

import numpy as np
import pytest
import torch

from libparanumal_b200 import api
from libparanumal_b200.api import InitialGuess

pytestmark = pytest.mark.gpu
G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ig_tridiag_n400.npz")))
N, K = int(G["N"]), int(G["K"])
CASES = [("zero", "ZERO", 0, 0, "MINNORM"), ("classic4", "CLASSIC", 4, 0, "MINNORM"), ("qr5", "QR", 5, 0, "MINNORM"),
         ("qr3", "QR", 3, 0, "MINNORM"), ("extrap_m2_M4_minnorm", "EXTRAP", 4, 2, "MINNORM"),
         ("extrap_m3_M6_minnorm", "EXTRAP", 6, 3, "MINNORM"), ("extrap_m2_M5_cpqr", "EXTRAP", 5, 2, "CPQR")]


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


@pytest.mark.parametrize("key,kind,hist,deg,method", CASES, ids=[c[0] for c in CASES])
def test_initial_guess_sequence_matches_reference(key, kind, hist, deg, method):
    d = torch.from_numpy(G["diag"]).cuda()
    rhs = torch.from_numpy(G["rhs"].reshape(K, N)).cuda()
    sol = torch.from_numpy(G["sol"].reshape(K, N)).cuda()
    ref = G["guess_" + key].reshape(K, N)

    def A(pin, pout):  # the dump driver's tridiagonal operator on raw device pointers
        # wrap the raw pointers as tensors through the CUDA array interface
        class _P:
            def __init__(self, p):
                self.__cuda_array_interface__ = {"shape": (N,), "typestr": "<f8", "data": (int(p), False), "version": 2}
        qi, qo = torch.as_tensor(_P(pin), device="cuda"), torch.as_tensor(_P(pout), device="cuda")
        out = d * qi
        out[1:] -= qi[:-1]
        out[:-1] -= qi[1:]
        qo.copy_(out)

    ig = InitialGuess(kind, N, 0, history=hist, extrap_degree=deg, coeffs_method=method)
    x = torch.zeros(N, dtype=torch.float64, device="cuda")
    for k in range(K):
        ig.FormInitialGuess(x, rhs[k].contiguous())
        got = x.cpu().numpy()
        scale = max(np.abs(ref[k]).max(), np.abs(G["sol"].reshape(K, N)[k]).max())
        assert np.abs(got - ref[k]).max() / scale < 1e-9, (key, k, np.abs(got - ref[k]).max() / scale)
        x.copy_(sol[k])
        ig.Update(A, x, rhs[k].contiguous())
    ig.Free()


def test_extrap_coefficients_reproduce_polynomials():
    """host-only part (Extrap::extrapCoeffs): sum_i c_i p(r_i) = p(1 + h) for every polynomial of degree <= m on the
    equispaced history points; MINNORM has the smaller 2-norm, CPQR at most m + 1 non-zeros"""
    for m, M in [(1, 2), (2, 4), (3, 6), (2, 8), (4, 5)]:
        h = 2.0 / (M - 1)
        r = -1.0 + h * np.arange(M)
        cm, cq = api.extrap_coeffs(m, M, "MINNORM"), api.extrap_coeffs(m, M, "CPQR")
        for deg in range(m + 1):
            for c in (cm, cq):
                assert abs(np.dot(c, r ** deg) - (1.0 + h) ** deg) < 1e-10 * max(1.0, (1 + h) ** deg)
        assert np.linalg.norm(cm) <= np.linalg.norm(cq) * (1 + 1e-12)
        assert np.count_nonzero(np.abs(cq) > 1e-14) <= m + 1
    with pytest.raises(Exception, match="too low"):
        api.extrap_coeffs(3, 3)
