"""Element-chain fused operator kernel (csrc/ax_chain.cu: TMA-staged geometric factors, owner-computes stores,
compressed connectivity) against the oracle restatement of elliptic_t::Operator
(solvers/elliptic/src/ellipticOperator.cpp:31-106) and against the previous fused kernel, for every order, chain
length, stage count, both lambda branches, Dirichlet and periodic boxes (periodic boxes of 1-2 elements have ids
repeated inside one element: those rows must never take the plain-store path)."""
import ctypes

import numpy as np
import pytest
import torch

from libparanumal_b200 import api
from libparanumal_b200.problem import EllipticProblem

pytestmark = pytest.mark.gpu
libc = ctypes.CDLL("libc.so.6")
TOL = 1e-12  # relative, north_star


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


def _rand(p, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = p.vec()
    q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    return q


def _oracle_operator(N, n, lam, flag, q):
    from oracle import elliptic_ref as er
    from oracle.mesh_box import build_box_hex_mesh, masked_global_ids
    from oracle.ogs_ref import SIGNED, ogs_setup_all
    libc.srand(1)
    m = build_box_hex_mesh(N, n, n, n, boundary_flag=flag) if flag != 1 else build_box_hex_mesh(N, n, n, n)
    _, ids = masked_global_ids(m)
    o = ogs_setup_all([ids], SIGNED, True)[0]
    G2L = o.global_to_local()
    return er.operator(N + 1, G2L, m.wJ, m.ggeo, m.D, lam, o.gatherLocal.rowStartsT, o.gatherLocal.colIdsT, q)


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("lam", [0.0, 1.3])
def test_chain_vs_oracle(N, lam):
    n = 5 if N <= 5 else 4
    libc.srand(1)
    p = EllipticProblem(N, n, lam=lam)
    q = _rand(p, 7 + N)
    ref = _oracle_operator(N, n, lam, 1, q[: p.Ndofs].cpu().numpy())
    scale = np.abs(ref).max()
    for L, S in [(16, 2), (1, 2), (3, 3), (7, 2), (64, 3)]:
        p.op.set_chain(L, S)
        Aq = p.vec(fill=float("nan"))  # clean sectors are never zero-filled: stale contents must not leak
        p.op.Operator(q, Aq)
        st = p.op.chain_stats(Aq)
        assert st["chain"] == L and st["stages"] == S
        err = np.abs(Aq[: p.Ndofs].cpu().numpy() - ref).max() / scale
        assert err < TOL, (N, lam, L, S, err)
    p.op.set_chain(0)
    Aq = p.vec()
    p.op.Operator(q, Aq)
    assert np.abs(Aq[: p.Ndofs].cpu().numpy() - ref).max() / scale < TOL


@pytest.mark.parametrize("N,n", [(7, 1), (7, 2), (3, 2), (4, 3), (7, 3)])
@pytest.mark.parametrize("lam", [0.0, 1.0])
def test_chain_periodic_boxes(N, n, lam):
    """periodic boxes: ids repeated inside an element (n = 1, 2) and wrap-around neighbours"""
    libc.srand(1)
    p = EllipticProblem(N, n, lam=lam, boundary_flag=-1)
    q = _rand(p, 11)
    p.op.set_chain(0)
    ref = p.operator(q)[: p.Ndofs].cpu().numpy()
    for L, S in [(16, 2), (2, 3), (1, 2)]:
        p.op.set_chain(L, S)
        Aq = p.vec(fill=float("nan"))
        p.op.Operator(q, Aq)
        err = np.abs(Aq[: p.Ndofs].cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < TOL, (N, n, lam, L, S, err)


def test_chain_plan_statistics_n7():
    """x-chains of a Dirichlet box: most sectors are chain-private, only boundary elements stay uncompressed"""
    libc.srand(1)
    n = 8
    p = EllipticProblem(7, n, lam=0.0)
    Aq = p.vec()
    p.op.set_chain(8, 2)
    st = p.op.chain_stats(Aq)
    assert st["sectors"] == (p.ogs.NlocalT + p.ogs.NhaloT + 3) // 4
    assert 0 < st["zero_sectors"] < 0.6 * st["sectors"], st
    assert st["raw_elements"] <= n ** 3 - (n - 2) ** 3, st


def test_chain_unaligned_accumulator_takes_previous_kernel():
    libc.srand(1)
    p = EllipticProblem(7, 3, lam=1.0)
    q = _rand(p, 3)
    ref = p.operator(q)[: p.Ndofs].cpu().numpy()
    buf = torch.zeros(p.Nall + 1, dtype=torch.float64, device="cuda")
    Aq = buf[1:]  # 8-byte aligned only
    assert p.op.chain_stats(Aq)["chain"] == 0
    p.op.Operator(q, Aq)
    assert np.abs(Aq[: p.Ndofs].cpu().numpy() - ref).max() / np.abs(ref).max() < TOL


@pytest.mark.parametrize("N", [4, 7])
def test_chain_pcg_same_iterations(N):
    """Jacobi-PCG (masked zero-fill folded into the p update) with and without the chain kernel"""
    libc.srand(1)
    p = EllipticProblem(N, 6, lam=1.0, coords=True)
    M = p.jacobi()
    out = []
    for L in (0, 16, 5):
        p.op.set_chain(L, 2)
        x, r = p.vec(), p.rhs_sine3d()
        solver = p.pcg()
        it = solver.Solve(p.op, M, x, r, tol=1e-8, maxit=2000)
        out.append((it, x[: p.Ndofs].cpu().numpy(), np.array(solver.residual_history())))
    for it, x, h in out[1:]:
        assert abs(it - out[0][0]) <= 1
        assert np.abs(x - out[0][1]).max() / np.abs(out[0][1]).max() < 1e-7
        k = min(len(h), len(out[0][2])) - 2
        assert np.allclose(h[:k], out[0][2][:k], rtol=1e-6)
