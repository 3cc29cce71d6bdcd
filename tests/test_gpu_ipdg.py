"""DISCRETIZATION = IPDG on hexahedra through the C ABI (csrc/ipdg.cu) against dumps of the UNMODIFIED reference
(tests/golden/ipdg_*.npz: oracle/refbuild/dump_ipdg_driver.cpp), the oracle (oracle/ipdg_ref.py) and the reference's own
regression cases (test/testElliptic.py:271-275, 355-358)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from libparanumal_b200 import api
from libparanumal_b200.api import EllipticIpdg, Pcg, Precon
from libparanumal_b200.problem import IpdgProblem
from oracle import ipdg_ref as ip

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SINGLE = ["ipdg_n1_e4_none", "ipdg_n2_e3_jacobi", "ipdg_n3_e3_periodic", "ipdg_n4_e3_none", "ipdg_n7_e2_jacobi"]


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


@pytest.mark.parametrize("name", SINGLE)
def test_ipdg_operator_on_reference_arrays(name):
    """gradient, operator, diagonal and the PCG solve on the reference's own vgeo / sgeo / vmapM / vmapP / EToB"""
    g = load(name)
    N = int(g["config_N"])
    Nq, Np = N + 1, (N + 1) ** 3
    E = int(g["r0_meta"][3])
    lam, tau = float(g["r0_dmeta"][0]), float(g["r0_dmeta"][1])
    vgeo, sgeo, D = dev(g["r0_vgeo"]), dev(g["r0_sgeo"]), dev(g["r0_D"])
    vmapM, vmapP, EToB = dev(g["r0_vmapM"]), dev(g["r0_vmapP"]), dev(g["r0_EToB"])
    op = EllipticIpdg(Nq, E, vmapM, vmapP, vgeo, sgeo, EToB, D, lam, tau)
    q = dev(g["r0_q"])
    Aq = torch.full_like(q, float("nan"))
    op.Operator(q, Aq)
    assert rel(Aq.cpu().numpy(), g["r0_Aq"]) < 1e-12
    assert rel(op.gradient().cpu().numpy().reshape(-1), g["r0_grad"]) < 1e-12
    oracle = ip.operator_single_rank(Nq, g["r0_vgeo"], g["r0_sgeo"], g["r0_vmapM"], g["r0_vmapP"], g["r0_EToB"], g["r0_D"],
                                     lam, tau, g["r0_q"])
    assert rel(Aq.cpu().numpy(), oracle) < 1e-12
    A = torch.empty_like(q)
    api.elliptic_build_diagonal_ipdg_hex3d(Nq, E, vgeo, sgeo, EToB, D, lam, tau, A)
    assert rel(A.cpu().numpy(), g["r0_diagA"]) < 1e-12
    M = Precon.Jacobi(E * Np, 1.0 / A) if str(g["config_precon"]) == "JACOBI" else Precon.Identity(E * Np)
    x = torch.zeros_like(q)
    solver = Pcg(E * Np, 0, api.Comm())
    it = solver.Solve(op, M, x, dev(g["r0_r"]), tol=1e-8, maxit=5000)
    assert abs(it - int(g["r0_iterations"][0])) <= 1
    h = np.array(solver.residual_history())
    k = min(len(h), len(g["pcg_history"]), 15)
    assert np.allclose(h[:k], g["pcg_history"][:k], rtol=1e-6)
    assert rel(x.cpu().numpy(), g["r0_xsol"]) < 1e-7
    op.Free()


@pytest.mark.parametrize("name", SINGLE)
def test_ipdg_self_built_problem_vs_reference(name):
    """nothing but the box from outside: device vgeo / sgeo, harness connectivity, boundary types, right-hand side of
    elliptic_t::Run, solve - against the reference's arrays and its printed iteration count / solution norm"""
    g = load(name)
    N, box, flag, lam = int(g["config_N"]), [int(v) for v in g["config_box"]], int(g["config_flag"]), float(g["config_lambda"])
    p = IpdgProblem(N, box[0], box[1], box[2], lam=lam, boundary_flag=flag)
    assert p.tau == float(g["r0_dmeta"][1])
    assert np.array_equal(p.vmapM.cpu().numpy().reshape(-1), g["r0_vmapM"])
    assert np.array_equal(p.vmapP.cpu().numpy().reshape(-1), g["r0_vmapP"])
    assert np.array_equal(p.mapP.cpu().numpy().reshape(-1), g["r0_mapP"])
    assert np.array_equal(p.EToB.cpu().numpy().reshape(-1), g["r0_EToB"])
    assert rel(p.vgeo.cpu().numpy(), g["r0_vgeo"]) < 1e-12
    assert rel(p.sgeo.cpu().numpy(), g["r0_sgeo"]) < 1e-12
    Aq = p.operator(dev(g["r0_q"]))
    assert rel(Aq.cpu().numpy(), g["r0_Aq"]) < 1e-11
    assert rel(p.diagonal().cpu().numpy(), g["r0_diagA"]) < 1e-11
    assert rel(p.rhs_sine3d().cpu().numpy(), g["r0_r"]) < 1e-11
    it, norm, x = p.run(precon=str(g["config_precon"]))
    assert abs(it - int(g["r0_iterations"][0])) <= 1
    assert abs(norm - float(g["r0_solnorm"][0])) < 1e-9
    assert rel(x.cpu().numpy(), g["r0_xsol"]) < 1e-7


@pytest.mark.parametrize("precon,listed,printed,its", [("NONE", 0.353553390458384, 0.353553386728646, 222),
                                                       ("JACOBI", 0.353553400508458, 0.353553389529891, 155)])
def test_reference_suite_hex_ipdg(precon, listed, printed, its):
    """testEllipticHex_Ipdg / testEllipticHex_Ipdg_Jacobi (test/testElliptic.py:271-275, 355-358): Hex N=4, 10^3 box,
    lambda = 1 (default), `Solution norm` within the suite's 1e-5; `printed` / `its` are what the unmodified build
    prints here (oracle/refbuild/make_golden_ipdg.py suite)"""
    p = IpdgProblem(4, 10, lam=1.0)
    it, norm, _ = p.run(precon=precon)
    assert abs(norm - listed) < 1e-5 * listed
    assert abs(norm - printed) < 1e-8
    assert abs(it - its) <= 2, (it, its)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("name", ["ipdg_n3_e4x4x4_p2", "ipdg_n2_e5x4x3_p4", "ipdg_n2_e4x4x4_periodic_p4"])
def test_ipdg_multirank_on_one_gpu_matches_multirank_reference(name):
    """P processes on cuda:0, trace halo of the gradient through the library's halo exchange (kind HALO, 4 entries
    per node), the reference's own per-rank arrays; see tests/mr_gpu_worker.py (run_rank_ipdg)"""
    size = int(name.rsplit("_p", 1)[1])
    port = _free_port()
    procs = []
    for r in range(size):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(size), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), LIBP_MR_ONE_GPU="1", LIBP_P2P_TIMEOUT_MS="20000", LIBP_P2P_WINDOW_MB="8",
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mr_gpu_worker.py"), name], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o[-3000:]}"
    assert "MR_GPU_OK" in outs[0], outs[0][-2000:]
