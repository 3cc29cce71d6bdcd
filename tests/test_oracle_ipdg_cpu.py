"""The IPDG oracle (oracle/ipdg_ref.py) against dumps of the UNMODIFIED reference run with DISCRETIZATION = IPDG
(tests/golden/ipdg_*.npz, oracle/refbuild/dump_ipdg_driver.cpp + make_golden_ipdg.py), single rank and 2 / 4 ranks."""
import os

import numpy as np
import pytest

from oracle import ipdg_ref as ip

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SINGLE = ["ipdg_n1_e4_none", "ipdg_n2_e3_jacobi", "ipdg_n3_e3_periodic", "ipdg_n4_e3_none", "ipdg_n7_e2_jacobi"]
MULTI = ["ipdg_n3_e4x4x4_p2", "ipdg_n2_e5x4x3_p4", "ipdg_n2_e4x4x4_periodic_p4"]


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


@pytest.mark.parametrize("name", SINGLE)
def test_ipdg_oracle_vs_reference_single_rank(name):
    g = load(name)
    N = int(g["config_N"])
    Nq, Np = N + 1, (N + 1) ** 3
    E = int(g["r0_meta"][3])
    lam, tau = float(g["r0_dmeta"][0]), float(g["r0_dmeta"][1])
    assert tau == ip.tau_hex(N)
    gr = ip.gradient(Nq, g["r0_vgeo"], g["r0_D"], g["r0_q"])
    assert rel(gr.reshape(-1), g["r0_grad"][: gr.size]) < 1e-13
    args = (Nq, g["r0_vgeo"], g["r0_sgeo"], g["r0_vmapM"], g["r0_vmapP"], g["r0_EToB"], g["r0_D"], lam, tau)
    assert rel(ip.ax_ipdg(*args, gr), g["r0_Aq"]) < 1e-13
    assert rel(ip.diagonal(Nq, g["r0_vgeo"], g["r0_sgeo"], g["r0_EToB"], g["r0_D"], lam, tau), g["r0_diagA"]) < 1e-13
    # connectivity conventions: vmapM = element * Np + face node table
    vm = (np.arange(E)[:, None, None] * Np + ip.face_nodes(Nq)[None]).reshape(-1)
    assert np.array_equal(vm, g["r0_vmapM"])
    if "r0_x" in g:
        sg, _ = ip.surface_factors(Nq, g["r0_x"], g["r0_y"], g["r0_z"], g["r0_D"], g["r0_gllw"], g["r0_mapP"])
        assert rel(sg.reshape(-1), g["r0_sgeo"]) < 1e-12
    inv = 1.0 / g["r0_diagA"] if str(g["config_precon"]) == "JACOBI" else None
    it, x, hist = ip.pcg(lambda p: ip.operator_single_rank(*args, p), inv, g["r0_r"])
    assert it == int(g["r0_iterations"][0])
    k = min(len(hist), len(g["pcg_history"]))
    m = min(k, 15)  # rounding differences grow along an unpreconditioned CG history: tight at the start, loose overall
    assert np.allclose(hist[:m], g["pcg_history"][:m], rtol=1e-6)
    assert np.allclose(hist[:k], g["pcg_history"][:k], rtol=0.2)
    assert rel(x, g["r0_xsol"]) < 1e-7  # both solves stop at the same 1e-8 residual; the iterates differ by rounding


@pytest.mark.parametrize("name", MULTI)
def test_ipdg_oracle_vs_reference_multi_rank(name):
    """per rank: gradient of the rank's own elements, trace values of the neighbours looked up through the ids
    mesh_t::HaloTraceSetup hands to the halo (|id| - 1 = global node), operator of the rank's elements"""
    g = load(name)
    P = int(g["config_P"])
    N = int(g["config_N"])
    Nq, Np = N + 1, (N + 1) ** 3
    own = {}
    grads = []
    for r in range(P):
        E = int(g[f"r{r}_meta"][3])
        gr = ip.gradient(Nq, g[f"r{r}_vgeo"], g[f"r{r}_D"], g[f"r{r}_q"])
        grads.append(gr.reshape(E * Np, 4))
        off = int(g[f"r{r}_elementOffset"][0])
        own[r] = (off * Np, grads[-1])
    glob = np.concatenate([grads[r] for r in range(P)])  # ranks hold consecutive global elements
    for r in range(P):
        E, Eh = int(g[f"r{r}_meta"][3]), int(g[f"r{r}_meta"][4])
        lam, tau = float(g[f"r{r}_dmeta"][0]), float(g[f"r{r}_dmeta"][1])
        ids = g[f"r{r}_traceGlobalIds"]
        full = np.zeros(((E + Eh) * Np, 4))
        full[: E * Np] = grads[r]
        halo = np.nonzero(ids[E * Np:] < 0)[0] + E * Np
        full[halo] = glob[-ids[halo] - 1]
        ref = g[f"r{r}_grad"].reshape(-1, 4)
        assert rel(full[halo], ref[halo]) < 1e-13
        Aq = ip.ax_ipdg(Nq, g[f"r{r}_vgeo"], g[f"r{r}_sgeo"], g[f"r{r}_vmapM"], g[f"r{r}_vmapP"], g[f"r{r}_EToB"],
                        g[f"r{r}_D"], lam, tau, full, Nelements=E)
        assert rel(Aq, g[f"r{r}_Aq"]) < 1e-13
        assert rel(ip.diagonal(Nq, g[f"r{r}_vgeo"], g[f"r{r}_sgeo"], g[f"r{r}_EToB"], g[f"r{r}_D"], lam, tau),
                   g[f"r{r}_diagA"]) < 1e-13


@pytest.mark.parametrize("name", SINGLE[:3] + MULTI)
def test_harness_dg_connectivity_matches_reference(name):
    """BoxMesh.dg_connectivity / dg_trace_halo_ids (the harness' stand-in for mesh_t::ConnectFaceNodes, HaloSetup and
    HaloTraceSetup on the box) against the reference's arrays on every rank: vmapM, vmapP, mapP, boundary flags, the
    element halo (slots in (element, face) order), internal / halo element lists and the trace-halo ids - bit for bit"""
    from libparanumal_b200.box_mesh import BoxMesh
    g = load(name)
    N, box, flag, P = int(g["config_N"]), [int(v) for v in g["config_box"]], int(g["config_flag"]), int(g["config_P"])
    for r in range(P):
        k = f"r{r}_"
        m = BoxMesh(N, box[0], box[1], box[2], rank=r, size=P, boundary_flag=flag, device="cpu", geometry=False)
        vM, vP, mP, EB = m.dg_connectivity()
        h = m.dg_halo
        assert np.array_equal(vM.numpy().reshape(-1), g[k + "vmapM"])
        assert np.array_equal(vP.numpy().reshape(-1), g[k + "vmapP"])
        assert np.array_equal(mP.numpy().reshape(-1), g[k + "mapP"])
        assert np.array_equal(EB.numpy().reshape(-1), g[k + "meshEToB"])
        assert h["totalHaloPairs"] == int(g[k + "meta"][4])
        assert np.array_equal(h["internalElementIds"].numpy(), g[k + "internalElementIds"])
        assert np.array_equal(h["haloElementIds"].numpy(), g[k + "haloElementIds"])
        assert np.array_equal(m.dg_trace_halo_ids().numpy(), g[k + "traceGlobalIds"])
        assert int(h["elementOffsets"][r]) == int(g[k + "elementOffset"][0])
