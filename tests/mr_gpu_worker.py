"""One rank of the multi-rank GPU parity check against the multi-rank REFERENCE fixtures (tests/golden/mr_*.npz).

Started P times by tests/test_gpu_multirank_onegpu.py (all ranks on cuda:0 - kernels of different processes
time-slice, the NVLink peer-window protocol runs over CUDA IPC between them exactly as between GPUs) and by
tests/multigpu_check.py under torchrun (one rank per GPU).  Host collectives over gloo; no NCCL needed: the hot path
(halo exchange, cross-rank combine, PCG scalar all-reduces) runs through the peer window, setup-time gathers are
staged through the host collectives.

Checks per rank, against dumps of the unmodified reference run on the same number of ranks:
  maps bit-exact (signed ids after setup, GlobalToLocal, counters); Operator(q) <= 1e-12 relative on the reference's
  own D / ggeo / wJ / q; the halo tail the exchange leaves in q exact; Jacobi diagonal 1e-12; gathered right-hand side;
  PCG iteration count +-1, residual history, solution 1e-7."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_rank(rank, size, name, device_index, modes=(1, 0), comm=None, group=None):
    from libparanumal_b200 import _lib as L
    from libparanumal_b200 import api
    from libparanumal_b200.api import Comm, Precon
    from libparanumal_b200.box_mesh import BoxMesh
    from libparanumal_b200.problem import EllipticProblem
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    N, box, flag, lam = int(g["config_N"]), [int(x) for x in g["config_box"]], int(g["config_flag"]), float(g["config_lambda"])
    assert int(g["config_P"]) == size
    k = f"r{rank}_"
    if comm is None:
        api.init(device_index)
        comm = Comm(rank, size)
        assert comm.init_p2p(required=True), "peer window (CUDA IPC) unavailable"
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = {}
    for mode in modes:
        mesh = BoxMesh(N, box[0], box[1], box[2], rank, size, flag, device="cuda", geometry=False)
        E, Np = mesh.Nelements, mesh.Np
        mesh.ggeo, mesh.wJ, mesh.D = dev(g[k + "ggeo"]).reshape(E, 6, Np), dev(g[k + "wJ"]).reshape(E, Np), dev(g[k + "D"])
        ctypes.CDLL("libc.so.6").srand(1)  # every MPI rank of the reference is a fresh process
        p = EllipticProblem(N, box[0], box[1], box[2], lam=lam, boundary_flag=flag, comm=comm, mesh=mesh, mode=mode)
        assert np.array_equal(p.maskedGlobalIds, g[k + "maskedGlobalIds"]), "signed ids (owner choice)"
        assert np.array_equal(p.G2L_host, g[k + "GlobalToLocal"]), "GlobalToLocal"
        c = [int(x) for x in g[k + "ogs_counts"]]
        assert [p.ogs.N, p.ogs.Ngather, p.ogs.NlocalT, p.ogs.NlocalP, p.ogs.NhaloT, p.ogs.NhaloP, p.ogs.NgatherGlobal,
                p.Nhalo] == c, "counters"
        q = p.vec()
        q[: p.Ndofs] = dev(g[k + "q"])
        Aq = p.vec(fill=float("nan"))
        p.op.Operator(q, Aq)
        ref = g[k + "Aq"]
        t = torch.tensor([float(np.abs(ref).max()) if ref.size else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        scale = float(t[0])
        err = float(np.abs(Aq[: p.Ndofs].cpu().numpy() - ref).max() / scale) if ref.size else 0.0
        assert err < 1e-12, ("Operator", mode, err)
        assert np.array_equal(q.cpu().numpy(), g[k + "q_after_halo"]), "halo exchange of q"
        # Jacobi diagonal (device gather with the cross-rank combine) and the solve on the reference's right-hand side
        inv = p.inv_diagonal()
        dg = g[k + "diagA"]
        derr = float(np.abs(1.0 / inv[: p.Ndofs].cpu().numpy() - dg).max() / np.abs(dg).max()) if dg.size else 0.0
        assert derr < 1e-12, ("diagA", derr)
        M = Precon.Identity(p.Ndofs) if str(g["config_precon"]) == "NONE" else Precon.Jacobi(p.Ndofs, inv)
        r = p.vec()
        r[: p.Ndofs] = dev(g[k + "r"])
        x = p.vec()
        solver = p.pcg()
        it = solver.Solve(p.op, M, x, r, tol=1e-8, maxit=5000)
        it_ref = int(g[k + "iterations"][0])
        assert abs(it - it_ref) <= 1, ("iterations", it, it_ref)
        xs = g[k + "xsol"]
        t = torch.tensor([float(np.abs(xs).max()) if xs.size else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        xerr = float(np.abs(x[: p.Ndofs].cpu().numpy() - xs).max() / float(t[0])) if xs.size else 0.0
        assert xerr < 1e-7, ("solution", xerr)
        h, href = np.array(solver.residual_history()), g["pcg_history"]
        kk = min(len(h), len(href)) - 1
        assert np.allclose(h[:kk], href[:kk], rtol=1e-5), ("history", h[:4], href[:4])
        assert not comm.p2p_timed_out()
        out[mode] = dict(operator_err=err, diag_err=derr, iterations=it, iterations_ref=it_ref, x_err=xerr)
        p.op.Free()
    return out


def run_rank_ipdg(rank, size, name, device_index, group=None):
    """DISCRETIZATION = IPDG on P ranks (tests/golden/ipdg_*_p<P>.npz): the reference's per-rank vgeo / sgeo / vmapM /
    vmapP / EToB / element lists, the trace halo set up through the C ABI from mesh_t::HaloTraceSetup's ids; checks the
    exchanged gradient traces, Operator(q), the diagonal and the Jacobi-PCG iteration count against the dumps."""
    from libparanumal_b200 import _lib as L
    from libparanumal_b200 import api
    from libparanumal_b200.api import Comm, EllipticIpdg, Ogs, Pcg, Precon
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    N, lam = int(g["config_N"]), float(g["config_lambda"])
    assert int(g["config_P"]) == size
    k = f"r{rank}_"
    Nq, Np = N + 1, (N + 1) ** 3
    api.init(device_index)
    comm = Comm(rank, size)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    meta = [int(v) for v in g[k + "meta"]]
    E, Eh = meta[3], meta[4]
    tau = float(g[k + "dmeta"][1])
    ids = g[k + "traceGlobalIds"].astype(np.int64).copy()
    halo = Ogs().Setup(ids.size, ids, comm, kind=L.HALO, unique=False)
    op = EllipticIpdg(Nq, E, dev(g[k + "vmapM"]), dev(g[k + "vmapP"]), dev(g[k + "vgeo"]), dev(g[k + "sgeo"]),
                      dev(g[k + "EToB"]), dev(g[k + "D"]), lam, tau, NhaloElementsTotal=Eh, traceHalo=halo,
                      internalElementIds=dev(g[k + "internalElementIds"]), haloElementIds=dev(g[k + "haloElementIds"]))
    q = torch.zeros((E + Eh) * Np, dtype=torch.float64, device="cuda")
    q[: E * Np] = dev(g[k + "q"])
    Aq = torch.full_like(q, float("nan"))
    op.Operator(q, Aq)

    def gmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t[0])
    ref = g[k + "Aq"]
    err = float(np.abs(Aq[: E * Np].cpu().numpy() - ref).max()) / gmax(np.abs(ref).max())
    assert err < 1e-12, ("Operator", err)
    gr, gref = op.gradient().cpu().numpy(), g[k + "grad"].reshape(-1, 4)
    tr = np.nonzero(ids < 0)[0]  # the trace nodes this rank receives
    gerr = float(np.abs(gr[tr] - gref[tr]).max()) / gmax(np.abs(gref).max()) if tr.size else 0.0
    assert gerr < 1e-12, ("trace halo of the gradient", gerr)
    A = torch.empty(E * Np, dtype=torch.float64, device="cuda")
    api.elliptic_build_diagonal_ipdg_hex3d(Nq, E, dev(g[k + "vgeo"]), dev(g[k + "sgeo"]), dev(g[k + "EToB"]), dev(g[k + "D"]),
                                           lam, tau, A)
    derr = float(np.abs(A.cpu().numpy() - g[k + "diagA"]).max() / np.abs(g[k + "diagA"]).max())
    assert derr < 1e-12, ("diagA", derr)
    M = Precon.Identity(E * Np) if str(g["config_precon"]) == "NONE" else Precon.Jacobi(E * Np, 1.0 / A)
    r = torch.zeros_like(q)
    r[: E * Np] = dev(g[k + "r"])
    x = torch.zeros_like(q)
    solver = Pcg(E * Np, Eh * Np, comm)
    it = solver.Solve(op, M, x, r, tol=1e-8, maxit=5000)
    it_ref = int(g[k + "iterations"][0])
    assert abs(it - it_ref) <= 1, ("iterations", it, it_ref)
    xs = g[k + "xsol"]
    xerr = float(np.abs(x[: E * Np].cpu().numpy() - xs).max()) / gmax(np.abs(xs).max())
    assert xerr < 1e-7, ("solution", xerr)
    op.Free()
    return dict(operator_err=err, trace_err=gerr, diag_err=derr, iterations=it, iterations_ref=it_ref, x_err=xerr)


def main():
    rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    name = sys.argv[1]
    one_gpu = os.environ.get("LIBP_MR_ONE_GPU", "0") == "1"
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        device = 0 if one_gpu else int(os.environ.get("LOCAL_RANK", rank))
        res = run_rank_ipdg(rank, size, name, device) if name.startswith("ipdg_") else run_rank(rank, size, name, device)
        dist.barrier()
        if rank == 0:
            print("MR_GPU_OK", name, res)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
