#!/usr/bin/env python
"""BASELINE configs[4]: Hex3D N=3..8 at ~30 M DOF per GPU (SURVEY section 8d, C5 box sizes), single rank or under
torchrun (weak scaling: every rank owns one such box, the global box is the Factor3 arrangement of them).
Prints one JSON line per degree: operator and Jacobi-PCG GDOF/s with their fractions of the HBM roofline."""
import argparse
import json
import os
import sys

# the CPU leg (--cpu-seconds) uses the host's physical cores; libgomp reads OMP_NUM_THREADS once, when it is first
# loaded, so bench.py (whose import sets it for single-process runs) comes before torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402,F401

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libparanumal_b200 import api  # noqa: E402
from libparanumal_b200.api import Comm  # noqa: E402
from libparanumal_b200.box_mesh import factor3  # noqa: E402
from libparanumal_b200.problem import EllipticProblem  # noqa: E402

BOX = {1: 310, 2: 155, 3: 104, 4: 78, 5: 62, 6: 52, 7: 44, 8: 39}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--degrees", default="3,4,5,6,7,8")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--pcg-iters", type=int, default=30)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the per-GPU box edge (smoke runs)")
    ap.add_argument("--cpu-seconds", type=float, default=0.0,
                    help="> 0: also time the reference's own CPU kernels of every degree on the host cores (rank 0, "
                         "single-GPU runs; about this many seconds of applies per degree, box of ~0.6 M DOF)")
    ap.add_argument("--chains", default="", help="comma list of chain:stages configurations of the fused operator "
                    "(0:1 = ax_hex3d_t_kernel); empty = library default")
    a = ap.parse_args()
    import torch.distributed as dist
    world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    api.init(lr)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        gloo = dist.new_group(backend="gloo")
    comm = Comm(rank, world, gloo)
    comm.init_nccl()
    if world > 1 and os.environ.get("LIBP_P2P", "1") != "0":
        comm.init_p2p()
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    sx, sy, sz = factor3(world)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for N in [int(x) for x in a.degrees.split(",")]:
        n = max(2, int(round(BOX[N] * a.scale)))
        p = None
        for lam, cfg in [(l, c) for l in (0.0, 1.0) for c in (a.chains.split(",") if a.chains else [""])]:
            if p is None:
                p = EllipticProblem(N, n * sx, n * sy, n * sz, lam=lam, comm=comm, coords=True)
            elif p.lam != lam:
                p.set_lambda(lam)   # same mesh, maps and ogs: only the operator handle changes
            E, Np = p.mesh.Nelements, p.mesh.Np
            out = {"N": N, "elements_per_gpu": [n, n, n], "n_gpus": world, "lambda": lam, "global_dofs": int(p.NglobalDofs),
                   "local_dofs": int(p.Ndofs)}
            if cfg:
                L_, S_ = (int(v) for v in cfg.split(":"))
                p.op.set_chain(L_, S_)
                out.update(chain=L_, stages=S_)
            ax_bytes = 8.0 * (6 + (lam != 0.0)) * E * Np + 16.0 * p.Ndofs
            if lam == 0.0:
                q, Aq = p.vec(), p.vec()
                q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda") * 2 - 1
                for _ in range(5):
                    p.op.Operator(q, Aq)
                ms = timed(lambda: p.op.Operator(q, Aq), a.steps) / a.steps
                out.update(kind="operator", ms=ms, gdofs=p.NglobalDofs / ms / 1e6, roofline_frac=ax_bytes / ms / 1e6 / peak)
                del q, Aq
            else:
                M, r0, solver = p.jacobi(), p.rhs_sine3d(), p.pcg()
                x, r = p.vec(), r0.clone()
                solver.Solve(p.op, M, x, r, tol=1e-30, maxit=3)
                x.zero_(); r.copy_(r0)
                res = {}
                def solve():
                    res["it"] = solver.Solve(p.op, M, x, r, tol=1e-30, maxit=a.pcg_iters)
                ms = timed(solve, 1) / max(res["it"], 1)
                it_bytes = ax_bytes + 88.0 * p.Ndofs
                out.update(kind="jacobi_pcg", ms=ms, gdofs=p.NglobalDofs / ms / 1e6, roofline_frac=it_bytes / ms / 1e6 / peak,
                           iterations=res["it"])
                del M, r0, solver, x, r
            if rank == 0:
                print(json.dumps(out), flush=True)
        p.op.Free()
        del p
        torch.cuda.empty_cache()
        if a.cpu_seconds > 0 and rank == 0 and world == 1:
            sys.path.insert(0, ROOT)
            import bench
            nc = max(3, int(round((6.0e5) ** (1.0 / 3.0) / N)))
            gd, dt, threads, ngc, nap, kind = bench.cpu_ax_sample(N, nc, 3, 1, seconds=a.cpu_seconds)
            print(json.dumps({"N": N, "kind": "cpu_reference_operator", "gdofs": gd, "s_per_apply": dt, "cores": threads,
                              "cpu_kind": kind, "elements": [nc] * 3, "dofs": ngc, "applies": nap}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
