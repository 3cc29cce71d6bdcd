#!/usr/bin/env python
"""Development tool: CUPTI timeline (via torch.profiler) of a few Jacobi-PCG iterations / operator applies.
Run under torchrun for multi-GPU; rank 0 writes gpurun_out/<tag>_trace.json (kernel name, start us, dur us, stream)."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libparanumal_b200 import api
from libparanumal_b200.api import Comm
from libparanumal_b200.problem import EllipticProblem

ap = argparse.ArgumentParser()
ap.add_argument("--elements", type=int, default=64)
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--tag", default="pcg")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
api.init(lr)
gloo = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    gloo = dist.new_group(backend="gloo")
comm = Comm(rank, world, gloo); comm.init_nccl()
if world > 1 and os.environ.get('LIBP_P2P', '1') != '0':
    comm.init_p2p()
p = EllipticProblem(a.degree, a.elements, lam=1.0, comm=comm, coords=True)
M = p.jacobi(); r0 = p.rhs_sine3d(); solver = p.pcg()
x, r = p.vec(), r0.clone()
solver.Solve(p.op, M, x, r, tol=1e-30, maxit=5)
q, Aq = p.vec(1.0), p.vec()
for _ in range(3):
    p.op.Operator(q, Aq)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    x.zero_(); r.copy_(r0)
    solver.Solve(p.op, M, x, r, tol=1e-30, maxit=a.iters)
    torch.cuda.synchronize()
    for _ in range(4):
        p.op.Operator(q, Aq)
    torch.cuda.synchronize()
if rank == 0:
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append({"name": e.name[:90], "start": e.time_range.start, "dur": e.time_range.end - e.time_range.start,
                       "stream": getattr(e, "device_resource_id", None) if hasattr(e, "device_resource_id") else None})
    ev.sort(key=lambda d: d["start"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(ev, open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_trace.json"), "w"))
    try:
        prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", f"{a.tag}_chrome.json"))
    except Exception as ex:
        print("chrome trace failed", ex)
    print("events", len(ev))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
