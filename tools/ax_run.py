#!/usr/bin/env python
"""Development tool: launch one configuration of the fused hex Ax kernel a few times (for ncu)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libparanumal_b200 import _lib as L, api
from libparanumal_b200.problem import EllipticProblem
ap = argparse.ArgumentParser()
ap.add_argument("--elements", type=int, default=32)
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--lam", type=float, default=0.0)
ap.add_argument("--variant", type=int, default=1)
ap.add_argument("--tune", default="")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--mode", default="fused", choices=["fused", "local", "operator"])
a = ap.parse_args()
api.init(0)
lib = L.load()
p = EllipticProblem(a.degree, a.elements, lam=a.lam)
api.register_D(p.Nq, p.mesh.D)  # GLL matrix: profile the even-odd kernel the operator handle launches
m = p.mesh
q = p.vec(); q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda") * 2 - 1
Aq = p.vec()
L.check(lib.libp_ax_hex3d_set_variant(a.variant))
if a.tune:
    L.check(lib.libp_ax_hex3d_tune(*[int(x) for x in a.tune.split(",")]))
AqL = torch.empty(m.Nelements * m.Np, dtype=torch.float64, device="cuda") if a.mode == "local" else None
for _ in range(a.reps):
    if a.mode == "fused":
        Aq.zero_()
        api.ax_hex3d_gather(p.Nq, m.Nelements, None, p.GlobalToLocal, m.wJ, m.ggeo, m.D, a.lam, q, Aq)
    elif a.mode == "local":
        api.ax_hex3d(p.Nq, m.Nelements, None, p.GlobalToLocal, m.wJ, m.ggeo, m.D, a.lam, q, AqL)
    else:
        p.op.Operator(q, Aq)
torch.cuda.synchronize()
print("ok")
