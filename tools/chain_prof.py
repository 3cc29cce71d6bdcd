"""Development tool: a few applies of the fused operator with a given chain configuration (for ncu)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libparanumal_b200 import api
from libparanumal_b200.problem import EllipticProblem
ap = argparse.ArgumentParser()
ap.add_argument("--elements", type=int, default=32)
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--lam", type=float, default=0.0)
ap.add_argument("--chain", type=int, default=16)
ap.add_argument("--stages", type=int, default=2)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
api.init(0)
p = EllipticProblem(a.degree, a.elements, lam=a.lam)
p.op.set_chain(a.chain, a.stages)
q = p.vec(); q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda") * 2 - 1
Aq = p.vec()
for _ in range(a.reps):
    p.op.Operator(q, Aq)
torch.cuda.synchronize()
print("ok", p.op.chain_stats(Aq))
