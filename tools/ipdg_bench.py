"""IPDG operator / Jacobi-PCG throughput on one B200 (jsonl on stdout):
   python tools/ipdg_bench.py [--degree 7] [--elements 48] [--lam 1] [--steps 30]
Byte model per node of one apply: gradient pass q 8 + vgeo 72 + grad 32 (write); element pass grad 32
+ vgeo 80 + Aq 8, plus per FACE node sgeo 40 + vmapM/P 8 + the neighbour's grad 32 (6 Nq^2 face nodes per Nq^3 nodes)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libparanumal_b200 import api  # noqa: E402
from libparanumal_b200.problem import IpdgProblem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--elements", type=int, default=48)
ap.add_argument("--lam", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--pcg-iters", type=int, default=30)
a = ap.parse_args()
api.init(0)
peak = 6545.3
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
p = IpdgProblem(a.degree, a.elements, lam=a.lam)
Nq = a.degree + 1
n = p.Ndofs
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
Aq = torch.empty_like(q)
for _ in range(3):
    p.op.Operator(q, Aq)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    p.op.Operator(q, Aq)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
face = 6.0 / Nq
bytes_apply = n * ((8 + 72 + 32) + (32 + 80 + 8) + face * (40 + 8 + 32))
print(json.dumps({"kind": "ipdg_operator", "N": a.degree, "elements": a.elements, "lambda": a.lam, "dofs": n,
                  "ms_per_apply": ms, "gdofs": n / ms / 1e6, "model_bytes_per_node": bytes_apply / n,
                  "frac_of_hbm_peak": bytes_apply / ms / 1e6 / peak}), flush=True)
M = p.jacobi()
r = p.rhs_sine3d()
x = p.vec()
solver = p.pcg()
solver.Solve(p.op, M, x, r, tol=0.0, maxit=5)
x.zero_()
torch.cuda.synchronize()
e0.record()
it = solver.Solve(p.op, M, x, r, tol=0.0, maxit=a.pcg_iters)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / max(it, 1)
print(json.dumps({"kind": "ipdg_jacobi_pcg", "N": a.degree, "elements": a.elements, "iterations": it,
                  "ms_per_iteration": ms, "gdofs": n / ms / 1e6}), flush=True)
