"""A/B of the fused operator kernels on one B200 (jsonl on stdout):
   python tools/chain_tune.py [--degree 7] [--elements 64] [--lam 0] [--steps 50]
For every (chain length, stages) - chain 0 = ax_hex3d_t_kernel (per-thread loads, memset + red.add everywhere) - one
line with the device-timed apply (whole step), its parts (zero-fill, Ax launches), the plan statistics and the
difference to the chain-0 result."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from libparanumal_b200 import api  # noqa: E402
from libparanumal_b200.problem import EllipticProblem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--elements", type=int, default=64)
ap.add_argument("--lam", type=float, default=0.0)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--grid", default="0:2,4:2,8:2,16:2,32:2,64:2,16:3,32:3")
ap.add_argument("--trilinear", action="store_true",
                help="ELEMENT MAP = TRILINEAR: geometry recomputed from the element vertices (libp_elliptic_set_trilinear); "
                     "fractions are of the trilinear byte model 16 B per DOF + 192 B per element")
a = ap.parse_args()
api.init(0)
peak = 6545.3
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
p = EllipticProblem(a.degree, a.elements, lam=a.lam)
g = torch.Generator(device="cuda").manual_seed(1234)
q = p.vec()
q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
E, Np = p.mesh.Nelements, p.mesh.Np
alg = 8.0 * (6 + (a.lam != 0)) * E * Np + 16.0 * p.Ndofs
if a.trilinear:
    import numpy as np
    ex, ey, ez = p.mesh.element_vertices()
    EXYZ = torch.stack([ex, ey, ez], dim=1).contiguous().reshape(-1)
    alg = 16.0 * p.Ndofs + 192.0 * E
ref = None
for item in a.grid.split(","):
    L, S = (int(v) for v in item.split(":"))
    p.op.set_chain(L, S)
    if a.trilinear and L > 0:
        p.op.set_trilinear(EXYZ, p.mesh.gllz, p.mesh.gllw)
    Aq = p.vec(fill=float("nan"))
    for _ in range(5):
        p.op.Operator(q, Aq)
    st = p.op.chain_stats(Aq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        p.op.Operator(q, Aq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    parts = [p.op.OperatorTimed(q, Aq) for _ in range(10)]
    zms = sum(x[0] for x in parts) / len(parts)
    kms = sum(x[1] for x in parts) / len(parts)
    if ref is None:
        ref = Aq.clone()
    diff = float((Aq[: p.Ndofs] - ref[: p.Ndofs]).abs().max() / ref[: p.Ndofs].abs().max())
    print(json.dumps({"N": a.degree, "elements": a.elements, "lambda": a.lam, "chain": L, "stages": S, "trilinear": bool(a.trilinear and L > 0),
                      "ms_per_apply": ms, "gdofs": p.NglobalDofs / ms / 1e6, "step_frac": alg / ms / 1e6 / peak,
                      "zero_fill_ms": zms, "ax_ms": kms, "kernel_frac": alg / kms / 1e6 / peak,
                      "plan": st, "rel_diff_vs_first": diff}), flush=True)
