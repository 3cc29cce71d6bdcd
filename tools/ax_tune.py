#!/usr/bin/env python
"""Development tool: times every compiled instantiation of the fused hex Ax kernel on one B200
(needs a library built with LIBP_AX_TUNE_GRID=1) and checks each against the pencil kernel."""
import argparse, itertools, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libparanumal_b200 import _lib as L, api
from libparanumal_b200.problem import EllipticProblem

ap = argparse.ArgumentParser()
ap.add_argument("--elements", type=int, default=48)
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--lam", type=float, default=0.0)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--grid", default="full")
args = ap.parse_args()
api.init(0)
lib = L.load()
p = EllipticProblem(args.degree, args.elements, lam=args.lam)
m = p.mesh
E, Np = m.Nelements, m.Np
q = p.vec(); q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda") * 2 - 1
Aq = p.vec()
alg = 8.0 * (6 + (args.lam != 0)) * E * Np + 16.0 * p.Ndofs

use_op = False
def run():
    if use_op:
        p.op.Operator(q, Aq)   # memset + Ax kernels through the operator handle (even-odd path when D is GLL)
    else:
        api.ax_hex3d_gather(p.Nq, E, None, p.GlobalToLocal, m.wJ, m.ggeo, m.D, args.lam, q, Aq)

def timeit():
    for _ in range(2):
        Aq.zero_(); run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        if not use_op:
            Aq.zero_()
        a.record(); run(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]

L.check(lib.libp_ax_hex3d_set_variant(0))
Aq.zero_(); run(); ref = Aq.clone()
med, mn = timeit()
print(json.dumps({"variant": "pencil", "ms_median": med, "ms_min": mn, "GBs": alg / med / 1e6}), flush=True)
L.check(lib.libp_ax_hex3d_set_variant(1))
med, mn = timeit()
print(json.dumps({"variant": "transposed-general-D(raw API, default cfg)", "ms_median": med, "ms_min": mn, "GBs": alg / med / 1e6}), flush=True)
use_op = True
grid = list(itertools.product([1, 2, 3, 4], [1], [4, 6, 8, 10])) + [(2, 0, 4), (2, 0, 6), (2, 0, 8), (2, 0, 10)]
if args.grid != "full":
    grid = [(2, 1, 6)]
for pf, hint, mb in grid:
    if lib.libp_ax_hex3d_tune(pf, hint, mb) != 0:
        print("tune grid not compiled in:", lib.libp_last_error().decode()); break
    Aq.zero_(); run()
    err = float((Aq - ref).abs().max() / ref.abs().max())
    med, mn = timeit()
    print(json.dumps({"variant": "operator(memset+even-odd kernels)", "pf": pf, "hint": hint, "minb": mb, "ms_median": med, "ms_min": mn,
                      "GBs": alg / med / 1e6, "frac_6555": alg / med / 1e6 / 6554.9, "relerr_vs_pencil": err}), flush=True)
