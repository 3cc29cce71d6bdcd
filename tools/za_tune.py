#!/usr/bin/env python
"""Development tool: operator step and Jacobi-PCG iteration time with the in-kernel zero-fill of the fused
accumulator (libp_elliptic_set_zero_ahead) off and on, on one B200."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libparanumal_b200 import api
from libparanumal_b200.problem import EllipticProblem

ap = argparse.ArgumentParser()
ap.add_argument("--elements", type=int, default=64)
ap.add_argument("--degree", type=int, default=7)
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--pcg-iters", type=int, default=30)
args = ap.parse_args()
api.init(0)
p = EllipticProblem(args.degree, args.elements, lam=0.0, coords=True)
q = p.vec(); q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda") * 2 - 1
Aq = p.vec()
E, Np = p.mesh.Nelements, p.mesh.Np


def time_op(reps):
    for _ in range(3):
        p.op.Operator(q, Aq)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        p.op.Operator(q, Aq)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ref = {}
for lam in (0.0, 1.0):
    p.set_lambda(lam)
    alg = 8.0 * (6 + (lam != 0)) * E * Np + 16.0 * p.Ndofs
    M = p.jacobi() if lam else None
    r0 = p.rhs_sine3d() if lam else None
    for za in (0, 1, 0, 1):
        p.op.set_zero_ahead(za)
        ms = time_op(args.reps)
        out = p.operator(q).clone()
        ref.setdefault(lam, out)
        err = float((out[: p.Ndofs] - ref[lam][: p.Ndofs]).abs().max() / ref[lam][: p.Ndofs].abs().max())
        line = {"lambda": lam, "zero_ahead": za, "op_ms": ms, "op_gdofs": p.NglobalDofs / ms / 1e6, "op_roofline_GBs": alg / ms / 1e6,
                "relerr_vs_first": err, "protocol_errors": p.op.zero_ahead_errors()}
        if lam:
            solver = p.pcg()
            x, r = p.vec(), r0.clone()
            solver.Solve(p.op, M, x, r, tol=1e-30, maxit=3)
            x.zero_(); r.copy_(r0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            it = solver.Solve(p.op, M, x, r, tol=1e-30, maxit=args.pcg_iters)
            b.record(); torch.cuda.synchronize()
            h = solver.residual_history()
            line.update(pcg_ms_per_it=a.elapsed_time(b) / it, pcg_gdofs=p.NglobalDofs * it / a.elapsed_time(b) / 1e6,
                        pcg_res_last=float(h[-1]), protocol_errors=p.op.zero_ahead_errors())
        print(json.dumps(line), flush=True)
