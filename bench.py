#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native elliptic hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm, rank 0 only)

Workload (BASELINE.json configs[1]): BP5 matrix-free operator apply, Hex3D N=7, 64^3-element box
(lambda=0, all-Dirichlet), FP64.  One "step" = one elliptic_t::Operator apply (halo exchange, Ax with the
gather fused into its epilogue, cross-rank combine) on the global 64^3 box, strong-scaled over N GPUs.
value = NglobalDofs * steps / seconds / 1e9 (the reference's own "nodes*iterations/time" metric,
solvers/elliptic/src/ellipticRun.cpp:212-221), device-timed with CUDA events, max over ranks.
Extra objects on the same JSON line: roofline (dominant kernel vs measured HBM peak), cpu_baseline
(the reference's own JIT-compiled kernels, or the oracle C port, on the host cores), e2e (same step through the C ABI with HOST buffers), pcg
(Jacobi-PCG iteration throughput on the lambda=1 problem, BASELINE configs[2]).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The reference's OpenMP-mode gather kernel is a parallel loop over K = 1 (OCCA parallelises the outermost loop), so
# all but one thread idle through it; with libgomp's default spin-wait those idle threads cost the working one
# 5x here.  Passive waiting is the setting that favours the CPU baseline; it must be in place before libgomp starts.
os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def _physical_cores():
    """physical cores this process may run on (the reference derives its OpenMP thread count from lscpu,
    libs/core/platformDeviceConfig.cpp:103-153)"""
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        allowed = set(range(os.cpu_count() or 1))
    cores = set()
    try:
        for c in allowed:
            base = f"/sys/devices/system/cpu/cpu{c}/topology/"
            with open(base + "physical_package_id") as f:
                pkg = f.read().strip()
            with open(base + "core_id") as f:
                cid = f.read().strip()
            cores.add((pkg, cid))
    except OSError:
        return max(1, len(allowed))
    return max(1, len(cores))


# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms (the reference arm, the cpu_baseline leg of a
# single-GPU run) must use the host's cores, and libgomp reads the variable once, when it is first loaded.
_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if "--impl" in sys.argv and "reference" in sys.argv or _WORLD == 1:
    if os.environ.get("LIBP_BENCH_KEEP_OMP") != "1":
        os.environ["OMP_NUM_THREADS"] = str(_physical_cores())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GDOF/s (FP64) hex N=7 Ax & PCG solve at 1/2/4/8 B200; % HBM roofline"
CPU_KIND_TEXT = {"reference": "the reference's own ellipticPartialAxHex3D + ogs gather kernels as JIT-compiled by its toolchain",
                 "port": "oracle C port of the reference operator"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--degree", type=int, default=7)
    ap.add_argument("--elements", type=int, default=64, help="global box is elements^3")
    ap.add_argument("--cpu-elements", type=int, default=24, help="box edge of the bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--pcg-iters", type=int, default=40)
    ap.add_argument("--mode", type=int, default=1, help="1 fused gather epilogue, 0 reference data flow")
    ap.add_argument("--no-pcg", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-trilinear", action="store_true", help="skip the ELEMENT MAP = TRILINEAR operator line")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: BP5 operator + Jacobi-PCG on the 64^3 box (BASELINE configs[1], [2]); "
                         "c4: MULTIGRID-PCG, Hex N=7, 96^3 box (configs[3], meant for --gpus 8)")
    ap.add_argument("--solve-tol", type=float, default=1e-8)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_ax_sample(N, n, steps, warmup, seconds=None):
    """The same operator apply on an n^3 box on the host cores.  kind "reference": the reference's own kernels
    (ellipticPartialAxHex3D + ogs gather, JIT-compiled by the reference's toolchain in OpenMP mode and harvested into
    oracle/_ref/kernels by oracle/refbuild/build_ref_kernels.sh), driven in the order of elliptic_t::Operator; kind
    "port": the oracle C port (OpenMP) when those binaries are absent.  With `seconds` the number of applies is chosen
    so that the timed loop lasts about that long.  Returns (GDOF/s, s per apply, threads, DOFs, applies, kind)."""
    from oracle import elliptic_ref as er
    from oracle import ref_kernels as rk
    from oracle.mesh_box import build_box_hex_mesh, masked_global_ids
    from oracle.ogs_ref import SIGNED, ogs_setup_all
    m = build_box_hex_mesh(N, n, n, n)
    _, ids = masked_global_ids(m)
    o = ogs_setup_all([ids], SIGNED, True)[0]
    G2L = o.global_to_local()
    q = er.splitmix_uniform(1234, o.Ngather)
    rs, ci = o.gatherLocal.rowStartsT, o.gatherLocal.colIdsT
    if rk.available(N):
        kind, threads = "reference", rk.max_threads()
        op = rk.RefOperator(N + 1, G2L, m.wJ, m.ggeo, m.D, 0.0, rs, ci)
        out = np.empty(o.Ngather)
        apply = lambda: op(q, out)
        port = er.operator(N + 1, G2L, m.wJ, m.ggeo, m.D, 0.0, rs, ci, q)
        assert np.abs(apply() - port).max() <= 1e-12 * np.abs(port).max(), "reference kernels disagree with the oracle"
    else:
        kind, threads = "port", er.num_threads()
        apply = lambda: er.operator(N + 1, G2L, m.wJ, m.ggeo, m.D, 0.0, rs, ci, q)
    for _ in range(warmup):
        apply()
    if seconds is not None:
        t0 = time.perf_counter()
        apply()
        one = max(time.perf_counter() - t0, 1e-6)
        steps = int(min(max(seconds / one, 3), 5000))
    t0 = time.perf_counter()
    for _ in range(steps):
        apply()
    dt = (time.perf_counter() - t0) / steps
    return o.Ngather / dt / 1e9, dt, threads, o.Ngather, steps, kind


def bench_config(N, n, mode=1):
    """identical on both arms (the driver compares them)"""
    return {"workload": f"bp5_operator_hex_n{N}_e{n}", "N": N, "elements": [n, n, n], "lambda": 0.0,
            "global_dofs": int((n * N - 1) ** 3), "mode": "fused-gather" if mode == 1 else "reference-flow",
            "l2": "inputs (6.4 GB of geometric factors per apply) are far larger than the 126 MB L2; no flush needed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, n = args.degree, args.cpu_elements
    # keep the whole run within a few minutes: ~0.3 s per apply at 16^3 on a few cores
    # bounded: at most ~60 s of applies however large --steps is (each "step" = one apply of the sample box)
    one = cpu_ax_sample(N, n, 1, 1)[1]
    steps = int(max(1, min(args.steps, 60.0 / max(one, 1e-6))))
    gd, dt, threads, ng, steps, kind = cpu_ax_sample(N, n, steps, min(max(args.warmup, 1), 3))
    sample = (f"{CPU_KIND_TEXT[kind]} (OpenMP, {threads} threads), elliptic_t::Operator, Hex N={N}, {n}^3 box, lambda=0, "
              f"{ng} DOFs per apply, {steps} timed applies of {dt:.4f} s")
    line = {"impl": "reference", "metric": METRIC, "value": gd, "unit": "GDOF/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(N, args.elements, args.mode),
            "cpu_sample_elements": [n] * 3, "cpu_flags": "g++ -O3 -march=x86-64-v3 -fopenmp (OCCA JIT of the reference's OKL)",
            "cpu_baseline": {"value": gd, "unit": "GDOF/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": gd, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "c4":
        # BASELINE configs[3]: MULTIGRID-PCG, Hex N=7, 96^3 box (run with --gpus 8); one "step" = one PCG iteration
        import importlib.util
        spec = importlib.util.spec_from_file_location("mg_bench", os.path.join(ROOT, "tools", "mg_bench.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        n = args.elements if args.elements != 64 else 96
        mod.main(["--degree", str(args.degree), "--elements", str(n), "--tol", str(args.solve_tol), "--bench-line"])
        return

    import torch
    import torch.distributed as dist
    from libparanumal_b200 import _lib as L
    from libparanumal_b200 import api
    from libparanumal_b200.api import Comm
    from libparanumal_b200.problem import EllipticProblem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    api.init(local_rank)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")
    comm = Comm(rank, world, gloo)
    comm.init_nccl()
    p2p = comm.init_p2p() if (world > 1 and os.environ.get("LIBP_P2P", "1") != "0") else False
    N, n = args.degree, args.elements
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        """device time of nsteps calls, max over ranks (ms)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(nsteps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ------------------------------------------------------------------ BP5 operator (lambda = 0)
    t_setup = time.perf_counter()
    p = EllipticProblem(N, n, lam=0.0, boundary_flag=1, comm=comm, mode=args.mode, coords=not args.no_pcg)
    t_setup = time.perf_counter() - t_setup
    Ng = p.NglobalDofs
    q = p.vec()
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    q[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    Aq = p.vec()
    step = lambda: p.op.Operator(q, Aq)
    for _ in range(warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # cudaProfilerStart/Stop bracket the timed regions so `ncu --profile-from-start off` lists exactly the
    # launches of the timed step (the problem setup runs ~1000 torch kernels first); no effect otherwise
    torch.cuda.profiler.start()
    ms = timed(step, steps)
    torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    value = Ng * steps / (ms * 1e-3) / 1e9

    # size-independent correctness properties at full size (symmetry, fused == reference data flow)
    checks = {}
    y = p.vec()
    y[: p.Ndofs] = torch.rand(p.Ndofs, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    Ay = p.operator(y)
    d1 = torch.dot(y[: p.Ndofs], Aq[: p.Ndofs]).reshape(1)
    d2 = torch.dot(q[: p.Ndofs], Ay[: p.Ndofs]).reshape(1)
    if world > 1:
        dist.all_reduce(d1); dist.all_reduce(d2)
    checks["symmetry_rel"] = abs(float(d1 - d2)) / max(abs(float(d1)), 1e-300)
    del y, Ay

    # ------------------------------------------------------------------ dominant kernel vs HBM roofline
    # Measured live, on the operator handle itself: libp_elliptic_operator_timed records CUDA events on the launching
    # stream around the two parts of one apply (zero-fill of the accumulator | exchange + Ax launches + combine).
    m = p.mesh
    E, Np = m.Nelements, m.Np
    peak, peak_src = measured_peaks()
    ksteps = min(steps, 50)
    parts = [p.op.OperatorTimed(q, Aq) for _ in range(ksteps)]
    zms = sum(x[0] for x in parts) / ksteps
    kms = sum(x[1] for x in parts) / ksteps
    alg_bytes = 8.0 * 6 * E * Np + 16.0 * p.Ndofs  # geofactors + read q + write Aq (lambda = 0), this rank
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    plan_stats = p.op.chain_stats(Aq)
    chain_on = plan_stats["chain"] > 0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ax_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            key = "chain" if chain_on else "previous"
            if N == 7 and n == 64 and world == 1 and key in tj:
                traffic, traffic_src = tj[key]["dram_bytes_per_apply"], tj[key]["source"]
        except Exception:
            traffic = None
    kname = (f"ax_hex3d_chain_kernel<Nq={p.Nq},stages={plan_stats['stages']}> x{plan_stats['chain']} elements per chain"
             if chain_on else f"ax_hex3d_t_kernel<Nq={p.Nq},gather,fused,even-odd>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kname, "kernel_ms": kms,
                "zero_fill_ms": zms, "step_frac": (alg_bytes / (ms / steps * 1e-3) / 1e9) / peak,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "elements_per_launch": E,
                "launches_per_apply": (1 if world == 1 else 2)}

    # ------------------------------------------------------------------ ELEMENT MAP = TRILINEAR
    # the same operator with the geometric factors recomputed from the element vertices instead of streamed
    # (libp_elliptic_set_trilinear); its own byte model: 16 B per DOF + 192 B of vertices per element
    trilinear = None
    if world == 1 and chain_on and args.mode == 1 and not args.no_trilinear:
        ex, ey, ez = m.element_vertices()
        EXYZ = torch.stack([ex, ey, ez], dim=1).contiguous().reshape(-1)
        p.op.set_trilinear(EXYZ, m.gllz, m.gllw)
        At = p.vec(fill=float("nan"))
        for _ in range(3):
            p.op.Operator(q, At)
        tdiff = float((At[: p.Ndofs] - Aq[: p.Ndofs]).abs().max() / Aq[: p.Ndofs].abs().max())
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tsteps = min(steps, 50)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(tsteps):
            p.op.Operator(q, At)
        t1.record()
        torch.cuda.synchronize()
        tms = t0.elapsed_time(t1) / tsteps
        tbytes = 16.0 * p.Ndofs + 192.0 * E
        trilinear = {"value": p.NglobalDofs / (tms * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_apply": tms,
                     "algorithmic_bytes_per_apply": tbytes, "hbm_frac": tbytes / (tms * 1e-3) / 1e9 / peak,
                     "bound": "fp64 pipe (geometry recomputed per node; affine elements take the constant-Jacobian path)",
                     "rel_diff_vs_stored_factors": tdiff}
        p.op.set_trilinear(None)
        del At, EXYZ

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    # Each step copies that step's q from pinned host memory, applies the operator through the C ABI and
    # copies Aq back to pinned host memory.  Two buffer sets alternate so that the D2H of step i overlaps
    # the H2D of step i+1 (PCIe is full duplex); every step still moves its own input and output.
    e2e = None
    if not args.no_e2e:
        hq = [torch.empty(p.Nall, dtype=torch.float64).pin_memory() for _ in range(2)]
        hA = [torch.empty(p.Nall, dtype=torch.float64).pin_memory() for _ in range(2)]
        for h in hq:
            h.copy_(q.cpu())
        dq, dA = [p.vec(), p.vec()], [p.vec(), p.vec()]
        main_s = torch.cuda.current_stream()
        h2d_s, d2h_s = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        state = {"i": 0}
        def e2e_step():
            b = state["i"] & 1
            state["i"] += 1
            h2d_s.wait_event(ev_cmp[b])          # dq[b] free again (compute of step i-2 done)
            with torch.cuda.stream(h2d_s):
                dq[b].copy_(hq[b], non_blocking=True)
                ev_in[b].record(h2d_s)
            main_s.wait_event(ev_in[b])
            main_s.wait_event(ev_out[b])         # dA[b] drained by the D2H of step i-2
            p.op.Operator(dq[b], dA[b])
            ev_cmp[b].record(main_s)
            d2h_s.wait_event(ev_cmp[b])
            with torch.cuda.stream(d2h_s):
                hA[b].copy_(dA[b], non_blocking=True)
                ev_out[b].record(d2h_s)
        def e2e_drain():
            main_s.wait_stream(d2h_s); main_s.wait_stream(h2d_s)
        nst = max(4, min(steps, 10))
        for _ in range(2):
            e2e_step()
        e2e_drain()
        ems = timed(lambda: None, 0)  # barrier + sync
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(nst):
            e2e_step()
        e2e_drain()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems = float(ems.item())
        ok = bool(torch.allclose(hA[0][: p.Ndofs], Aq[: p.Ndofs].cpu(), rtol=1e-9, atol=1e-12))
        e2e = {"value": Ng * nst / (ems * 1e-3) / 1e9, "unit": "GDOF/s",
               "h2d_bytes_per_step": int(8 * p.Nall), "d2h_bytes_per_step": int(8 * p.Nall),
               "ms_per_step": ems / nst, "result_matches_device_run": ok,
               "path": "pinned host q -> H2D -> libp_elliptic_operator -> D2H Aq (double-buffered over steps)"}
        del hq, hA, dq, dA

    # launches inside the timed region per step: single rank = zero-fill kernel + 1 Ax kernel;
    # sharded = zero-fill + 2 Ax (local | boundary elements) + halo pack/send + wait/unpack + combine pack + combine unpack
    launches_per_step = 2 if world == 1 else 7

    # ------------------------------------------------------------------ Jacobi-PCG on the screened problem
    pcg = None
    if not args.no_pcg:
        p.set_lambda(1.0)           # same mesh / ogs / maps, new operator handle (BASELINE configs[2])
        M = p.jacobi()
        r0 = p.rhs_sine3d()
        solver = p.pcg()
        iters = args.pcg_iters
        x, r = p.vec(), r0.clone()
        solver.Solve(p.op, M, x, r, tol=1e-30, maxit=3)  # warm-up
        def solve():
            x.zero_(); r.copy_(r0)
            return solver.Solve(p.op, M, x, r, tol=1e-30, maxit=iters)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.profiler.start()
        e0.record()
        it = solve()
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
        pms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        pms = float(pms.item())
        hist = solver.residual_history()
        it_bytes = (8.0 * 7 * p.mesh.Nelements * p.mesh.Np + 16.0 * p.Ndofs) + 88.0 * p.Ndofs
        # solve-level end to end: gathered right-hand side in pinned HOST memory -> H2D -> libp_pcg_solve to
        # args.solve_tol -> D2H of the solution (what a caller of elliptic_t::Solve sees, vectors resident in between)
        solve_e2e = None
        if not args.no_e2e:
            h_r = r0.cpu().pin_memory()
            h_x = torch.empty(p.Nall, dtype=torch.float64).pin_memory()
            xs, rs_ = p.vec(), p.vec()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            rs_.copy_(h_r, non_blocking=True)
            xs.zero_()
            its = solver.Solve(p.op, M, xs, rs_, tol=args.solve_tol, maxit=5000)
            h_x.copy_(xs, non_blocking=True)
            e1.record()
            barrier()
            sms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(sms, op=dist.ReduceOp.MAX)
            sms = float(sms.item())
            sh = solver.residual_history()
            solve_e2e = {"value": p.NglobalDofs * its / (sms * 1e-3) / 1e9, "unit": "GDOF/s", "iterations": its,
                         "seconds": sms * 1e-3, "tol": args.solve_tol, "h2d_bytes": int(8 * p.Nall),
                         "d2h_bytes": int(8 * p.Nall), "residual_first_last": [float(sh[0]), float(sh[-1])],
                         "path": "pinned host rhs -> H2D -> libp_pcg_solve (Jacobi) -> D2H x"}
            del h_r, h_x, xs, rs_
        pcg = {"config": f"screened Poisson (lambda=1) Jacobi-PCG, Hex N={N}, {n}^3 box", "iterations": it,
               "solve_e2e": solve_e2e,
               "ms_per_iteration": pms / max(it, 1), "value": p.NglobalDofs * it / (pms * 1e-3) / 1e9,
               "unit": "GDOF/s", "residual_first_last": [float(hist[0]), float(hist[-1])],
               "roofline_frac": (it_bytes / (pms / max(it, 1) * 1e-3) / 1e9) / peak,
               "algorithmic_bytes_per_iteration": it_bytes}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            gd, dt, threads, ngc, nap, kind = cpu_ax_sample(N, args.cpu_elements, 3, 1, seconds=args.cpu_seconds)
            cpu = {"value": gd, "unit": "GDOF/s", "cores": threads, "kind": kind,
                   "sample": f"{CPU_KIND_TEXT[kind]} (OpenMP, {threads} threads): the same operator apply on a "
                             f"{args.cpu_elements}^3 box ({ngc} DOFs): {nap} applies of {dt:.4f} s "
                             f"(~{nap * dt:.0f} s of CPU work)"}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "GDOF/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": bench_config(N, n, args.mode),
                "exchange": ("nvlink-peer-window" if p2p else "nccl") if world > 1 else "none",
                "setup_seconds": round(t_setup, 1), "chain_plan": plan_stats,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * steps,
                "roofline": roofline, "cpu_baseline": cpu, "pcg": pcg, "trilinear": trilinear, "checks": checks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
