"""Generates the IPDG fixtures tests/golden/ipdg_*.npz from the UNMODIFIED reference (DISCRETIZATION = IPDG on
hexahedra; oracle/refbuild/dump_ipdg_driver.cpp, one process per rank under mpirun_stub.sh over the multi-process MPI
stand-in).  Runs only in the build container; the fixtures are committed.

Every fixture holds, per rank r (keys "r<r>_<array>"): D, gllw, vgeo, sgeo, vmapM, vmapP, mapP, EToE/EToF/EToP, EToB
(mesh flag and translated type), element halo lists, the trace-halo ids of mesh_t::HaloTraceSetup, tau, lambda, a
seeded q, grad(q) after the trace exchange, Operator(q), the diagonal, the right-hand side, the PCG solution,
iteration count and residual history.

usage: python oracle/refbuild/make_golden_ipdg.py [name ...]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
WORK = os.environ.get("LIBP_REF_WORK", "/tmp/libp_ref")
DRIVER = os.path.join(REPO, "oracle", "_ref", "dump_ipdg_driver")
MPIRUN = os.path.join(HERE, "mpirun_stub.sh")
GOLD = os.path.join(REPO, "tests", "golden")

CONFIGS = {
    "ipdg_n2_e3_jacobi": dict(N=2, box=(3, 3, 3), flag=1, lam=1.0, precon="JACOBI", P=1),
    "ipdg_n4_e3_none": dict(N=4, box=(3, 3, 3), flag=1, lam=1.0, precon="NONE", P=1),
    "ipdg_n7_e2_jacobi": dict(N=7, box=(2, 2, 2), flag=1, lam=0.0, precon="JACOBI", P=1),
    "ipdg_n3_e3_periodic": dict(N=3, box=(3, 3, 3), flag=-1, lam=1.0, precon="JACOBI", P=1),
    "ipdg_n1_e4_none": dict(N=1, box=(4, 3, 2), flag=1, lam=0.5, precon="NONE", P=1),
    # multi-rank: trace halo of the gradient across ranks
    "ipdg_n3_e4x4x4_p2": dict(N=3, box=(4, 4, 4), flag=1, lam=1.0, precon="JACOBI", P=2),
    "ipdg_n2_e5x4x3_p4": dict(N=2, box=(5, 4, 3), flag=1, lam=0.5, precon="JACOBI", P=4),
    "ipdg_n2_e4x4x4_periodic_p4": dict(N=2, box=(4, 4, 4), flag=-1, lam=1.0, precon="JACOBI", P=4),
}
DT = {"f64": np.float64, "i32": np.int32, "i64": np.int64}


def rc_text(c):
    s = {"FORMAT": "2.0", "DATA FILE": "data/ellipticSine3D.h", "MESH FILE": "BOX", "MESH DIMENSION": 3,
         "ELEMENT TYPE": 12, "BOX GLOBAL NX": c["box"][0], "BOX GLOBAL NY": c["box"][1], "BOX GLOBAL NZ": c["box"][2],
         "BOX DIMX": 1, "BOX DIMY": 1, "BOX DIMZ": 1, "BOX BOUNDARY FLAG": c["flag"], "POLYNOMIAL DEGREE": c["N"],
         "THREAD MODEL": "Serial", "PLATFORM NUMBER": 0, "DEVICE NUMBER": 0, "LAMBDA": c["lam"],
         "DISCRETIZATION": "IPDG", "LINEAR SOLVER": "PCG", "PRECONDITIONER": c["precon"],
         "OUTPUT TO FILE": "FALSE", "VERBOSE": "TRUE"}
    return "".join(f"[{k}]\n{v}\n" for k, v in s.items())


def run(name, c, keep=True):
    with tempfile.TemporaryDirectory() as td:
        rc = os.path.join(td, "setup.rc")
        open(rc, "w").write(rc_text(c))
        out = os.path.join(td, "out")
        os.makedirs(out)
        env = dict(os.environ, LIBP_CACHE_DIR=os.path.join(WORK, ".occa_cache_ipdg"), OCCA_CXX="g++",
                   OCCA_CXXFLAGS="-O3 -march=native", OMP_NUM_THREADS="1")
        p = subprocess.run(["bash", MPIRUN, str(c["P"]), DRIVER, rc, out], cwd=os.path.join(WORK, "solvers", "elliptic"),
                           env=env, capture_output=True, text=True, timeout=2400)
        if p.returncode != 0:
            sys.stderr.write(p.stdout[-3000:] + p.stderr[-3000:])
            raise SystemExit(f"{name}: reference run failed")
        hist = [float(l.split("r norm")[1].split(",")[0]) for l in p.stdout.splitlines() if l.startswith("CG: it")]
        init = [float(l.split()[-1]) for l in p.stdout.splitlines() if "initial res norm" in l]
        norm = [l for l in p.stdout.splitlines() if l.startswith("Solution norm")]
        its = [l for l in p.stdout.splitlines() if l.startswith("ITERATIONS")]
        print(name, its, norm)
        if not keep:
            return
        d = {"config_N": c["N"], "config_box": np.array(c["box"]), "config_flag": c["flag"], "config_lambda": c["lam"],
             "config_precon": c["precon"], "config_P": c["P"], "pcg_history": np.array(init + hist)}
        for r in range(c["P"]):
            rd = os.path.join(out, f"r{r}")
            for fn in sorted(os.listdir(rd)):
                key, dt, _ = fn.rsplit(".", 2)
                if key in ("x", "y", "z") and c["N"] > 4:
                    continue
                d[f"r{r}_{key}"] = np.fromfile(os.path.join(rd, fn), dtype=DT[dt])
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)
        print("   size", os.path.getsize(os.path.join(GOLD, name + ".npz")))


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        if n == "suite":  # the reference's own regression cases (test/testElliptic.py:271-275, 355-358): numbers only
            run("testEllipticHex_Ipdg", dict(N=4, box=(10, 10, 10), flag=1, lam=1.0, precon="NONE", P=1), keep=False)
            run("testEllipticHex_Ipdg_Jacobi", dict(N=4, box=(10, 10, 10), flag=1, lam=1.0, precon="JACOBI", P=1), keep=False)
            continue
        run(n, CONFIGS[n])
