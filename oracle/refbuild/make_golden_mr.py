"""Generates the MULTI-RANK fixtures tests/golden/mr_*.npz from the UNMODIFIED reference run on P ranks
(oracle/refbuild/dump_mr_driver.cpp under mpirun_stub.sh: one process per rank over the multi-process MPI stand-in,
oracle/refbuild/mpistub/).  Runs only in the build container; the fixtures are committed.

Every fixture holds, per rank r (keys "r<r>_<array>"): the box decomposition outputs, the signed ids after
ogsBase_t::Setup (owner choice across ranks), GlobalToLocal, all counters, gatherLocal / gatherHalo maps, the
ogsPairwise_t lists (send ids, ranks, counts, offsets, postmpi operator), D / ggeo / wJ, a seeded q, Operator(q), the
halo-filled input, the gathered right-hand side, the PCG solution and its iteration count.

usage: python oracle/refbuild/make_golden_mr.py [name ...]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
WORK = os.environ.get("LIBP_REF_WORK", "/tmp/libp_ref")
DRIVER = os.path.join(REPO, "oracle", "_ref", "dump_mr_driver")
MPIRUN = os.path.join(HERE, "mpirun_stub.sh")
GOLD = os.path.join(REPO, "tests", "golden")

CONFIGS = {
    # name: degree, global box, boundary flag, lambda, preconditioner, ranks
    "mr_n3_e4x4x4_p2": dict(N=3, box=(4, 4, 4), flag=1, lam=1.0, precon="JACOBI", P=2),
    "mr_n3_e4x4x4_p4": dict(N=3, box=(4, 4, 4), flag=1, lam=1.0, precon="JACOBI", P=4),
    "mr_n3_e4x4x4_p8": dict(N=3, box=(4, 4, 4), flag=1, lam=1.0, precon="JACOBI", P=8),
    # uneven splits (remainders go to the low ranks, libs/core/rankDecomp.cpp)
    "mr_n2_e5x4x3_p2": dict(N=2, box=(5, 4, 3), flag=1, lam=0.5, precon="NONE", P=2),
    "mr_n2_e5x4x3_p4": dict(N=2, box=(5, 4, 3), flag=1, lam=0.5, precon="NONE", P=4),
    "mr_n2_e5x4x3_p8": dict(N=2, box=(5, 4, 3), flag=1, lam=0.5, precon="NONE", P=8),
    # N = 7 (the headline order): one element per rank, every surface node is shared
    "mr_n7_e2x2x2_p8": dict(N=7, box=(2, 2, 2), flag=1, lam=1.0, precon="JACOBI", P=8),
    "mr_n7_e4x2x2_p2": dict(N=7, box=(4, 2, 2), flag=1, lam=0.0, precon="JACOBI", P=2),
    # periodic box across ranks (wrap-around neighbours on the same / another rank), lambda = 1
    "mr_n2_e4x4x4_periodic_p4": dict(N=2, box=(4, 4, 4), flag=-1, lam=1.0, precon="JACOBI", P=4),
}
DT = {"f64": np.float64, "i32": np.int32, "i64": np.int64}


def rc_text(c):
    s = {"FORMAT": "2.0", "DATA FILE": "data/ellipticSine3D.h", "MESH FILE": "BOX", "MESH DIMENSION": 3,
         "ELEMENT TYPE": 12, "BOX GLOBAL NX": c["box"][0], "BOX GLOBAL NY": c["box"][1], "BOX GLOBAL NZ": c["box"][2],
         "BOX DIMX": 1, "BOX DIMY": 1, "BOX DIMZ": 1, "BOX BOUNDARY FLAG": c["flag"], "POLYNOMIAL DEGREE": c["N"],
         "THREAD MODEL": "Serial", "PLATFORM NUMBER": 0, "DEVICE NUMBER": 0, "LAMBDA": c["lam"],
         "DISCRETIZATION": "CONTINUOUS", "LINEAR SOLVER": "PCG", "PRECONDITIONER": c["precon"],
         "OUTPUT TO FILE": "FALSE", "VERBOSE": "TRUE"}
    return "".join(f"[{k}]\n{v}\n" for k, v in s.items())


def run(name, c):
    with tempfile.TemporaryDirectory() as td:
        rc = os.path.join(td, "setup.rc")
        open(rc, "w").write(rc_text(c))
        out = os.path.join(td, "out")
        env = dict(os.environ, LIBP_CACHE_DIR=os.path.join(WORK, ".occa_cache_mr"), OCCA_CXX="g++",
                   OCCA_CXXFLAGS="-O3 -march=native", OMP_NUM_THREADS="1")
        p = subprocess.run(["bash", MPIRUN, str(c["P"]), DRIVER, rc, out], cwd=os.path.join(WORK, "solvers", "elliptic"),
                           env=env, capture_output=True, text=True, timeout=1200)
        if p.returncode != 0:
            sys.stderr.write(p.stdout[-3000:] + p.stderr[-3000:])
            raise SystemExit(f"{name}: reference run failed")
        hist = [float(l.split("r norm")[1].split(",")[0]) for l in p.stdout.splitlines() if l.startswith("CG: it")]
        init = [float(l.split()[-1]) for l in p.stdout.splitlines() if "initial res norm" in l]
        d = {"config_N": c["N"], "config_box": np.array(c["box"]), "config_flag": c["flag"], "config_lambda": c["lam"],
             "config_precon": c["precon"], "config_P": c["P"], "pcg_history": np.array(init + hist)}
        for r in range(c["P"]):
            rd = os.path.join(out, f"r{r}")
            for fn in sorted(os.listdir(rd)):
                key, dt, _ = fn.rsplit(".", 2)
                d[f"r{r}_{key}"] = np.fromfile(os.path.join(rd, fn), dtype=DT[dt])
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)
        its = [int(d[f"r{r}_iterations"][0]) for r in range(c["P"])]
        print(name, "ranks", c["P"], "iterations", its, "Ngather", [int(d[f"r{r}_ogs_counts"][1]) for r in range(c["P"])],
              "size", os.path.getsize(os.path.join(GOLD, name + ".npz")))


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        run(n, CONFIGS[n])
