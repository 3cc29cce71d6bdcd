"""Generates tests/golden/mg_*.npz from the UNMODIFIED reference (oracle/_ref/dump_mg_driver, built by
build_ref.sh): the p-multigrid + parAlmond hierarchy as the reference's apply path sees it, one V-cycle on a
seeded vector, and the MULTIGRID-PCG iteration count / residual history.  Build-container only.

usage: python oracle/refbuild/make_golden_mg.py [name ...]
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
WORK = os.environ.get("LIBP_REF_WORK", "/tmp/libp_ref")
DRIVER = os.path.join(REPO, "oracle", "_ref", "dump_mg_driver")
GOLD = os.path.join(REPO, "tests", "golden")
DT = {"f64": np.float64, "i32": np.int32, "i64": np.int64}

CONFIGS = {
    # all defaults of the reference (Chebyshev degree 2, HALFDOFS, V-cycle) unless stated
    "mg_n3_e3": dict(N=3, n=3, flag=1, lam=1.0),
    "mg_n7_e2": dict(N=7, n=2, flag=1, lam=1.0),
    "mg_n2_e12": dict(N=2, n=12, flag=1, lam=1.0, drop_geo=True),          # one CSR (AMG) level above the exact solve
    "mg_n4_e10": dict(N=4, n=10, flag=1, lam=1.0, drop_geo=True),          # BASELINE configs[0] with MULTIGRID
    "mg_n3_e4_jacobi": dict(N=3, n=4, flag=1, lam=0.3, smoother="DAMPEDJACOBI"),
    # two CSR (AMG) levels above the exact solve: pins the host-side AMG setup (libparanumal_b200/amg_setup.py);
    # only the algebraic hierarchy, the Chebyshev bounds and the solve history are kept
    "amg_n2_e24": dict(N=2, n=24, flag=1, lam=1.0, amg_only=True),
}
AMG_KEEP = re.compile(r"^(L\d+_(A|P|R)_(meta|rowStarts|cols|vals)|L\d+_(lambda|meta)|coarse_A_.*|coarse_meta|level_kinds|"
                      r"iterations|meta)$")
# "drop_geo" (digest) configs: arrays the product-side harness regenerates itself are dropped (geometry,
# inverse diagonals, the dense coarse inverse = inv(coarse_A)), integer maps are kept as sha256 digests and
# long vectors as strided samples
GEO = re.compile(r".*_(ggeo|wJ|invDiagA|weightG)$|^coarse_diagInvAT$|^vc_r$")
MAPS = re.compile(r".*_(GlobalToLocal|maskedGlobalIds)$")
SAMPLED = re.compile(r"^(vc_z|xsol|r|l0_.*)$")
STRIDE = 7


def rc_text(c):
    s = {"FORMAT": "2.0", "DATA FILE": "data/ellipticSine3D.h", "MESH FILE": "BOX", "MESH DIMENSION": 3,
         "ELEMENT TYPE": 12, "BOX NX": c["n"], "BOX NY": c["n"], "BOX NZ": c["n"], "BOX DIMX": 1, "BOX DIMY": 1,
         "BOX DIMZ": 1, "BOX BOUNDARY FLAG": c["flag"], "POLYNOMIAL DEGREE": c["N"], "THREAD MODEL": "Serial",
         "PLATFORM NUMBER": 0, "DEVICE NUMBER": 0, "LAMBDA": c["lam"], "DISCRETIZATION": "CONTINUOUS",
         "LINEAR SOLVER": "PCG", "PRECONDITIONER": "MULTIGRID", "MULTIGRID SMOOTHER": c.get("smoother", "CHEBYSHEV"),
         "PARALMOND SMOOTHER": c.get("smoother", "CHEBYSHEV"),
         "OUTPUT TO FILE": "FALSE", "VERBOSE": "TRUE"}
    return "".join(f"[{k}]\n{v}\n" for k, v in s.items())


def run(name, c):
    with tempfile.TemporaryDirectory() as td:
        rc = os.path.join(td, "setup.rc")
        open(rc, "w").write(rc_text(c))
        out = os.path.join(td, "out")
        os.makedirs(out)
        env = dict(os.environ, LIBP_CACHE_DIR=os.path.join(WORK, ".occa_cache"), OCCA_CXX="g++",
                   OCCA_CXXFLAGS="-O3 -march=native -fopenmp", OMP_NUM_THREADS="1")
        p = subprocess.run([DRIVER, rc, out], cwd=os.path.join(WORK, "solvers", "elliptic"), env=env,
                           capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stdout[-3000:], p.stderr[-3000:])
            raise SystemExit(1)
        hist = [float(m.group(1)) for m in re.finditer(r"CG: it \d+, r norm ([0-9.eE+-]+)", p.stdout)]
        init = re.search(r"PCG: initial res norm ([0-9.eE+-]+)", p.stdout)
        arrays = {}
        for fn in sorted(os.listdir(out)):
            nm, dt, _ = fn.rsplit(".", 2)
            a = np.fromfile(os.path.join(out, fn), dtype=DT[dt])
            if c.get("amg_only"):
                if AMG_KEEP.match(nm):
                    arrays[nm] = a
                continue
            if c.get("drop_geo"):
                if GEO.match(nm):
                    continue
                if MAPS.match(nm):
                    arrays[nm + "_sha256"] = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8)
                    continue
                if SAMPLED.match(nm):
                    arrays[nm + "_sample"] = a[::STRIDE].copy()
                    arrays[nm + "_norm2"] = np.array([np.sqrt(np.sum(a * a))])
                    continue
            arrays[nm] = a
        arrays["res_history"] = np.array(hist)
        arrays["res_initial"] = np.array([float(init.group(1))])
        arrays["config"] = np.array([c["N"], c["n"], c["flag"]], dtype=np.int64)
        arrays["lambda"] = np.array([c["lam"]])
        arrays["sample_stride"] = np.array([STRIDE if c.get("drop_geo") else 1], dtype=np.int64)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
        report = [l for l in p.stdout.splitlines() if re.search(r"Degree|AMG|Exact|Matrix-free|Level|\|", l)]
        print(name, "iterations", arrays["iterations"], "kinds", arrays["level_kinds"], "size",
              os.path.getsize(os.path.join(GOLD, name + ".npz")) // 1024, "KiB")
        print("\n".join(report[-14:]))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CONFIGS)):
        run(n, CONFIGS[n])
