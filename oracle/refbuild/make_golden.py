"""Generates tests/golden/*.npz from the UNMODIFIED reference built by build_ref.sh.

Runs only in the build container (needs /root/reference and oracle/_ref/dump_driver).
The fixtures it writes are committed; nothing at test/bench run time reads the reference.

  full fixtures   every array the dump driver writes (small meshes)
  digest fixtures sha256 of the integer maps + strided samples of the floating arrays
                  (BASELINE config 1: N=4, 10^3 box; too large to commit in full)

usage: python oracle/refbuild/make_golden.py [name ...]
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
DRIVER = os.path.join(REPO, "oracle", "_ref", "dump_driver")
GOLD = os.path.join(REPO, "tests", "golden")

CONFIGS = {
    # name: (N, nx, boundary flag, lambda, precon, mode)
    "hex_n3_e3_jacobi": dict(N=3, n=3, flag=1, lam=1.0, precon="JACOBI", mode="full"),
    "hex_n7_e2_jacobi": dict(N=7, n=2, flag=1, lam=1.0, precon="JACOBI", mode="full"),
    "hex_n2_e4_periodic": dict(N=2, n=4, flag=-1, lam=1.0, precon="JACOBI", mode="full"),
    "hex_n1_e5_none": dict(N=1, n=5, flag=1, lam=1.0, precon="NONE", mode="full"),
    # the remaining orders (their kernels use other element packings / shared-memory strides than N=7)
    "hex_n5_e2_jacobi": dict(N=5, n=2, flag=1, lam=1.0, precon="JACOBI", mode="full"),
    "hex_n6_e2_jacobi": dict(N=6, n=2, flag=1, lam=0.7, precon="JACOBI", mode="full"),
    "hex_n8_e2_none": dict(N=8, n=2, flag=1, lam=1.0, precon="NONE", mode="full"),
    # edge cases: a periodic box of ONE element per direction (every face node collides with its opposite copy inside
    # the same element) and of two elements (each node pair is shared twice)
    "hex_n3_e1_periodic": dict(N=3, n=1, flag=-1, lam=1.0, precon="JACOBI", mode="full"),
    "hex_n2_e2_periodic": dict(N=2, n=2, flag=-1, lam=0.5, precon="NONE", mode="full"),
    "hex_n7_e3_bp5": dict(N=7, n=3, flag=1, lam=0.0, precon="JACOBI", mode="digest"),
    "hex_n4_e10_jacobi": dict(N=4, n=10, flag=1, lam=1.0, precon="JACOBI", mode="digest"),
    "hex_n4_e10_none": dict(N=4, n=10, flag=1, lam=1.0, precon="NONE", mode="digest"),
}

DT = {"f64": np.float64, "i32": np.int32, "i64": np.int64}
INT_MAPS = ["globalIds", "maskedGlobalIds", "GlobalToLocal", "mapB", "meshMapB",
            "gatherLocal_rowStartsN", "gatherLocal_rowStartsT", "gatherLocal_colIdsN", "gatherLocal_colIdsT",
            "gatherHalo_rowStartsN", "gatherHalo_rowStartsT", "gatherHalo_colIdsN", "gatherHalo_colIdsT"]
SAMPLED = ["Aq", "diagA", "r", "xsol", "weightG", "ggeo", "wJ", "rL"]
STRIDE = 97


def rc_text(c):
    s = {"FORMAT": "2.0", "DATA FILE": "data/ellipticSine3D.h", "MESH FILE": "BOX", "MESH DIMENSION": 3,
         "ELEMENT TYPE": 12, "BOX NX": c["n"], "BOX NY": c["n"], "BOX NZ": c["n"], "BOX DIMX": 1, "BOX DIMY": 1,
         "BOX DIMZ": 1, "BOX BOUNDARY FLAG": c["flag"], "POLYNOMIAL DEGREE": c["N"], "THREAD MODEL": "Serial",
         "PLATFORM NUMBER": 0, "DEVICE NUMBER": 0, "LAMBDA": c["lam"], "DISCRETIZATION": "CONTINUOUS",
         "LINEAR SOLVER": "PCG", "PRECONDITIONER": c["precon"], "OUTPUT TO FILE": "FALSE", "VERBOSE": "TRUE"}
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
                           capture_output=True, text=True, check=True)
        hist = [float(m.group(1)) for m in re.finditer(r"CG: it \d+, r norm ([0-9.eE+-]+)", p.stdout)]
        init = re.search(r"PCG: initial res norm ([0-9.eE+-]+)", p.stdout)
        arrays = {}
        for fn in sorted(os.listdir(out)):
            nm, dt, _ = fn.rsplit(".", 2)
            arrays[nm] = np.fromfile(os.path.join(out, fn), dtype=DT[dt])
        arrays["res_history"] = np.array(hist)
        arrays["res_initial"] = np.array([float(init.group(1))])
        arrays["config"] = np.array([c["N"], c["n"], c["flag"]], dtype=np.int64)
        arrays["lambda"] = np.array([c["lam"]])
        if c["mode"] == "digest":
            d = {}
            for k, v in arrays.items():
                if k in INT_MAPS:
                    d[k + "_sha256"] = np.frombuffer(hashlib.sha256(v.tobytes()).digest(), dtype=np.uint8)
                    d[k + "_len"] = np.array([v.size], dtype=np.int64)
                elif k in SAMPLED:
                    d[k + "_sample"] = v[::STRIDE].copy()
                    d[k + "_norm2"] = np.array([np.sqrt(np.sum(v * v))])
                elif v.size <= 4096 and k not in ("q", "x", "y", "z"):
                    d[k] = v
            d["sample_stride"] = np.array([STRIDE], dtype=np.int64)
            arrays = d
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
        print(name, "iterations", arrays["iterations"], "size",
              os.path.getsize(os.path.join(GOLD, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        run(n, CONFIGS[n])
