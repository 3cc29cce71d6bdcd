"""Generates tests/golden/ig_tridiag_n400.npz: the initial guesses the UNMODIFIED reference's InitialGuess classes
(libs/linearSolver/initialGuess.cpp: ZERO, CLASSIC, QR, EXTRAP with MINNORM / CPQR coefficients) form over a sequence
of 14 solves of a fixed SPD tridiagonal system (oracle/refbuild/dump_ig_driver.cpp, built by build_ref.sh).
Runs only in the build container."""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
WORK = os.environ.get("LIBP_REF_WORK", "/tmp/libp_ref")

with tempfile.TemporaryDirectory() as td:
    env = dict(os.environ, LIBP_CACHE_DIR=os.path.join(WORK, ".occa_cache"), OCCA_CXX="g++", OCCA_CXXFLAGS="-O3 -march=native",
               OMP_NUM_THREADS="1")
    subprocess.run([os.path.join(REPO, "oracle", "_ref", "dump_ig_driver"), td], cwd=WORK, env=env, check=True,
                   capture_output=True)
    out = {f.split(".")[0]: np.fromfile(os.path.join(td, f)) for f in sorted(os.listdir(td))}
    out["N"], out["K"] = 400, 14
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "ig_tridiag_n400.npz"), **out)
    s = out["sol"].reshape(14, 400)
    for k, v in out.items():
        if k.startswith("guess"):
            g = v.reshape(14, 400)
            print(k, ["%.1e" % (np.abs(g[i] - s[i]).max() / np.abs(s[i]).max()) for i in range(14)])
