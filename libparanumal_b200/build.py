"""Builds libparanumal_b200/lib/libparanumal_b200.so (hand-written sm_100a CUDA + C++ host code)
with nvcc, in-tree, so the shared object travels to the GPU box with the repository snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development builds can live beside the product library: LIBP_B200_VARIANT=pf2 LIBP_NVCC_EXTRA=-DLIBP_AX_PF=2
# builds lib/libparanumal_b200_pf2.so, which LIBP_B200_VARIANT=pf2 also makes _lib.load() pick up
_VAR = os.environ.get("LIBP_B200_VARIANT", "")
OBJ = os.path.join(HERE, "lib", "obj" + ("_" + _VAR if _VAR else ""))
LIB = os.path.join(HERE, "lib", "libparanumal_b200" + ("_" + _VAR if _VAR else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp",
          "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]
# development builds: LIBP_AX_TUNE_GRID=1 compiles the (prefetch, L2 hint, occupancy) grid of the Ax kernel
EXTRA = (["-DLIBP_AX_TUNE_GRID"] if os.environ.get("LIBP_AX_TUNE_GRID") == "1" else []) + \
    os.environ.get("LIBP_NVCC_EXTRA", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "libp_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr = _newest_header()
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            cmd = [NVCC] + ARCH + COMMON + EXTRA + (["-Xptxas", "-v"] if ptxas_info else []) + ["-x", "cu", "-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), p.stdout, p.stderr))
        return p.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose or ptxas_info:
                    sys.stderr.write(out)
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-ldl", "-cudart", "static"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
