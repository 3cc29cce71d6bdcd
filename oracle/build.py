"""Builds the oracle's C restatement (oracle/csrc/oracle.c) into oracle/_build/liboracle.so.

ORACLE = test infrastructure.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may load this library.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False, march="x86-64-v3"):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    # -march=x86-64-v3 (AVX2/FMA) rather than native: the .so travels to the GPU box whose host
    # CPU may differ from the build container's.
    cmd = ["gcc", "-O3", f"-march={march}", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
