#!/bin/bash
# Harvests the reference's own JIT-compiled CPU kernels for the hot path so that they can be timed on the GPU box's
# host cores without OCCA, the OKL sources or /root/reference (cpu_baseline.kind = "reference" in bench.py).
#
# How: the unmodified reference (oracle/_ref/ellipticMain, built by build_ref.sh) is run once per degree in
# THREAD MODEL = OpenMP with a fresh kernel cache; OCCA translates the reference's .okl files and compiles them with
#   g++ -O3 -march=x86-64-v3 -fopenmp        (x86-64-v3 rather than native: the binaries travel to another host)
# into self-contained shared objects (they depend on libgomp / libstdc++ only) with plain C entry points, e.g.
#   ellipticPartialAxHex3D(const int& Nelements, const int* elementList, const int* GlobalToLocal, const double* wJ,
#                          const double* ggeo, const double* DT, const double* S, const double* MM,
#                          const double& lambda, const double* q, double* Aq)
# Only those BINARIES are copied, into oracle/_ref/kernels/ (git-ignored, not gpurun-ignored).  No reference source
# enters the repository.  oracle/ref_kernels.py loads them with ctypes.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
W=${LIBP_REF_WORK:-/tmp/libp_ref}
OUT="$REPO/oracle/_ref/kernels"
MAIN="$REPO/oracle/_ref/ellipticMain"
[ -x "$MAIN" ] || { echo "run oracle/refbuild/build_ref.sh first"; exit 1; }
mkdir -p "$OUT"
for N in ${LIBP_REF_KERNEL_DEGREES-1 2 3 4 5 6 7 8}; do
  C=$(mktemp -d /tmp/occa_refk.XXXXXX)
  RC="$C/setup.rc"
  {
    for kv in "FORMAT=2.0" "DATA FILE=data/ellipticSine3D.h" "MESH FILE=BOX" "MESH DIMENSION=3" "ELEMENT TYPE=12" \
              "BOX NX=2" "BOX NY=2" "BOX NZ=2" "BOX DIMX=1" "BOX DIMY=1" "BOX DIMZ=1" "BOX BOUNDARY FLAG=1" \
              "POLYNOMIAL DEGREE=$N" "THREAD MODEL=OpenMP" "PLATFORM NUMBER=0" "DEVICE NUMBER=0" "LAMBDA=1.0" \
              "DISCRETIZATION=CONTINUOUS" "LINEAR SOLVER=PCG" "PRECONDITIONER=JACOBI" "OUTPUT TO FILE=FALSE" "VERBOSE=FALSE"; do
      echo "[${kv%%=*}]"; echo "${kv#*=}"
    done
  } > "$RC"
  (cd "$W/solvers/elliptic" && LIBP_CACHE_DIR="$C" OCCA_CXX=g++ OCCA_CXXFLAGS="-O3 -march=x86-64-v3 -fopenmp" \
     OMP_NUM_THREADS=2 "$MAIN" "$RC" > "$C/run.log" 2>&1)
  grep -q "Solution norm" "$C/run.log" || { cat "$C/run.log"; exit 1; }
  AX=$(grep -l 'extern "C" void ellipticPartialAxHex3D' "$C"/cache/*/source.cpp | head -1)
  cp "$(dirname "$AX")/binary" "$OUT/ellipticAxHex3D_N$N.so"
  if [ ! -f "$OUT/ogsKernels_double_add.so" ]; then
    GS=$(grep -l 'extern "C" void gatherScatter' "$C"/cache/*/source.cpp | head -1)
    cp "$(dirname "$GS")/binary" "$OUT/ogsKernels_double_add.so"
    UP=$(grep -l 'extern "C" void updatePCG' "$C"/cache/*/source.cpp | head -1)
    cp "$(dirname "$UP")/binary" "$OUT/linearSolverUpdatePCG.so"
  fi
  rm -rf "$C"
done
# p-multigrid transfer kernels (ellipticPreconCoarsenHex3D.okl / ellipticPreconProlongateHex3D.okl), one binary per
# (fine, coarse) pair of the HALFDOFS ladders of N = 7 and N = 8: (8,6) (6,4) (4,3) (3,2) and (9,7) (7,5) (5,3)
for N in ${LIBP_REF_KERNEL_MG_DEGREES-7 8}; do
  C=$(mktemp -d /tmp/occa_refk.XXXXXX)
  RC="$C/setup.rc"
  {
    for kv in "FORMAT=2.0" "DATA FILE=data/ellipticSine3D.h" "MESH FILE=BOX" "MESH DIMENSION=3" "ELEMENT TYPE=12" \
              "BOX NX=2" "BOX NY=2" "BOX NZ=2" "BOX DIMX=1" "BOX DIMY=1" "BOX DIMZ=1" "BOX BOUNDARY FLAG=1" \
              "POLYNOMIAL DEGREE=$N" "THREAD MODEL=OpenMP" "PLATFORM NUMBER=0" "DEVICE NUMBER=0" "LAMBDA=1.0" \
              "DISCRETIZATION=CONTINUOUS" "LINEAR SOLVER=PCG" "PRECONDITIONER=MULTIGRID" "MULTIGRID SMOOTHER=CHEBYSHEV" \
              "OUTPUT TO FILE=FALSE" "VERBOSE=FALSE"; do
      echo "[${kv%%=*}]"; echo "${kv#*=}"
    done
  } > "$RC"
  (cd "$W/solvers/elliptic" && LIBP_CACHE_DIR="$C" OCCA_CXX=g++ OCCA_CXXFLAGS="-O3 -march=x86-64-v3 -fopenmp" \
     OMP_NUM_THREADS=2 "$MAIN" "$RC" > "$C/run.log" 2>&1)
  grep -q "Solution norm" "$C/run.log" || { cat "$C/run.log"; exit 1; }
  for kind in Coarsen Prolongate; do
    for src in $(grep -l "extern \"C\" void ellipticPartialPrecon${kind}Hex3D" "$C"/cache/*/source.cpp); do
      d=$(dirname "$src")
      F=$(grep -m1 "define  *p_NqFine " "$d/raw_source.cpp" | awk '{print $3}')
      Cq=$(grep -m1 "define  *p_NqCoarse " "$d/raw_source.cpp" | awk '{print $3}')
      cp "$d/binary" "$OUT/ellipticPrecon${kind}Hex3D_F${F}_C${Cq}.so"
    done
  done
  rm -rf "$C"
done
ls -la "$OUT"
