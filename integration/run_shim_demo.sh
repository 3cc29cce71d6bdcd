#!/bin/bash
# GPU box: runs integration/elliptic_b200_main (the unmodified reference in THREAD MODEL = CUDA with the hot path bound
# to libparanumal_b200.so through the shim) for the Hex3D cases given as "N:NX" pairs, e.g.
#   bash integration/run_shim_demo.sh 4:10 7:32
set -u
REPO="$(cd "$(dirname "$0")/.." && pwd)"
T="$REPO/baseline/_ref/libp_ref_cuda"
[ -x "$T/elliptic_b200_main" ] || { echo "run integration/build_shim_demo.sh in the build container first"; exit 1; }
ln -sfn "$T" /tmp/libp_ref_cuda   # LIBP_DIR / OCCA_BUILD_DIR baked into the binaries
export LIBP_SHIM_TRACE=1 LIBP_CACHE_DIR=/tmp/occa_cache_shim OCCA_CACHE_DIR=/tmp/occa_cache_shim OCCA_CXX=g++
for case in "$@"; do
  N=${case%%:*}; NX=${case##*:}
  for PC in ${PRECONS:-JACOBI}; do
    RC=/tmp/shim_${N}_${NX}_${PC}.rc
    { for kv in "FORMAT=2.0" "DATA FILE=data/ellipticSine3D.h" "MESH FILE=BOX" "MESH DIMENSION=3" "ELEMENT TYPE=12" \
        "BOX NX=$NX" "BOX NY=$NX" "BOX NZ=$NX" "BOX DIMX=1" "BOX DIMY=1" "BOX DIMZ=1" "BOX BOUNDARY FLAG=1" \
        "POLYNOMIAL DEGREE=$N" "THREAD MODEL=CUDA" "PLATFORM NUMBER=0" "DEVICE NUMBER=0" "LAMBDA=1.0" \
        "DISCRETIZATION=CONTINUOUS" "LINEAR SOLVER=PCG" "PRECONDITIONER=$PC" "OUTPUT TO FILE=FALSE" "VERBOSE=FALSE"; do
        echo "[${kv%%=*}]"; echo "${kv#*=}"; done; } > $RC
    echo "=== Hex3D N=$N ${NX}^3 PRECONDITIONER=$PC (THREAD MODEL = CUDA)"
    (cd "$T/solvers/elliptic" && "$T/elliptic_b200_main" $RC ${APPLIES:-50} > /tmp/shim_run.log 2>&1; grep -E "Operator vs|OPERATOR TIMING|path|solutions|shim\]" /tmp/shim_run.log; grep -q "solutions" /tmp/shim_run.log || tail -25 /tmp/shim_run.log )
  done
done
