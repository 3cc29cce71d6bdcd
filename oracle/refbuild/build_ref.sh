#!/bin/bash
# Builds the UNMODIFIED reference (libParanumal + OCCA, Serial/OpenMP modes only) in a
# scratch copy under /tmp and the dump driver oracle/refbuild/dump_driver.cpp against it.
# Only binaries land in oracle/_ref/ (git-ignored). Used in THIS container only, to
# generate tests/golden/* (see oracle/refbuild/make_golden.py). Never used by the product.
#
# Why a scratch copy: /root/reference is read-only and both makefiles write objects into
# the source tree; LIBP_DIR (the OKL kernel search path) is baked in at compile time.
# Missing in the image: MPI (-> oracle/refbuild/mpistub/mpi.h), system LAPACK (-> scipy's
# bundled OpenBLAS whose symbols carry a scipy_ prefix -> lapack_rename.h).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
SRC=${LIBP_REFERENCE:-/root/reference}
W=${LIBP_REF_WORK:-/tmp/libp_ref}
OUT="$REPO/oracle/_ref"
J=${J:-8}
mkdir -p "$W" "$OUT"
if [ ! -d "$W/libs" ]; then
  cp -r "$SRC"/{include,libs,solvers,make.top,makefile} "$W"/
  cp -r "$SRC"/occa "$W"/occa
  chmod -R u+w "$W"
fi
# LAPACK symbol renames
BL=$(python3 -c "import scipy,os;print(os.path.realpath(os.path.join(os.path.dirname(scipy.__file__),'..','scipy.libs')))")
BLSO=$(basename "$BL"/libscipy_openblas-*.so)
{
  for p in d s; do for f in gecon geev gels geqp3 gesv getrf getri lange ormqr trsm; do
    echo "#define ${p}${f}_ scipy_${p}${f}_"; done; done
  echo "#define dsyev_ scipy_dsyev_"
} > "$W/lapack_rename.h"
# 1. OCCA
if [ ! -f "$W/occa/lib/libocca.so" ]; then
  make -C "$W/occa" -j"$J" CXX=g++ CC=gcc CXXFLAGS="-O2 -include cstdint" \
     OCCA_CUDA_ENABLED=0 OCCA_OPENCL_ENABLED=0 OCCA_HIP_ENABLED=0 OCCA_DPCPP_ENABLED=0 OCCA_METAL_ENABLED=0 \
     > "$W/occa_build.log" 2>&1
fi
# 1b. the MPI stand-in (single rank: memcpy collectives; MPISTUB_SIZE > 1: one process per rank, see mpistub.c)
gcc -O2 -fPIC -c "$HERE/mpistub/mpistub.c" -I"$HERE/mpistub" -o "$W/mpistub.o"
ar rcs "$W/libmpistub.a" "$W/mpistub.o"
# 2. libParanumal libs + elliptic (std::sort semantics: no -DGLIBCXX_PARALLEL)
INC="-I$HERE/mpistub -include $W/lapack_rename.h -I$W/include -I$W/occa/include"
make -C "$W" -j"$J" elliptic LIBP_CC=gcc LIBP_CXX=g++ LIBP_LD=g++ \
  LIBP_INCLUDES="$INC" \
  LIBP_CXXFLAGS="-fopenmp -O3 -Wall -Wno-unused-function -std=c++17 -mavx2 -march=native" \
  LIBP_CFLAGS="-fopenmp -O3 -Wall -Wno-unused-function -mavx2 -march=native" \
  LIBP_BLAS_DIR="$BL" LIBP_BLAS_LIB="-L$BL -l:$BLSO $W/libmpistub.a" > "$W/libp_build.log" 2>&1
make -C "$W/solvers/elliptic" lib LIBP_DIR="$W" LIBP_CC=gcc LIBP_CXX=g++ LIBP_LD=g++ \
  LIBP_INCLUDES="$INC" \
  LIBP_CXXFLAGS="-fopenmp -O3 -Wall -Wno-unused-function -std=c++17 -mavx2 -march=native" \
  LIBP_BLAS_DIR="$BL" LIBP_BLAS_LIB="-L$BL -l:$BLSO" >> "$W/libp_build.log" 2>&1 || true
cp "$W/solvers/elliptic/ellipticMain" "$OUT/ellipticMain"
echo "reference built: $OUT/ellipticMain (work tree $W)"
# 3. dump driver (our own code, oracle/refbuild/dump_driver.cpp) against libelliptic.a
g++ -fopenmp -O2 -std=c++17 -march=native -Wno-unused-function \
  -DLIBP_DIR="\"$W\"" \
  -I"$HERE/mpistub" -include "$W/lapack_rename.h" -I"$W/include" -I"$W/occa/include" -I"$W/solvers/elliptic" \
  -o "$OUT/dump_driver" "$HERE/dump_driver.cpp" \
  "$W/solvers/elliptic/libelliptic.a" \
  -L"$W/libs" -lparAlmond -llinearSolver -lmesh -lparAdogs -logs -llinAlg -lcore \
  "$W/libmpistub.a" -Wl,-rpath,"$BL" -L"$BL" -l:"$BLSO" -Wl,-rpath,"$W/occa/lib" -L"$W/occa/lib" -locca
echo "dump driver built: $OUT/dump_driver"
# 4. multigrid dump driver (our own code, oracle/refbuild/dump_mg_driver.cpp)
g++ -fopenmp -O2 -std=c++17 -march=native -Wno-unused-function \
  -DLIBP_DIR="\"$W\"" \
  -I"$HERE/mpistub" -include "$W/lapack_rename.h" -I"$W/include" -I"$W/occa/include" -I"$W/solvers/elliptic" \
  -o "$OUT/dump_mg_driver" "$HERE/dump_mg_driver.cpp" \
  "$W/solvers/elliptic/libelliptic.a" \
  -L"$W/libs" -lparAlmond -llinearSolver -lmesh -lparAdogs -logs -llinAlg -lcore \
  "$W/libmpistub.a" -Wl,-rpath,"$BL" -L"$BL" -l:"$BLSO" -Wl,-rpath,"$W/occa/lib" -L"$W/occa/lib" -locca
echo "mg dump driver built: $OUT/dump_mg_driver"
# 5. multi-rank dump driver (our own code, oracle/refbuild/dump_mr_driver.cpp; run under mpirun_stub.sh)
g++ -fopenmp -O2 -std=c++17 -march=native -Wno-unused-function \
  -DLIBP_DIR="\"$W\"" \
  -I"$HERE/mpistub" -include "$W/lapack_rename.h" -I"$W/include" -I"$W/occa/include" -I"$W/solvers/elliptic" \
  -o "$OUT/dump_mr_driver" "$HERE/dump_mr_driver.cpp" \
  "$W/solvers/elliptic/libelliptic.a" \
  -L"$W/libs" -lparAlmond -llinearSolver -lmesh -lparAdogs -logs -llinAlg -lcore \
  "$W/libmpistub.a" -Wl,-rpath,"$BL" -L"$BL" -l:"$BLSO" -Wl,-rpath,"$W/occa/lib" -L"$W/occa/lib" -locca
echo "multi-rank dump driver built: $OUT/dump_mr_driver"
# 6. initial-guess dump driver (our own code, oracle/refbuild/dump_ig_driver.cpp)
g++ -fopenmp -O2 -std=c++17 -march=native -Wno-unused-function \
  -DLIBP_DIR="\"$W\"" \
  -I"$HERE/mpistub" -include "$W/lapack_rename.h" -I"$W/include" -I"$W/occa/include" \
  -o "$OUT/dump_ig_driver" "$HERE/dump_ig_driver.cpp" \
  -L"$W/libs" -llinearSolver -lmesh -lparAdogs -logs -llinAlg -lcore \
  "$W/libmpistub.a" -Wl,-rpath,"$BL" -L"$BL" -l:"$BLSO" -Wl,-rpath,"$W/occa/lib" -L"$W/occa/lib" -locca
echo "initial-guess dump driver built: $OUT/dump_ig_driver"
# 7. IPDG dump driver (our own code, oracle/refbuild/dump_ipdg_driver.cpp; single rank or under mpirun_stub.sh)
g++ -fopenmp -O2 -std=c++17 -march=native -Wno-unused-function \
  -DLIBP_DIR="\"$W\"" \
  -I"$HERE/mpistub" -include "$W/lapack_rename.h" -I"$W/include" -I"$W/occa/include" -I"$W/solvers/elliptic" \
  -o "$OUT/dump_ipdg_driver" "$HERE/dump_ipdg_driver.cpp" \
  "$W/solvers/elliptic/libelliptic.a" \
  -L"$W/libs" -lparAlmond -llinearSolver -lmesh -lparAdogs -logs -llinAlg -lcore \
  "$W/libmpistub.a" -Wl,-rpath,"$BL" -L"$BL" -l:"$BLSO" -Wl,-rpath,"$W/occa/lib" -L"$W/occa/lib" -locca
echo "IPDG dump driver built: $OUT/dump_ipdg_driver"
