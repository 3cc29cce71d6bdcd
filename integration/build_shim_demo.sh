#!/bin/bash
# Builds the reference (libParanumal + OCCA with CUDA enabled) in a second scratch copy and links
# integration/elliptic_b200_main.cpp (reference objects + the B200 shim) against it; packs what the binary needs at run
# time (OCCA JIT: OKL sources, headers) under baseline/_ref/libp_ref_cuda/ (git-ignored, travels to the GPU box).
# Build container only.  On the GPU box: integration/run_shim_demo.sh.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/.." && pwd)"
SRC=${LIBP_REFERENCE:-/root/reference}
W=/tmp/libp_ref_cuda           # baked into the binaries as LIBP_DIR / OCCA_BUILD_DIR: the run script symlinks it
RB="$REPO/oracle/refbuild"
OUT="$REPO/baseline/_ref/libp_ref_cuda"
if [ ! -d "$W/libs" ]; then
  mkdir -p "$W"
  cp -r "$SRC"/{include,libs,solvers,make.top,makefile} "$W"/
  cp -r "$SRC"/occa "$W"/occa
  chmod -R u+w "$W"
fi
BL=$(python3 -c "import scipy,os;print(os.path.realpath(os.path.join(os.path.dirname(scipy.__file__),'..','scipy.libs')))")
BLSO=$(basename "$BL"/libscipy_openblas-*.so)
{ for p in d s; do for f in gecon geev gels geqp3 gesv getrf getri lange ormqr trsm; do echo "#define ${p}${f}_ scipy_${p}${f}_"; done; done
  echo "#define dsyev_ scipy_dsyev_"; } > "$W/lapack_rename.h"
if [ ! -f "$W/occa/lib/libocca.so" ]; then
  make -C "$W/occa" -j8 CXX=g++ CC=gcc CXXFLAGS="-O2 -include cstdint -I/usr/local/cuda/include" \
    OCCA_CUDA_ENABLED=1 OCCA_OPENCL_ENABLED=0 OCCA_HIP_ENABLED=0 OCCA_DPCPP_ENABLED=0 OCCA_METAL_ENABLED=0 \
    LDFLAGS="-L/usr/local/cuda/lib64/stubs -L/usr/local/cuda/lib64" > "$W/occa_build.log" 2>&1 || true  # `occa info` needs libcuda.so.1
  [ -f "$W/occa/lib/libocca.so" ]
fi
gcc -O2 -fPIC -c "$RB/mpistub/mpistub.c" -I"$RB/mpistub" -o "$W/mpistub.o" && ar rcs "$W/libmpistub.a" "$W/mpistub.o"
INC="-I$RB/mpistub -include $W/lapack_rename.h -I$W/include -I$W/occa/include"
FL="-fopenmp -O3 -Wall -Wno-unused-function -std=c++17 -mavx2 -march=x86-64-v3"
if [ ! -f "$W/solvers/elliptic/libelliptic.a" ]; then
  make -C "$W" -j8 libp_libs LIBP_CC=gcc LIBP_CXX=g++ LIBP_LD=g++ LIBP_INCLUDES="$INC" LIBP_CXXFLAGS="$FL" \
    LIBP_CFLAGS="-fopenmp -O3 -mavx2 -march=x86-64-v3" LIBP_BLAS_DIR="$BL" LIBP_BLAS_LIB="-L$BL -l:$BLSO" > "$W/libp_build.log" 2>&1
  make -C "$W/solvers/elliptic" lib LIBP_DIR="$W" LIBP_CC=gcc LIBP_CXX=g++ LIBP_LD=g++ LIBP_INCLUDES="$INC" LIBP_CXXFLAGS="$FL" \
    LIBP_BLAS_DIR="$BL" LIBP_BLAS_LIB="-L$BL -l:$BLSO" >> "$W/libp_build.log" 2>&1
fi
mkdir -p "$OUT/occa/lib"
g++ -fopenmp -O2 -std=c++17 -march=x86-64-v3 -Wno-unused-function -DLIBP_DIR="\"$W\"" \
  -I"$RB/mpistub" -include "$W/lapack_rename.h" -I"$W/include" -I"$W/occa/include" -I"$W/solvers/elliptic" \
  -I"$REPO/include" -I"$HERE" -I/usr/local/cuda/include \
  -o "$OUT/elliptic_b200_main" "$HERE/elliptic_b200_main.cpp" "$W/solvers/elliptic/libelliptic.a" \
  -L"$W/libs" -lparAlmond -llinearSolver -lmesh -lparAdogs -logs -llinAlg -lcore "$W/libmpistub.a" \
  -Wl,-rpath,"$BL" -L"$BL" -l:"$BLSO" -Wl,-rpath,'$ORIGIN/occa/lib' -L"$W/occa/lib" -locca \
  -Wl,-rpath,'$ORIGIN/../../../libparanumal_b200/lib' -L"$REPO/libparanumal_b200/lib" -lparanumal_b200 \
  -L/usr/local/cuda/lib64 -lcudart -Wl,--allow-shlib-undefined
cp "$W/occa/lib/libocca.so" "$OUT/occa/lib/"
# what OCCA's JIT reads at run time: OKL sources and the headers they include
for d in libs solvers/elliptic; do
  (cd "$W/$d" && find . \( -name '*.okl' -o -name '*.h' -o -name '*.hpp' \) -print0 | cpio -0 -pdm --quiet "$OUT/$d") 2>/dev/null || \
  (cd "$W/$d" && find . \( -name '*.okl' -o -name '*.h' -o -name '*.hpp' \) | while read f; do mkdir -p "$OUT/$d/$(dirname "$f")"; cp "$f" "$OUT/$d/$f"; done)
done
mkdir -p "$OUT/include" "$OUT/occa/include"
cp -r "$W/include/." "$OUT/include/"
cp -r "$W/occa/include/." "$OUT/occa/include/"
du -sh "$OUT"
echo "shim demo built: $OUT/elliptic_b200_main"
