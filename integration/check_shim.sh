#!/bin/bash
# Compile check of integration/libp_b200_shim.hpp against the reference's own headers (build container only:
# needs the scratch copy made by oracle/refbuild/build_ref.sh for OCCA's generated headers).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/.." && pwd)"
W=${LIBP_REF_WORK:-/tmp/libp_ref}
g++ -std=c++17 -fsyntax-only -Wall -Wno-unused-function \
  -I"$REPO/oracle/refbuild/mpistub" -I"$W/include" -I"$W/occa/include" -I"$W/solvers/elliptic" \
  -I"$REPO/include" -I/usr/local/cuda/include -I"$HERE" -DLIBP_DIR="\"$W\"" "$HERE/check_shim.cpp"
echo "shim compiles against the reference headers"
