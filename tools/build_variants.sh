#!/bin/bash
# Kernel tuning aid: builds lib/var/libpnol_<tag>.so with extra -D flags for ONE translation unit, the other objects are reused
# from build/. Load a variant with PNOL_B200_LIB=<path> (capi.py).
#   tools/build_variants.sh residual_kernels E2 -DLORENTZ_FENCE_EVERY=2
set -e
cd "$(dirname "$0")/.."
tu=$1; tag=$2; shift 2
PKG=parallelnonlinearoptimizationlibrary_b200
mkdir -p build/var $PKG/lib/var
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off \
     -Iinclude -I$PKG/csrc "$@" -c $PKG/csrc/$tu.cu -o build/var/${tu}_$tag.o 2> build/var/${tu}_$tag.log
objs=""
for o in build/*.o; do
  case "$o" in build/host_*) continue;; build/$tu.o) continue;; esac
  objs="$objs $o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/lib/var/libpnol_$tag.so $objs build/var/${tu}_$tag.o -ldl
echo "built $PKG/lib/var/libpnol_$tag.so"
