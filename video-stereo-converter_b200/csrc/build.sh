#!/bin/bash
# Builds libvsc_b200.so (C ABI + sm_100a kernels) in-tree.  No torch / Python dependency.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib/libvsc_b200.so"
mkdir -p "$HERE/../lib"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 \
  --shared -o "$OUT" "$HERE/vsc_api.cu" -lcudart_static -ldl -lrt -lpthread "$@"
echo "built $OUT"
