#!/bin/bash
# Builds the C-ABI library in-tree for sm_100a (B200).  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../pai_b200/libpai_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=$(ls "$HERE"/*.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -shared -cudart static \
    --expt-relaxed-constexpr "$@" -o "$OUT" $SRCS
echo "built $OUT"
