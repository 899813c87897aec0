#!/bin/bash
# Builds the C-ABI library in-tree for sm_100a (B200).  Usage: csrc/build.sh [extra nvcc flags]
# Every .cu is its own translation unit (no relocatable device code): they compile in parallel into
# csrc/build/*.o and are re-compiled only when the source, a header or the flags changed.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../pai_b200/libpai_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJ="$HERE/build"
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 128 $*"
STAMP="$OBJ/.flags"
if [ ! -f "$STAMP" ] || [ "$(cat "$STAMP")" != "$FLAGS" ]; then
    rm -f "$OBJ"/*.o
    echo "$FLAGS" > "$STAMP"
fi
NEWEST_HDR=$(ls -t "$HERE"/*.cuh "$HERE"/*.h "$HERE"/../../include/*.h | head -1)
pids=()
for src in "$HERE"/*.cu; do
    o="$OBJ/$(basename "${src%.cu}").o"
    if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ "$NEWEST_HDR" -nt "$o" ]; then
        "$NVCC" $FLAGS -c -o "$o" "$src" &
        pids+=($!)
    fi
done
for p in "${pids[@]:-}"; do
    [ -n "$p" ] && wait "$p"
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
