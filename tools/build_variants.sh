#!/usr/bin/env bash
# Builds alternative libsobfu_b200.so files with different compile-time tuning macros into sobfu_b200/_lib/var/lib_<NAME>.so
# (git-ignored; they travel to the GPU box).  tools/run_variants.sh then benchmarks and parity-tests each of them in ONE gpurun
# call through SOBFU_B200_LIB.  Usage: tools/build_variants.sh NAME "<nvcc -D flags>" [NAME "<flags>" ...]
#   tools/build_variants.sh A "" B "-DPA2_AHEAD=2" C "-DPA2_CTAS=2"
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/sobfu_b200/csrc"
OUT="$ROOT/sobfu_b200/_lib/var"
mkdir -p "$OUT"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --ftz=true --prec-div=false --prec-sqrt=false -Xcompiler -fPIC
       -I"$ROOT/include" -I"$CSRC")
while [ $# -ge 2 ]; do
    name="$1"; defs="$2"; shift 2
    objdir="$OUT/obj_$name"; mkdir -p "$objdir"
    objs=()
    for src in capi.cu solver_generic.cu solver_tiled.cu field_ops.cu tsdf_ops.cu marching_cubes.cu io_capi.cu; do
        extra=(); [ "$src" = io_capi.cu ] && extra=(-I"$ROOT/include/compat")
        # only the tiled kernels carry tuning macros; the other objects are compiled once per variant for simplicity
        nvcc "${FLAGS[@]}" "${extra[@]}" $defs -c "$CSRC/$src" -o "$objdir/${src%.cu}.o" &
        objs+=("$objdir/${src%.cu}.o")
    done
    wait
    nvcc -shared -o "$OUT/lib_$name.so" "${objs[@]}" -lcudart
    echo "built $OUT/lib_$name.so  [$defs]"
done
