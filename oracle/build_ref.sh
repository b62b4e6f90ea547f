#!/usr/bin/env bash
# Builds the UNMODIFIED reference hot path (dgrzech/sobfu) for sm_100a from the sources where they lie
# (/root/reference, read-only) into oracle/_ref/libsobfu_ref.so.  Test infrastructure only; no reference
# source is copied into the repo: objects and the two texture-patched translation units live in a mktemp dir.
# Flags follow the reference's CMakeLists.txt:25,40-45 (numerics flags kept, only the gencode differs).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SOBFU_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/sobfu" ]; then echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0; fi
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
INC="-I$HERE/../include/compat -I$REF/include"
NVFLAGS="-gencode arch=compute_100a,code=sm_100a --ftz=true --prec-div=false --prec-sqrt=false -O3 -std=c++14 -Xcompiler -fPIC -w $INC"
CXXFLAGS="-std=c++14 -O2 -fPIC -fpermissive -w -include cstdio -include cmath $INC -I/usr/local/cuda/include"
python3 "$HERE/patch_textures.py" "$REF" "$TMP"
pids=()
for f in src/sobfu/cuda/solver.cu src/sobfu/cuda/vector_fields.cu src/sobfu/cuda/reductor.cu \
         src/sobfu/cuda/scalar_fields.cu src/kfusion/cuda/imgproc.cu; do
  nvcc $NVFLAGS -c "$REF/$f" -o "$TMP/$(echo $f | tr / _).o" & pids+=($!)
done
for f in tsdf_volume.cu marching_cubes.cu; do
  nvcc $NVFLAGS -c "$TMP/$f" -o "$TMP/patched_$f.o" & pids+=($!)
done
# second oracle for marching cubes only: the same translation unit + the __syncwarp its compaction lacks (patch_textures.py)
mkdir -p "$TMP/alt"
nvcc $NVFLAGS -c "$TMP/marching_cubes_syncwarp.cu" -o "$TMP/alt/patched_marching_cubes.cu.o" & pids+=($!)
for f in src/sobfu/solver.cpp src/sobfu/reductor.cpp src/sobfu/vector_fields.cpp src/sobfu/scalar_fields.cpp \
         src/sobfu/precomp.cpp src/kfusion/device_memory.cpp src/kfusion/tsdf_volume.cpp src/kfusion/precomp.cpp \
         src/kfusion/marching_cubes.cpp src/kfusion/imgproc.cpp src/kfusion/core.cpp; do
  g++ $CXXFLAGS -c "$REF/$f" -o "$TMP/$(echo $f | tr / _).o" & pids+=($!)
done
g++ $CXXFLAGS -c "$HERE/ref_harness.cpp" -o "$TMP/ref_harness.o" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libsobfu_ref.so" "$TMP"/*.o -lcudart -L/usr/local/cuda/lib64/stubs -lcuda
echo "build_ref: wrote $OUT/libsobfu_ref.so"
objs=()
for o in "$TMP"/*.o; do [ "$(basename "$o")" = "patched_marching_cubes.cu.o" ] || objs+=("$o"); done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libsobfu_ref_mcsync.so" "${objs[@]}" "$TMP/alt/patched_marching_cubes.cu.o" -lcudart -L/usr/local/cuda/lib64/stubs -lcuda
echo "build_ref: wrote $OUT/libsobfu_ref_mcsync.so (marching cubes with the missing __syncwarp; test oracle only)"
