#!/usr/bin/env bash
# Compiles the reference's OWN gtest sources (test/main.cpp, deformation_field_test.cpp, reductions_test.cpp,
# solver_test.cpp) and its application (src/apps/demo.cpp), unchanged and where they lie under /root/reference, against the drop-in headers of this repo
# (include/) and links them with libsobfu_b200.so.  Output: oracle/_ref/sobfu_test_dropin (git-ignored; travels to the GPU
# box).  Test infrastructure: proves the drop-in boundary; -fpermissive as in the reference's CMakeLists.txt:25.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SOBFU_REFERENCE_DIR:-/root/reference}"
ROOT="$HERE/.."
if [ ! -d "$REF/test" ]; then echo "build_dropin_tests: $REF not present; keeping prebuilt binary (if any)"; exit 0; fi
mkdir -p "$HERE/_ref"
g++ -std=c++14 -O1 -fpermissive -w -I"$ROOT/include" -I"$ROOT/include/compat" -I/usr/local/cuda/include \
    "$REF/test/main.cpp" "$REF/test/deformation_field_test.cpp" "$REF/test/reductions_test.cpp" "$REF/test/solver_test.cpp" \
    -o "$HERE/_ref/sobfu_test_dropin" -L"$ROOT/sobfu_b200/_lib" -lsobfu_b200 -L/usr/local/cuda/lib64 -lcudart \
    -Wl,-rpath,"$ROOT/sobfu_b200/_lib" -Wl,-rpath,/usr/local/cuda/lib64
echo "build_dropin_tests: wrote $HERE/_ref/sobfu_test_dropin"
# the reference's application (src/apps/demo.cpp), unchanged, against the same headers: Boost.ProgramOptions, OpenCV highgui,
# PCL io / visualization and VTK come from the dependency-free stand-ins in include/compat (the viewer calls are no-ops)
g++ -std=c++14 -O1 -fpermissive -w -I"$ROOT/include" -I"$ROOT/include/compat" -I/usr/local/cuda/include \
    "$REF/src/apps/demo.cpp" -o "$HERE/_ref/sobfu_app_dropin" -L"$ROOT/sobfu_b200/_lib" -lsobfu_b200 -L/usr/local/cuda/lib64 -lcudart \
    -Wl,-rpath,"$ROOT/sobfu_b200/_lib" -Wl,-rpath,/usr/local/cuda/lib64
echo "build_dropin_tests: wrote $HERE/_ref/sobfu_app_dropin"
