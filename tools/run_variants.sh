#!/usr/bin/env bash
# On the GPU box: for every sobfu_b200/_lib/var/lib_<NAME>.so (tools/build_variants.sh) run a parity subset and a short bench.
# Usage (inside one gpurun call): tools/run_variants.sh [bench.py flags, e.g. --variant 4]
for lib in "$PWD"/sobfu_b200/_lib/var/lib_*.so; do
  export SOBFU_B200_LIB="$lib"
  echo "== variant $(basename "$lib") $*"
  timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiled" 2>&1 | tail -1
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-traffic "$@" 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['solver_iters_per_s'], d['kernel_ms'])"
done
