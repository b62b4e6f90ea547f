for v in A B C D E; do
  export SOBFU_B200_LIB=$PWD/sobfu_b200/_lib/var/lib_$v.so
  echo "== variant $v"
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['solver_iters_per_s'], d['kernel_ms'])"
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiled" 2>&1 | tail -1
done
