mkdir -p gpurun_out/r2c8
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tiled or full_size" 2>&1 | tail -2
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r2c8/bench_v0.json 2> gpurun_out/r2c8/bench_v0.err
tools/run_variants.sh > gpurun_out/r2c8/variants.log 2>&1; cat gpurun_out/r2c8/variants.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c8/bench_v0.json").read().strip().splitlines()[-1]); print("v0", d["solver_iters_per_s"], d["value"], d["kernel_ms"], d["e2e"]["frames_per_s"])
PY
