mkdir -p gpurun_out/r2c5
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2c5/pytest_all.log 2>&1; tail -6 gpurun_out/r2c5/pytest_all.log
for v in 0 4; do timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-traffic --variant $v > gpurun_out/r2c5/bench_v$v.json 2> gpurun_out/r2c5/bench_v$v.err; done
tools/run_variants.sh > gpurun_out/r2c5/variants.log 2>&1; cat gpurun_out/r2c5/variants.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-traffic --no-parity --dim 512 --iters 100 > gpurun_out/r2c5/bench_v0_512.json 2> gpurun_out/r2c5/bench_v0_512.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-traffic --no-parity --dim 128 > gpurun_out/r2c5/bench_v0_128.json 2> gpurun_out/r2c5/bench_v0_128.err
python - <<'PY'
import json
for f in ("v0","v4","v0_512","v0_128"):
    try:
        d=json.loads(open("gpurun_out/r2c5/bench_%s.json"%f).read().strip().splitlines()[-1]); print(f, d["solver_iters_per_s"], d["value"], d["kernel_ms"], d["e2e"]["frames_per_s"])
    except Exception as e: print(f, "failed", e)
PY
tail -c 300 gpurun_out/r2c5/*.err
