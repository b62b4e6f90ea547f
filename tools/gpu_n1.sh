# one-GPU validation + measurements of a round: tools/gpurun_retry.sh --timeout 2400 -- 'bash tools/gpu_n1.sh'
mkdir -p gpurun_out/n1
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/n1/pytest_all.log 2>&1; tail -4 gpurun_out/n1/pytest_all.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/n1/bench_n1.json 2> gpurun_out/n1/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 --extra-dim 512 > gpurun_out/n1/bench_ref.json 2> gpurun_out/n1/bench_ref.err
timeout 300 python bench.py --workload pipeline --frames 50 > gpurun_out/n1/bench_pipe_n1.json 2> gpurun_out/n1/bench_pipe_n1.err
timeout 600 python bench.py --impl reference --workload pipeline --frames 50 > gpurun_out/n1/bench_pipe_ref.json 2> gpurun_out/n1/bench_pipe_ref.err
timeout 300 python bench.py --steps 3 --warmup 3 --dim 512 --no-cpu-baseline --no-traffic > gpurun_out/n1/bench_512_n1.json 2> gpurun_out/n1/bench_512_n1.err
python - <<'PY'
import json
for f in ("bench_n1","bench_ref","bench_pipe_n1","bench_pipe_ref","bench_512_n1"):
    try:
        d=json.loads(open("gpurun_out/n1/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("solver_iters_per_s"), d["value"], d["ms_per_step"], d.get("kernel_ms"), d["e2e"], (d.get("parity") or {}).get("bit_exact"), d.get("roofline"), d.get("cpu_baseline",{}).get("value"))
        if "extra_512" in d: print("  extra_512", d["extra_512"])
    except Exception as e: print(f, "failed", e)
PY
tail -c 300 gpurun_out/n1/*.err
