# 4-GPU validation + measurements (one gpurun --gpus 4 call)
mkdir -p gpurun_out/n4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
(time timeout 500 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q) > gpurun_out/n4/pytest_mg.log 2>&1; tail -4 gpurun_out/n4/pytest_mg.log
timeout 300 $TR --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/n4/bench_n4_peer.json 2> gpurun_out/n4/bench_n4_peer.err
timeout 300 $TR --master-port 29515 bench.py --gpus 4 --workload pipeline --frames 50 > gpurun_out/n4/bench_pipe_n4.json 2> gpurun_out/n4/bench_pipe_n4.err
SOBFU_B200_TRACE=1 PEER_CHECK_MODES=peer timeout 200 $TR --master-port 29511 tests/peer_check_worker.py 256 200 > gpurun_out/n4/peer_check_n4.log 2>&1; grep -a "^{" gpurun_out/n4/peer_check_n4.log | cut -c1-3000
python - <<'PY'
import json
for f in ("bench_n4_peer","bench_pipe_n4"):
    try:
        d=json.loads(open("gpurun_out/n4/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("solver_iters_per_s"), d["value"], d["ms_per_step"], d.get("loop_ms_per_iter"), d.get("kernel_ms"), d["e2e"], (d.get("parity") or {}).get("bit_exact"))
    except Exception as e: print(f, "failed", e)
PY
tail -c 400 gpurun_out/n4/*.err | grep -v "UserWarning\|return func\|^$\|OMP_NUM\|\*\*\*\*" | tail -10
