# 8-GPU measurements (one gpurun --gpus 8 call)
mkdir -p gpurun_out/n8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
SOBFU_B200_TRACE=1 timeout 240 $TR --master-port 29511 tests/peer_check_worker.py 256 200 > gpurun_out/n8/peer_check_n8.log 2>&1; grep -a "bit for bit\|^{" gpurun_out/n8/peer_check_n8.log | cut -c1-6000
timeout 400 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8/bench_n8_peer.json 2> gpurun_out/n8/bench_n8_peer.err
SOBFU_B200_NO_PEER=1 timeout 300 $TR --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --extra-dim 0 > gpurun_out/n8/bench_n8_nccl.json 2> gpurun_out/n8/bench_n8_nccl.err
timeout 300 $TR --master-port 29515 bench.py --gpus 8 --workload pipeline --frames 50 > gpurun_out/n8/bench_pipe_n8.json 2> gpurun_out/n8/bench_pipe_n8.err
python - <<'PY'
import json
for f in ("bench_n8_peer","bench_n8_nccl","bench_pipe_n8"):
    try:
        d=json.loads(open("gpurun_out/n8/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("solver_iters_per_s"), d["value"], d["ms_per_step"], d.get("loop_ms_per_iter"), d.get("kernel_ms"), d["e2e"], (d.get("parity") or {}).get("bit_exact"))
        if "extra_512" in d:
            e=d["extra_512"]; print("  extra_512", e["value"], e["ms_per_step"], e["loop_ms_per_iter"], e["kernel_ms"], e["e2e"], (e.get("parity") or {}).get("bit_exact"))
    except Exception as e: print(f, "failed", e)
PY
tail -c 600 gpurun_out/n8/*.err | grep -v "UserWarning\|return func\|^$\|OMP_NUM\|\*\*\*\*" | tail -20
