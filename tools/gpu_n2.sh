# 2-GPU validation + measurements (one gpurun --gpus 2 call)
mkdir -p gpurun_out/n2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time timeout 400 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q) > gpurun_out/n2/pytest_mg.log 2>&1; tail -4 gpurun_out/n2/pytest_mg.log
SOBFU_B200_TRACE=1 timeout 200 $TR --master-port 29511 tests/peer_check_worker.py 256 200 > gpurun_out/n2/peer_check_n2.log 2>&1; tail -1 gpurun_out/n2/peer_check_n2.log | cut -c1-2500
SOBFU_B200_PDL=1 PEER_CHECK_MODES=peer SOBFU_B200_TRACE=1 timeout 200 $TR --master-port 29513 tests/peer_check_worker.py 256 200 > gpurun_out/n2/peer_check_n2_pdl.log 2>&1; tail -1 gpurun_out/n2/peer_check_n2_pdl.log | cut -c1-1500
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --extra-dim 512 > gpurun_out/n2/bench_n2_peer.json 2> gpurun_out/n2/bench_n2_peer.err
SOBFU_B200_NO_PEER=1 timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2/bench_n2_nccl.json 2> gpurun_out/n2/bench_n2_nccl.err
timeout 300 $TR --master-port 29515 bench.py --gpus 2 --workload pipeline --frames 24 > gpurun_out/n2/bench_pipe_n2.json 2> gpurun_out/n2/bench_pipe_n2.err
python - <<'PY'
import json
for f in ("bench_n2_peer","bench_n2_nccl","bench_pipe_n2"):
    try:
        d=json.loads(open("gpurun_out/n2/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("solver_iters_per_s"), d["value"], d.get("kernel_ms"), d["e2e"], (d.get("parity") or {}).get("bit_exact"))
        if "extra_512" in d: print("  extra_512", d["extra_512"]["value"], d["extra_512"]["e2e"], (d["extra_512"].get("parity") or {}).get("bit_exact"))
    except Exception as e: print(f, "failed", e)
PY
tail -c 400 gpurun_out/n2/*.err
