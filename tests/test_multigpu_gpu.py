"""z-slab solve on 2 GPUs == single-GPU solve, bit for bit (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_slab_solve_matches_single_gpu(built, mode):
    """mode peer: halo stores + maxima over NVLink peer memory (CUDA IPC); mode nccl: the same iteration over NCCL"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu)")
    n = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("SOBFU_B200_NO_PEER", None)
    env.pop("SOBFU_B200_PEER", None)
    if mode == "nccl":
        env["SOBFU_B200_NO_PEER"] = "1"      # peer mode is the default; this keeps the exchange over NCCL
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n, "--master-addr", "127.0.0.1",
                        "--master-port", "29541" if mode == "peer" else "29543", os.path.join(ROOT, "tests", "multigpu_worker.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert r.stdout.count("slab == single GPU, bit for bit") == 5
    assert r.stdout.count("slab meshes == single GPU, bit for bit") == 1
    assert ("peer mode: True" in r.stdout) == (mode == "peer"), r.stdout[-2000:]
