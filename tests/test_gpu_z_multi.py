"""(Named to sort after the single-GPU parity tests.)  Multi-GPU parity (needs >= 2 B200 on the box; skipped on a 1-GPU box, where scripts/gpu_multi.sh is the manual route):
torchrun + NCCL, (1) the ray-sharded render gathered from N GPUs equals the 1-GPU render bit for bit, (2) a data-parallel
training step with one bucketed gradient all-reduce equals the whole-batch step (tests/_nccl_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_render_and_gradient_sync_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    with socket.socket() as sk:                      # a free rendezvous port on this box
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_SHARD_OK" in r.stdout and "NCCL_GRADSYNC_OK" in r.stdout
