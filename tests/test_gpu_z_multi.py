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


def test_multi_device_renderer_matches_single_device():
    """`renderer.bind_parallel(net, [0, 1])` (the reference's `--gpu_id "0 1"` / DataParallel(dim=1) switch, nerf.py:373-377):
    with the same full-batch noise the 2-device render equals the 1-device render bit for bit (tuple and dict outputs), the
    replicas follow a re-encode and a weight update, and a call that would need gradients is refused."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import pixel_nerf_yolo_b200.synth as synth
    from pixel_nerf_yolo_b200.dist import MultiDeviceRenderer
    from pixel_nerf_yolo_b200.render import NeRFRenderer
    scene = H.make_scene_dict(feat=32)
    net = H.build_net(scene, device="cuda:0", precision="bf16")
    renderer = NeRFRenderer(64, 32, 16, white_bkgd=True).eval().to("cuda:0")
    rays = synth.target_rays(128)[:, :3001].contiguous().to("cuda:0")
    noise = {k: v.to("cuda:0") for k, v in H.make_noise(rays.shape[1], seed=3).items()}
    single = renderer.bind_parallel(net, None, simple_output=True).eval()
    multi = renderer.bind_parallel(net, [0, 1], simple_output=True).eval()
    assert isinstance(multi, MultiDeviceRenderer)
    multi_dict = renderer.bind_parallel(net, [0, 1], simple_output=False).eval()

    def both():
        renderer.noise_override = noise
        with torch.no_grad():
            a = single(rays)
            b = multi(rays)
            d = multi_dict(rays, want_weights=True)
        torch.cuda.synchronize()
        assert b[0].device == rays.device
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert torch.equal(d["fine"]["rgb"], a[0]) and d["fine"]["weights"].shape == (1, rays.shape[1], 96)
        return a[0].clone()
    first = both()
    net.encoder.set_latent(net.encoder.latent * 0.7)                   # re-encode: the replicas must follow
    second = both()
    assert not torch.equal(first, second)
    with torch.no_grad():
        net.mlp_fine.lin_out.bias.add_(0.5)                            # weight update
    third = both()
    assert not torch.equal(second, third)
    renderer.noise_override = None
    net.requires_grad_(True)
    with pytest.raises(NotImplementedError, match="inference driver"):
        multi(rays)
    with torch.no_grad():
        multi(rays)                                                    # fine under no_grad
