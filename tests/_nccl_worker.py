"""Multi-GPU check (torchrun, NCCL): the ray-sharded render gathered from N GPUs equals the 1-GPU render bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import pixel_nerf_yolo_b200.synth as synth  # noqa: E402
from pixel_nerf_yolo_b200.dist import ShardedRenderer  # noqa: E402
from pixel_nerf_yolo_b200.render import NeRFRenderer  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    scene = H.make_scene_dict(feat=32)
    net = H.build_net(scene, device=dev, precision="bf16")
    renderer = NeRFRenderer(64, 32, 16, white_bkgd=True).eval().to(dev)
    wrapped = renderer.bind_parallel(net, None, simple_output=True).eval()
    rays = synth.target_rays(128)[:, :5003].contiguous().to(dev)          # not a multiple of the tile or the world size
    noise = {k: v.to(dev) for k, v in H.make_noise(rays.shape[1], seed=3).items()}

    def render_fn(r, nz):
        renderer.noise_override = nz
        with torch.no_grad():
            return wrapped(r)

    rgb, depth = ShardedRenderer(render_fn)(rays, noise)
    ref_rgb, ref_depth = render_fn(rays, noise)                          # every rank also renders everything
    assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth), "sharded render differs from single-GPU render"
    dist.barrier()
    if dist.get_rank() == 0:
        print("NCCL_SHARD_OK world", dist.get_world_size(), "rays", rays.shape[1])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
