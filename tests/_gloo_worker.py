"""Worker for test_sharded_render_two_ranks_gloo_matches_single (launched under torchrun, gloo, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200.dist import ShardedRenderer  # noqa: E402


def fake_render(rays, noise):
    # any per-ray function of (ray, its noise rows): stands in for the single-GPU render
    sb, b = rays.shape[:2]
    n = noise["coarse"].reshape(sb, b, -1)
    rgb = torch.stack((rays[..., 0] * 2 + n.sum(-1), rays[..., 3] - n[..., 0], rays[..., 6] * n[..., 1]), dim=-1)
    return rgb, rays[..., 7] + n.mean(-1)


def main():
    dist.init_process_group("gloo")
    g = torch.Generator().manual_seed(3)
    for sb, b in ((1, 1000), (2, 77), (1, 5)):
        rays = torch.randn(sb, b, 8, generator=g)
        noise = {"coarse": torch.rand(sb * b, 4, generator=g), "depth": None}
        rgb, depth = ShardedRenderer(fake_render)(rays, noise)
        ref_rgb, ref_depth = fake_render(rays, noise)
        assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth), (sb, b)
    dist.barrier()
    if dist.get_rank() == 0:
        print("GLOO_SHARD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
