"""Worker for test_sharded_render_two_ranks_gloo_matches_single (launched under torchrun, gloo, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200.dist import GradientSync, ShardedRenderer  # noqa: E402


def fake_render(rays, noise):
    # any per-ray function of (ray, its noise rows): stands in for the single-GPU render
    sb, b = rays.shape[:2]
    n = noise["coarse"].reshape(sb, b, -1)
    rgb = torch.stack((rays[..., 0] * 2 + n.sum(-1), rays[..., 3] - n[..., 0], rays[..., 6] * n[..., 1]), dim=-1)
    return rgb, rays[..., 7] + n.mean(-1)


def check_gradient_sync():
    """Two ranks, each with its own half of a batch: after GradientSync.sync() every rank holds the gradient of the
    mean loss over the WHOLE batch (what a single process computes), including a parameter one rank never touched."""
    rank, world = dist.get_rank(), dist.get_world_size()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8 * world, 6, generator=g, dtype=torch.float64)
    y = torch.randn(8 * world, 3, generator=g, dtype=torch.float64)

    def make():
        torch.manual_seed(5)
        m = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
        m.extra = torch.nn.Parameter(torch.ones(4))          # used by rank 0 only
        return m

    def loss_of(m, xs, ys, use_extra):
        out = m(xs.float())
        l = ((out - ys.float()) ** 2).mean()
        return l + (m.extra.sum() * 0.5 if use_extra else 0.0)

    whole = make()
    # the whole-batch loss = mean over ranks of the per-rank losses (rank 0's includes the extra term)
    l = sum(loss_of(whole, x[r * 8:(r + 1) * 8], y[r * 8:(r + 1) * 8], r == 0) for r in range(world)) / world
    l.backward()
    mine = make()
    loss_of(mine, x[rank * 8:(rank + 1) * 8], y[rank * 8:(rank + 1) * 8], rank == 0).backward()
    assert (mine.extra.grad is None) == (rank != 0)
    sync = GradientSync(mine.parameters())
    n = sync.sync()
    assert n == sum(p.numel() for p in mine.parameters())
    for (name, a), b in zip(whole.named_parameters(), mine.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7), name
    n2 = sync.sync()                                         # idempotent on already-equal gradients (mean of equals)
    assert n2 == n
    for a, b in zip(whole.parameters(), mine.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7)


def main():
    dist.init_process_group("gloo")
    g = torch.Generator().manual_seed(3)
    for sb, b in ((1, 1000), (2, 77), (1, 5)):
        rays = torch.randn(sb, b, 8, generator=g)
        noise = {"coarse": torch.rand(sb * b, 4, generator=g), "depth": None}
        rgb, depth = ShardedRenderer(fake_render)(rays, noise)
        ref_rgb, ref_depth = fake_render(rays, noise)
        assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth), (sb, b)
    # no noise passed: the renderer's noise_fn draws the FULL batch on every rank from equally seeded generators (the four
    # draws of one NeRFRenderer.forward, nerf.py:117,141,147,164) and each rank takes its rows
    from pixel_nerf_yolo_b200.dist import draw_render_noise
    rays = torch.randn(1, 333, 8, generator=g)
    noise_fn = lambda n: draw_render_noise(n, 4, 8, 4, "cpu")
    torch.manual_seed(99)
    rgb, depth = ShardedRenderer(fake_render, noise_fn=noise_fn)(rays)
    torch.manual_seed(99)
    full = noise_fn(333)
    assert set(full) == {"coarse", "fine_u", "fine_jitter", "depth"} and full["fine_u"].shape == (333, 4) and full["depth"].shape == (333, 4)
    ref_rgb, ref_depth = fake_render(rays, full)
    assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
    check_gradient_sync()
    dist.barrier()
    if dist.get_rank() == 0:
        print("GLOO_SHARD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
