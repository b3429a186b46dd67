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
from pixel_nerf_yolo_b200.dist import GradientSync, ShardedRenderer  # noqa: E402
from pixel_nerf_yolo_b200.render import NeRFRenderer  # noqa: E402


def check_data_parallel_train_step(dev):
    """Data-parallel training: every rank renders + back-propagates its own equal slice of the ray batch, one
    bucketed all-reduce averages the gradients; the result must equal the whole-batch step's gradients computed on one GPU
    (fp32 training path: only the summation order differs, <= 1e-3 of each tensor's max)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per = 16
    scene = H.make_scene_dict(num_objs=1, num_views=3, feat=16)
    rays = H.rays_subset(1, per * world, seed=3).to(dev)
    noise = {k: v.to(dev) for k, v in H.make_noise(per * world, seed=4).items()}
    gt = torch.rand(1, per * world, 3, generator=torch.Generator().manual_seed(5)).to(dev)
    mse = torch.nn.functional.mse_loss

    def step(s, e):
        net = H.build_net(scene, device=dev, precision="bf16", train=True)
        r = NeRFRenderer(64, 32, 16, white_bkgd=True).train().to(dev)
        r.noise_override = {k: v[s:e].contiguous() for k, v in noise.items()}
        res = r(net, rays[:, s:e].contiguous())
        (mse(res.coarse.rgb, gt[:, s:e]) + mse(res.fine.rgb, gt[:, s:e])).backward()
        return net

    whole = step(0, per * world)
    mine = step(rank * per, (rank + 1) * per)
    params = list(mine.mlp_coarse.parameters()) + list(mine.mlp_fine.parameters())
    n = GradientSync(params).sync()
    assert n == sum(p.numel() for p in params)
    ref = list(whole.mlp_coarse.parameters()) + list(whole.mlp_fine.parameters())
    worst = 0.0
    for a, b in zip(ref, params):
        scale = a.grad.abs().max().item()
        err = (a.grad - b.grad).abs().max().item()
        assert err <= 1e-3 * scale + 1e-12, (tuple(a.shape), err, scale)
        worst = max(worst, err / (scale + 1e-30))
    return n, worst


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
    n, worst = check_data_parallel_train_step(dev)
    dist.barrier()
    if dist.get_rank() == 0:
        print("NCCL_SHARD_OK world", dist.get_world_size(), "rays", rays.shape[1])
        print("NCCL_GRADSYNC_OK elements", n, "worst rel err", f"{worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
