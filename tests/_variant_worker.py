"""Renders a small ray batch with the bf16 tensor-core path under whatever PNR_* kernel-variant knobs the parent test set in the
environment (the knobs are read once per process) and checks it against the CPU oracle; prints VARIANT_OK."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import pixel_nerf_yolo_b200.synth as synth  # noqa: E402
from oracle import pixelnerf_oracle as O  # noqa: E402
from pixel_nerf_yolo_b200.render import NeRFRenderer  # noqa: E402


def main():
    worst = 0.0
    for num_views, n_rays in ((3, 150), (2, 70), (1, 40)):
        scene = H.make_scene_dict(num_views=num_views)
        net = H.build_net(scene, precision="bf16")
        r = NeRFRenderer(64, 32, 16, white_bkgd=True).eval().cuda()
        rays = H.rays_subset(1, n_rays, seed=num_views)
        noise = H.make_noise(n_rays, seed=10 + num_views)
        r.noise_override = {k: v.cuda() for k, v in noise.items()}
        with torch.no_grad():
            res = r(net, rays.cuda())
        ref = O.render(H.oracle_scene(scene), synth.mlp_state(1), synth.mlp_state(2), rays,
                       O.RenderNoise(noise["coarse"], noise["fine_u"], noise["fine_jitter"], noise["depth"]))
        for lvl in ("coarse", "fine"):
            e = max((res[lvl].rgb.cpu() - ref[lvl]["rgb"]).abs().max().item(), (res[lvl].depth.cpu() - ref[lvl]["depth"]).abs().max().item())
            worst = max(worst, e)
            assert e < 1e-2, (num_views, lvl, e)
    print("VARIANT_OK", {k: v for k, v in os.environ.items() if k.startswith("PNR_")}, f"max err {worst:.2e}")


if __name__ == "__main__":
    main()
