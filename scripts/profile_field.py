"""Small driver for ncu: a few renders of N rays (default 4096) of the config-2 scene."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
net, renderer, rays, scene, images, state = bench.build_ours(dev)
rays = rays[:, :: max(1, rays.shape[1] // n)][:, :n].contiguous().to(dev)
wrapped = renderer.bind_parallel(net, None, simple_output=True).eval()
with torch.no_grad():
    for _ in range(reps):
        rgb, depth = wrapped(rays)
torch.cuda.synchronize()
print("ok", rgb.shape, float(rgb.mean()), float(depth.mean()))
