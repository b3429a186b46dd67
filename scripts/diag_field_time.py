"""Time the fused field kernel alone on the bench workload (16 384 rays; coarse launch 64 samples/ray + fine launch 96+64) with
FIXED sample depths, so that timing-diagnostic builds whose outputs are garbage (scripts/build_diag_variants.sh) can be measured
without their outputs feeding back into the fine pass.   python scripts/diag_field_time.py [path/to/variant.so] [steps]

Diagnostics only: the variant library is selected by overriding the loader's path inside THIS script; the product loader has no
such switch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixel_nerf_yolo_b200 import _lib  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] not in ("", "product"):
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
import bench  # noqa: E402

dev = torch.device("cuda", 0)
net, renderer, rays, scene, images, state = bench.build_ours(dev)
rays = rays.reshape(-1, 8).contiguous().to(dev)
B = rays.shape[0]
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    z_c = renderer.sample_coarse(rays, torch.rand(B, 64, device=dev, generator=g))
    extra = 0.8 + torch.rand(B, 32, device=dev, generator=g)                        # stand-in for the 32 fine samples in [near, far]
    z_f = torch.sort(torch.cat((z_c, extra), -1), -1).values.contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for it in range(3 + steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net.field_from_rays(rays, z_c, coarse=True, sb=1)
        net.field_from_rays(rays, z_f, coarse=False, sb=1)
        e1.record()
        e1.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
ms = sum(times) / len(times)
import subprocess
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                     capture_output=True, text=True).stdout.strip()
print(f"{os.path.basename(_lib.LIB_PATH):32s} field ms/step {ms:7.2f}  (min {min(times):.2f})  algorithmic TFLOP/s {bench.FLOP_PER_RAY * B / ms / 1e9:7.1f}  clocks/power after: {clk}")
