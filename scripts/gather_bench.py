"""Stand-alone feature gather (pnr_gather_encode: projection + 4-tap bilinear gather + positional encoding) against the
HBM roofline (4 taps x C x 2 B read + C x 2 B written per (point, view) row).

  case l2 : config-2 maps, 3 x 64 x 64 x 512 bf16 = 12.6 MB (L2-resident; only the output stream touches HBM)
  case hbm: 3 x 320 x 320 x 512 bf16 = 315 MB (resnet34 on 640^2 inputs, SURVEY.md 8d config 4 variant) with points
            spread over the whole frustum, so taps miss the 126 MB L2
Prints one JSON line per case: rows/s, algorithmic GB/s and its fraction of MEASURED_PEAKS.json's hbm_gbs.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pixel_nerf_yolo_b200 import _lib  # noqa: E402
import helpers as H  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
peak = 6529.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(tag, feat_hw, size, n_points, iters=5):
    scene = H.make_scene_dict(num_objs=1, num_views=3, feat=16, size=size)
    net = H.build_net(scene, precision="bf16")
    C = 512
    g = torch.Generator(device="cpu").manual_seed(1)
    base = (torch.randn(3, C, 64, 64, generator=g) * 0.5).to(dev)
    latent = base if feat_hw == 64 else torch.nn.functional.interpolate(base, (feat_hw, feat_hw), mode="bilinear", align_corners=True)
    net.encoder.set_latent(latent.contiguous())           # NCHW fp32, as the encoder trunk leaves it
    net._cam_cache = None
    sc, keep = net._scene(fp32_maps=False)                # channels-last bf16 via pnr_pack_features
    maps = net.encoder.packed_latent(fp32=False)
    # points uniformly inside the unit cube around the origin (all three cameras look at it)
    xyz = ((torch.rand(1, n_points, 3, generator=g) - 0.5) * 0.9).to(dev).contiguous()
    dirs = torch.nn.functional.normalize(torch.randn(1, n_points, 3, generator=g), dim=-1).to(dev).contiguous()
    pts = _lib.points_xyz(xyz, dirs)
    rows = 3 * n_points
    lat = torch.empty(rows, C, dtype=torch.bfloat16, device=dev)
    zf = torch.empty(rows, 42, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def call():
        _lib.check(lib.pnr_gather_encode(sc, pts, lat.data_ptr(), zf.data_ptr(), 0, 6, 1.5, _lib.stream_ptr(dev)), "gather")

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.fill_(1)                      # evict the maps / outputs from L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[len(ms) // 2] * 1e-3
    algo = rows * (4 * C * 2 + C * 2 + 42 * 4)
    inb = (rows * (lat[:1].numel() * 0 + 1))  # noqa: F841
    frac_in = float((lat.float().abs().sum(dim=1) > 0).float().mean())
    print(json.dumps({"case": tag, "maps": f"3x{feat_hw}x{feat_hw}x{C} bf16 ({maps.numel() * 2 / 1e6:.1f} MB)", "rows": rows,
                      "ms": round(t * 1e3, 3), "rows_per_s": round(rows / t), "algorithmic_GBps": round(algo / t / 1e9, 1),
                      "hbm_peak_GBps": peak, "frac_of_hbm_peak": round(algo / t / 1e9 / peak, 3),
                      "rows_inside_a_map": round(frac_in, 3),
                      "bytes_per_row": 4 * C * 2 + C * 2 + 42 * 4}))


run("l2", 64, 128, 1 << 20)
run("hbm", 320, 640, 1 << 20)
