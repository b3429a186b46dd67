"""One-off measurements of the BASELINE.json configs that are not the bench line (config 2 is bench.py):
  config3: training step, 4 objects x 128 rays, forward + backward (fp32 training path), 3 source views
  config4: 640x640 target view (409 600 rays), 3 source views, YOLO-sized maps (3 x 1792 x 80 x 80, synthetic)
  config5: samples-per-ray sweep (64/128/256 coarse) x source views (1/3/5), one 128x128 target view each
Prints one JSON line per measurement (CUDA events, median of `reps`).  Usage: python scripts/config_bench.py [3] [4] [5]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import pixel_nerf_yolo_b200.synth as synth  # noqa: E402
from pixel_nerf_yolo_b200.render import NeRFRenderer  # noqa: E402

dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


def flop_per_ray(ns, c, kc, kf, h=512, d_in=42, d_out=4):
    per_pt = 2 * (ns * (d_in * h + 3 * (c * h + 2 * h * h)) + 2 * 2 * h * h + h * d_out)
    return per_pt * (kc + kc + kf)


def config3(train_precision="fp32"):
    scene = H.make_scene_dict(num_objs=4, num_views=3, feat=64, size=128)
    net = H.build_net(scene, precision="bf16", train=True)
    net.train_precision = train_precision
    lat = scene["latent"].to(dev).clone().requires_grad_(True)
    net.encoder.set_latent(lat)
    r = NeRFRenderer(64, 32, 16, white_bkgd=True).train().to(dev)
    rays = H.rays_subset(4, 128, seed=1).to(dev)
    gt = torch.rand(4, 128, 3, device=dev)
    params = [p for p in net.parameters() if p.requires_grad]

    def fwd_only():
        with torch.enable_grad():
            res = r(net, rays)
        return res

    def step():
        for p in params:
            p.grad = None
        lat.grad = None
        res = fwd_only()
        loss = ((res.coarse.rgb - gt) ** 2).mean() + ((res.fine.rgb - gt) ** 2).mean()
        loss.backward()

    t_f, t_fb = timed(fwd_only), timed(step)
    net.eval()
    r.eval()

    def infer():
        with torch.no_grad():
            r(net, rays)
    t_inf = timed(infer)
    fl = flop_per_ray(3, 512, 64, 32) * 512
    print(json.dumps({"config": 3, "workload": f"train step: 4 objects x 128 rays, 3 views, 64+32 samples, fwd+bwd ({train_precision} training path)",
                      "fwd_ms": round(t_f, 2), "fwd_bwd_ms": round(t_fb, 2), "rays_per_s_fwd_bwd": round(512 / (t_fb * 1e-3)),
                      "algorithmic_TFLOPs_fwd_bwd": round(3 * fl / (t_fb * 1e-3) / 1e12, 1),
                      "inference_same_rays_ms (bf16 tcgen05)": round(t_inf, 3)}))


def render_case(ns, c, feat, size, n_rays, kc, kf=32, kfd=16, reps=3):
    scene = H.make_scene_dict(num_objs=1, num_views=ns, feat=feat, size=size, C=c)
    conf = dict(H.MODEL_CONF)
    if c != 512:
        conf = json.loads(json.dumps(H.MODEL_CONF))
        conf["encoder"] = {"backbone": "custom", "pretrained": False, "num_layers": 4, "index_padding": "zeros"}
    net = H.build_net(scene, precision="bf16", model_conf=conf)
    r = NeRFRenderer(kc, kf, kfd, white_bkgd=True).eval().to(dev)
    rays = synth.target_rays(size)[:, :n_rays].contiguous().to(dev)

    def go():
        with torch.no_grad():
            r(net, rays)
    ms = timed(go, reps=reps)
    fl = flop_per_ray(ns, c, kc, kf) * rays.shape[1]
    return ms, rays.shape[1], fl


def config4():
    ms, n, fl = render_case(3, 1792, 80, 640, 640 * 640, 64, reps=2)
    print(json.dumps({"config": 4, "workload": "640x640 target view (409600 rays), 3 views, synthetic 3x1792x80x80 maps, 64+32 samples, 1 GPU",
                      "ms": round(ms, 1), "rays_per_s": round(n / (ms * 1e-3)), "algorithmic_TFLOPs": round(fl / (ms * 1e-3) / 1e12, 1)}))


def config5():
    for kc in (64, 128, 256):
        for ns in (1, 3, 5):
            ms, n, fl = render_case(ns, 512, 64, 128, 128 * 128, kc)
            print(json.dumps({"config": 5, "n_coarse": kc, "source_views": ns, "rays": n, "ms": round(ms, 2),
                              "rays_per_s": round(n / (ms * 1e-3)), "algorithmic_TFLOPs": round(fl / (ms * 1e-3) / 1e12, 1)}))


def config5_sweep_views():
    """BASELINE config 5 as stated: 24 target views (eval_real.py's turntable, theta = linspace(-180, 180, 25)[:-1], phi 0) of
    128 x 128 rays = 393 216 rays per (n_coarse, source views) cell.  Under torchrun every rank renders views rank::world (the
    reference's loop over views, sharded; no data-path collective); time = max over ranks (device events, barrier on both sides)."""
    import numpy as np
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    thetas = np.linspace(-180, 180, 25)[:-1]
    mine = [float(t) for t in thetas[rank::world]]
    for kc in (64, 128, 256):
        for ns in (1, 3, 5):
            scene = H.make_scene_dict(num_objs=1, num_views=ns, feat=64, size=128, C=512)
            net = H.build_net(scene, precision="bf16")
            r = NeRFRenderer(kc, 32, 16, white_bkgd=True).eval().to(dev)
            views = [synth.target_rays(128, theta=t, phi=0.0).contiguous().to(dev) for t in mine]

            def go():
                with torch.no_grad():
                    for v in views:
                        r(net, v)
            go()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); go(); e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            n = 24 * 128 * 128
            if rank == 0:
                print(json.dumps({"config": 5, "workload": "24 turntable views x 128x128", "n_gpus": world, "n_coarse": kc, "source_views": ns,
                                  "rays": n, "ms": round(ms, 2), "rays_per_s": round(n / (ms * 1e-3)),
                                  "algorithmic_TFLOPs": round(flop_per_ray(ns, 512, kc, 32) * n / (ms * 1e-3) / 1e12, 1)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def yolo():
    """conf/exp/yolo.conf shape: YoloRenderer, 128 samples per ray, 3 anchors x 7 values, 3 x 1792 x 80 x 80 maps (synthetic), 640x640 image."""
    import copy
    from pixel_nerf_yolo_b200.conf import ConfigTree
    from pixel_nerf_yolo_b200.model import make_model
    from pixel_nerf_yolo_b200.render import YoloRenderer
    conf = copy.deepcopy(H.MODEL_CONF)
    conf["mlp_coarse"].update({"d_out": 7, "num_scales": 1, "num_anchors_per_scale": 3, "yolo": True})
    conf["mlp_fine"] = {"type": "empty"}
    conf["encoder"] = {"backbone": "custom", "pretrained": False, "num_layers": 4, "index_padding": "zeros"}
    scene = H.make_scene_dict(num_objs=1, num_views=3, feat=80, size=640, C=1792)
    net = make_model(ConfigTree.from_dict(conf)).eval()
    net.mlp_coarse.load_state_dict(synth.mlp_state(31, d_out=21, d_latent=1792))
    net = net.to(dev)
    net.num_objs, net.num_views_per_obj = 1, 3
    net.encoder.set_latent(scene["latent"].to(dev))
    net.set_cameras(torch.linalg.inv(scene["poses"][0]).to(dev), scene["focal"].to(dev), scene["image_wh"])
    r = YoloRenderer(128, 128, 1, 3)
    r.bind_parallel(net)
    rays = synth.target_rays(640)[0, ::10].contiguous().to(dev)          # 40 960 rays
    ms = timed(lambda: r(rays), reps=3)
    h, c, d_in = 512, 1792, 42
    per_pt = 2 * (3 * (d_in * h + 3 * (c * h + 2 * h * h)) + 2 * 2 * h * h + h * 21)
    print(json.dumps({"config": "yolo", "workload": "YoloRenderer, 40960 rays x 128 samples, 3 views, 3x1792x80x80 maps, d_out 21",
                      "ms": round(ms, 1), "rays_per_s": round(rays.shape[0] / (ms * 1e-3)),
                      "algorithmic_TFLOPs": round(per_pt * 128 * rays.shape[0] / (ms * 1e-3) / 1e12, 1)}))


which = [a if a in ("yolo", "5views") else int(a) for a in sys.argv[1:]] or [3, 4, 5, "yolo"]
for w in which:
    {3: config3, 4: config4, 5: config5, "yolo": yolo, "5views": config5_sweep_views}[w]()
    if w == 3:
        config3("tf32")
        config3("bf16")
