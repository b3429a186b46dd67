"""Golden vectors of the YOLO mode (PixelNeRFNet.forward with mlp_coarse.yolo, YoloRenderer.forward) from the
UNMODIFIED reference.  Run in the build container only:   python tests/golden/make_golden_yolo.py

The YOLOv7 trunk (src/model/custom_encoder.py, un-vendored NeRF-YOLO checkout + checkpoint) is not available, so the
network is built with the resnet34 encoder class and synthetic 512-channel maps are injected as its output: everything
downstream of the encoder -- the code this golden pins -- is the reference's YOLO path (conf/exp/yolo.conf head:
d_out = 7 x 3 anchors, n_coarse 128, no fine network).
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG._install_shims()
    import pixel_nerf_yolo_b200.synth as synth
    from model import make_model
    from render.yolo import YoloRenderer

    conf = copy.deepcopy(MG.MODEL_CONF)
    conf["mlp_coarse"].update({"d_out": 7, "num_scales": 1, "num_anchors_per_scale": 3, "yolo": True})
    conf["mlp_fine"] = {"type": "empty"}
    torch.manual_seed(0)
    net = make_model(MG._Conf(conf)).eval()
    assert net.yolo and net.d_out == 21
    net.mlp_coarse.load_state_dict(synth.mlp_state(31, d_out=21))
    scene = synth.scene_config1(seed=5, num_views=3, C=512, size=128, feat=16, num_objs=1)
    # YOLO mode takes world->camera poses as given (models.py:119-120); use the inverse of the synthetic c2w poses with
    # the y/z axes flipped so that points in front of the camera have NEGATIVE z on one side and POSITIVE on the other
    c2w = scene["poses"][0]
    w2c = torch.linalg.inv(c2w)
    w2c[1, :] *= -1.0                                  # view 1 looks the other way: its z >= 0 rows get masked
    w2c[1, 2, :] *= 1.0
    images = torch.zeros(1, 3, 3, 128, 128)
    net.encode(images, w2c[None], scene["focal"])
    net.encoder.latent = scene["latent"]
    ls = torch.tensor([16.0, 16.0])
    net.encoder.latent_scaling = ls / (ls - 1) * 2.0
    rng = np.random.default_rng(17)
    pts = torch.from_numpy(rng.uniform(-0.4, 0.4, (1, 23, 3)).astype(np.float32))
    dirs = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((1, 23, 3)).astype(np.float32)), dim=-1)
    out = {"w2c": w2c.numpy(), "field_xyz": pts.numpy(), "field_dirs": dirs.numpy()}
    with torch.no_grad():
        out["field_out"] = net(pts, coarse=True, viewdirs=dirs).numpy()
    all_rays = synth.target_rays(128, 15.0, -10.0)
    pick = torch.from_numpy(np.random.default_rng(7).choice(128 * 128, 20, replace=False)).long()
    rays = all_rays[0, pick]
    noise = torch.from_numpy(rng.random((20, 128), dtype=np.float32))
    out["ray_idx"], out["noise"] = pick.numpy(), noise.numpy()
    r = YoloRenderer(128, 1024, 1, 3)
    r.bind_net(net)
    saved = torch.rand_like
    torch.rand_like = lambda *a, **k: noise.clone()
    try:
        with torch.no_grad():
            out["render"] = r(rays).numpy()
    finally:
        torch.rand_like = saved
    np.savez_compressed(os.path.join(HERE, "reference_yolo.npz"), **out)
    print("wrote reference_yolo.npz", {k: v.shape for k, v in out.items()}, "max prob", out["render"][..., 0].max())


if __name__ == "__main__":
    main()
