"""Goldens for the camera-intrinsics formats ``PixelNeRFNet.encode`` accepts (models.py:124-148: focal as a scalar, (SB,) or
(SB, 2); principal point ``c`` absent, scalar, (SB,) or (SB, 2); per-object rows are repeated over the source views in
``forward``, models.py:225-230), produced by executing the UNMODIFIED reference -- same recipe and shims as make_golden.py:

    python tests/golden/make_golden_cameras.py        # build container only -> tests/golden/reference_cameras.npz

SB = 2 objects x 3 views; ``PixelNeRFNet.forward`` (coarse MLP) at 23 points per object for each format.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402

# name -> (focal, c); None = argument omitted
CASES = {
    "scalar_focal_scalar_c": (torch.tensor(120.0), torch.tensor(60.0)),
    "per_object_focal_vec_c": (torch.tensor([131.25, 118.0]), torch.tensor([61.5, 66.0])),
    "per_object_fxfy_cxcy": (torch.tensor([[131.25, 125.0], [110.0, 140.0]]), torch.tensor([[60.0, 70.0], [68.5, 58.25]])),
    "single_row_fxfy_no_c": (torch.tensor([[140.0, 120.0]]), None),
}


def main():
    G._install_shims()
    import pixel_nerf_yolo_b200.synth as synth
    from model import make_model
    torch.set_num_threads(8)
    out = {}
    num_objs, size = 2, 128
    scene = synth.scene_config1(seed=5, num_views=3, C=512, size=size, feat=16, num_objs=num_objs)
    rng = np.random.default_rng(23)
    pts = torch.from_numpy(rng.uniform(-0.4, 0.4, (num_objs, 23, 3)).astype(np.float32))
    dirs = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((num_objs, 23, 3)).astype(np.float32)), dim=-1)
    out["xyz"], out["dirs"] = pts.numpy(), dirs.numpy()
    for name, (focal, c) in CASES.items():
        torch.manual_seed(0)
        net = make_model(G._Conf(G.MODEL_CONF)).eval()
        net.mlp_coarse.load_state_dict(synth.mlp_state(1))
        net.encode(torch.zeros(num_objs, 3, 3, size, size), scene["poses"], focal.clone(), c=None if c is None else c.clone())
        lat = scene["latent"]
        net.encoder.latent = lat
        ls = torch.tensor([float(lat.shape[-1]), float(lat.shape[-2])])
        net.encoder.latent_scaling = ls / (ls - 1) * 2.0
        with torch.no_grad():
            out[name] = net(pts, coarse=True, viewdirs=dirs).numpy()
        out[name + "_focal"] = focal.numpy()
        if c is not None:
            out[name + "_c"] = c.numpy()
    path = os.path.join(HERE, "reference_cameras.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    base = out["scalar_focal_scalar_c"]
    print({k: float(np.abs(out[k] - base).max()) for k in CASES})      # the formats really give different images


if __name__ == "__main__":
    main()
