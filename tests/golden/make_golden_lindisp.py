"""Goldens for the linear-in-disparity sampler variant (``lindisp=True``: nerf.py:121,153; the DTU configuration renders with
it) produced by executing the UNMODIFIED reference, same recipe and shims as make_golden.py:

    python tests/golden/make_golden_lindisp.py        # build container only -> tests/golden/reference_lindisp.npz

One SB=1 scene, 40 rays, black background (``white_bkgd=False``, as for DTU), rays with per-ray near/far bounds that differ
from ray to ray (so 1/near and 1/far are genuinely per ray).  Stored: the inputs that cannot be regenerated (ray indices,
near/far, noise) and the reference's outputs (three samplers stage by stage, full NeRFRenderer.forward).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    G._install_shims()
    import pixel_nerf_yolo_b200.synth as synth
    from render import NeRFRenderer
    torch.set_num_threads(8)
    out = {}
    nrays, size = 40, 128
    scene = synth.scene_config1(seed=5, num_views=3, C=512, size=size, feat=16, num_objs=1)
    net = G.build_reference_net(synth, 1, 2, scene, 1)
    pick = torch.from_numpy(np.random.default_rng(17).choice(size * size, nrays, replace=False)).long()
    rays = synth.target_rays(size, 15.0, -10.0)[:, pick].clone()
    rng = np.random.default_rng(18)
    rays[0, :, 6] = torch.from_numpy(rng.uniform(0.6, 0.9, nrays).astype(np.float32))      # near
    rays[0, :, 7] = torch.from_numpy(rng.uniform(1.6, 2.2, nrays).astype(np.float32))      # far
    out["ray_idx"], out["near"], out["far"] = pick.numpy(), rays[0, :, 6].numpy(), rays[0, :, 7].numpy()
    noise = G.np_noise(19, nrays)
    for k, v in noise.items():
        out[f"noise_{k}"] = v.numpy()
    conf = dict(G.RENDER_CONF)
    conf["white_bkgd"] = False
    renderer = NeRFRenderer.from_conf(G._Conf(conf), lindisp=True, eval_batch_size=50000).eval()
    assert renderer.lindisp and not renderer.white_bkgd
    with torch.no_grad(), G._NoisePatch(noise):
        res = renderer(net, rays, want_weights=True)
    for lvl in ("coarse", "fine"):
        out[f"{lvl}_rgb"] = res[lvl].rgb.numpy()
        out[f"{lvl}_depth"] = res[lvl].depth.numpy()
        out[f"{lvl}_weights"] = res[lvl].weights.numpy()
    r = rays.reshape(-1, 8)
    with G._NoisePatch(noise):
        zc = renderer.sample_coarse(r)
        zf = renderer.sample_fine(r, torch.from_numpy(out["coarse_weights"]).reshape(-1, 64))
        zd = renderer.sample_fine_depth(r, torch.from_numpy(out["coarse_depth"]).reshape(-1))
    out["z_coarse"], out["z_fine"], out["z_depth"] = zc.numpy(), zf.numpy(), zd.numpy()
    path = os.path.join(HERE, "reference_lindisp.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    print("coarse sum-w", out["coarse_weights"].sum(-1).mean(), "fine sum-w", out["fine_weights"].sum(-1).mean(),
          "fine depth mean", out["fine_depth"].mean(), "z_coarse spacing first/last",
          float(zc[0, 1] - zc[0, 0]), float(zc[0, -1] - zc[0, -2]))


if __name__ == "__main__":
    main()
