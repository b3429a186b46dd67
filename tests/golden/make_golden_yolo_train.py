"""Golden GRADIENTS of the YOLO detection head's training step from the UNMODIFIED reference's autograd
(YoloRenderer.forward -> loss.backward(), train/trainlib/YoloTrainer.py:140-190).
Run in the build container only:   python tests/golden/make_golden_yolo_train.py

Same network / scene / rays / noise as make_golden_yolo.py.  YoloLoss itself is outside the rendering path, so the scalar is a
fixed random linear functional of the rendered (B, anchors, 7) tensor: loss = sum(render * gw); it exercises every output
(the max-probability channel and the six probability-weighted box values).  Stored like make_golden_grads.py: loss, render,
every parameter gradient's norm/sum (+ full tensor or a slice), the gradient reaching the encoder output."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
from make_golden_grads import grad_summary  # noqa: E402


def main():
    MG._install_shims()
    import pixel_nerf_yolo_b200.synth as synth
    from model import make_model
    from render.yolo import YoloRenderer

    conf = copy.deepcopy(MG.MODEL_CONF)
    conf["mlp_coarse"].update({"d_out": 7, "num_scales": 1, "num_anchors_per_scale": 3, "yolo": True})
    conf["mlp_fine"] = {"type": "empty"}
    torch.manual_seed(0)
    net = make_model(MG._Conf(conf)).train()
    net.mlp_coarse.load_state_dict(synth.mlp_state(31, d_out=21))
    scene = synth.scene_config1(seed=5, num_views=3, C=512, size=128, feat=16, num_objs=1)
    w2c = torch.linalg.inv(scene["poses"][0])
    w2c[1, :] *= -1.0                                  # as in make_golden_yolo.py: view 1's z >= 0 rows get masked
    net.encode(torch.zeros(1, 3, 3, 128, 128), w2c[None], scene["focal"])
    lat = scene["latent"].clone().requires_grad_(True)
    net.encoder.latent = lat
    ls = torch.tensor([16.0, 16.0])
    net.encoder.latent_scaling = ls / (ls - 1) * 2.0
    rng = np.random.default_rng(23)
    all_rays = synth.target_rays(128, 15.0, -10.0)
    pick = torch.from_numpy(np.random.default_rng(7).choice(128 * 128, 20, replace=False)).long()
    rays = all_rays[0, pick]
    noise = torch.from_numpy(rng.random((20, 128), dtype=np.float32))
    gw = torch.from_numpy(rng.standard_normal((20, 3, 7)).astype(np.float32))
    r = YoloRenderer(128, 1024, 1, 3)
    r.bind_net(net)
    saved = torch.rand_like
    torch.rand_like = lambda *a, **k: noise.clone()
    try:
        render = r(rays)
    finally:
        torch.rand_like = saved
    loss = (render * gw).sum()
    loss.backward()
    out = {"ray_idx": pick.numpy(), "noise": noise.numpy(), "gw": gw.numpy(), "loss": np.float64(loss.item()),
           "render": render.detach().numpy(), "w2c": w2c.numpy()}
    grad_summary(out, "coarse.", [(n, p.grad) for n, p in net.mlp_coarse.named_parameters()])
    out["latent.norm"] = np.float64(lat.grad.double().norm().item())
    out["latent.slice"] = lat.grad[:, :16].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "reference_yolo_train.npz"), **out)
    print("wrote reference_yolo_train.npz; loss", loss.item(), "latent grad norm", out["latent.norm"],
          "lin_out.weight grad norm", out["coarse.lin_out.weight.norm"])


if __name__ == "__main__":
    main()
