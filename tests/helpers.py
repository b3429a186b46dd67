"""Shared builders for the parity tests: the same synthetic scene/weights for the CUDA path and the oracle."""
import torch

import pixel_nerf_yolo_b200.synth as synth
from pixel_nerf_yolo_b200.conf import ConfigTree

MODEL_CONF = {
    "use_encoder": True, "use_global_encoder": False, "use_xyz": True, "canon_xyz": False,
    "use_code": True, "code": {"num_freqs": 6, "freq_factor": 1.5, "include_input": True},
    "use_viewdirs": True, "use_code_viewdirs": False,
    "mlp_coarse": {"type": "resnet", "n_blocks": 5, "d_hidden": 512, "d_out": 4, "combine_layer": 3,
                   "combine_type": "average"},
    "mlp_fine": {"type": "resnet", "n_blocks": 5, "d_hidden": 512, "d_out": 4, "combine_layer": 3,
                 "combine_type": "average"},
    "encoder": {"backbone": "resnet34", "pretrained": False, "num_layers": 4, "index_padding": "zeros"},
}
RENDER_CONF = {"n_coarse": 64, "n_fine": 32, "n_fine_depth": 16, "depth_std": 0.01, "sched": [], "white_bkgd": True}


def make_scene_dict(num_objs=1, num_views=3, feat=16, size=128, seed=5, C=512):
    return synth.scene_config1(seed=seed, num_views=num_views, C=C, size=size, feat=feat, num_objs=num_objs)


def build_net(scene, device="cuda", coarse_seed=1, fine_seed=2, precision="bf16", model_conf=None, train=False):
    """pixel_nerf_yolo_b200 PixelNeRFNet with synthetic weights and an injected (synthetic) encoder output.
    ``train=False`` freezes the parameters (an inference build: the drop-in, like the reference, records a backward pass
    whenever gradient mode is on and something on the path requires grad, whatever the train/eval mode)."""
    from pixel_nerf_yolo_b200.model import make_model
    net = make_model(ConfigTree.from_dict(model_conf or MODEL_CONF)).eval()
    C = scene["latent"].shape[1]
    net.mlp_coarse.load_state_dict(synth.mlp_state(coarse_seed, d_latent=C))
    net.mlp_fine.load_state_dict(synth.mlp_state(fine_seed, d_latent=C))
    net = net.to(device)
    poses = scene["poses"]
    net.num_objs, net.num_views_per_obj = poses.shape[0], poses.shape[1]
    net.encoder.set_latent(scene["latent"].to(device))
    net.set_cameras(poses.reshape(-1, 4, 4).to(device), scene["focal"].to(device), scene["image_wh"])
    net.precision = precision
    if train:
        net.train()
    else:
        net.requires_grad_(False)
    return net


def oracle_scene(scene):
    from oracle import pixelnerf_oracle as O
    return O.encode_cameras(scene["latent"], scene["poses"], scene["focal"], scene["image_wh"])


def rays_subset(num_objs, n, size=128, seed=7):
    import numpy as np
    allr = torch.cat([synth.target_rays(size, 15.0 + 20 * s, -10.0) for s in range(num_objs)])
    pick = torch.from_numpy(np.random.default_rng(seed).choice(size * size, n, replace=False)).long()
    return allr[:, pick].contiguous()


def make_noise(B, seed=9, kc=64, kf=16, kfd=16):
    import numpy as np
    rng = np.random.default_rng(seed)
    return {"coarse": torch.from_numpy(rng.random((B, kc), dtype=np.float32)),
            "fine_u": torch.from_numpy(rng.random((B, kf), dtype=np.float32)),
            "fine_jitter": torch.from_numpy(rng.random((B, kf), dtype=np.float32)),
            "depth": torch.from_numpy(rng.standard_normal((B, kfd)).astype(np.float32))}


def oracle_train_step(scene, rays, noise, gt, coarse_seed=1, fine_seed=2, depth_loss=0.0, **render_kw):
    """The reference's training step on the CPU oracle under autograd (PixelNerfTrainer.py:141-157):
    loss = MSE(coarse.rgb, gt) + MSE(fine.rgb, gt) [+ depth_loss * mean(fine.depth)], backward.
    Returns (loss, result, {"coarse": grads, "fine": grads, "latent": grad})."""
    from oracle import pixelnerf_oracle as O
    C = scene["latent"].shape[1]
    sc = oracle_scene(scene)
    sc.latent = sc.latent.clone().requires_grad_(True)
    mc = {k: v.clone().requires_grad_(True) for k, v in synth.mlp_state(coarse_seed, d_latent=C).items()}
    mf = {k: v.clone().requires_grad_(True) for k, v in synth.mlp_state(fine_seed, d_latent=C).items()}
    n = O.RenderNoise(noise["coarse"], noise.get("fine_u"), noise.get("fine_jitter"), noise.get("depth"))
    res = O.render(sc, mc, mf, rays, n, grad=True, **render_kw)
    mse = torch.nn.functional.mse_loss
    loss = mse(res["coarse"]["rgb"], gt)
    if "fine" in res:
        loss = loss + mse(res["fine"]["rgb"], gt) + depth_loss * res["fine"]["depth"].mean()
    loss.backward()
    z = lambda d: {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in d.items()}
    return loss.detach(), res, {"coarse": z(mc), "fine": z(mf), "latent": sc.latent.grad}
