"""Golden vectors for the steps either side of the hot path, from the UNMODIFIED
reference: util.gen_rays (src/util/util.py:240-278) and SpatialEncoder.forward's pyramid tail
(src/model/encoder.py:138-172).  Run in the build container only:   python tests/golden/make_golden_aux.py
"""
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
    import util
    from model.encoder import SpatialEncoder

    out = {}
    # ---- gen_rays: 2 cameras, non-square image, explicit principal point and the default one
    poses = torch.stack([synth.pose_spherical(25.0, -15.0, 1.3), synth.pose_spherical(-70.0, 10.0, 1.7)])
    out["rays_poses"] = poses.numpy()
    out["rays_default_c"] = util.gen_rays(poses, 40, 24, torch.tensor(41.5), 0.8, 1.8).numpy()
    out["rays_with_c"] = util.gen_rays(poses, 40, 24, torch.tensor([41.5, 39.0]), 0.5, 2.5, c=torch.tensor([17.25, 13.5])).numpy()
    # ---- encoder pyramid: seeded random-init resnet34, one 64x64 image; per-level tensors are captured from the
    # reference's own forward (self.latents holds them after upsampling, so hook the trunk stages instead)
    torch.manual_seed(3)
    enc = SpatialEncoder("resnet34", pretrained=False, num_layers=4, index_padding="zeros").eval()
    levels = []
    enc.model.relu.register_forward_hook(lambda m, i, o: levels.append(o.detach().clone()) if len(levels) == 0 else None)
    for stage in (enc.model.layer1, enc.model.layer2, enc.model.layer3):
        stage.register_forward_hook(lambda m, i, o: levels.append(o.detach().clone()))
    img = torch.from_numpy(np.random.default_rng(4).uniform(-1, 1, (1, 3, 64, 64)).astype(np.float32))
    with torch.no_grad():
        latent = enc(img)
    assert len(levels) == 4 and latent.shape == (1, 512, 32, 32), (len(levels), latent.shape)
    for i, l in enumerate(levels):
        out[f"pyr_level{i}"] = l.numpy()
    out["pyr_latent_sub"] = latent[:, ::4].numpy().copy()        # every 4th channel (all four levels are covered)
    out["pyr_latent_sum"] = np.float64(latent.double().sum().item())
    out["pyr_latent_scaling"] = enc.latent_scaling.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_aux.npz"), **out)
    print("wrote reference_aux.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
