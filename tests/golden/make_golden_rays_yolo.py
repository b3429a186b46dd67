"""Golden vectors of util.gen_rays_yolo (src/util/util.py:808-876) from the UNMODIFIED reference.
Run in the build container only:   python tests/golden/make_golden_rays_yolo.py
Cameras are world-to-camera extrinsics (what the YOLO dataset stores, models.py:119-120); grids as YoloTrainer.calc_losses
builds them (image / cell size, YoloTrainer.py:108-121): a 20 x 15 grid with separate fx/fy and an off-centre principal point."""
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
    c2w = torch.stack([synth.pose_spherical(t, -20.0 + 5 * i, 1.3 + 0.1 * i) for i, t in enumerate((0.0, 40.0, -75.0))])
    w2c = torch.linalg.inv(c2w)
    out = {"w2c": w2c.numpy()}
    focal, c = torch.tensor([20.5, 19.75]), torch.tensor([9.6, 7.9])
    out["focal"], out["c"] = focal.numpy(), c.numpy()
    out["rays_20x15"] = util.gen_rays_yolo(w2c, 20, 15, focal, c, 0.5, 6.0).numpy()
    out["rays_1x1"] = util.gen_rays_yolo(w2c[:1], 1, 1, focal, c, 0.1, 2.0).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_rays_yolo.npz"), **out)
    print("wrote reference_rays_yolo.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
