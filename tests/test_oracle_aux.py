"""Oracle restatements of util.gen_rays and the encoder's pyramid tail vs the unmodified reference
(tests/golden/make_golden_aux.py)."""
import os

import numpy as np
import torch

from oracle import pixelnerf_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_aux.npz")
T = torch.from_numpy


def test_gen_rays_matches_reference():
    g = np.load(GOLD)
    poses = T(g["rays_poses"])
    np.testing.assert_allclose(O.gen_rays(poses, 40, 24, 41.5, 0.8, 1.8).numpy(), g["rays_default_c"], atol=1e-6, rtol=0)
    r = O.gen_rays(poses, 40, 24, [41.5, 39.0], 0.5, 2.5, c=[17.25, 13.5])
    np.testing.assert_allclose(r.numpy(), g["rays_with_c"], atol=1e-6, rtol=0)


def test_pyramid_latent_matches_reference():
    g = np.load(GOLD)
    lat = O.pyramid_latent([T(g[f"pyr_level{i}"]) for i in range(4)])
    assert lat.shape == (1, 512, 32, 32)
    np.testing.assert_allclose(lat[:, ::4].numpy(), g["pyr_latent_sub"], atol=1e-6, rtol=1e-6)
    assert abs(lat.double().sum().item() - float(g["pyr_latent_sum"])) < 1e-2


def test_gen_rays_yolo_matches_reference():
    """util.gen_rays_yolo (src/util/util.py:808-876), goldens from tests/golden/make_golden_rays_yolo.py."""
    g = np.load(os.path.join(os.path.dirname(GOLD), "reference_rays_yolo.npz"))
    w2c = T(g["w2c"])
    r = O.gen_rays_yolo(w2c, 20, 15, g["focal"], g["c"], 0.5, 6.0)
    assert r.shape == (3, 15, 20, 8)
    np.testing.assert_allclose(r.numpy(), g["rays_20x15"], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(O.gen_rays_yolo(w2c[:1], 1, 1, g["focal"], g["c"], 0.1, 2.0).numpy(), g["rays_1x1"], atol=1e-6, rtol=1e-6)
