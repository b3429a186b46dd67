"""Oracle restatement of the YOLO mode (models.py:119-120,220-224,254-264,309-310; render/yolo.py) vs the unmodified
reference (tests/golden/make_golden_yolo.py)."""
import os

import numpy as np
import torch

import helpers as H
import pixel_nerf_yolo_b200.synth as synth
from oracle import pixelnerf_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_yolo.npz")
T = torch.from_numpy


def yolo_case():
    g = np.load(GOLD)
    scene = H.make_scene_dict(num_objs=1, num_views=3, feat=16, size=128, seed=5)
    sc = O.encode_cameras(scene["latent"], T(g["w2c"]), scene["focal"], scene["image_wh"], num_views=3, yolo=True)
    mlp = synth.mlp_state(31, d_out=21)
    rays = synth.target_rays(128, 15.0, -10.0)[0, T(g["ray_idx"]).long()]
    return g, scene, sc, mlp, rays


def test_yolo_field_and_render_match_reference():
    g, scene, sc, mlp, rays = yolo_case()
    out = O.field_forward(sc, mlp, T(g["field_xyz"]), T(g["field_dirs"]), yolo=True)
    assert out.shape == (1, 23, 21)
    np.testing.assert_allclose(out.numpy(), g["field_out"], atol=3e-5, rtol=1e-5)
    res = O.yolo_render(sc, mlp, rays, T(g["noise"]))
    np.testing.assert_allclose(res.numpy(), g["render"], atol=3e-5, rtol=1e-4)
    # the z >= 0 mask is exercised (view 1 looks away) and the other views do sample their maps
    x = T(g["field_xyz"])[0]
    zc = (sc.poses[:, None, :3, :3] @ x[None, :, :, None])[..., 0][..., 2] + sc.poses[:, None, 2, 3]
    assert (zc[1] >= 0).all() and (zc[0] < 0).all()
