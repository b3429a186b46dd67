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


def yolo_train_case():
    g = np.load(os.path.join(os.path.dirname(GOLD), "reference_yolo_train.npz"))
    scene = H.make_scene_dict(num_objs=1, num_views=3, feat=16, size=128, seed=5)
    rays = synth.target_rays(128, 15.0, -10.0)[0, T(g["ray_idx"]).long()]
    return g, scene, rays


def check_yolo_grads(g, grads, lat_grad, loss, render, rtol, what, lat_rtol=None):
    """Gradients of the YOLO head's training step vs the reference's autograd (tests/golden/make_golden_yolo_train.py)."""
    assert abs(float(loss) - float(g["loss"])) <= rtol * max(abs(float(g["loss"])), 1.0), f"{what}: loss {loss} vs {g['loss']}"
    np.testing.assert_allclose(render, g["render"], atol=max(rtol, 3e-5) * 3, rtol=rtol)
    for name, gr in grads.items():
        gr = gr.detach().cpu().double()
        ref_norm = float(g[f"coarse.{name}.norm"])
        kf, ks = f"coarse.{name}.full", f"coarse.{name}.slice"
        ref = T(g[kf]).double() if kf in g else T(g[ks]).double()
        got = gr if kf in g else gr[:8, :96]
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item()
        assert err <= rtol * max(scale, ref_norm / gr.numel() ** 0.5 * 30) + 1e-12, f"{what}: d {name} max err {err:.3e} scale {scale:.3e}"
        assert abs(gr.norm().item() - ref_norm) <= rtol * ref_norm + 1e-12, f"{what}: d {name} norm"
    lat = lat_grad.detach().cpu().double()
    ref_norm = float(g["latent.norm"])
    assert abs(lat.norm().item() - ref_norm) <= rtol * ref_norm, f"{what}: latent grad norm"
    err = (lat[:, :16] - T(g["latent.slice"]).double()).abs().max().item()
    scale = T(g["latent.slice"]).abs().max().item()
    assert err <= (lat_rtol or rtol) * scale + 1e-9, f"{what}: latent grad slice err {err:.3e} scale {scale:.3e}"


def test_yolo_train_step_oracle_matches_reference_autograd():
    g, scene, rays = yolo_train_case()
    sc = O.encode_cameras(scene["latent"], T(g["w2c"]), scene["focal"], scene["image_wh"], num_views=3, yolo=True)
    sc.latent = sc.latent.clone().requires_grad_(True)
    mlp = {k: v.clone().requires_grad_(True) for k, v in synth.mlp_state(31, d_out=21).items()}
    res = O.yolo_render(sc, mlp, rays, T(g["noise"]), grad=True)
    loss = (res * T(g["gw"])).sum()
    loss.backward()
    check_yolo_grads(g, {k: v.grad for k, v in mlp.items()}, sc.latent.grad, loss.item(), res.detach().numpy(), 2e-4, "oracle")
