"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/make_golden.py).

The reference arithmetic is fp32 ATen; the oracle restates it with the same ops, so agreement is at
fp32 round-off (matmul blocking may differ between host CPUs, hence 2e-5 rather than bit-exact on the
MLP outputs; integer/sampler stages are compared exactly).
"""
import numpy as np
import torch

import pixel_nerf_yolo_b200.synth as synth
from oracle import pixelnerf_oracle as O

T = torch.from_numpy


def _scene(num_objs):
    s = synth.scene_config1(seed=5, num_views=3, C=512, size=128, feat=16, num_objs=num_objs)
    return O.encode_cameras(s["latent"], s["poses"], s["focal"], s["image_wh"])


def test_positional_encoding_matches_reference(golden):
    out = O.positional_encoding(T(golden["pe_x"]))
    ref = T(golden["pe_out"])
    # the oracle rounds sin() once from double (host-independent); ATen's fp32 sin in the golden run is within 1 ulp of that
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=1.2e-7, rtol=0)
    assert (out == ref).float().mean() > 0.9


def test_bilinear_index_matches_grid_sample(golden):
    sc = _scene(1)
    out = O.bilinear_index(sc.latent, T(golden["index_uv"]), sc.latent_scaling, sc.image_shape)
    np.testing.assert_allclose(out.numpy(), golden["index_out"], atol=2e-6, rtol=0)
    # the out-of-bounds (zero padding) branch is exercised
    assert (golden["index_out"] == 0).all(axis=1).any()


def test_resnetfc_matches_reference(golden):
    out = O.resnetfc_forward(synth.mlp_state(21), T(golden["mlp_zx"]), 512, 5, 3, (3, 5))
    np.testing.assert_allclose(out.reshape(2, 5, 4).numpy(), golden["mlp_out"], atol=2e-5, rtol=1e-5)


def test_field_forward_matches_reference(golden):
    sc = _scene(1)
    for name, seed in (("field_coarse", 1), ("field_fine", 2)):
        out = O.field_forward(sc, synth.mlp_state(seed), T(golden["field_xyz"]), T(golden["field_dirs"]))
        np.testing.assert_allclose(out.numpy(), golden[name], atol=2e-5, rtol=1e-5)


def _rays(golden, tag, num_objs):
    allr = torch.cat([synth.target_rays(128, 15.0 + 20 * s, -10.0) for s in range(num_objs)])
    return allr[:, T(golden[f"{tag}_ray_idx"]).long()]


def _noise(golden, tag):
    g = lambda k: T(golden[f"{tag}_noise_{k}"])
    return O.RenderNoise(g("coarse"), g("fine_u"), g("fine_jitter"), g("depth"))


def test_samplers_bit_exact(golden):
    for tag, nobj in (("sb1", 1), ("sb2", 2)):
        r = _rays(golden, tag, nobj).reshape(-1, 8)
        n = _noise(golden, tag)
        assert torch.equal(O.sample_coarse(r, n.coarse, 64), T(golden[f"{tag}_z_coarse"]))
        w = T(golden[f"{tag}_coarse_weights"]).reshape(-1, 64)
        assert torch.equal(O.sample_fine(r, w, n.fine_u, n.fine_jitter, 64), T(golden[f"{tag}_z_fine"]))
        d = T(golden[f"{tag}_coarse_depth"]).reshape(-1)
        assert torch.equal(O.sample_fine_depth(r, d, n.depth, 0.01), T(golden[f"{tag}_z_depth"]))


def test_render_matches_reference(golden):
    for tag, nobj in (("sb1", 1), ("sb2", 2)):
        sc = _scene(nobj)
        res = O.render(sc, synth.mlp_state(1), synth.mlp_state(2), _rays(golden, tag, nobj), _noise(golden, tag))
        for lvl in ("coarse", "fine"):
            for k in ("rgb", "depth", "weights"):
                np.testing.assert_allclose(res[lvl][k].numpy(), golden[f"{tag}_{lvl}_{k}"], atol=3e-5, rtol=1e-4,
                                           err_msg=f"{tag} {lvl} {k}")
        # both passes have non-trivial opacity, so the MLP output really reaches the pixels
        assert 0.3 < golden[f"{tag}_fine_weights"].sum(-1).mean() < 0.99


# ---------------------------------------------------------------------------------- lindisp (nerf.py:121,153; DTU renders with it)
def lindisp_case():
    """Inputs of tests/golden/make_golden_lindisp.py rebuilt from the stored indices/bounds/noise."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_lindisp.npz"))
    rays = synth.target_rays(128, 15.0, -10.0)[:, T(g["ray_idx"]).long()].clone()
    rays[0, :, 6], rays[0, :, 7] = T(g["near"]), T(g["far"])
    noise = {k: T(g[f"noise_{k}"]) for k in ("coarse", "fine_u", "fine_jitter", "depth")}
    return g, rays, noise


def test_lindisp_samplers_bit_exact():
    g, rays, nz = lindisp_case()
    r = rays.reshape(-1, 8)
    assert torch.equal(O.sample_coarse(r, nz["coarse"], 64, lindisp=True), T(g["z_coarse"]))
    w = T(g["coarse_weights"]).reshape(-1, 64)
    assert torch.equal(O.sample_fine(r, w, nz["fine_u"], nz["fine_jitter"], 64, lindisp=True), T(g["z_fine"]))
    assert torch.equal(O.sample_fine_depth(r, T(g["coarse_depth"]).reshape(-1), nz["depth"], 0.01), T(g["z_depth"]))
    # the variant is really exercised: sample spacing grows along the ray (uniform in 1/z)
    zc = g["z_coarse"]
    assert ((zc[:, -1] - zc[:, -9]) > 2 * (zc[:, 8] - zc[:, 0])).all()


def test_lindisp_render_matches_reference():
    g, rays, nz = lindisp_case()
    res = O.render(_scene(1), synth.mlp_state(1), synth.mlp_state(2), rays,
                   O.RenderNoise(nz["coarse"], nz["fine_u"], nz["fine_jitter"], nz["depth"]), white_bkgd=False, lindisp=True)
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "depth", "weights"):
            np.testing.assert_allclose(res[lvl][k].numpy(), g[f"{lvl}_{k}"], atol=3e-5, rtol=1e-4, err_msg=f"{lvl} {k}")
    assert 0.3 < g["fine_weights"].sum(-1).mean() < 0.99


# ---------------------------------------------------------------------------------- intrinsics formats (models.py:124-148, 225-230)
CAMERA_CASES = ("scalar_focal_scalar_c", "per_object_focal_vec_c", "per_object_fxfy_cxcy", "single_row_fxfy_no_c")


def camera_goldens():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cameras.npz"))


def test_camera_formats_match_reference():
    g = camera_goldens()
    s = synth.scene_config1(seed=5, num_views=3, C=512, size=128, feat=16, num_objs=2)
    for name in CAMERA_CASES:
        c = T(g[name + "_c"]) if name + "_c" in g else None
        sc = O.encode_cameras(s["latent"], s["poses"], T(g[name + "_focal"]), s["image_wh"], c=c)
        out = O.field_forward(sc, synth.mlp_state(1), T(g["xyz"]), T(g["dirs"]))
        np.testing.assert_allclose(out.numpy(), g[name], atol=2e-5, rtol=1e-5, err_msg=name)
