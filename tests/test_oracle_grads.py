"""Pin the oracle's GRADIENTS (torch autograd over the functional restatement) against the gradients the unmodified
reference's autograd produced for the same training step (tests/golden/make_golden_grads.py, BASELINE config 3
shape: 2 objects x 24 rays, 3 source views, loss = MSE(coarse) + MSE(fine))."""
import os

import numpy as np
import torch

import helpers as H
import pixel_nerf_yolo_b200.synth as synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_grads.npz")
T = torch.from_numpy


def check_against_golden(grads, loss, gold, rtol, what):
    """Shared by the CPU (oracle) and GPU (CUDA backward) tests."""
    assert abs(float(loss) - float(gold["loss"])) <= rtol * abs(float(gold["loss"])), f"{what}: loss"
    for lvl in ("coarse", "fine"):
        for name, g in grads[lvl].items():
            g = g.detach().cpu().double()
            ref_norm = float(gold[f"{lvl}.{name}.norm"])
            key_full, key_slice = f"{lvl}.{name}.full", f"{lvl}.{name}.slice"
            ref = T(gold[key_full]).double() if key_full in gold else T(gold[key_slice]).double()
            got = g if key_full in gold else g[:8, :96]
            # error relative to the tensor's own scale (norm / sqrt(numel) = rms)
            rms = ref_norm / max(g.numel(), 1) ** 0.5
            err = (got - ref).abs().max().item()
            assert err <= rtol * max(rms * 30, 1e-12), f"{what}: {lvl}.{name} max err {err:.3e} vs rms {rms:.3e}"
            assert abs(g.norm().item() - ref_norm) <= rtol * ref_norm + 1e-12, f"{what}: {lvl}.{name} norm"
    lat = grads["latent"].detach().cpu().double()
    ref_norm = float(gold["latent.norm"])
    assert abs(lat.norm().item() - ref_norm) <= rtol * ref_norm, f"{what}: latent grad norm"
    err = (lat[:, :16] - T(gold["latent.slice"]).double()).abs().max().item()
    assert err <= rtol * 30 * ref_norm / lat.numel() ** 0.5 + 1e-9, f"{what}: latent grad slice, max err {err:.3e}"


def golden_case():
    gold = np.load(GOLD)
    scene = H.make_scene_dict(num_objs=2, num_views=3, feat=16, size=128, seed=5)
    allr = torch.cat([synth.target_rays(128, 15.0 + 20 * s, -10.0) for s in range(2)])
    rays = allr[:, T(gold["ray_idx"]).long()].contiguous()
    noise = H.make_noise(2 * 24, seed=9)
    return gold, scene, rays, noise, T(gold["gt"])


def test_oracle_gradients_match_reference_autograd():
    gold, scene, rays, noise, gt = golden_case()
    loss, res, grads = H.oracle_train_step(scene, rays, noise, gt)
    np.testing.assert_allclose(res["fine"]["rgb"].detach().numpy(), gold["fine_rgb"], atol=3e-5, rtol=1e-4)
    check_against_golden(grads, loss, gold, 2e-4, "oracle")
    # the depth-sample branch really carries gradient into the coarse network (nerf.py:296-298)
    assert float(gold["coarse.lin_in.weight.norm"]) > 0
