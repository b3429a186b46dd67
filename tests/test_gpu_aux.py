"""GPU parity of the kernels either side of the hot path, through the C-ABI:
pnr_gen_rays vs util.gen_rays goldens / oracle, pnr_pyramid_pack (fused upsample + concat + channels-last) vs the
oracle and the reference's own encoder output."""
import os

import numpy as np
import pytest
import torch

import helpers as H
import pixel_nerf_yolo_b200.synth as synth
from oracle import pixelnerf_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_aux.npz")
T = torch.from_numpy


def test_gen_rays_matches_reference_golden():
    from pixel_nerf_yolo_b200.util import gen_rays
    g = np.load(GOLD)
    poses = T(g["rays_poses"]).cuda()
    r = gen_rays(poses, 40, 24, torch.tensor(41.5), 0.8, 1.8)
    np.testing.assert_allclose(r.cpu().numpy(), g["rays_default_c"], atol=1e-6, rtol=0)
    r = gen_rays(poses, 40, 24, torch.tensor([41.5, 39.0]), 0.5, 2.5, c=torch.tensor([17.25, 13.5]))
    np.testing.assert_allclose(r.cpu().numpy(), g["rays_with_c"], atol=1e-6, rtol=0)
    # origins, near and far are copies: bit-exact
    assert torch.equal(r.cpu()[..., :3], T(g["rays_with_c"])[..., :3]) and torch.equal(r.cpu()[..., 6:], T(g["rays_with_c"])[..., 6:])


@pytest.mark.parametrize("size", [128, 640])
def test_gen_rays_full_image_and_selection(size):
    from pixel_nerf_yolo_b200.util import gen_rays
    poses = torch.stack([synth.pose_spherical(t, -10.0, 1.3) for t in (15.0, 100.0, -60.0)])
    focal = 131.25 * size / 128
    ref = O.gen_rays(poses, size, size, focal, 0.8, 1.8)
    r = gen_rays(poses.cuda(), size, size, focal, 0.8, 1.8)
    np.testing.assert_allclose(r.cpu().numpy(), ref.numpy(), atol=1e-6, rtol=0)
    assert abs(float(r[..., 3:6].norm(dim=-1).mean()) - 1.0) < 1e-6
    pick = torch.randint(0, 3 * size * size, (777,), generator=torch.Generator().manual_seed(1))
    sel = gen_rays(poses.cuda(), size, size, focal, 0.8, 1.8, pix_inds=pick)
    assert torch.equal(sel.cpu(), r.cpu().reshape(-1, 8)[pick])          # the trainer's ray sampling: same bits
    assert gen_rays(poses.cuda(), size, size, focal, 0.8, 1.8, pix_inds=pick[:0]).shape == (0, 8)


def test_pyramid_pack_matches_reference_golden():
    """Levels captured from the reference's resnet34 forward -> fused kernel == the reference's latent."""
    from pixel_nerf_yolo_b200.model.encoder import SpatialEncoder
    g = np.load(GOLD)
    enc = SpatialEncoder("resnet34", pretrained=False, num_layers=4, index_padding="zeros").eval().cuda()
    enc.set_levels([T(g[f"pyr_level{i}"]).cuda() for i in range(4)])
    np.testing.assert_allclose(enc.latent_scaling.cpu().numpy(), g["pyr_latent_scaling"], rtol=1e-7)
    nhwc = enc.packed_latent(fp32=True)                                   # (1, 32, 32, 512) fp32 channels-last
    got = nhwc.permute(0, 3, 1, 2)[:, ::4].cpu().numpy()
    np.testing.assert_allclose(got, g["pyr_latent_sub"], atol=2e-6, rtol=1e-6)
    bf = enc.packed_latent(fp32=False)
    assert bf.dtype == torch.bfloat16
    assert torch.equal(bf, nhwc.to(torch.bfloat16))                       # one rounding, of the same fp32 value
    # the lazily materialised reference attribute agrees too
    np.testing.assert_allclose(enc.latent[:, ::4].cpu().numpy(), g["pyr_latent_sub"], atol=2e-6, rtol=1e-6)


@pytest.mark.parametrize("n,h,w", [(3, 64, 64), (2, 33, 47), (1, 1, 5)])
def test_pyramid_pack_matches_oracle(n, h, w):
    """Ragged sizes (non power-of-two, odd channel counts, 1-pixel levels)."""
    from pixel_nerf_yolo_b200.model.encoder import SpatialEncoder
    gen = torch.Generator().manual_seed(h)
    dims = [(40, h, w), (24, max(h // 2, 1), max(w // 2, 1)), (72, max(h // 4, 1), max(w // 3, 1)), (8, 1, 1)]
    levels = [torch.randn(n, c, hh, ww, generator=gen) for c, hh, ww in dims]
    ref = O.pyramid_latent(levels)
    enc = SpatialEncoder("resnet34", pretrained=False, num_layers=4, index_padding="zeros").eval().cuda()
    enc.set_levels([l.cuda() for l in levels])
    got = enc.packed_latent(fp32=True).permute(0, 3, 1, 2).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-6, rtol=1e-6)


def test_encode_uses_fused_pyramid_and_render_agrees():
    """PixelNeRFNet.encode in inference installs levels (no fp32 NCHW latent); rendering from them equals rendering
    from the materialised latent bit for bit."""
    scene = H.make_scene_dict(num_objs=1, num_views=3)
    net = H.build_net(scene)
    img = (torch.rand(1, 3, 3, 128, 128, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
    with torch.no_grad():
        net.encode(img, scene["poses"].cuda(), scene["focal"].cuda())
    assert net.encoder._latent is None and net.encoder._levels is not None
    assert net.encoder.latent_shape() == (3, 512, 64, 64)
    a = net.encoder.packed_latent(fp32=False).clone()
    lat = net.encoder.latent                                  # materialise (torch interpolate + cat)
    net.encoder.set_latent(lat)
    b = net.encoder.packed_latent(fp32=False)
    diff = (a.float() - b.float()).abs()
    # same fp32 value up to the interpolation's rounding, then one bf16 rounding: at most 1 bf16 ulp apart
    assert (diff <= 2.0 ** -7 * b.float().abs().clamp_min(1e-3)).all()
    assert (a == b).float().mean() > 0.99


def test_image_output_matches_eval_script_arithmetic():
    """eval/eval.py:283-290 restated with the same torch/numpy calls."""
    from pixel_nerf_yolo_b200.util import image_output
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(5000, 3, generator=g) * 1.4 - 0.2            # some values outside [0, 1]
    depth = 0.8 + torch.rand(5000, generator=g)
    u8, dn = image_output(rgb.cuda(), depth.cuda(), 0.8, 1.8)
    ref_u8 = (torch.clamp(rgb, 0.0, 1.0).numpy() * 255).astype(np.uint8)
    ref_d = ((depth - 0.8) / (1.8 - 0.8)).numpy()
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    np.testing.assert_allclose(dn.cpu().numpy(), ref_d, atol=1e-6, rtol=0)


@pytest.mark.parametrize("use_l1", [False, True])
def test_rgb_loss_matches_torch_loss_and_grad(use_l1):
    """src/model/loss.py:92-104: MSELoss / L1Loss (the reference's own criterion objects) forward and backward."""
    from pixel_nerf_yolo_b200.conf import ConfigTree
    from pixel_nerf_yolo_b200.model.loss import get_rgb_loss
    g = torch.Generator().manual_seed(1)
    rgb, gt = torch.rand(4, 128, 3, generator=g), torch.rand(4, 128, 3, generator=g)
    a = rgb.clone().requires_grad_(True)
    crit = torch.nn.L1Loss() if use_l1 else torch.nn.MSELoss()
    ref = crit(a, gt)
    (ref * 1.7).backward()
    b = rgb.cuda().requires_grad_(True)
    ours = get_rgb_loss(ConfigTree.from_dict({"use_l1": use_l1}))(b, gt.cuda())
    (ours * 1.7).backward()
    assert abs(ours.item() - ref.item()) < 1e-6
    np.testing.assert_allclose(b.grad.cpu().numpy(), a.grad.numpy(), atol=1e-8, rtol=1e-6)
