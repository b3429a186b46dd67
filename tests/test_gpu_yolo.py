"""GPU parity of the YOLO mode through the C-ABI: the fused field kernel with d_out = 21,
the z >= 0 latent mask and raw outputs, and YoloRenderer's per-ray reduction, vs the unmodified reference's outputs
(tests/golden/reference_yolo.npz) and the oracle.  Tolerances as for the NeRF path: 1e-4 fp32 check path, bf16
tensor-core path 2e-2 on the raw (unbounded) head values / 1e-2 on the reduced probabilities."""
import copy

import numpy as np
import pytest
import torch

import helpers as H
import pixel_nerf_yolo_b200.synth as synth
from oracle import pixelnerf_oracle as O
from test_oracle_yolo import yolo_case

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def build_yolo_net(scene, w2c, precision, C=512, train=False):
    from pixel_nerf_yolo_b200.conf import ConfigTree
    from pixel_nerf_yolo_b200.model import make_model
    conf = copy.deepcopy(H.MODEL_CONF)
    conf["mlp_coarse"].update({"d_out": 7, "num_scales": 1, "num_anchors_per_scale": 3, "yolo": True})
    conf["mlp_fine"] = {"type": "empty"}
    if C != 512:
        conf["encoder"] = {"backbone": "custom", "pretrained": False, "num_layers": 4, "index_padding": "zeros"}
    net = make_model(ConfigTree.from_dict(conf)).eval()
    assert net.yolo and net.d_out == 21 and net.mlp_fine is None
    net.mlp_coarse.load_state_dict(synth.mlp_state(31, d_out=21, d_latent=C))
    net = net.cuda()
    net.num_objs, net.num_views_per_obj = 1, w2c.shape[0]
    net.encoder.set_latent(scene["latent"].cuda())
    net.set_cameras(w2c.cuda(), scene["focal"].cuda(), scene["image_wh"])
    net.precision = precision
    if train:
        net.train()
    else:
        net.requires_grad_(False)        # an inference build (see helpers.build_net)
    return net


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_yolo_field_matches_reference_golden(precision, tol):
    g, scene, sc, mlp, rays = yolo_case()
    net = build_yolo_net(scene, T(g["w2c"]), precision)
    with torch.no_grad():
        out = net(T(g["field_xyz"]).cuda(), coarse=True, viewdirs=T(g["field_dirs"]).cuda())
    assert out.shape == (1, 23, 21)
    ref = g["field_out"]
    err = np.abs(out.cpu().numpy() - ref).max()
    # raw head values are unbounded linear outputs (no sigmoid / relu): the bound is stated relative to their scale
    assert err <= tol * max(1.0, np.abs(ref).max()), f"{precision}: max err {err:.3e}, |ref| max {np.abs(ref).max():.3f}"


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_yolo_render_matches_reference_golden(precision, tol):
    from pixel_nerf_yolo_b200.render import YoloRenderer
    g, scene, sc, mlp, rays = yolo_case()
    net = build_yolo_net(scene, T(g["w2c"]), precision)
    r = YoloRenderer(128, 1024, 1, 3)
    r.bind_parallel(net)
    r.noise_override = T(g["noise"]).cuda()
    res = r(rays.cuda())
    assert res.shape == (20, 3, 7)
    np.testing.assert_allclose(res.cpu().numpy()[..., 0], g["render"][..., 0], atol=tol, rtol=0)       # probabilities
    np.testing.assert_allclose(res.cpu().numpy()[..., 1:], g["render"][..., 1:], atol=2 * tol, rtol=0)  # box values


def test_yolo_render_yolo_sized_maps_matches_oracle():
    """conf/exp/yolo.conf shape: 1792-channel maps (custom backbone), 128 samples per ray, 3 anchors."""
    from pixel_nerf_yolo_b200.conf import ConfigTree
    from pixel_nerf_yolo_b200.render import make_renderer
    scene = H.make_scene_dict(num_objs=1, num_views=3, feat=20, size=128, seed=8, C=1792)
    w2c = torch.linalg.inv(scene["poses"][0])
    sc = O.encode_cameras(scene["latent"], w2c, scene["focal"], scene["image_wh"], num_views=3, yolo=True)
    mlp = synth.mlp_state(31, d_out=21, d_latent=1792)
    rays = H.rays_subset(1, 300, seed=2)[0]
    noise = torch.rand(300, 128, generator=torch.Generator().manual_seed(3))
    ref = O.yolo_render(sc, mlp, rays, noise)
    net = build_yolo_net(scene, w2c, "bf16", C=1792)
    r = make_renderer(ConfigTree.from_dict({"renderer": {"type": "yolo", "n_coarse": 128, "eval_batch_size": 128},
                                            "model": {"mlp_coarse": {"num_scales": 1, "num_anchors_per_scale": 3}}}))
    r.bind_parallel(net)
    r.noise_override = noise.cuda()
    res = r(rays.cuda()).cpu()
    np.testing.assert_allclose(res.numpy()[..., 0], ref.numpy()[..., 0], atol=1e-2, rtol=0)
    np.testing.assert_allclose(res.numpy()[..., 1:], ref.numpy()[..., 1:], atol=3e-2, rtol=0)
    assert r(rays[:0].cuda()).shape == (0, 3, 7)


@pytest.mark.parametrize("train_precision,rtol", [("fp32", 1e-3), ("tf32", 2e-2), ("bf16", 2e-2)])
def test_yolo_train_step_matches_reference_autograd(train_precision, rtol):
    """The fork's product is a TRAINED detector: YoloRenderer.forward -> loss.backward() (train/trainlib/YoloTrainer.py:140-190)
    through the raw 21-wide head, the z >= 0 latent mask and the per-ray reduction, vs the gradients the unmodified reference's
    autograd produced (tests/golden/make_golden_yolo_train.py): fp32 SIMT path 1e-3, tensor-core paths 2e-2 of each tensor's scale."""
    from test_oracle_yolo import check_yolo_grads, yolo_train_case
    from pixel_nerf_yolo_b200.render import YoloRenderer
    g, scene, rays = yolo_train_case()
    net = build_yolo_net(scene, T(g["w2c"]), "bf16", train=True)
    net.train_precision = train_precision
    lat = scene["latent"].cuda().clone().requires_grad_(True)
    net.encoder.set_latent(lat)
    r = YoloRenderer(128, 1024, 1, 3)
    r.bind_parallel(net)
    r.noise_override = T(g["noise"]).cuda()
    res = r(rays.cuda())
    assert res.requires_grad and res.shape == (20, 3, 7)
    loss = (res * T(g["gw"]).cuda()).sum()
    loss.backward()
    grads = {n: p.grad for n, p in net.mlp_coarse.named_parameters()}
    # element-wise, the sparse encoder-output gradient of the bf16-operand forward differs from the fp32 forward's by ReLU mask
    # flips near zero (see test_field_backward_bf16_matches_autograd); its norm is held to rtol inside check_yolo_grads
    check_yolo_grads(g, grads, lat.grad, loss.item(), res.detach().cpu().numpy(), rtol, f"cuda {train_precision}",
                     lat_rtol=1e-1 if train_precision == "bf16" else None)
    # inference on the same (training-mode) network under no_grad takes the fused tcgen05 kernel and agrees
    with torch.no_grad():
        inf = r(rays.cuda())
    assert not inf.requires_grad
    np.testing.assert_allclose(inf.cpu().numpy(), g["render"], atol=3e-2, rtol=0)


def test_yolo_reduce_backward_matches_autograd():
    """pnr_yolo_reduce_backward vs torch autograd of the reduction itself (yolo.py:96-114), incl. ties in the max."""
    from pixel_nerf_yolo_b200.render.yolo import _YoloReduceFn
    B, K, A = 37, 128, 3
    gen = torch.Generator().manual_seed(5)
    out = torch.randn(B, K, A * 7, generator=gen)
    out[3, 10, 0] = out[3, 20, 0] = 9.0                      # a tie: torch.max sends the gradient to the first index
    gw = torch.randn(B, A, 7, generator=gen)
    o = out.clone().requires_grad_(True)
    v = o.reshape(B, K, A, 7)
    p = torch.sigmoid(v[..., 0])
    ref = torch.cat([p.max(dim=1)[0].unsqueeze(-1), (v[..., 1:] * p.unsqueeze(-1)).sum(1) / (p.sum(1).unsqueeze(-1) + 1e-5)], dim=-1)
    (ref * gw).sum().backward()
    oc = out.cuda().requires_grad_(True)
    res = _YoloReduceFn.apply(oc, B, K, A)
    (res * gw.cuda()).sum().backward()
    np.testing.assert_allclose(res.detach().cpu().numpy(), ref.detach().numpy(), atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(oc.grad.cpu().numpy(), o.grad.numpy(), atol=1e-6, rtol=1e-4)
