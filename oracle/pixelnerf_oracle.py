"""CPU oracle for the pixelNeRF-YOLO rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU baseline, never as the thing that is shipped.

It is a functional restatement, in fp32 torch-CPU arithmetic, of what the
reference computes on the path

    NeRFRenderer.forward            src/render/nerf.py:257-309
      sample_coarse                 src/render/nerf.py:104-124
      composite                     src/render/nerf.py:169-255
      sample_fine                   src/render/nerf.py:126-154
      sample_fine_depth             src/render/nerf.py:156-167
    PixelNeRFNet.encode (cameras)   src/model/models.py:92-151
    PixelNeRFNet.forward            src/model/models.py:153-318
      PositionalEncoding.forward    src/model/code.py:30-42
      SpatialEncoder.index          src/model/encoder.py:79-108
      ResnetFC.forward              src/model/resnetfc.py:134-186
      ResnetBlockFC.forward         src/model/resnetfc.py:53-62
      util.combine_interleaved      src/util/util.py:489-499
      util.repeat_interleave        src/util/util.py:60-67

The arithmetic itself lives in PyTorch ATen (torch is unpinned by the reference,
this image pins 2.11.0), so the oracle uses the same ATen ops where the exact
rounding matters (``cumsum``, ``cumprod``, ``searchsorted``, ``sort``, ``sin``,
``exp``) and restates ``F.grid_sample`` as an explicit 4-tap bilinear blend.

Parity pin: the reference's own tests hold no assertions or golden vectors for
this path, so the oracle is pinned against outputs of the
reference itself, executed in the build container by
``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``
(checked by ``tests/test_oracle_golden.py``).

Random numbers: the reference draws ``rand_like(B,Kc)``, ``rand(B,Kf-Kfd)``,
``rand_like(B,Kf-Kfd)``, ``randn_like(B,Kfd)`` in this order inside one
``NeRFRenderer.forward``.  The oracle takes those four tensors as explicit
inputs (``RenderNoise``) so that it and the CUDA path see identical noise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch


# --------------------------------------------------------------------------- #
# containers
# --------------------------------------------------------------------------- #
@dataclass
class RenderNoise:
    """The four random tensors one NeRFRenderer.forward consumes (nerf.py:117,141,147,164)."""

    coarse: torch.Tensor            # (B, Kc)      U[0,1)   stratified jitter
    fine_u: Optional[torch.Tensor]  # (B, Kf-Kfd)  U[0,1)   inverse-CDF draw
    fine_jitter: Optional[torch.Tensor]  # (B, Kf-Kfd) U[0,1) in-bin jitter
    depth: Optional[torch.Tensor]   # (B, Kfd)     N(0,1)   depth-sample noise

    @staticmethod
    def draw(B: int, n_coarse: int, n_fine: int, n_fine_depth: int, generator=None,
             device="cpu") -> "RenderNoise":
        g = generator
        kf = n_fine - n_fine_depth
        c = torch.rand(B, n_coarse, generator=g, device=device)
        u = torch.rand(B, kf, generator=g, device=device) if kf > 0 else None
        j = torch.rand(B, kf, generator=g, device=device) if kf > 0 else None
        d = torch.randn(B, n_fine_depth, generator=g, device=device) if n_fine_depth > 0 else None
        return RenderNoise(c, u, j, d)

    @staticmethod
    def fixed(B: int, n_coarse: int, n_fine: int, n_fine_depth: int, device="cpu") -> "RenderNoise":
        """The 'deterministic (non-perturbed) sampler': mid-bin jitter, zero depth noise."""
        kf = n_fine - n_fine_depth
        c = torch.full((B, n_coarse), 0.5, device=device)
        u = ((torch.arange(kf, device=device, dtype=torch.float32) + 0.5) / max(kf, 1)).expand(B, kf).contiguous() if kf > 0 else None
        j = torch.full((B, kf), 0.5, device=device) if kf > 0 else None
        d = torch.zeros(B, n_fine_depth, device=device) if n_fine_depth > 0 else None
        return RenderNoise(c, u, j, d)


@dataclass
class Scene:
    """What PixelNeRFNet.encode leaves behind (models.py:116-148, encoder.py:170-172)."""

    latent: torch.Tensor          # (SB*NS, C, Hl, Wl) fp32, NCHW
    poses: torch.Tensor           # (SB*NS, 3, 4) world -> camera
    focal: torch.Tensor           # (1|SB*NS... , 2) with fy negated
    c: torch.Tensor               # (1|.., 2)
    image_shape: torch.Tensor     # (2,) = [W, H]
    latent_scaling: torch.Tensor  # (2,)
    num_views: int                # NS


# --------------------------------------------------------------------------- #
# PixelNeRFNet.encode: camera bookkeeping only (the CNN is out of scope)
# --------------------------------------------------------------------------- #
def encode_cameras(latent: torch.Tensor, poses_c2w: torch.Tensor, focal: torch.Tensor,
                   image_wh: Tuple[int, int], c: Optional[torch.Tensor] = None,
                   num_views: Optional[int] = None, yolo: bool = False) -> Scene:
    """models.py:92-151 minus ``self.encoder(images)``; ``latent`` is the encoder output.
    ``yolo``: poses are used as given and fy keeps its sign (models.py:119-120, 136-137)."""
    if poses_c2w.dim() == 4:  # (SB, NS, 4, 4)
        num_views = poses_c2w.shape[1]
        poses_c2w = poses_c2w.reshape(-1, 4, 4)
    elif num_views is None:
        num_views = 1
    rot = poses_c2w[:, :3, :3].transpose(1, 2)              # models.py:116
    trans = -torch.bmm(rot, poses_c2w[:, :3, 3:])           # models.py:117
    w2c = torch.cat((rot, trans), dim=-1)                   # models.py:118
    if yolo:
        w2c = poses_c2w[:, :3, :4]                          # models.py:120
    image_shape = torch.tensor([float(image_wh[0]), float(image_wh[1])])  # models.py:122-123
    focal = torch.as_tensor(focal, dtype=torch.float32)
    if focal.dim() == 0:                                    # models.py:126-134
        focal = focal[None, None].repeat((1, 2))
    elif focal.dim() == 1:
        focal = focal.unsqueeze(-1).repeat((1, 2))
    else:
        focal = focal.clone()
    focal = focal.float()
    if not yolo:
        focal[..., 1] *= -1.0                               # models.py:137
    if c is None:                                           # models.py:139-148
        c = (image_shape * 0.5).unsqueeze(0)
    else:
        c = torch.as_tensor(c, dtype=torch.float32)
        if c.dim() == 0:
            c = c[None, None].repeat((1, 2))
        elif c.dim() == 1:
            c = c.unsqueeze(-1).repeat((1, 2))
    Hl, Wl = latent.shape[-2:]
    ls = torch.tensor([float(Wl), float(Hl)])
    latent_scaling = ls / (ls - 1) * 2.0                    # encoder.py:170-172
    return Scene(latent.float(), w2c.float(), focal, c.float(), image_shape, latent_scaling, num_views)


# --------------------------------------------------------------------------- #
# ops
# --------------------------------------------------------------------------- #
def _f64(fn, x: torch.Tensor) -> torch.Tensor:
    """Transcendental functions (sin, exp, sigmoid) are evaluated in double and rounded once to fp32.
    ATen's fp32 CPU kernels are vectorised libm calls whose accuracy depends on the host's SIMD level (one GPU
    box's host returned exp() values ~100 ulp off); the correctly rounded value is host-independent and within
    1 ulp of what the reference's torch.sin / torch.exp / torch.sigmoid return on a healthy host."""
    return fn(x.double()).to(x.dtype) if x.dtype == torch.float32 else fn(x)


def positional_encoding(x: torch.Tensor, num_freqs: int = 6, freq_factor: float = 1.5,
                        include_input: bool = True) -> torch.Tensor:
    """code.py:11-42.  Layout: [x, sin(f0 x), cos(f0 x), sin(f1 x), ...] with cos as sin(.+pi/2)."""
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)                 # code.py:15
    f2 = torch.repeat_interleave(freqs, 2).view(1, -1, 1).float()           # code.py:21-23
    ph = torch.zeros(2 * num_freqs)
    ph[1::2] = math.pi * 0.5                                                # code.py:26-27
    ph = ph.view(1, -1, 1)
    e = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)
    e = _f64(torch.sin, torch.addcmul(ph, e, f2))                           # code.py:38
    e = e.view(x.shape[0], -1)
    if include_input:
        e = torch.cat((x, e), dim=-1)
    return e


def bilinear_index(latent: torch.Tensor, uv: torch.Tensor, latent_scaling: torch.Tensor,
                   image_shape: torch.Tensor, padding: str = "zeros") -> torch.Tensor:
    """SpatialEncoder.index (encoder.py:79-108) with F.grid_sample(bilinear, align_corners=True)
    restated as an explicit 4-tap blend (ATen GridSampler.h: grid_sampler_unnormalize /
    within_bounds_2d).  latent (V,C,Hl,Wl); uv (V,P,2) pixels -> (V,C,P)."""
    V, C, Hl, Wl = latent.shape
    scale = latent_scaling / image_shape                    # encoder.py:97
    g = uv * scale - 1.0                                    # encoder.py:98
    # align_corners=True unnormalise: ((g + 1) / 2) * (size - 1)
    ix = ((g[..., 0] + 1) / 2) * (Wl - 1)
    iy = ((g[..., 1] + 1) / 2) * (Hl - 1)
    if padding == "border":
        ix = ix.clamp(0, Wl - 1)
        iy = iy.clamp(0, Hl - 1)
    elif padding != "zeros":
        raise NotImplementedError(padding)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = latent.reshape(V, C, Hl * Wl)
    out = torch.zeros(V, C, uv.shape[1], dtype=latent.dtype)

    def tap(xx, yy, w):
        ok = (xx >= 0) & (xx <= Wl - 1) & (yy >= 0) & (yy <= Hl - 1)
        lin = (yy.clamp(0, Hl - 1) * Wl + xx.clamp(0, Wl - 1)).long()
        vals = torch.gather(flat, 2, lin.unsqueeze(1).expand(-1, C, -1))
        return vals * (w * ok).unsqueeze(1)

    out = tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)
    return out


def _bf16_ste(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 (round-to-nearest-even) with a straight-through gradient."""
    return t + (t.bfloat16().float() - t).detach()


def resnetfc_forward(p: Dict[str, torch.Tensor], zx: torch.Tensor, d_latent: int,
                     n_blocks: int, combine_layer: int, inner_dims: Tuple[int, int], emulate_bf16: bool = False) -> torch.Tensor:
    """ResnetFC.forward (resnetfc.py:134-186), ReLU activations, 'average' combine.
    ``p`` holds the module's state_dict entries (lin_in.weight, blocks.0.fc_0.weight, ...).
    ``emulate_bf16`` (NOT reference behaviour; a test instrument for the tensor-core paths): every Linear of the trunk sees
    its input and weight rounded to bf16 and accumulates in fp32 -- the arithmetic the tcgen05 paths define (fp32 residual
    stream, fp32 lin_out).  Gradients flow straight through the roundings, so autograd gives the gradient of THAT forward:
    the same ReLU masks as the CUDA path, which a comparison against the fp32 forward does not have for pre-activations
    within the forward tolerance of zero."""
    if emulate_bf16:
        def lin(x, w, b, trunk=True):
            return torch.nn.functional.linear(_bf16_ste(x), _bf16_ste(w), b) if trunk else torch.nn.functional.linear(x, w, b)
    else:
        def lin(x, w, b, trunk=True):
            return torch.nn.functional.linear(x, w, b)
    z = zx[..., :d_latent]
    x = zx[..., d_latent:]
    x = lin(x, p["lin_in.weight"], p["lin_in.bias"])                        # resnetfc.py:149
    for b in range(n_blocks):
        if b == combine_layer:                                              # resnetfc.py:154,172-174
            ns, pts = inner_dims
            x = x.reshape(-1, ns, pts, x.shape[-1]).mean(dim=1)             # util.py:489-499
            x = x.reshape(-1, x.shape[-1])
        if d_latent > 0 and b < combine_layer:
            x = x + lin(z, p[f"lin_z.{b}.weight"], p[f"lin_z.{b}.bias"])    # resnetfc.py:176-182
        net = lin(torch.relu(x), p[f"blocks.{b}.fc_0.weight"], p[f"blocks.{b}.fc_0.bias"])
        dx = lin(torch.relu(net), p[f"blocks.{b}.fc_1.weight"], p[f"blocks.{b}.fc_1.bias"])
        x = x + dx                                                          # resnetfc.py:53-62
    xr = _bf16_ste(torch.relu(x)) if emulate_bf16 else torch.relu(x)        # the training path's lin_out reads the bf16 operand
    return lin(xr, p["lin_out.weight"], p["lin_out.bias"], trunk=False)     # resnetfc.py:185


def field_forward(scene: Scene, mlp: Dict[str, torch.Tensor], xyz: torch.Tensor,
                  viewdirs: torch.Tensor, *, num_freqs: int = 6, freq_factor: float = 1.5,
                  n_blocks: int = 5, combine_layer: int = 3, padding: str = "zeros",
                  return_raw: bool = False, yolo: bool = False, emulate_bf16: bool = False) -> torch.Tensor:
    """PixelNeRFNet.forward (models.py:153-318) for the default_mv.conf switches
    (use_xyz, normalize_z, use_code, not use_code_viewdirs, use_viewdirs, no global encoder).
    xyz, viewdirs (SB, P, 3) -> (SB, P, 4) = [sigmoid rgb, relu sigma]."""
    SB, P, _ = xyz.shape
    NS = scene.num_views
    R = scene.poses[:, None, :3, :3]
    rep = lambda t: t.unsqueeze(1).expand(-1, NS, *t.shape[1:]).reshape(-1, *t.shape[1:])  # util.py:60-67
    x_w = rep(xyz)                                                          # (SB*NS, P, 3)
    x_rot = torch.matmul(R, x_w.unsqueeze(-1))[..., 0]                      # models.py:169-171
    x_cam = x_rot + scene.poses[:, None, :3, 3]                             # models.py:172
    zf = positional_encoding(x_rot.reshape(-1, 3), num_freqs, freq_factor)  # models.py:183-195
    vd = rep(viewdirs.reshape(SB, P, 3, 1))
    vd = torch.matmul(R, vd).reshape(-1, 3)                                 # models.py:201-206
    zf = torch.cat((zf, vd), dim=1)                                         # models.py:207-209
    if not yolo:
        uv = -x_cam[:, :, :2] / x_cam[:, :, 2:]                             # models.py:220
    else:
        uv = x_cam[:, :, :2] / x_cam[:, :, 2:]                              # models.py:222
    foc = scene.focal.unsqueeze(1)
    cc = scene.c.unsqueeze(1)
    if foc.shape[0] > 1:                                                    # models.py:225-227
        foc = rep(foc)
    if cc.shape[0] > 1:                                                     # models.py:228-230
        cc = rep(cc)
    uv = uv * foc + cc                                                      # models.py:225-230
    lat = bilinear_index(scene.latent, uv, scene.latent_scaling, scene.image_shape, padding)
    C = lat.shape[1]
    lat = lat.transpose(1, 2).reshape(-1, C)                                # models.py:244-246
    if yolo:                                                                # models.py:223, 254-264
        nonneg = (x_cam[:, :, 2:] >= 0).reshape(-1, 1)
        lat = torch.where(nonneg | lat.isnan(), torch.zeros_like(lat), lat)
    mlp_in = torch.cat((lat, zf), dim=-1)                                   # models.py:276
    out = resnetfc_forward(mlp, mlp_in, C, n_blocks, combine_layer, (NS, P), emulate_bf16=emulate_bf16)
    out = out.reshape(-1, P, out.shape[-1])
    if return_raw or yolo:                                                  # models.py:309-310
        return out.reshape(SB, P, -1)
    rgb = _f64(torch.sigmoid, out[..., :3])                                 # models.py:312-317
    sigma = torch.relu(out[..., 3:4])
    return torch.cat((rgb, sigma), dim=-1).reshape(SB, P, -1)


def sample_coarse(rays: torch.Tensor, noise: torch.Tensor, n_coarse: int, lindisp: bool = False):
    """nerf.py:104-124.  rays (B,8), noise (B,Kc) -> z (B,Kc)."""
    near, far = rays[:, -2:-1], rays[:, -1:]
    step = 1.0 / n_coarse
    zs = torch.linspace(0, 1 - step, n_coarse).unsqueeze(0).repeat(rays.shape[0], 1)
    zs = zs + noise * step
    if not lindisp:
        return near * (1 - zs) + far * zs
    return 1 / (1 / near * (1 - zs) + 1 / far * zs)


def sample_fine(rays: torch.Tensor, weights: torch.Tensor, u: torch.Tensor, jitter: torch.Tensor,
                n_coarse: int, lindisp: bool = False, return_inds: bool = False):
    """nerf.py:126-154.  weights (B,Kc), u/jitter (B,Kf-Kfd) -> z (B,Kf-Kfd)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True).float() - 1.0
    inds = torch.clamp_min(inds, 0.0)
    zs = (inds + jitter) / n_coarse
    near, far = rays[:, -2:-1], rays[:, -1:]
    if not lindisp:
        z = near * (1 - zs) + far * zs
    else:
        z = 1 / (1 / near * (1 - zs) + 1 / far * zs)
    return (z, inds.long()) if return_inds else z


def sample_fine_depth(rays: torch.Tensor, depth: torch.Tensor, gauss: torch.Tensor, depth_std: float):
    """nerf.py:156-167."""
    z = depth.unsqueeze(1).repeat((1, gauss.shape[1]))
    z = z + gauss * depth_std
    return torch.max(torch.min(z, rays[:, -1:]), rays[:, -2:-1])


def ray_points(rays: torch.Tensor, z: torch.Tensor):
    """nerf.py:191,208: sample positions and per-sample view directions."""
    pts = rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]
    dirs = rays[:, None, 3:6].expand(-1, z.shape[1], -1)
    return pts, dirs


def alpha_composite(out: torch.Tensor, z: torch.Tensor, rays: torch.Tensor, white_bkgd: bool):
    """nerf.py:184-188, 229-255.  out (B,K,4) = [rgb, sigma] -> weights (B,K), rgb (B,3), depth (B)."""
    deltas = z[:, 1:] - z[:, :-1]
    deltas = torch.cat([deltas, rays[:, -1:] - z[:, -1:]], -1)
    rgbs, sig = out[..., :3], out[..., 3]
    alphas = 1 - _f64(torch.exp, -deltas * torch.relu(sig))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    T = torch.cumprod(shifted, -1)
    w = alphas * T[:, :-1]
    rgb = torch.sum(w.unsqueeze(-1) * rgbs, -2)
    depth = torch.sum(w * z, -1)
    if white_bkgd:
        rgb = rgb + 1 - w.sum(dim=1).unsqueeze(-1)
    return w, rgb, depth


def _composite(scene, mlp, rays, z, sb, white_bkgd, field_kw):
    B, K = z.shape
    pts, dirs = ray_points(rays, z)
    pts = pts.reshape(sb, -1, 3)                                            # nerf.py:197-199
    dirs = dirs.reshape(sb, -1, 3)
    out = field_forward(scene, mlp, pts, dirs, **field_kw).reshape(B, K, -1)
    return alpha_composite(out, z, rays, white_bkgd)


def render(scene: Scene, mlp_coarse: Dict[str, torch.Tensor], mlp_fine: Optional[Dict[str, torch.Tensor]],
           rays: torch.Tensor, noise: RenderNoise, *, n_coarse: int = 64, n_fine: int = 32,
           n_fine_depth: int = 16, depth_std: float = 0.01, white_bkgd: bool = True,
           lindisp: bool = False, grad: bool = False, **field_kw):
    """NeRFRenderer.forward (nerf.py:257-309) with explicit noise.  rays (SB,B',8).
    Returns dict(coarse=dict(rgb,depth,weights,z), fine=...).  ``grad=True`` records the autograd graph exactly
    as the reference's training step does (importance sampler on detached weights nerf.py:136,293; the depth
    samples stay attached to the coarse depth nerf.py:296-298), so ``loss.backward()`` on the result gives the
    gradients the CUDA backward kernels are checked against."""
    assert rays.dim() == 3
    SB = rays.shape[0]
    r = rays.reshape(-1, 8)
    with torch.set_grad_enabled(grad):
        z_c = sample_coarse(r, noise.coarse, n_coarse, lindisp)
        wc, rgbc, dc = _composite(scene, mlp_coarse, r, z_c, SB, white_bkgd, field_kw)
        res = {"coarse": dict(rgb=rgbc.reshape(SB, -1, 3), depth=dc.reshape(SB, -1),
                              weights=wc.reshape(SB, -1, wc.shape[-1]), z=z_c)}
        if n_fine > 0:
            parts = [z_c]
            if n_fine - n_fine_depth > 0:
                parts.append(sample_fine(r, wc.detach(), noise.fine_u, noise.fine_jitter, n_coarse, lindisp))
            if n_fine_depth > 0:
                parts.append(sample_fine_depth(r, dc, noise.depth, depth_std))
            z_all, _ = torch.sort(torch.cat(parts, dim=-1), dim=-1)         # nerf.py:300-301
            mf = mlp_fine if mlp_fine is not None else mlp_coarse           # models.py:291
            wf, rgbf, df = _composite(scene, mf, r, z_all, SB, white_bkgd, field_kw)
            res["fine"] = dict(rgb=rgbf.reshape(SB, -1, 3), depth=df.reshape(SB, -1),
                               weights=wf.reshape(SB, -1, wf.shape[-1]), z=z_all)
    return res


# --------------------------------------------------------------------------- #
# the steps either side of the path
# --------------------------------------------------------------------------- #
def pyramid_latent(levels, upsample_interp: str = "bilinear") -> torch.Tensor:
    """SpatialEncoder.forward's tail (encoder.py:159-168): upsample every level to the first level's size
    (align_corners=True) and concatenate along channels -> latent (N, sum C_l, H_0, W_0)."""
    size = levels[0].shape[-2:]
    ups = [torch.nn.functional.interpolate(l, size, mode=upsample_interp, align_corners=True) for l in levels]
    return torch.cat(ups, dim=1)


def unproj_map(width: int, height: int, f, c=None) -> torch.Tensor:
    """util.py:115-145: unit camera ray per pixel, (H, W, 3)."""
    if c is None:
        c = [width * 0.5, height * 0.5]
    f = [float(f), float(f)] if not hasattr(f, "__len__") else [float(f[0]), float(f[-1])]
    Y, X = torch.meshgrid(torch.arange(height, dtype=torch.float32) - float(c[1]),
                          torch.arange(width, dtype=torch.float32) - float(c[0]), indexing="ij")
    X = X / f[0]
    Y = Y / f[1]
    unproj = torch.stack((X, -Y, -torch.ones_like(X)), dim=-1)
    return unproj / torch.norm(unproj, dim=-1).unsqueeze(-1)


def gen_rays(poses: torch.Tensor, width: int, height: int, focal, z_near: float, z_far: float, c=None) -> torch.Tensor:
    """util.py:240-278 (ndc=False): (N, H, W, 8) = [camera centre, R d, near, far]."""
    n = poses.shape[0]
    um = unproj_map(width, height, focal, c).unsqueeze(0).repeat(n, 1, 1, 1)
    centers = poses[:, None, None, :3, 3].expand(-1, height, width, -1)
    dirs = torch.matmul(poses[:, None, None, :3, :3], um.unsqueeze(-1))[:, :, :, :, 0]
    near = torch.tensor(z_near).view(1, 1, 1, 1).expand(n, height, width, -1)
    far = torch.tensor(z_far).view(1, 1, 1, 1).expand(n, height, width, -1)
    return torch.cat((centers, dirs, near, far), dim=-1)


def gen_rays_yolo(poses_w2c: torch.Tensor, width: int, height: int, focal, c, z_near: float, z_far: float) -> torch.Tensor:
    """util.py:808-876: rays through (x + 0.49, y + 0.49) of a width x height cell grid for world-to-camera extrinsics:
    direction = inv(E)[:3,:3] . inv(K) . [x+0.49, y+0.49, 1] (not normalised), origin = inv(E)[:3,3].  (N, H, W, 8)."""
    K = torch.tensor([[float(focal[0]), 0.0, float(c[0])], [0.0, float(focal[1]), float(c[1])], [0.0, 0.0, 1.0]])
    Kinv = torch.inverse(K)                                                  # :818-821
    gx, gy = torch.meshgrid(torch.linspace(0, width - 1, width), torch.linspace(0, height - 1, height), indexing="ij")
    pix = torch.stack([gx + 0.49, gy + 0.49, torch.ones_like(gx)], dim=2).view(-1, 3)      # :824-835, x-major order
    d_cam = torch.matmul(Kinv, pix.t()).t()                                  # :838
    near = torch.full((height * width, 1), float(z_near))
    far = torch.full((height * width, 1), float(z_far))
    rays = []
    for i in range(poses_w2c.shape[0]):
        Einv = torch.inverse(poses_w2c[i])                                   # :853
        d_world = torch.matmul(Einv[:3, :3], d_cam.t()).t()                  # :857
        o = Einv[:3, 3].repeat(height * width, 1)                            # :860-863
        ray = torch.cat([o, d_world, near, far], dim=1).view(width, height, 8).permute(1, 0, 2)   # :866-872
        rays.append(ray)
    return torch.stack(rays)


def yolo_render(scene: Scene, mlp: Dict[str, torch.Tensor], rays: torch.Tensor, noise: torch.Tensor, *,
                n_coarse: int = 128, num_anchors: int = 3, grad: bool = False, **field_kw) -> torch.Tensor:
    """YoloRenderer.forward (src/render/yolo.py:37-114).  rays (B, 8), noise (B, Kc) -> (B, anchors, 7).  ``grad=True`` records
    the autograd graph the reference's YoloTrainer back-propagates through (train/trainlib/YoloTrainer.py:140-190)."""
    with torch.set_grad_enabled(grad):
        r = rays.reshape(-1, 8)
        z = sample_coarse(r, noise, n_coarse, False)                        # yolo.py:15-27
        B, K = z.shape
        pts, dirs = ray_points(r, z)                                        # yolo.py:58-66
        out = field_forward(scene, mlp, pts.reshape(1, -1, 3), dirs.reshape(1, -1, 3), yolo=True, **field_kw)
        out = out.reshape(B, K, num_anchors, 7)                             # yolo.py:93
        prob = _f64(torch.sigmoid, out[..., 0])                             # yolo.py:96
        summed = prob.sum(dim=1)
        vals = (out[..., 1:] * prob.unsqueeze(-1)).sum(dim=1)
        vals = vals / (summed.unsqueeze(-1) + 1e-5)                         # yolo.py:107
        return torch.cat([prob.max(dim=1)[0].unsqueeze(-1), vals], dim=-1)  # yolo.py:109-114
