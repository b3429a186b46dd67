"""Deterministic synthetic inputs for tests and benchmarks.

There is no dataset and no checkpoint on the GPU box, so weights, feature maps and cameras are
generated from numpy's PCG64 stream (stable across numpy versions and machines) rather than from
torch's initialisers.  The MLP weights follow the "exercised" recipe of SURVEY.md section 8(c):
the reference initialises every ``fc_1`` to zero (``src/model/resnetfc.py:39``), which would make
each residual block the identity, so ``fc_1`` gets a scaled kaiming draw and the density bias is
raised so that both passes have non-trivial opacity.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch


def mlp_state(seed: int, d_in: int = 42, d_latent: int = 512, d_hidden: int = 512, d_out: int = 4,
              n_blocks: int = 5, combine_layer: int = 3, fc1_scale: float = 0.05,
              sigma_bias: float = 2.0, bias_scale: float = 0.02) -> Dict[str, torch.Tensor]:
    """state_dict of one ResnetFC (names of ``src/model/resnetfc.py:103-132``)."""
    rng = np.random.default_rng(seed)

    def kaiming(out_f, in_f, scale=1.0):
        return (rng.standard_normal((out_f, in_f)) * (math.sqrt(2.0 / in_f) * scale)).astype(np.float32)

    def bias(n):
        return (rng.standard_normal(n) * bias_scale).astype(np.float32)

    sd = {"lin_in.weight": kaiming(d_hidden, d_in), "lin_in.bias": bias(d_hidden),
          "lin_out.weight": kaiming(d_out, d_hidden), "lin_out.bias": bias(d_out)}
    sd["lin_out.bias"][3] += sigma_bias
    for b in range(n_blocks):
        sd[f"blocks.{b}.fc_0.weight"] = kaiming(d_hidden, d_hidden)
        sd[f"blocks.{b}.fc_0.bias"] = bias(d_hidden)
        sd[f"blocks.{b}.fc_1.weight"] = kaiming(d_hidden, d_hidden, fc1_scale)
        sd[f"blocks.{b}.fc_1.bias"] = bias(d_hidden)
    for b in range(min(combine_layer, n_blocks)):
        sd[f"lin_z.{b}.weight"] = kaiming(d_hidden, d_latent)
        sd[f"lin_z.{b}.bias"] = bias(d_hidden)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def feature_maps(seed: int, n_maps: int, C: int, Hl: int, Wl: int, scale: float = 0.5) -> torch.Tensor:
    """Stand-in for the encoder output ``SpatialEncoder.latent`` (NCHW fp32)."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy((rng.standard_normal((n_maps, C, Hl, Wl)) * scale).astype(np.float32))


def _rot(axis: str, a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4, dtype=np.float64)
    if axis == "phi":      # about x
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    else:                  # theta, about y
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    return m


def pose_spherical(theta: float, phi: float, radius: float) -> torch.Tensor:
    """Camera-to-world pose on a sphere (same convention as ``src/util/util.py:323-337``)."""
    t = np.eye(4, dtype=np.float64)
    t[2, 3] = radius
    m = _rot("theta", theta / 180.0 * math.pi).astype(np.float32) @ (
        _rot("phi", phi / 180.0 * math.pi).astype(np.float32) @ t.astype(np.float32))
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return torch.from_numpy((flip @ m).astype(np.float32))


def gen_rays(poses: torch.Tensor, width: int, height: int, focal: float, z_near: float, z_far: float,
             c: Tuple[float, float] | None = None) -> torch.Tensor:
    """Pinhole rays [origin(3), dir(3), near, far] per pixel, (N, H, W, 8)
    (semantics of ``src/util/util.py:115-145,240-278``)."""
    cx, cy = (width * 0.5, height * 0.5) if c is None else c
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32) - float(cy),
                            torch.arange(width, dtype=torch.float32) - float(cx), indexing="ij")
    d = torch.stack((xs / float(focal), -(ys / float(focal)), -torch.ones_like(xs)), dim=-1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    n = poses.shape[0]
    dirs = torch.matmul(poses[:, None, None, :3, :3], d[None, ..., None])[..., 0]
    orig = poses[:, None, None, :3, 3].expand(-1, height, width, -1)
    near = torch.full((n, height, width, 1), float(z_near))
    far = torch.full((n, height, width, 1), float(z_far))
    return torch.cat((orig, dirs, near, far), dim=-1)


def scene_config1(seed: int = 0, num_views: int = 3, C: int = 512, size: int = 128,
                  feat: int | None = None, num_objs: int = 1):
    """Cameras/feature maps of BASELINE.json config 1/2: source views
    pose_spherical(theta,-20,1.3), theta in {0,40,-40,80,-80}; focal 131.25 at 128 px; z in [0.8,1.8]."""
    thetas = [0.0, 40.0, -40.0, 80.0, -80.0, 120.0, -120.0, 160.0][:num_views]
    feat = size // 2 if feat is None else feat
    poses = torch.stack([torch.stack([pose_spherical(t + 10.0 * s, -20.0, 1.3) for t in thetas])
                         for s in range(num_objs)])                    # (SB, NS, 4, 4)
    latent = feature_maps(seed + 1000, num_objs * num_views, C, feat, feat)
    focal = 131.25 * size / 128.0
    return dict(latent=latent, poses=poses, focal=torch.tensor(focal), image_wh=(size, size),
                z_near=0.8, z_far=1.8)


def target_rays(size: int = 128, theta: float = 15.0, phi: float = -10.0, z_near: float = 0.8,
                z_far: float = 1.8) -> torch.Tensor:
    """(1, size*size, 8) rays of one target view."""
    focal = 131.25 * size / 128.0
    pose = pose_spherical(theta, phi, 1.3)[None]
    return gen_rays(pose, size, size, focal, z_near, z_far).reshape(1, -1, 8)
