"""Ray generation: drop-in for ``util.gen_rays`` / ``util.unproj_map`` (src/util/util.py:115-145, 240-278) and
``util.gen_rays_yolo`` (:808-876), the step right before the renderer, as kernels of the C-ABI
library (``pnr_gen_rays``, ``pnr_gen_rays_yolo``).

``pix_inds`` is an extension for the trainer's ray sampling (PixelNerfTrainer.py:100-117): instead of generating all
N*H*W rays and indexing them, only the selected pixels' rays are produced.
"""
import warnings

import torch

from . import _lib


def _pair(v, default=None):
    if v is None:
        return default
    if isinstance(v, (int, float)):
        return float(v), float(v)
    v = torch.as_tensor(v).detach().float().reshape(-1).cpu()
    return (float(v[0]), float(v[0])) if v.numel() == 1 else (float(v[0]), float(v[1]))


def gen_rays(poses, width, height, focal, z_near, z_far, c=None, ndc=False, pix_inds=None):
    """poses (N, 4, 4) camera-to-world on a CUDA device -> rays (N, H, W, 8) = [origin, direction, near, far]
    (or (len(pix_inds), 8) for flat pixel indices into (N, H, W))."""
    if ndc:
        raise NotImplementedError("gen_rays (B200 path): ndc rays are not built (no shipped conf renders in NDC)")
    _lib.require_cuda(poses, "poses")
    _lib.require_device(poses.device)
    dev = poses.device
    n = poses.shape[0]
    fx, fy = _pair(focal.squeeze() if torch.is_tensor(focal) else focal)
    cx, cy = _pair(c.squeeze() if torch.is_tensor(c) else c, (width * 0.5, height * 0.5))
    p = poses.detach().contiguous().float()
    if pix_inds is None:
        n_out = n * height * width
        rays = torch.empty(n, height, width, 8, device=dev, dtype=torch.float32)
        idx_ptr = None
    else:
        pix_inds = pix_inds.to(device=dev, dtype=torch.int64).contiguous()
        n_out = pix_inds.numel()
        rays = torch.empty(n_out, 8, device=dev, dtype=torch.float32)
        if n_out == 0:
            return rays
        idx_ptr = pix_inds.data_ptr()
    with torch.cuda.device(dev):
        rc = _lib.load().pnr_gen_rays(p.data_ptr(), idx_ptr, rays.data_ptr(), n_out, n, height, width, fx, fy, cx, cy,
                                      float(z_near), float(z_far), _lib.stream_ptr(dev))
    _lib.check(rc, "pnr_gen_rays")
    return rays


def gen_rays_yolo(poses, width, height, focal, c, z_near, z_far):
    """Drop-in for ``util.gen_rays_yolo`` (src/util/util.py:808-876): poses (N, 4, 4) WORLD-TO-CAMERA extrinsics on a CUDA
    device -> rays (N, H, W, 8) through the cell centres (+0.49) of a ``width x height`` detection grid; directions are not
    normalised (as in the reference).  The two small matrix inverses are taken with ``torch.inverse`` like the reference
    does; the per-ray arithmetic is one kernel (``pnr_gen_rays_yolo``)."""
    _lib.require_cuda(poses, "poses")
    _lib.require_device(poses.device)
    dev = poses.device
    fx, fy = _pair(focal)
    cx, cy = _pair(c)
    intr = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=torch.float32).to(dev)
    inv_intr = torch.inverse(intr).contiguous()
    inv_extr = torch.inverse(poses.detach().float()).contiguous()
    n = poses.shape[0]
    rays = torch.empty(n, height, width, 8, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().pnr_gen_rays_yolo(inv_intr.data_ptr(), inv_extr.data_ptr(), rays.data_ptr(), n, height, width,
                                           float(z_near), float(z_far), _lib.stream_ptr(dev))
    _lib.check(rc, "pnr_gen_rays_yolo")
    return rays


def image_output(rgb, depth, z_near, z_far):
    """eval/eval.py:283-290 in one kernel: rgb (..., 3) -> uint8 (clamp to [0, 1], x255, truncate), depth (...) ->
    (depth - z_near) / (z_far - z_near).  Returns (rgb_u8, depth_norm) on the device."""
    _lib.require_cuda(rgb, "rgb")
    _lib.require_device(rgb.device)
    dev = rgb.device
    r, d = rgb.contiguous().float(), depth.contiguous().float()
    n = d.numel()
    assert r.numel() == 3 * n, "rgb must hold 3 values per depth value"
    out_u8 = torch.empty(r.shape, device=dev, dtype=torch.uint8)
    out_d = torch.empty(d.shape, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().pnr_image_output(r.data_ptr(), d.data_ptr(), out_u8.data_ptr(), out_d.data_ptr(), n, float(z_near),
                                          float(z_far), _lib.stream_ptr(dev))
    _lib.check(rc, "pnr_image_output")
    return out_u8, out_d
