"""Multi-GPU rendering: rays shard, nothing else moves.

The reference's only multi-GPU mechanism is ``torch.nn.DataParallel(dim=1)`` (src/render/nerf.py:373-377),
which re-broadcasts every parameter and the encoded feature maps to every GPU on EVERY call and gathers
on GPU 0.  Rays are independent, so here:

* weights and the encoded scene are replicated once (each rank encodes / packs for itself);
* the ray batch is split into contiguous, tile-aligned slices (``shard_bounds``), one per rank;
* the per-ray outputs (rgb, depth[, weights]) are gathered ONCE per call with one NCCL all-gather
  (``ShardedRenderer``, one process per GPU under torchrun), or with peer copies to the first device
  (``MultiDeviceRenderer``, the single-process equivalent behind ``bind_parallel(net, gpus)``).

Noise is drawn for the FULL batch and sliced, so an N-GPU render is bit-identical to the 1-GPU render.
The host logic (bounds, padding, gather) is device-agnostic and is tested on CPU with the gloo backend.
"""
from __future__ import annotations

import copy
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RAY_TILE = 32      # slices are aligned to a multiple of this many rays


def shard_bounds(n: int, world: int, tile: int = RAY_TILE) -> List[Tuple[int, int]]:
    """Contiguous [start, end) slice of ``n`` rays for each of ``world`` ranks, sizes differing by at most
    one tile, every boundary (except the last) a multiple of ``tile``."""
    tiles = (n + tile - 1) // tile
    base, extra = divmod(tiles, world)
    out, start = [], 0
    for r in range(world):
        cnt = (base + (1 if r < extra else 0)) * tile
        end = min(n, start + cnt)
        out.append((start, end))
        start = end
    return out


def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
    if t.shape[0] == rows:
        return t.contiguous()
    pad = torch.zeros((rows - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    return torch.cat((t, pad), dim=0)


def all_gather_rows(local: torch.Tensor, bounds: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
    """All-gather row slices of unequal length: one collective on a padded buffer, then trim."""
    world = len(bounds)
    longest = max(e - s for s, e in bounds)
    buf = _pad_rows(local, longest)
    out = torch.empty((world,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)
    if buf.is_cuda:
        dist.all_gather_into_tensor(out, buf, group=group)      # one NCCL all-gather over NVLink
    else:
        dist.all_gather(list(out.unbind(0)), buf, group=group)  # gloo (CPU tests of the host logic)
    return torch.cat([out[r, : e - s] for r, (s, e) in enumerate(bounds)], dim=0)


def slice_noise(noise: Optional[Dict[str, torch.Tensor]], sb: int, b_total: int, s: int, e: int):
    """Noise tensors are (SB*B, k); take rays [s, e) of every object."""
    if noise is None:
        return None
    out = {}
    for k, v in noise.items():
        out[k] = None if v is None else v.reshape(sb, b_total, v.shape[-1])[:, s:e].reshape(sb * (e - s), v.shape[-1]).contiguous()
    return out


def draw_render_noise(n_rays: int, n_coarse: int, n_fine: int, n_fine_depth: int, device) -> Dict[str, Optional[torch.Tensor]]:
    """The four noise draws of one ``NeRFRenderer.forward`` in the reference's order (nerf.py:117,141,147,164) for the FULL
    ray batch.  Ranks that seed their generators equally draw equal tensors, so slicing them makes the N-GPU render equal the
    1-GPU render of the same seed bit for bit."""
    kf, kfd = n_fine - n_fine_depth, n_fine_depth
    out: Dict[str, Optional[torch.Tensor]] = {"coarse": torch.rand(n_rays, n_coarse, device=device, dtype=torch.float32),
                                              "fine_u": None, "fine_jitter": None, "depth": None}
    if n_fine > 0:
        if kf > 0:
            out["fine_u"] = torch.rand(n_rays, kf, dtype=torch.float32, device=device)
            out["fine_jitter"] = torch.rand_like(out["fine_u"])
        if kfd > 0:
            out["depth"] = torch.randn(n_rays, kfd, dtype=torch.float32, device=device)
    return out


class ShardedRenderer(torch.nn.Module):
    """One process per GPU (torchrun).  ``render_fn(rays_slice, noise_slice) -> (rgb (SB,b,3), depth (SB,b))``
    is the local single-GPU render; every rank receives the full (SB, B, 3)/(SB, B) result.  ``noise_fn(n_rays)``
    (optional) draws the full-batch noise when the caller passes none (see ``draw_render_noise``)."""

    def __init__(self, render_fn: Callable, group=None, tile: int = RAY_TILE, noise_fn: Optional[Callable] = None):
        super().__init__()
        self.render_fn = render_fn
        self.group = group
        self.tile = tile
        self.noise_fn = noise_fn

    @classmethod
    def for_renderer(cls, renderer, net, group=None):
        """The sharded equivalent of ``renderer.bind_parallel(net, gpus, simple_output=True)`` under torchrun."""
        wrapped = renderer.bind_parallel(net, None, simple_output=True).eval()

        def render_fn(rays, noise):
            renderer.noise_override = noise
            try:
                with torch.no_grad():
                    return wrapped(rays)
            finally:
                renderer.noise_override = None

        def noise_fn(n):
            return draw_render_noise(n, renderer.n_coarse, renderer.n_fine if renderer.using_fine else 0,
                                     renderer.n_fine_depth if renderer.using_fine else 0, next(net.parameters()).device)

        return cls(render_fn, group=group, noise_fn=noise_fn)

    def forward(self, rays: torch.Tensor, noise: Optional[Dict[str, torch.Tensor]] = None):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        sb, b_total = rays.shape[0], rays.shape[1]
        if noise is None and self.noise_fn is not None:
            noise = self.noise_fn(sb * b_total)
        bounds = shard_bounds(b_total, world, self.tile)
        s, e = bounds[rank]
        if e > s:
            rgb, depth = self.render_fn(rays[:, s:e].contiguous(), slice_noise(noise, sb, b_total, s, e))
        else:   # more ranks than ray tiles: this rank only takes part in the gather
            rgb, depth = rays.new_zeros(sb, 0, 3), rays.new_zeros(sb, 0)
        # pack (r, g, b, depth) = 16 B/ray, ray-major, and gather once
        packed = torch.cat((rgb, depth.unsqueeze(-1)), dim=-1).permute(1, 0, 2).contiguous()      # (b, SB, 4)
        full = all_gather_rows(packed, bounds, self.group).permute(1, 0, 2)                        # (SB, B, 4)
        return full[..., :3].contiguous(), full[..., 3].contiguous()


class MultiDeviceRenderer(torch.nn.Module):
    """Single-process multi-device driver behind ``NeRFRenderer.bind_parallel(net, gpus)`` (nerf.py:373-377, the
    reference's ``DataParallel(dim=1)``): the NETWORK is replicated to every device once and re-copied only when its
    state changed (new ``encode`` / new weights, tracked by explicit generation counters: ``net.render_state_key()``);
    the renderer object is shared, so its sampling settings are always current.  Each call launches one ray slice per
    device asynchronously and gathers the outputs on the first device with peer copies over NVLink.

    Inference only: the reference's DataParallel also reduces the replicas' gradients onto GPU 0 in backward; this
    driver does not, so a call that would record a backward pass raises (train with one process per GPU and
    ``dist.GradientSync``)."""

    def __init__(self, wrapped, gpus: Sequence[int]):
        super().__init__()
        self.module = wrapped
        self.devices = [torch.device("cuda", int(g)) for g in gpus]
        self._replicas = None
        self._key = None

    def state_key(self):
        """What the replicas depend on; compared by value, never by pointer identity."""
        return self.module.net.render_state_key()

    def invalidate(self):
        """Force re-replication on the next call (after editing weights through ``p.data`` etc.)."""
        self._key = None

    def _sync(self):
        key = self.state_key()
        if self._replicas is not None and key == self._key:
            return
        from .render.nerf import _RenderWrapper
        src = self.module.net
        reps = [self.module]
        for d in self.devices[1:]:
            with torch.cuda.device(d), torch.no_grad():
                latent = src.encoder._latent
                src.encoder._latent = None if latent is None else latent.detach()     # deepcopy only copies graph leaves
                try:
                    net = copy.deepcopy(src)
                finally:
                    src.encoder._latent = latent
                net = net.to(d)
                net.requires_grad_(False)
                net.num_objs, net.num_views_per_obj = src.num_objs, src.num_views_per_obj
                net._cam_cache = None
                net._field_ws = None
            reps.append(_RenderWrapper(net, self.module.renderer, self.module.simple_output))
        self._replicas, self._key = reps, key

    def forward(self, rays, want_weights=False):
        net = self.module.net
        if torch.is_grad_enabled() and (net._wants_grad(True) or net._wants_grad(False)):
            raise NotImplementedError(
                "MultiDeviceRenderer (bind_parallel(net, gpus) with several GPUs) is an inference driver: gradients of the "
                "replicas are not reduced onto the bound network.  Train with one process per GPU (torchrun) and "
                "dist.GradientSync, or call under torch.no_grad().")
        self._sync()
        renderer = self.module.renderer
        sb, b_total = rays.shape[0], rays.shape[1]
        bounds = shard_bounds(b_total, len(self.devices))
        full_noise = renderer.noise_override
        outs = []
        try:
            for rep, d, (s, e) in zip(self._replicas, self.devices, bounds):
                with torch.cuda.device(d):
                    if full_noise is not None:
                        nz = slice_noise(full_noise, sb, b_total, s, e)
                        renderer.noise_override = {k: (None if v is None else v.to(d, non_blocking=True)) for k, v in nz.items()}
                    outs.append(rep(rays[:, s:e].to(d, non_blocking=True), want_weights=want_weights))
        finally:
            renderer.noise_override = full_noise
        d0 = self.devices[0]
        if isinstance(outs[0], tuple):
            return tuple(torch.cat([o[i].to(d0, non_blocking=True) for o in outs], dim=1) for i in range(2))
        merged: Dict = {}
        for lvl in outs[0]:
            merged[lvl] = {k: torch.cat([o[lvl][k].to(d0, non_blocking=True) for o in outs], dim=1) for k in outs[0][lvl]}
        return merged


class GradientSync:
    """Data-parallel training: every rank renders its own ray batch of the same step and runs its own
    backward; the gradients of the MLPs (and of the encoder, unless ``stop_encoder_grad``) are then averaged with ONE
    all-reduce over a persistent flat fp32 bucket -- 28 M parameters = 113 MB, one NCCL call per step instead of one per
    tensor.  The reference gets the same sum from ``DataParallel``'s backward (src/render/nerf.py:373-377 replicates the
    module per call and reduces the replicas' gradients onto GPU 0), i.e. the mean over ranks of per-rank mean losses
    equals the loss of the concatenated batch when every rank holds equally many rays.

    Parameters whose ``.grad`` is None on this rank (unused this step) contribute zeros, so all ranks issue the same
    collective.  ``sync()`` returns the number of elements reduced."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self._bucket: Optional[torch.Tensor] = None

    def _flat(self) -> torch.Tensor:
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        if self._bucket is None or self._bucket.numel() != n or self._bucket.device != dev:
            self._bucket = torch.zeros(n, dtype=torch.float32, device=dev)
        return self._bucket

    @torch.no_grad()
    def sync(self) -> int:
        if not self.params:
            return 0
        flat = self._flat()
        views, off = [], 0
        for p in self.params:
            v = flat[off: off + p.numel()].view_as(p)
            views.append(v)
            off += p.numel()
        have = [p.grad is not None for p in self.params]
        if all(have):
            torch._foreach_copy_(views, [p.grad for p in self.params])
        else:
            flat.zero_()
            for v, p in zip(views, self.params):
                if p.grad is not None:
                    v.copy_(p.grad)
        world = dist.get_world_size(self.group)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)          # the training path's one collective
        flat.mul_(1.0 / world)
        for v, p in zip(views, self.params):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return flat.numel()
