"""SpatialEncoder: the feature-map producer (``src/model/encoder.py``).

The convolutional trunk is outside the hot path: it runs once per scene
through torchvision/cuDNN exactly as in the reference.  What this class adds is the layout the hot path
wants: after ``forward`` the fp32 NCHW ``latent`` buffer is repacked once into channels-last bf16 (and,
on demand, fp32) by ``pnr_pack_features`` so that a bilinear tap is one contiguous C-vector.
In inference ``forward`` goes one step further: the pyramid levels are upsampled,
concatenated and written channels-last by ONE kernel (``pnr_pyramid_pack``); the reference's fp32 NCHW ``latent``
is then only materialised if somebody reads the attribute.
``index`` is the stand-alone 4-tap gather operator (``pnr_index_features``).
"""
import functools

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib


def _norm_layer(norm_type):
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == "group":
        return functools.partial(nn.GroupNorm, 32)
    if norm_type == "none":
        return None
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


class SpatialEncoder(nn.Module):
    def __init__(self, backbone="resnet34", pretrained=True, num_layers=4, index_interp="bilinear",
                 index_padding="border", upsample_interp="bilinear", feature_scale=1.0, use_first_pool=True,
                 norm_type="batch"):
        super().__init__()
        if norm_type != "batch":
            assert not pretrained
        if index_interp != "bilinear":
            raise NotImplementedError(f"SpatialEncoder: index_interp={index_interp!r} is not built (bilinear only)")
        if index_padding != "zeros":
            raise NotImplementedError(
                f"SpatialEncoder: index_padding={index_padding!r} is not built; every shipped conf uses 'zeros'")
        self.use_custom_resnet = backbone == "custom"
        self.feature_scale = feature_scale
        self.use_first_pool = use_first_pool
        if self.use_custom_resnet:
            # The reference wraps the YOLOv7 backbone of the un-vendored NeRF-YOLO repo here
            # (src/model/custom_encoder.py:7-22, dims=[1792]).  That trunk is not available, so only its
            # output contract is kept: 1792-channel maps installed with set_latent().
            self.model = None
            self.latent_size = 1792
        else:
            import torchvision
            self.model = getattr(torchvision.models, backbone)(weights="DEFAULT" if pretrained else None,
                                                               norm_layer=_norm_layer(norm_type))
            self.model.fc = nn.Sequential()
            self.model.avgpool = nn.Sequential()
            self.latent_size = [0, 64, 128, 256, 512, 1024][num_layers]
        self.num_layers = num_layers
        self.index_interp, self.index_padding, self.upsample_interp = index_interp, index_padding, upsample_interp
        self._latent = torch.empty(1, 1, 1, 1)
        self._levels = None          # pyramid levels of the last inference forward (latent not materialised yet)
        self.register_buffer("latent_scaling", torch.empty(2, dtype=torch.float32), persistent=False)
        self._packed = {}
        self.generation = 0          # bumped whenever the encoded maps change (set_latent / set_levels / .to()): cache keys use
                                     # it instead of (data_ptr, _version), which the caching allocator can hand out again

    # ``latent`` (N, C, Hl, Wl) fp32 NCHW, the reference's attribute (encoder.py:77,168).  After an inference forward it
    # is built lazily from the pyramid levels (the hot path never needs it).
    @property
    def latent(self):
        if self._latent is None:
            size = self._levels[0].shape[-2:]
            self._latent = torch.cat([F.interpolate(l, size, mode=self.upsample_interp, align_corners=True)
                                      for l in self._levels], dim=1)
        return self._latent

    @latent.setter
    def latent(self, value):
        self.set_latent(value)

    def _apply(self, fn, *args, **kwargs):          # .to()/.cuda() also move the (non-buffer) latent
        super()._apply(fn, *args, **kwargs)
        if self._latent is not None:
            self._latent = fn(self._latent)
        if self._levels is not None:
            self._levels = [fn(l) for l in self._levels]
        self._packed = {}
        self.generation += 1
        return self

    # ---- channels-last caches --------------------------------------------------------------------------
    def set_latent(self, latent: torch.Tensor):
        """Install an encoder output (N, C, Hl, Wl) and derive latent_scaling (encoder.py:170-172)."""
        self._latent = latent
        self._levels = None
        self.latent_size = latent.shape[1]
        ls = torch.tensor([float(latent.shape[-1]), float(latent.shape[-2])], device=latent.device)
        self.latent_scaling = ls / (ls - 1) * 2.0
        self._packed = {}
        self.generation += 1

    def set_levels(self, levels):
        """Install pyramid levels [(N, C_l, H_l, W_l)] without materialising the concatenated latent."""
        self._levels = [l.detach().contiguous().float() for l in levels]
        self._latent = None
        self.latent_size = sum(l.shape[1] for l in levels)
        h, w = levels[0].shape[-2:]
        ls = torch.tensor([float(w), float(h)], device=levels[0].device)
        self.latent_scaling = ls / (ls - 1) * 2.0
        self._packed = {}
        self.generation += 1

    def latent_shape(self):
        """(N, C, Hl, Wl) without forcing the lazy latent into existence."""
        if self._latent is not None:
            return tuple(self._latent.shape)
        l0 = self._levels[0]
        return (l0.shape[0], self.latent_size, l0.shape[2], l0.shape[3])

    def packed_latent(self, fp32: bool = False) -> torch.Tensor:
        """(N, Hl, Wl, C) channels-last copy of ``latent`` (bf16, or fp32 for the check path)."""
        if self._latent is None:
            return self._packed_from_levels(fp32)
        lat = self.latent
        key = (lat.data_ptr(), lat._version, tuple(lat.shape), bool(fp32))
        hit = self._packed.get(bool(fp32))
        if hit is not None and hit[0] == key:
            return hit[1]
        _lib.require_cuda(lat, "SpatialEncoder.latent")
        _lib.require_device(lat.device)
        n, c, h, w = lat.shape
        src = lat.detach().contiguous().float()
        dst = torch.empty(n, h, w, c, device=lat.device, dtype=torch.float32 if fp32 else torch.bfloat16)
        with torch.cuda.device(lat.device):
            rc = _lib.load().pnr_pack_features(src.data_ptr(), dst.data_ptr(), n, c, h, w, int(fp32),
                                               _lib.stream_ptr(lat.device))
        _lib.check(rc, "pnr_pack_features")
        self._packed[bool(fp32)] = (key, dst)
        return dst

    def _packed_from_levels(self, fp32: bool) -> torch.Tensor:
        """Fused upsample + concat + channels-last repack of the pyramid levels (``pnr_pyramid_pack``)."""
        import ctypes as C
        lv = self._levels
        key = (tuple(l.data_ptr() for l in lv), tuple(l._version for l in lv), bool(fp32))
        hit = self._packed.get(bool(fp32))
        if hit is not None and hit[0] == key:
            return hit[1]
        dev = lv[0].device
        _lib.require_cuda(lv[0], "SpatialEncoder pyramid levels")
        _lib.require_device(dev)
        n, _, h, w = lv[0].shape
        nl = len(lv)
        ptrs = (C.c_void_p * nl)(*[l.data_ptr() for l in lv])
        i32 = lambda vals: (C.c_int32 * nl)(*vals)
        dst = torch.empty(n, h, w, self.latent_size, device=dev, dtype=torch.float32 if fp32 else torch.bfloat16)
        with torch.cuda.device(dev):
            rc = _lib.load().pnr_pyramid_pack(ptrs, i32([l.shape[1] for l in lv]), i32([l.shape[2] for l in lv]),
                                              i32([l.shape[3] for l in lv]), nl, n, dst.data_ptr(), int(fp32),
                                              _lib.stream_ptr(dev))
        _lib.check(rc, "pnr_pyramid_pack")
        self._packed[bool(fp32)] = (key, dst)
        return dst

    def index(self, uv, cam_z=None, image_size=(), z_bounds=None):
        """uv (B, N, 2) pixel coordinates -> (B, L, N) bilinear features (align_corners, zero padding)."""
        _lib.require_cuda(uv, "uv")
        if len(image_size) == 0:
            raise NotImplementedError("SpatialEncoder.index: image_size is required (normalised uv input is not built)")
        if len(image_size) == 1:
            image_size = (image_size[0], image_size[0])
        feat = self.packed_latent(fp32=True)
        n, h, w, c = feat.shape
        sc = _lib.Scene()
        sc.feat, sc.SB, sc.NS, sc.C, sc.Hl, sc.Wl, sc.feat_fp32 = feat.data_ptr(), 1, n, c, h, w, 1
        sc.image_w, sc.image_h = float(image_size[0]), float(image_size[1])
        sc.lat_scale_x, sc.lat_scale_y = float(self.latent_scaling[0]), float(self.latent_scaling[1])
        uvc = uv.contiguous().float()
        out = torch.empty(n, c, uvc.shape[1], device=uv.device, dtype=torch.float32)
        with torch.cuda.device(uv.device):
            rc = _lib.load().pnr_index_features(sc, uvc.data_ptr(), uvc.shape[0], uvc.shape[1], out.data_ptr(),
                                                _lib.stream_ptr(uv.device))
        _lib.check(rc, "pnr_index_features")
        return out

    def forward(self, x, return_latent=True):
        """Image batch (B, 3, H, W) -> multi-level feature pyramid upsampled to the first level's size and
        concatenated along channels (B, latent_size, H/2, W/2).  ``return_latent=False`` (PixelNeRFNet.encode, which
        ignores the return value, models.py:114) skips materialising the fp32 NCHW tensor in inference."""
        if self.model is None:
            raise NotImplementedError("SpatialEncoder(backbone='custom'): the YOLOv7 trunk is not vendored; "
                                      "install its feature maps with set_latent()")
        if self.feature_scale != 1.0:
            up = self.feature_scale > 1.0
            x = F.interpolate(x, scale_factor=self.feature_scale, mode="bilinear" if up else "area",
                              align_corners=True if up else None, recompute_scale_factor=True)
        x = x.to(device=self.latent_scaling.device)
        m = self.model
        x = m.relu(m.bn1(m.conv1(x)))
        levels = [x]
        stages = [m.layer1, m.layer2, m.layer3, m.layer4]
        for li in range(1, self.num_layers):
            if li == 1 and self.use_first_pool:
                x = m.maxpool(x)
            x = stages[li - 1](x)
            levels.append(x)
        fused = (x.is_cuda and self.upsample_interp == "bilinear"
                 and not (torch.is_grad_enabled() and any(l.requires_grad for l in levels)))
        if fused:       # inference: one kernel writes the channels-last maps straight from the levels
            self.set_levels(levels)
            return self.latent if return_latent else None
        size = levels[0].shape[-2:]      # training: stay on the autograd graph (encoder.py:159-168)
        levels = [F.interpolate(l, size, mode=self.upsample_interp, align_corners=True) for l in levels]
        self.set_latent(torch.cat(levels, dim=1))
        return self.latent

    @classmethod
    def from_conf(cls, conf):
        return cls(conf.get_string("backbone"), pretrained=conf.get_bool("pretrained", True),
                   num_layers=conf.get_int("num_layers", 4), index_interp=conf.get_string("index_interp", "bilinear"),
                   index_padding=conf.get_string("index_padding", "border"),
                   upsample_interp=conf.get_string("upsample_interp", "bilinear"),
                   feature_scale=conf.get_float("feature_scale", 1.0), use_first_pool=conf.get_bool("use_first_pool", True))
