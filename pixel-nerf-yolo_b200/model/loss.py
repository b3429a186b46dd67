"""RGB loss of the training step: drop-in for ``get_rgb_loss`` (src/model/loss.py:92-104) whose forward and backward are
one kernel (``pnr_rgb_loss``) instead of sub / pow / mean and their autograd nodes."""
import torch

from .. import _lib


class _RgbLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, gt, use_l1):
        _lib.require_cuda(rgb, "rgb")
        dev = rgb.device
        rgb_c, gt_c = rgb.contiguous().float(), gt.to(dev).contiguous().float()
        assert rgb_c.shape == gt_c.shape, "rgb and ground truth disagree in shape"
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        d_rgb = torch.empty_like(rgb_c) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(dev):
            rc = _lib.load().pnr_rgb_loss(rgb_c.data_ptr(), gt_c.data_ptr(), loss.data_ptr(), _lib.ptr(d_rgb), rgb_c.numel(),
                                          int(use_l1), _lib.stream_ptr(dev))
        _lib.check(rc, "pnr_rgb_loss")
        ctx.d_rgb = d_rgb
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.d_rgb * g if ctx.d_rgb is not None else None), None, None


class RgbLoss(torch.nn.Module):
    def __init__(self, use_l1=False):
        super().__init__()
        self.use_l1 = bool(use_l1)

    def forward(self, rgb, gt):
        return _RgbLossFn.apply(rgb, gt, self.use_l1)


def get_rgb_loss(conf, coarse=True, using_bg=False, reduction="mean"):
    if conf.get_bool("use_uncertainty", False) and not coarse:
        raise NotImplementedError("get_rgb_loss (B200 path): RGBWithUncertainty is not built")
    if reduction != "mean":
        raise NotImplementedError("get_rgb_loss (B200 path): only reduction='mean' is built")
    return RgbLoss(conf.get_bool("use_l1", False))
