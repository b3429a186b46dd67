"""Generate golden vectors by executing the UNMODIFIED reference (kofinandi/pixel-nerf-yolo).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference needs three import shims because ``pyhocon``/``dotmap`` are not installed and
``src/model/custom_encoder.py:7-12`` imports an un-vendored sibling repo at import time; the shims
live in this file (they are not part of the product).  Weights and feature maps come from
``pixel_nerf_yolo_b200.synth`` (numpy PCG64), so only inputs that cannot be regenerated and the
reference's OUTPUTS are stored -> ``tests/golden/*.npz`` stay small.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------ shims
class _Conf(dict):
    def _walk(self, key):
        cur = self
        for part in key.split("."):
            cur = dict.__getitem__(cur, part)
        return cur

    def __getitem__(self, key):
        v = self._walk(key)
        return _Conf(v) if isinstance(v, dict) else v

    def __contains__(self, key):
        try:
            self._walk(key)
            return True
        except KeyError:
            return False

    def get(self, key, default=None):
        return self[key] if key in self else default

    get_int = get_float = get_bool = get_string = get_list = get


def _install_shims():
    ph = types.ModuleType("pyhocon")
    ph.ConfigFactory = types.SimpleNamespace(from_dict=lambda d: _Conf(d), parse_file=None)
    ph.ConfigTree = _Conf
    sys.modules["pyhocon"] = ph

    class DotMap(dict):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            if k not in self:
                self[k] = DotMap()
            return self[k]

        __setattr__ = dict.__setitem__

        def toDict(self):
            return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}

    dm = types.ModuleType("dotmap")
    dm.DotMap = DotMap
    sys.modules["dotmap"] = dm
    models = types.ModuleType("models")
    yolo = types.ModuleType("models.yolo")
    yolo.Model = type("Model", (), {})
    models.yolo = yolo
    sys.modules["models"] = models
    sys.modules["models.yolo"] = yolo
    sys.path.insert(0, REF)


MODEL_CONF = {
    "use_encoder": True, "use_global_encoder": False, "use_xyz": True, "canon_xyz": False,
    "use_code": True, "code": {"num_freqs": 6, "freq_factor": 1.5, "include_input": True},
    "use_viewdirs": True, "use_code_viewdirs": False,
    "mlp_coarse": {"type": "resnet", "n_blocks": 5, "d_hidden": 512, "d_out": 4,
                   "combine_layer": 3, "combine_type": "average"},
    "mlp_fine": {"type": "resnet", "n_blocks": 5, "d_hidden": 512, "d_out": 4,
                 "combine_layer": 3, "combine_type": "average"},
    "encoder": {"backbone": "resnet34", "pretrained": False, "num_layers": 4, "index_padding": "zeros"},
}
RENDER_CONF = {"n_coarse": 64, "n_fine": 32, "n_fine_depth": 16, "depth_std": 0.01, "sched": [],
               "white_bkgd": True}


class _NoisePatch:
    """Feed explicit noise to the reference's four RNG calls (nerf.py:117,141,147,164)."""

    def __init__(self, noise):
        self.q_rand_like = [noise["coarse"], noise["fine_jitter"]]
        self.q_rand = [noise["fine_u"]]
        self.q_randn_like = [noise["depth"]]

    def __enter__(self):
        self.saved = (torch.rand, torch.rand_like, torch.randn_like)
        torch.rand = lambda *a, **k: self.q_rand.pop(0).clone()
        torch.rand_like = lambda *a, **k: self.q_rand_like.pop(0).clone()
        torch.randn_like = lambda *a, **k: self.q_randn_like.pop(0).clone()
        return self

    def __exit__(self, *exc):
        torch.rand, torch.rand_like, torch.randn_like = self.saved


def np_noise(seed, B, kc=64, kf=16, kfd=16):
    rng = np.random.default_rng(seed)
    return {"coarse": torch.from_numpy(rng.random((B, kc), dtype=np.float32)),
            "fine_u": torch.from_numpy(rng.random((B, kf), dtype=np.float32)),
            "fine_jitter": torch.from_numpy(rng.random((B, kf), dtype=np.float32)),
            "depth": torch.from_numpy(rng.standard_normal((B, kfd)).astype(np.float32))}


def build_reference_net(synth, coarse_seed, fine_seed, scene, num_objs):
    from model import make_model
    torch.manual_seed(0)
    net = make_model(_Conf(MODEL_CONF)).eval()
    net.mlp_coarse.load_state_dict(synth.mlp_state(coarse_seed))
    net.mlp_fine.load_state_dict(synth.mlp_state(fine_seed))
    # encode(): run the reference's camera bookkeeping, then inject the synthetic feature maps in
    # place of the CNN output (the encoder trunk is out of scope for the path under test).
    NS = scene["poses"].shape[1]
    W, H = scene["image_wh"]
    images = torch.zeros(num_objs, NS, 3, H, W)
    net.encode(images, scene["poses"], scene["focal"])
    lat = scene["latent"]
    net.encoder.latent = lat
    ls = torch.tensor([float(lat.shape[-1]), float(lat.shape[-2])])
    net.encoder.latent_scaling = ls / (ls - 1) * 2.0
    return net


def main():
    _install_shims()
    import pixel_nerf_yolo_b200.synth as synth
    from render import NeRFRenderer
    from model.code import PositionalEncoding
    from model.resnetfc import ResnetFC

    out = {}
    torch.set_num_threads(8)

    # ---------------- op level: positional encoding (code.py:30-42)
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.uniform(-2, 2, (37, 3)).astype(np.float32))
    out["pe_x"] = x.numpy()
    out["pe_out"] = PositionalEncoding(6, 3, 1.5, True)(x).numpy()

    # ---------------- op level: ResnetFC.forward (resnetfc.py:134-186), NS=3, 5 rows per view
    mlp = ResnetFC(42, 4, n_blocks=5, d_latent=512, d_hidden=512, combine_layer=3)
    mlp.load_state_dict(synth.mlp_state(21))
    zx = torch.from_numpy(rng.standard_normal((2 * 3 * 5, 554)).astype(np.float32) * 0.5)
    out["mlp_zx"] = zx.numpy()
    with torch.no_grad():
        out["mlp_out"] = mlp(zx, combine_inner_dims=(3, 5)).numpy()

    # ---------------- model + renderer, SB=1 (configs 1/2 shape) and SB=2 (config 3 shape)
    for tag, num_objs, size, feat, nrays in (("sb1", 1, 128, 16, 48), ("sb2", 2, 128, 16, 24)):
        scene = synth.scene_config1(seed=5, num_views=3, C=512, size=size, feat=feat, num_objs=num_objs)
        net = build_reference_net(synth, 1, 2, scene, num_objs)
        all_rays = torch.cat([synth.target_rays(size, 15.0 + 20 * s, -10.0) for s in range(num_objs)])
        pick = torch.from_numpy(np.random.default_rng(7).choice(size * size, nrays, replace=False)).long()
        rays = all_rays[:, pick]                                        # (SB, nrays, 8)
        out[f"{tag}_ray_idx"] = pick.numpy()
        noise = np_noise(9, num_objs * nrays)
        for k, v in noise.items():
            out[f"{tag}_noise_{k}"] = v.numpy()
        renderer = NeRFRenderer.from_conf(_Conf(RENDER_CONF), eval_batch_size=50000).eval()
        # index(): SpatialEncoder.index with zeros padding incl. out-of-bounds uv (encoder.py:79-108)
        if tag == "sb1":
            uv = torch.from_numpy(rng.uniform(-20, size + 20, (3, 29, 2)).astype(np.float32))
            out["index_uv"] = uv.numpy()
            out["index_out"] = net.encoder.index(uv, None, net.image_shape).numpy()
            # PixelNeRFNet.forward at a handful of points (models.py:153-318)
            pts = torch.from_numpy(rng.uniform(-0.4, 0.4, (1, 19, 3)).astype(np.float32))
            dirs = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((1, 19, 3)).astype(np.float32)), dim=-1)
            out["field_xyz"], out["field_dirs"] = pts.numpy(), dirs.numpy()
            with torch.no_grad():
                out["field_coarse"] = net(pts, coarse=True, viewdirs=dirs).numpy()
                out["field_fine"] = net(pts, coarse=False, viewdirs=dirs).numpy()
        with torch.no_grad(), _NoisePatch(noise):
            res = renderer(net, rays, want_weights=True)
        for lvl in ("coarse", "fine"):
            out[f"{tag}_{lvl}_rgb"] = res[lvl].rgb.numpy()
            out[f"{tag}_{lvl}_depth"] = res[lvl].depth.numpy()
            out[f"{tag}_{lvl}_weights"] = res[lvl].weights.numpy()
        # stage-wise sampler goldens (nerf.py:104-167) with the same noise
        r = rays.reshape(-1, 8)
        with _NoisePatch(noise) as p:
            zc = renderer.sample_coarse(r)
            wts = torch.from_numpy(out[f"{tag}_coarse_weights"]).reshape(-1, 64)
            zf = renderer.sample_fine(r, wts)
            zd = renderer.sample_fine_depth(r, torch.from_numpy(out[f"{tag}_coarse_depth"]).reshape(-1))
        out[f"{tag}_z_coarse"], out[f"{tag}_z_fine"], out[f"{tag}_z_depth"] = zc.numpy(), zf.numpy(), zd.numpy()

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_outputs.npz"),
          {k: v.shape for k, v in out.items()})
    print("sb1 coarse sum-w mean", out["sb1_coarse_weights"].sum(-1).mean(),
          "fine sum-w mean", out["sb1_fine_weights"].sum(-1).mean(),
          "fine depth mean", out["sb1_fine_depth"].mean())


if __name__ == "__main__":
    main()
