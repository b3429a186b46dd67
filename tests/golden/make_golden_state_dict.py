"""Checkpoint-compatibility manifest: names, shapes and dtypes of ``state_dict()`` of the UNMODIFIED reference's
``make_model(conf["model"])`` (default_mv-style conf, resnet34 encoder) and ``NeRFRenderer.from_conf`` -- what a
``pixel_nerf_latest`` / ``_renderer`` checkpoint written by the reference's trainer contains (models.py:320-370,
train/trainlib/trainer.py) -- plus the values of the two positional-encoding buffers.

    python tests/golden/make_golden_state_dict.py      # build container only -> tests/golden/reference_state_dict.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    G._install_shims()
    from model import make_model
    from render import NeRFRenderer
    torch.manual_seed(0)
    net = make_model(G._Conf(G.MODEL_CONF))
    ren = NeRFRenderer.from_conf(G._Conf(G.RENDER_CONF))
    man = {
        "model": {k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()},
        "renderer": {k: [list(v.shape), str(v.dtype)] for k, v in ren.state_dict().items()},
        "code._freqs": net.state_dict()["code._freqs"].flatten().tolist(),
        "code._phases": net.state_dict()["code._phases"].flatten().tolist(),
        "n_params": sum(p.numel() for p in net.parameters()),
    }
    path = os.path.join(HERE, "reference_state_dict.json")
    with open(path, "w") as fh:
        json.dump(man, fh, indent=0, sort_keys=True)
    print("wrote", path, len(man["model"]), "model entries,", len(man["renderer"]), "renderer entries,", man["n_params"], "parameters")


if __name__ == "__main__":
    main()
