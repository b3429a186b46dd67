"""``src/render/render_util.py``."""
from .nerf import NeRFRenderer
from .yolo import YoloRenderer


def make_renderer(conf, lindisp=False):
    renderer_type = conf.get_string("renderer.type", "nerf")   # nerf | yolo
    if renderer_type == "nerf":
        return NeRFRenderer.from_conf(conf["renderer"], lindisp=lindisp)
    if renderer_type == "yolo":
        return YoloRenderer.from_conf(conf)
    raise NotImplementedError("Unsupported renderer type")
