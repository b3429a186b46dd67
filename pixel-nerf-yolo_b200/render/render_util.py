"""``src/render/render_util.py``."""
from .nerf import NeRFRenderer


def make_renderer(conf, lindisp=False):
    renderer_type = conf.get_string("renderer.type", "nerf")
    if renderer_type == "nerf":
        return NeRFRenderer.from_conf(conf["renderer"], lindisp=lindisp)
    if renderer_type == "yolo":
        raise NotImplementedError("YoloRenderer is the next row after the NeRF path (SURVEY.md section 8f); not built yet")
    raise NotImplementedError("Unsupported renderer type")
