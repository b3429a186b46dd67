from .nerf import NeRFRenderer
from .yolo import YoloRenderer
from .render_util import make_renderer
