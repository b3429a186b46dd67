"""Importable alias of the ``pixel-nerf-yolo_b200/`` package directory.

The package directory carries the repository's name (with hyphens), which Python cannot import
directly; this module points ``__path__`` at it so that ``import pixel_nerf_yolo_b200.render`` works.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pixel-nerf-yolo_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
