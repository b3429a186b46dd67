"""Import the unmodified reference from baseline/_ref/src (or /root/reference/src in the build container).

The reference cannot be imported as-is in this image: `pyhocon` and `dotmap` are not installed and
`src/model/custom_encoder.py:7-12` imports `models.yolo` from an un-vendored sibling repository at import time.  The three
stand-ins below are the only non-reference code on its import path; they carry no arithmetic.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


class Conf(dict):
    """Minimal pyhocon.ConfigTree: dotted lookups and the typed getters the reference calls."""

    def _walk(self, key):
        cur = self
        for part in key.split("."):
            cur = dict.__getitem__(cur, part)
        return cur

    def __getitem__(self, key):
        v = self._walk(key)
        return Conf(v) if isinstance(v, dict) else v

    def __contains__(self, key):
        try:
            self._walk(key)
            return True
        except KeyError:
            return False

    def get(self, key, default=None):
        return self[key] if key in self else default

    get_int = get_float = get_bool = get_string = get_list = get


def reference_src():
    for cand in (os.path.join(HERE, "_ref", "src"), "/root/reference/src"):
        if os.path.isdir(os.path.join(cand, "render")):
            return cand
    return None


_loaded = None


def load_reference():
    """-> namespace(make_model, NeRFRenderer, util, Conf, src) of the unmodified reference, or None if it is not available."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = reference_src()
    if src is None:
        return None
    ph = types.ModuleType("pyhocon")
    ph.ConfigFactory = types.SimpleNamespace(from_dict=lambda d: Conf(d), parse_file=None)
    ph.ConfigTree = Conf
    sys.modules.setdefault("pyhocon", ph)

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
    sys.modules.setdefault("dotmap", dm)
    if "models.yolo" not in sys.modules:
        models = types.ModuleType("models")
        yolo = types.ModuleType("models.yolo")
        yolo.Model = type("Model", (), {})
        models.yolo = yolo
        sys.modules["models"] = models
        sys.modules["models.yolo"] = yolo
    sys.path.insert(0, src)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from model import make_model          # noqa: E402  (reference: src/model/__init__.py:4-11)
        from render import NeRFRenderer        # noqa: E402  (reference: src/render/nerf.py)
        import util                            # noqa: E402  (reference: src/util)
    _loaded = types.SimpleNamespace(make_model=make_model, NeRFRenderer=NeRFRenderer, util=util, Conf=Conf, src=src)
    return _loaded
