"""Configuration objects with the pyhocon surface the reference uses, plus a HOCON-subset parser.

The reference reads ``conf/*.conf`` with pyhocon (``src/util/args.py:89-99``) and then only uses
``conf["a.b"]``, ``"key" in conf``, ``get_int/get_float/get_bool/get_string/get_list(key, default)``.
pyhocon is not a dependency of this package; ``ConfigTree`` offers exactly that surface (and accepts a
real pyhocon tree unchanged, since the model/renderer classes only call those methods), and
``parse_file`` understands the syntax the reference's conf files use: ``#`` / ``//`` comments,
``key = value`` / ``key : value``, nested ``key { ... }`` (with or without a space), lists (possibly
nested and multi-line), ``True/False/true/false``, numbers, bare and quoted strings and
``include required("relative/path.conf")`` with later keys deep-merging over included ones.
"""
from __future__ import annotations

import os
import re
from typing import Any, List


class ConfigTree(dict):
    def _walk(self, key: str):
        cur: Any = self
        for part in key.split("."):
            if not isinstance(cur, dict) or not dict.__contains__(cur, part):
                raise KeyError(key)
            cur = dict.__getitem__(cur, part)
        return cur

    def __getitem__(self, key):
        v = self._walk(key)
        return ConfigTree(v) if isinstance(v, dict) and not isinstance(v, ConfigTree) else v

    def __contains__(self, key) -> bool:
        try:
            self._walk(key)
            return True
        except KeyError:
            return False

    def get(self, key, default=None):
        return self[key] if key in self else default

    def get_int(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else int(v)

    def get_float(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else float(v)

    def get_bool(self, key, default=None):
        v = self.get(key, default)
        if isinstance(v, str):
            return v.lower() in ("true", "yes", "on")
        return None if v is None else bool(v)

    def get_string(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else str(v)

    def get_list(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else list(v)

    @staticmethod
    def from_dict(d: dict) -> "ConfigTree":
        return ConfigTree({k: (ConfigTree.from_dict(v) if isinstance(v, dict) else v) for k, v in d.items()})


def _deep_merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = v
    return dst


_TOKEN = re.compile(r"""
    (?P<ws>[ \t\r]+) | (?P<nl>\n) | (?P<comment>(\#|//)[^\n]*) |
    (?P<str>"(?:[^"\\]|\\.)*") | (?P<punct>[{}\[\]=:,()]) |
    (?P<word>[^\s{}\[\]=:,()"#]+)
""", re.X)


def _tokens(text: str) -> List[tuple]:
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f"HOCON: cannot tokenise at {text[pos:pos + 20]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind in ("ws", "comment"):
            continue
        out.append((kind, m.group()))
    return out


def _scalar(word: str):
    low = word.lower()
    if low in ("true", "false"):
        return low == "true"
    if low == "null":
        return None
    try:
        return int(word)
    except ValueError:
        pass
    try:
        return float(word)
    except ValueError:
        return word


class _Parser:
    def __init__(self, toks, base_dir):
        self.t, self.i, self.base = toks, 0, base_dir

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def skip_sep(self):
        while self.peek()[0] == "nl" or self.peek() == ("punct", ","):
            self.i += 1

    def obj(self, closing: bool) -> dict:
        out: dict = {}
        while True:
            self.skip_sep()
            kind, val = self.peek()
            if kind == "eof":
                if closing:
                    raise ValueError("HOCON: missing '}'")
                return out
            if (kind, val) == ("punct", "}"):
                if not closing:
                    raise ValueError("HOCON: stray '}'")
                self.i += 1
                return out
            if kind == "word" and val == "include":
                self.i += 1
                _deep_merge(out, self.include())
                continue
            key = self.next()[1]
            if key.startswith('"'):
                key = key[1:-1]
            kind, val = self.peek()
            if (kind, val) == ("punct", "{"):
                self.i += 1
                v = self.obj(True)
            elif kind == "punct" and val in "=:":
                self.i += 1
                v = self.value()
            else:
                raise ValueError(f"HOCON: expected '=', ':' or '{{' after key {key!r}, got {val!r}")
            cur = out
            parts = key.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            if isinstance(v, dict) and isinstance(cur.get(parts[-1]), dict):
                _deep_merge(cur[parts[-1]], v)
            else:
                cur[parts[-1]] = v

    def include(self) -> dict:
        required = False
        if self.peek() == ("word", "required"):
            required = True
            self.i += 1
            assert self.next() == ("punct", "(")
        kind, val = self.next()
        if kind != "str":
            raise ValueError("HOCON: include expects a quoted path")
        if required:
            assert self.next() == ("punct", ")")
        path = os.path.join(self.base, val[1:-1])
        if not os.path.exists(path):
            if required:
                raise FileNotFoundError(path)
            return {}
        return _parse_path(path)

    def value(self):
        while self.peek()[0] == "nl":
            self.i += 1
        kind, val = self.next()
        if (kind, val) == ("punct", "{"):
            return self.obj(True)
        if (kind, val) == ("punct", "["):
            items = []
            while True:
                self.skip_sep()
                if self.peek() == ("punct", "]"):
                    self.i += 1
                    return items
                items.append(self.value())
        if kind == "str":
            return val[1:-1]
        if kind == "word":
            words = [val]                      # unquoted strings may contain spaces up to end of line
            while self.peek()[0] == "word":
                words.append(self.next()[1])
            return _scalar(words[0]) if len(words) == 1 else " ".join(words)
        raise ValueError(f"HOCON: unexpected token {val!r}")


def _parse_path(path: str) -> dict:
    with open(path) as fh:
        text = fh.read()
    return _Parser(_tokens(text), os.path.dirname(os.path.abspath(path))).obj(False)


def parse_file(path: str) -> ConfigTree:
    return ConfigTree.from_dict(_parse_path(path))


def parse_string(text: str, base_dir: str = ".") -> ConfigTree:
    return ConfigTree.from_dict(_Parser(_tokens(text), base_dir).obj(False))


class DotMap(dict):
    """Attribute-access dict with auto-vivified children, the subset of ``dotmap.DotMap`` the renderer's
    callers rely on (``outputs.fine.rgb``, ``len(outputs.fine) > 0``, ``.toDict()``)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        if k not in self:
            self[k] = DotMap()
        return self[k]

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}
