"""``MODELS`` registry with the mmengine surface the reference uses
(``@MODELS.register_module()`` / ``MODELS.build(dict(type=...))``; reference:
unidet3d/spconv_unet.py:94, unidet3d/encoder.py:113, unidet3d/unidet3d.py:20,80-82).

When mmdet3d is importable the classes are ALSO registered into ``mmdet3d.registry.MODELS``
(``force=True``) so that the reference configs resolve ``type='SpConvUNet'`` etc. to these
implementations after swapping ``custom_imports`` to ``unidet3d_b200``.
"""
from __future__ import annotations

import inspect


class Registry:
    def __init__(self, name: str):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, **default_args):
        if cfg is None:
            return None
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        t = args.pop("type")
        cls = self._modules.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        for k, v in default_args.items():
            args.setdefault(k, v)
        return cls(**args)

    def __contains__(self, key):
        return key in self._modules


MODELS = Registry("unidet3d_b200::model")

try:  # pragma: no cover - mmdet3d is absent in the build container
    from mmdet3d.registry import MODELS as _MM_MODELS
except Exception:  # noqa: BLE001
    _MM_MODELS = None


def register_model(cls):
    """Register into the local registry and, if present, into mmdet3d's (overriding the reference)."""
    MODELS.register_module(module=cls, force=True)
    if _MM_MODELS is not None:
        _MM_MODELS.register_module(module=cls, force=True)
    return cls
