"""mmdet / mmcv registry glue.

The reference registers the head with ``@HEADS.register_module()`` on ``mmcv.runner.BaseModule``
(relation_transformer_head_v4.py:11-13,20-21) and the detector builds it with
``mmdet.models.builder.build_head`` (detectors/openseed_relation_v2.py:69).  When mmcv / mmdet (1.x / 2.x
era) are importable we register into the real registries, so ``configs/psg/baseline_v4_ov.py`` works
unchanged through ``custom_imports``; otherwise a minimal local registry with the same two entry points
keeps the class constructible from the same config dict.
"""
from __future__ import annotations

import torch.nn as nn

try:  # pragma: no cover - exercised only where mmdet 2.x is installed
    from mmcv.runner import BaseModule  # type: ignore
    from mmdet.models.builder import HEADS  # type: ignore
    HAVE_MMDET = True
except Exception:  # noqa: BLE001
    HAVE_MMDET = False
    BaseModule = nn.Module

    class _Registry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                self.module_dict[name or cls.__name__] = cls
                return cls
            return deco(module) if module is not None else deco

        def get(self, key):
            return self.module_dict.get(key)

        def build(self, cfg):
            cfg = dict(cfg)
            return self.module_dict[cfg.pop("type")](**cfg)

    HEADS = _Registry("head")


def build_head(cfg):
    """``mmdet.models.builder.build_head`` equivalent."""
    if HAVE_MMDET:  # pragma: no cover
        from mmdet.models.builder import build_head as _bh  # type: ignore
        return _bh(cfg)
    return HEADS.build(cfg)
