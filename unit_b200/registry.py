"""``Registry`` and ``@configurable`` with the Detectron2 semantics the reference relies on
(every reference class is created as ``REGISTRY.get(cfg...NAME)(cfg, input_shape)``, e.g. fast_rcnn.py:587-589,
and declares ``@configurable __init__`` + ``from_config``, e.g. roi_heads.py:136,173).

When the real ``detectron2`` is importable its registries are reused, so the classes of this package become
visible to ``detectron2.modeling.build_model`` under the reference's names (see d2compat.py).
"""
from __future__ import annotations

import functools
import inspect
from typing import Any, Dict, Optional

from .config import CfgNode


class Registry:
    def __init__(self, name: str):
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def _do_register(self, name: str, obj: Any) -> None:
        if name in self._obj_map:
            raise KeyError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj: Any = None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name: str) -> Any:
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def names(self):
        return sorted(self._obj_map)


def _called_with_cfg(*args, **kwargs) -> bool:
    if len(args) and isinstance(args[0], CfgNode):
        return True
    return isinstance(kwargs.get("cfg", None), CfgNode)


def _args_from_config(from_config_func, *args, **kwargs):
    sig = inspect.signature(from_config_func)
    if list(sig.parameters.keys())[0] != "cfg":
        raise TypeError(f"{from_config_func.__qualname__}.from_config must take 'cfg' as the first argument!")
    var = any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in sig.parameters.values())
    if var:
        return from_config_func(*args, **kwargs)
    supported = set(sig.parameters.keys())
    extra = {k: kwargs.pop(k) for k in list(kwargs.keys()) if k not in supported}
    ret = from_config_func(*args, **kwargs)
    ret.update(extra)
    return ret


def configurable(init_func):
    """Call ``__init__`` either with explicit arguments or with ``(cfg, ...)`` routed through ``from_config``."""

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        from_config_func = getattr(type(self), "from_config", None)
        if from_config_func is None or not inspect.ismethod(from_config_func):
            raise AttributeError("Class with @configurable must have a 'from_config' classmethod.")
        if _called_with_cfg(*args, **kwargs):
            init_func(self, **_args_from_config(from_config_func, *args, **kwargs))
        else:
            init_func(self, *args, **kwargs)

    return wrapped


ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
ROI_BOX_HEAD_REGISTRY = Registry("ROI_BOX_HEAD")
ROI_MASK_HEAD_REGISTRY = Registry("ROI_MASK_HEAD")
FAST_RCNN_REGISTRY = Registry("FAST_RCNN_REGISTRY")
WEAK_DETECTOR_FAST_RCNN_REGISTRY = Registry("WEAK_DETECTOR_FAST_RCNN")


class _EventStorage:
    """Scalar sink with the ``put_scalar`` signature of Detectron2's EventStorage (stats only, never on the hot path)."""

    def __init__(self):
        self.scalars: Dict[str, float] = {}

    def put_scalar(self, name: str, value, smoothing_hint: bool = True) -> None:
        self.scalars[name] = float(value)


_storage = _EventStorage()


def get_event_storage() -> _EventStorage:
    try:  # use Detectron2's storage when a trainer has opened one
        from detectron2.utils.events import get_event_storage as d2_get  # type: ignore

        return d2_get()
    except Exception:
        return _storage
