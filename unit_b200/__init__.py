"""unit_b200 -- B200-native (sm_100a) implementation of UniT's per-image RoI stage.

Drop-in for the path ``self.roi_heads(images, features, proposals, targets)`` of ubc-vision/UniT
(modeling/meta_arch/rcnn.py:482,538): the heads, predictors, ``ROIPooler``, ``Matcher`` and inference helpers keep
the reference's names and call conventions; the arithmetic runs in hand-written CUDA kernels of
``libunit_b200.so`` (C ABI in include/unit_b200.h).  There is no CPU path and no fallback.
"""
from .config import CfgNode, add_config, get_cfg, load_cfg  # noqa: F401

__version__ = "0.1.0"


def _lazy():
    from . import heads_aux, layers, ops, predictors, roi_heads, structures  # noqa: F401
    from . import d2compat  # noqa: F401

    return ops


def __getattr__(name):
    import importlib

    if name in ("ops", "layers", "structures", "predictors", "roi_heads", "heads_aux", "registry", "distributed",
                "d2compat"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
