"""Registration behind the reference's registries.

With the real Detectron2 importable, every class of this package is registered into
``detectron2.modeling.ROI_HEADS_REGISTRY`` / ``ROI_BOX_HEAD_REGISTRY`` / ``ROI_MASK_HEAD_REGISTRY`` under the
reference's names, so ``build_model(cfg)`` on an unchanged UniT YAML instantiates the B200 heads instead of
modeling/roi_heads/roi_heads.py's.  Without Detectron2 (this image) the package's own registries serve
``unit_b200.roi_heads.build_roi_heads(cfg, input_shape)``.
"""
from __future__ import annotations

from . import heads_aux, predictors, roi_heads  # noqa: F401  (populate the package registries)
from .registry import ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY, ROI_MASK_HEAD_REGISTRY


def register_into_detectron2(overwrite: bool = True) -> bool:
    try:
        from detectron2.modeling import ROI_HEADS_REGISTRY as D2_HEADS  # type: ignore
        from detectron2.modeling.roi_heads.box_head import ROI_BOX_HEAD_REGISTRY as D2_BOX  # type: ignore
        from detectron2.modeling.roi_heads.mask_head import ROI_MASK_HEAD_REGISTRY as D2_MASK  # type: ignore
    except Exception:
        return False
    for ours, theirs in ((ROI_HEADS_REGISTRY, D2_HEADS), (ROI_BOX_HEAD_REGISTRY, D2_BOX),
                         (ROI_MASK_HEAD_REGISTRY, D2_MASK)):
        for name in ours.names():
            if name in theirs._obj_map:
                if not overwrite:
                    continue
                del theirs._obj_map[name]
            theirs._do_register(name, ours.get(name))
    return True
