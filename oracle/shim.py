"""Run the reference's own RoI-stage files verbatim.  TEST INFRASTRUCTURE (build container only).

The UniT side of the hot path (modeling/roi_heads/{roi_heads,fast_rcnn,weak_detector_fast_rcnn,mask_head}.py,
modeling/matcher.py) is plain PyTorch once ``import detectron2`` / ``import fvcore`` resolve.  Detectron2 cannot be
installed here (no wheel, no network), so ``install()`` registers stand-in ``detectron2.*`` / ``fvcore.*`` modules
whose ~45 symbols are the restatements in :mod:`oracle.d2`; ``load_reference()`` then imports the reference files
from ``/root/reference`` UNMODIFIED and UNCOPIED under a synthetic parent package (so the reference's
``modeling/__init__.py`` star-import of meta-arch/backbone is not executed).

Nothing here runs on the GPU box (``/root/reference`` does not exist there); it is used by
``tests/golden/make_golden.py`` and by CPU tests that are skipped when the reference is absent.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
from typing import Dict

import numpy as np
import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("UNIT_REFERENCE_ROOT", "/root/reference")
_PKG = "_unit_reference"

VOC_CLASSES = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
               "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train",
               "tvmonitor"]
# data/data_utils/cfg.py:24 -- COCO-80 order with the VOC spellings for the 20 shared categories
COCO_CLASSES = ['person', 'bicycle', 'car', 'motorbike', 'aeroplane', 'bus', 'train', 'truck', 'boat',
                'traffic light', 'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse',
                'sheep', 'cow', 'elephant', 'bear', 'zebra', 'giraffe', 'backpack', 'umbrella', 'handbag', 'tie',
                'suitcase', 'frisbee', 'skis', 'snowboard', 'sports ball', 'kite', 'baseball bat', 'baseball glove',
                'skateboard', 'surfboard', 'tennis racket', 'bottle', 'wine glass', 'cup', 'fork', 'knife', 'spoon',
                'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut',
                'cake', 'chair', 'sofa', 'pottedplant', 'bed', 'diningtable', 'toilet', 'tvmonitor', 'laptop', 'mouse',
                'remote', 'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink', 'refrigerator', 'book',
                'clock', 'vase', 'scissors', 'teddy bear', 'hair drier', 'toothbrush']


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modeling", "roi_heads", "roi_heads.py"))


class _Metadata:
    def __init__(self, name):
        self.name = name

    def set(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        return self


class _MetadataCatalog:
    def __init__(self):
        self._m: Dict[str, _Metadata] = {}

    def get(self, name):
        if name not in self._m:
            m = _Metadata(name)
            # data/datasets/voc/base_training.py:53-54, data/datasets/coco/base_training.py:97-98
            m.thing_classes = VOC_CLASSES if name.startswith("voc") or "pascal" in name else COCO_CLASSES
            self._m[name] = m
        return self._m[name]


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave as a package so that submodule imports resolve through sys.modules
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


class _StandInBoxHead(nn.Module):
    """Stands for Res5BoxHead (res5 + global mean, box_head.py:47-89), which is OUT OF SCOPE (dense cuDNN conv).
    A fixed random projection of the spatially averaged RoI feature keeps [R,C,P,P] -> [R,out] deterministic."""

    OUT = 2048

    def __init__(self, cfg, input_shape):
        super().__init__()
        from .d2.structures import ShapeSpec

        self._in = input_shape.channels
        self.proj = nn.Linear(self._in, self.OUT)
        g = torch.Generator().manual_seed(12345)
        with torch.no_grad():
            self.proj.weight.copy_(torch.randn(self.OUT, self._in, generator=g) * (1.0 / self._in ** 0.5))
            self.proj.bias.zero_()
        self._shape = ShapeSpec(channels=self.OUT, height=1, width=1)

    def forward(self, x):
        return torch.relu(self.proj(x.mean(dim=[2, 3])))

    @property
    def output_shape(self):
        return self._shape


class _StandInBoxHeadWithMask(_StandInBoxHead):
    """Stands for Res5BoxHeadWithMask (box_head.py:137-141): keeps a 7x7 map."""

    def __init__(self, cfg, input_shape):
        super().__init__(cfg, input_shape)
        from .d2.structures import ShapeSpec

        self._shape = ShapeSpec(channels=self.OUT, height=7, width=7)

    def forward(self, x):
        x = torch.nn.functional.avg_pool2d(x, 2)  # [R,C,7,7]
        return torch.relu(torch.einsum("oc,rchw->rohw", self.proj.weight, x))


_INSTALLED = False


def install() -> None:
    """Register the stand-in ``detectron2`` / ``fvcore`` module trees (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    from unit_b200.config import CfgNode  # config plumbing only

    from .d2 import modeling as M
    from .d2 import ops as O
    from .d2 import structures as S

    if not hasattr(np, "float"):  # fast_rcnn.py:250,428 use np.float("inf"), removed in NumPy >= 1.24
        np.float = float  # type: ignore[attr-defined]

    d2 = _mod("detectron2")
    _mod("detectron2.config", configurable=M.configurable, CfgNode=CfgNode)
    _mod("detectron2.utils")
    _mod("detectron2.utils.registry", Registry=M.Registry)
    _mod("detectron2.utils.events", get_event_storage=M.get_event_storage)
    _mod("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
    _mod("detectron2.utils.logger", log_first_n=lambda *a, **k: None)
    _mod("detectron2.utils.visualizer", Visualizer=object)
    _mod("detectron2.structures", Boxes=S.Boxes, Instances=S.Instances, ImageList=S.ImageList,
         pairwise_iou=S.pairwise_iou)
    _mod("detectron2.layers", Linear=M.Linear, Conv2d=M.Conv2d, ConvTranspose2d=M.ConvTranspose2d,
         ShapeSpec=S.ShapeSpec, cat=O.cat, nonzero_tuple=O.nonzero_tuple, batched_nms=O.batched_nms,
         get_norm=M.get_norm)
    _mod("detectron2.layers.batch_norm", FrozenBatchNorm2d=nn.Identity)

    class GeneralizedRCNN(nn.Module):
        pass

    _mod("detectron2.modeling", ROI_HEADS_REGISTRY=M.ROI_HEADS_REGISTRY, META_ARCH_REGISTRY=M.META_ARCH_REGISTRY,
         GeneralizedRCNN=GeneralizedRCNN)
    _mod("detectron2.modeling.matcher", Matcher=O.Matcher)
    _mod("detectron2.modeling.sampling", subsample_labels=O.subsample_labels)
    _mod("detectron2.modeling.box_regression", Box2BoxTransform=O.Box2BoxTransform)
    _mod("detectron2.modeling.poolers", ROIPooler=O.ROIPooler)
    _mod("detectron2.modeling.postprocessing", detector_postprocess=O.detector_postprocess)
    _mod("detectron2.modeling.roi_heads", StandardROIHeads=M.StandardROIHeads, Res5ROIHeads=M.Res5ROIHeads,
         build_box_head=M.build_box_head, ROI_HEADS_REGISTRY=M.ROI_HEADS_REGISTRY)
    _mod("detectron2.modeling.roi_heads.roi_heads", select_foreground_proposals=O.select_foreground_proposals,
         StandardROIHeads=M.StandardROIHeads)
    _mod("detectron2.modeling.roi_heads.fast_rcnn", FastRCNNOutputLayers=M.FastRCNNOutputLayers,
         FastRCNNOutputs=M.FastRCNNOutputs, fast_rcnn_inference=O.fast_rcnn_inference)
    _mod("detectron2.modeling.roi_heads.mask_head", ROI_MASK_HEAD_REGISTRY=M.ROI_MASK_HEAD_REGISTRY,
         MaskRCNNConvUpsampleHead=M.MaskRCNNConvUpsampleHead, build_mask_head=M.build_mask_head,
         mask_rcnn_inference=O.mask_rcnn_inference, mask_rcnn_loss=O.mask_rcnn_loss)
    _mod("detectron2.modeling.roi_heads.box_head", ROI_BOX_HEAD_REGISTRY=M.ROI_BOX_HEAD_REGISTRY)
    _mod("detectron2.modeling.proposal_generator", PROPOSAL_GENERATOR_REGISTRY=M.PROPOSAL_GENERATOR_REGISTRY,
         RPN=nn.Module, build_proposal_generator=None)
    _mod("detectron2.modeling.proposal_generator.proposal_utils",
         add_ground_truth_to_proposals=O.add_ground_truth_to_proposals)
    _mod("detectron2.modeling.backbone", Backbone=nn.Module)
    _mod("detectron2.modeling.backbone.build", BACKBONE_REGISTRY=M.BACKBONE_REGISTRY)
    _mod("detectron2.modeling.backbone.resnet", BottleneckBlock=nn.Module, ResNet=nn.Module,
         build_resnet_backbone=None)
    _mod("detectron2.evaluation")
    _mod("detectron2.evaluation.evaluator", inference_context=contextlib.nullcontext)
    _mod("detectron2.data", MetadataCatalog=_MetadataCatalog(), DatasetCatalog=object())
    _mod("fvcore")
    _mod("fvcore.nn", smooth_l1_loss=O.smooth_l1_loss, giou_loss=O.giou_loss)
    _mod("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)

    # stand-ins for the out-of-scope res5 box heads, registered under the reference's names
    for name, klass in (("Res5BoxHead", _StandInBoxHead), ("Res5BoxHeadWithMask", _StandInBoxHeadWithMask)):
        if name not in M.ROI_BOX_HEAD_REGISTRY:
            M.ROI_BOX_HEAD_REGISTRY._do_register(name, klass)
    _INSTALLED = True


def load_reference() -> types.SimpleNamespace:
    """Import the reference's RoI-stage modules verbatim; returns a namespace of the loaded modules."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install()
    if _PKG not in sys.modules:
        root = types.ModuleType(_PKG)
        root.__path__ = [REFERENCE_ROOT]
        sys.modules[_PKG] = root
        modeling = types.ModuleType(_PKG + ".modeling")
        modeling.__path__ = [os.path.join(REFERENCE_ROOT, "modeling")]
        sys.modules[_PKG + ".modeling"] = modeling
        rh = types.ModuleType(_PKG + ".modeling.roi_heads")
        rh.__path__ = [os.path.join(REFERENCE_ROOT, "modeling", "roi_heads")]
        sys.modules[_PKG + ".modeling.roi_heads"] = rh
    ns = types.SimpleNamespace()
    ns.matcher = importlib.import_module(_PKG + ".modeling.matcher")
    ns.weak = importlib.import_module(_PKG + ".modeling.roi_heads.weak_detector_fast_rcnn")
    ns.fast_rcnn = importlib.import_module(_PKG + ".modeling.roi_heads.fast_rcnn")
    ns.roi_heads = importlib.import_module(_PKG + ".modeling.roi_heads.roi_heads")
    ns.mask_head = importlib.import_module(_PKG + ".modeling.roi_heads.mask_head")
    return ns


def reference_cfg(yaml_rel: str, overrides=()):
    """Load one of the reference's YAMLs through the reference's own ``add_config`` (configs/default_config.py)."""
    install()
    from unit_b200.config import get_cfg

    spec = importlib.util.spec_from_file_location(_PKG + "_default_config",
                                                  os.path.join(REFERENCE_ROOT, "configs", "default_config.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = get_cfg()
    mod.add_config(cfg)
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, "configs", yaml_rel))
    cfg.merge_from_list(list(overrides))
    return cfg
