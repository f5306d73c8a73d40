"""yacs-compatible config node + the Detectron2 default subtree the RoI stage reads.

The reference builds its config as ``get_cfg()`` (Detectron2 defaults) +
``add_config(cfg)`` (configs/default_config.py:4-105) + ``merge_from_file(yaml)``
(scripts/train_VOC.py:26-31).  Neither yacs nor detectron2 is installed in this
image, so this module provides the same three calls with the same semantics
(``_BASE_`` inheritance, unknown keys rejected, ``KEY VALUE`` overrides, tuple
literals) so that the reference's configs/VOC and configs/COCO YAMLs load
unchanged.  Only plumbing lives here; nothing in this file touches the GPU.
"""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Iterable

import yaml

BASE_KEY = "_BASE_"


class CfgNode(dict):
    """Attribute-style nested dict with yacs merge semantics."""

    IMMUTABLE = "__immutable__"
    NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict: dict | None = None, new_allowed: bool = False):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        self.__dict__[CfgNode.NEW_ALLOWED] = new_allowed
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v, new_allowed) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # -- attribute protocol ------------------------------------------------
    def __getattr__(self, name: str) -> Any:
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        if self.__dict__[CfgNode.IMMUTABLE]:
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value)
        self[name] = value

    # -- yacs API ----------------------------------------------------------
    def is_frozen(self) -> bool:
        return self.__dict__[CfgNode.IMMUTABLE]

    def _set_immutable(self, flag: bool) -> None:
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self) -> None:
        self._set_immutable(True)

    def defrost(self) -> None:
        self._set_immutable(False)

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode(new_allowed=self.__dict__[CfgNode.NEW_ALLOWED])
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.__dict__[CfgNode.IMMUTABLE]
        return out

    def dump(self, **kwargs) -> str:
        def plain(n):
            if isinstance(n, CfgNode):
                return {k: plain(v) for k, v in n.items()}
            if isinstance(n, tuple):
                return list(n)
            return n

        return yaml.safe_dump(plain(self), **kwargs)

    @staticmethod
    def load_yaml_with_base(filename: str) -> dict:
        """Load a YAML file, resolving ``_BASE_`` recursively (relative to the file)."""
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) or {}

        def merge_a_into_b(a: dict, b: dict) -> None:
            for k, v in a.items():
                if isinstance(v, dict) and k in b:
                    assert isinstance(b[k], dict), f"Cannot inherit key '{k}' from base!"
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if BASE_KEY in cfg:
            base_file = cfg.pop(BASE_KEY)
            if base_file.startswith("~"):
                base_file = os.path.expanduser(base_file)
            if not base_file.startswith("/"):
                base_file = os.path.join(os.path.dirname(filename), base_file)
            base = CfgNode.load_yaml_with_base(base_file)
            merge_a_into_b(cfg, base)
            return base
        return cfg

    def merge_from_file(self, cfg_filename: str) -> None:
        loaded = CfgNode.load_yaml_with_base(cfg_filename)
        _merge_into(CfgNode(loaded), self, [])

    def merge_from_other_cfg(self, other: "CfgNode") -> None:
        _merge_into(other, self, [])

    def merge_from_list(self, cfg_list: Iterable[Any]) -> None:
        cfg_list = list(cfg_list)
        assert len(cfg_list) % 2 == 0, f"Override list has odd length: {cfg_list}"
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            parts = full_key.split(".")
            d = self
            for sub in parts[:-1]:
                if sub not in d:
                    raise KeyError(f"Non-existent key: {full_key}")
                d = d[sub]
            last = parts[-1]
            if last not in d:
                raise KeyError(f"Non-existent key: {full_key}")
            value = _decode(v)
            d[last] = _coerce(value, d[last], full_key)


def _decode(v: Any) -> Any:
    """yacs decodes string overrides with literal_eval ("(1,2)" -> tuple)."""
    if isinstance(v, dict) and not isinstance(v, CfgNode):
        return CfgNode(v)
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _coerce(replacement: Any, original: Any, full_key: str) -> Any:
    """yacs type rule: same type, or the tuple<->list / None / int->float casts."""
    ot, rt = type(original), type(replacement)
    if rt is ot or original is None or replacement is None:
        return replacement
    if ot is tuple and rt is list:
        return tuple(replacement)
    if ot is list and rt is tuple:
        return list(replacement)
    if ot is float and rt is int:
        return float(replacement)
    if ot is str and rt is str:
        return replacement
    raise ValueError(
        f"Type mismatch ({ot} vs. {rt}) with values ({original} vs. {replacement}) for config key: {full_key}"
    )


def _merge_into(a: CfgNode, b: CfgNode, key_list: list) -> None:
    for k, v_ in a.items():
        full_key = ".".join(key_list + [k])
        v = _decode(copy.deepcopy(v_))
        if k in b:
            if isinstance(v, CfgNode) and isinstance(b[k], CfgNode):
                _merge_into(v, b[k], key_list + [k])
            else:
                b[k] = _coerce(v, b[k], full_key)
        elif b.__dict__[CfgNode.NEW_ALLOWED]:
            b[k] = v
        else:
            raise KeyError(f"Non-existent config key: {full_key}")


CN = CfgNode


def get_cfg() -> CfgNode:
    """Detectron2 ``get_cfg()`` defaults for every key the reference's YAMLs or the RoI stage read.

    Values are Detectron2 v0.3/v0.4 ``config/defaults.py`` defaults ("Config / flags").
    """
    _C = CN()
    _C.VERSION = 2

    _C.MODEL = CN()
    _C.MODEL.LOAD_PROPOSALS = False
    _C.MODEL.MASK_ON = False
    _C.MODEL.KEYPOINT_ON = False
    _C.MODEL.DEVICE = "cuda"
    _C.MODEL.META_ARCHITECTURE = "GeneralizedRCNN"
    _C.MODEL.WEIGHTS = ""
    _C.MODEL.PIXEL_MEAN = [103.530, 116.280, 123.675]
    _C.MODEL.PIXEL_STD = [1.0, 1.0, 1.0]

    _C.INPUT = CN()
    _C.INPUT.MIN_SIZE_TRAIN = (800,)
    _C.INPUT.MIN_SIZE_TRAIN_SAMPLING = "choice"
    _C.INPUT.MAX_SIZE_TRAIN = 1333
    _C.INPUT.MIN_SIZE_TEST = 800
    _C.INPUT.MAX_SIZE_TEST = 1333
    _C.INPUT.RANDOM_FLIP = "horizontal"
    _C.INPUT.CROP = CN({"ENABLED": False, "TYPE": "relative_range", "SIZE": [0.9, 0.9]})
    _C.INPUT.FORMAT = "BGR"
    _C.INPUT.MASK_FORMAT = "polygon"

    _C.DATASETS = CN()
    _C.DATASETS.TRAIN = ()
    _C.DATASETS.PROPOSAL_FILES_TRAIN = ()
    _C.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TRAIN = 2000
    _C.DATASETS.TEST = ()
    _C.DATASETS.PROPOSAL_FILES_TEST = ()
    _C.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST = 1000

    _C.DATALOADER = CN()
    _C.DATALOADER.NUM_WORKERS = 4
    _C.DATALOADER.ASPECT_RATIO_GROUPING = True
    _C.DATALOADER.SAMPLER_TRAIN = "TrainingSampler"
    _C.DATALOADER.REPEAT_THRESHOLD = 0.0
    _C.DATALOADER.FILTER_EMPTY_ANNOTATIONS = True

    _C.MODEL.BACKBONE = CN()
    _C.MODEL.BACKBONE.NAME = "build_resnet_backbone"
    _C.MODEL.BACKBONE.FREEZE_AT = 2

    _C.MODEL.FPN = CN({"IN_FEATURES": [], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"})

    _C.MODEL.PROPOSAL_GENERATOR = CN()
    _C.MODEL.PROPOSAL_GENERATOR.NAME = "RPN"
    _C.MODEL.PROPOSAL_GENERATOR.MIN_SIZE = 0

    _C.MODEL.ANCHOR_GENERATOR = CN()
    _C.MODEL.ANCHOR_GENERATOR.NAME = "DefaultAnchorGenerator"
    _C.MODEL.ANCHOR_GENERATOR.SIZES = [[32, 64, 128, 256, 512]]
    _C.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS = [[0.5, 1.0, 2.0]]
    _C.MODEL.ANCHOR_GENERATOR.ANGLES = [[-90, 0, 90]]
    _C.MODEL.ANCHOR_GENERATOR.OFFSET = 0.0

    _C.MODEL.RPN = CN()
    _C.MODEL.RPN.HEAD_NAME = "StandardRPNHead"
    _C.MODEL.RPN.IN_FEATURES = ["res4"]
    _C.MODEL.RPN.BOUNDARY_THRESH = -1
    _C.MODEL.RPN.IOU_THRESHOLDS = [0.3, 0.7]
    _C.MODEL.RPN.IOU_LABELS = [0, -1, 1]
    _C.MODEL.RPN.BATCH_SIZE_PER_IMAGE = 256
    _C.MODEL.RPN.POSITIVE_FRACTION = 0.5
    _C.MODEL.RPN.BBOX_REG_LOSS_TYPE = "smooth_l1"
    _C.MODEL.RPN.BBOX_REG_LOSS_WEIGHT = 1.0
    _C.MODEL.RPN.BBOX_REG_WEIGHTS = (1.0, 1.0, 1.0, 1.0)
    _C.MODEL.RPN.SMOOTH_L1_BETA = 0.0
    _C.MODEL.RPN.LOSS_WEIGHT = 1.0
    _C.MODEL.RPN.PRE_NMS_TOPK_TRAIN = 12000
    _C.MODEL.RPN.PRE_NMS_TOPK_TEST = 6000
    _C.MODEL.RPN.POST_NMS_TOPK_TRAIN = 2000
    _C.MODEL.RPN.POST_NMS_TOPK_TEST = 1000
    _C.MODEL.RPN.NMS_THRESH = 0.7

    _C.MODEL.ROI_HEADS = CN()
    _C.MODEL.ROI_HEADS.NAME = "Res5ROIHeads"
    _C.MODEL.ROI_HEADS.NUM_CLASSES = 80
    _C.MODEL.ROI_HEADS.IN_FEATURES = ["res4"]
    _C.MODEL.ROI_HEADS.IOU_THRESHOLDS = [0.5]
    _C.MODEL.ROI_HEADS.IOU_LABELS = [0, 1]
    _C.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE = 512
    _C.MODEL.ROI_HEADS.POSITIVE_FRACTION = 0.25
    _C.MODEL.ROI_HEADS.SCORE_THRESH_TEST = 0.05
    _C.MODEL.ROI_HEADS.NMS_THRESH_TEST = 0.5
    _C.MODEL.ROI_HEADS.PROPOSAL_APPEND_GT = True

    _C.MODEL.ROI_BOX_HEAD = CN()
    _C.MODEL.ROI_BOX_HEAD.NAME = ""
    _C.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE = "smooth_l1"
    _C.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT = 1.0
    _C.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS = (10.0, 10.0, 5.0, 5.0)
    _C.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA = 0.0
    _C.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION = 14
    _C.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO = 0
    _C.MODEL.ROI_BOX_HEAD.POOLER_TYPE = "ROIAlignV2"
    _C.MODEL.ROI_BOX_HEAD.NUM_FC = 0
    _C.MODEL.ROI_BOX_HEAD.FC_DIM = 1024
    _C.MODEL.ROI_BOX_HEAD.NUM_CONV = 0
    _C.MODEL.ROI_BOX_HEAD.CONV_DIM = 256
    _C.MODEL.ROI_BOX_HEAD.NORM = ""
    _C.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG = False
    _C.MODEL.ROI_BOX_HEAD.TRAIN_ON_PRED_BOXES = False

    _C.MODEL.ROI_MASK_HEAD = CN()
    _C.MODEL.ROI_MASK_HEAD.NAME = "MaskRCNNConvUpsampleHead"
    _C.MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION = 14
    _C.MODEL.ROI_MASK_HEAD.POOLER_SAMPLING_RATIO = 0
    _C.MODEL.ROI_MASK_HEAD.NUM_CONV = 0
    _C.MODEL.ROI_MASK_HEAD.CONV_DIM = 256
    _C.MODEL.ROI_MASK_HEAD.NORM = ""
    _C.MODEL.ROI_MASK_HEAD.CLS_AGNOSTIC_MASK = False
    _C.MODEL.ROI_MASK_HEAD.POOLER_TYPE = "ROIAlignV2"

    _C.MODEL.ROI_KEYPOINT_HEAD = CN({"NAME": "KRCNNConvDeconvUpsampleHead", "NUM_KEYPOINTS": 17})

    _C.MODEL.RESNETS = CN()
    _C.MODEL.RESNETS.DEPTH = 50
    _C.MODEL.RESNETS.OUT_FEATURES = ["res4"]
    _C.MODEL.RESNETS.NUM_GROUPS = 1
    _C.MODEL.RESNETS.NORM = "FrozenBN"
    _C.MODEL.RESNETS.WIDTH_PER_GROUP = 64
    _C.MODEL.RESNETS.STRIDE_IN_1X1 = True
    _C.MODEL.RESNETS.RES5_DILATION = 1
    _C.MODEL.RESNETS.RES2_OUT_CHANNELS = 256
    _C.MODEL.RESNETS.STEM_OUT_CHANNELS = 64
    _C.MODEL.RESNETS.DEFORM_ON_PER_STAGE = [False, False, False, False]
    _C.MODEL.RESNETS.DEFORM_MODULATED = False
    _C.MODEL.RESNETS.DEFORM_NUM_GROUPS = 1

    _C.SOLVER = CN()
    _C.SOLVER.LR_SCHEDULER_NAME = "WarmupMultiStepLR"
    _C.SOLVER.MAX_ITER = 40000
    _C.SOLVER.BASE_LR = 0.001
    _C.SOLVER.MOMENTUM = 0.9
    _C.SOLVER.NESTEROV = False
    _C.SOLVER.WEIGHT_DECAY = 0.0001
    _C.SOLVER.WEIGHT_DECAY_NORM = 0.0
    _C.SOLVER.GAMMA = 0.1
    _C.SOLVER.STEPS = (30000,)
    _C.SOLVER.WARMUP_FACTOR = 1.0 / 1000
    _C.SOLVER.WARMUP_ITERS = 1000
    _C.SOLVER.WARMUP_METHOD = "linear"
    _C.SOLVER.CHECKPOINT_PERIOD = 5000
    _C.SOLVER.IMS_PER_BATCH = 16
    _C.SOLVER.BIAS_LR_FACTOR = 1.0
    _C.SOLVER.WEIGHT_DECAY_BIAS = 0.0001
    _C.SOLVER.CLIP_GRADIENTS = CN({"ENABLED": False, "CLIP_TYPE": "value", "CLIP_VALUE": 1.0, "NORM_TYPE": 2.0})

    _C.TEST = CN()
    _C.TEST.EXPECTED_RESULTS = []
    _C.TEST.EVAL_PERIOD = 0
    _C.TEST.DETECTIONS_PER_IMAGE = 100
    _C.TEST.AUG = CN({"ENABLED": False, "MIN_SIZES": (400, 500, 600, 700, 800, 900, 1000, 1100, 1200),
                      "MAX_SIZE": 4000, "FLIP": True})
    _C.TEST.PRECISE_BN = CN({"ENABLED": False, "NUM_ITER": 200})

    _C.OUTPUT_DIR = "./output"
    _C.SEED = -1
    _C.CUDNN_BENCHMARK = False
    _C.VIS_PERIOD = 0
    _C.GLOBAL = CN({"HACK": 1.0})
    return _C


def add_config(cfg: CfgNode) -> None:
    """UniT's extra keys, same names and defaults as the reference (configs/default_config.py:4-105)."""
    _C = cfg
    _C.MODEL.BACKBONE.DILATED = False
    _C.MODEL.BACKBONE.FREEZE_CONVS = 0

    _C.MODEL.FREEZE_LAYERS = CN()
    for k in ("ROI_HEADS", "META_ARCH", "FAST_RCNN", "BOX_HEAD", "MASK_HEAD"):
        _C.MODEL.FREEZE_LAYERS[k] = []

    rh = _C.MODEL.ROI_HEADS
    rh.EMBEDDING_PATH = ""
    rh.FINETUNE_TERMS = CN()
    rh.FINETUNE_TERMS.CLASSIFIER = ["lingual", "visual"]
    rh.FINETUNE_TERMS.BBOX = ["lingual", "visual"]
    rh.FINETUNE_TERMS.MASK = ["lingual", "visual"]
    rh.WEAK_CLASSIFIER_PROPOSAL_DIVISOR = 1
    rh.MULTI_BOX_HEAD = False
    _C.MODEL.PROPOSAL_GENERATOR.WEAK_RPN_SCORE_TRESHOLD = 0.99
    rh.TRAIN_USING_WEAK = False
    rh.TRAIN_PROPOSAL_REGRESSOR = True
    rh.WEAK_PROPOSAL_DIVISOR = 1.0

    rh.FAST_RCNN = CN()
    rh.FAST_RCNN.NAME = "SupervisedDetectorOutputsBase"
    rh.FAST_RCNN.MODE = "Pre_Softmax"
    wd = rh.FAST_RCNN.WEAK_DETECTOR = CN()
    wd.NAME = "WeakDetectorOutputsBase"
    wd.NUM_KMEANS_CLUSTER = 3
    wd.GRAPH_IOU_THRESHOLD = 0.4
    wd.MAX_PC_NUM = 5
    wd.WEAK_LOSS_MULTIPLIER = 1.0
    wd.OICR_ITER = 3
    wd.FG_THRESHOLD = 0.5
    wd.BG_THRESHOLD = 0.1
    wd.MIL_MULTIPLIER = 1.0
    wd.DETECTOR_TEMP = 1.0
    wd.CLASSIFIER_TEMP = 1.0
    wd.REGRESSION_BRANCH = False
    wd.TYPE = "OICR"
    wd.OICR_REGRESSION_BRANCH = False

    va = rh.VISUAL_ATTENTION_HEAD = CN()
    va.NAME = "MeanSimilarity"
    va.IN_FEATURES = ["res4"]
    va.POOLER_RESOLUTION = 14
    va.POOLER_SAMPLING_RATIO = 0
    va.POOLER_TYPE = "ROIAlignV2"
    va.VISUAL_SIMILARITY_THRESHOLD = 0.02
    va.SIMILARITY_COMBINATION = "Sum"
    va.TOPK = 5

    ds = _C.DATASETS
    ds.META_TRAIN = ""
    ds.META_VAL = ""
    ds.META_SHOTS = []
    ds.META_VAL_SHOTS = 1
    ds.BASE_META = ""
    ds.BASE_META_SHOTS = 50
    ds.MODE = "base"
    ds.CLASSIFIER_DATAROOT = ""
    ds.CLASSIFIER_TRAIN = ()
    ds.ONLY_NOVEL_CLASSIFIER_DATA = False
    _C.INPUT.META_MIN_SIZE = 224
    _C.INPUT.META_MAX_SIZE = 480
    _C.INPUT.RESIZE_META = True
    _C.INPUT.NORMALIZE_IMAGES = False

    fs = ds.FEWSHOT = CN()
    fs.TYPE = "VOC"
    fs.NUM_SHOTS = 5
    fs.IS_ZERO_SHOT = False
    fs.SPLIT_ID = 1
    fs.BASE_CLASSES_ID = [0, 1, 3, 4, 6, 7, 8, 10, 11, 12, 14, 15, 16, 18, 19]
    fs.NOVEL_CLASSES_ID = [2, 5, 9, 13, 17]
    ds.WEAK_CLASSIFIER_MUTLIPLIER = 1.0
    ds.WEAK_CLASSIFIER_SAMPLE_NUM = -1
    ds.NUM_SAMPLES = 120
    ds.BASE_MULTIPLIER = -1.0
    ds.NOVEL_MULTIPLER = 0.0
    ds.SAMPLE_MULTIPLIER = 3
    ds.OVER_SAMPLE = False
    ds.SAMPLE_WITH_REPLACEMENT = False
    ds.SAMPLE_SEED = 0
    ds.PROPOSAL_FILES_CLASSIFIER_TRAIN = ()

    _C.TEST.MIN_EVAL_PERIOD = 0
    _C.TEST.AUG = CN()
    _C.TEST.AUG.ENABLED = True
    _C.TEST.AUG.MIN_SIZES = (480, 576, 688, 864, 1200)
    _C.TEST.AUG.MAX_SIZE = 2000
    _C.TEST.AUG.FLIP = True

    _C.SOLVER.REFINEMENT_LR_FACTOR = 1.0
    _C.SOLVER.DELTA_LR_FACTOR = 1.0
    _C.SOLVER.MIL_LR_FACTOR = 1.0
    _C.SOLVER.TRAIN_ONLY_WEAK = -1


def load_cfg(yaml_path: str | None = None, overrides: Iterable[Any] = ()) -> CfgNode:
    """``get_cfg(); add_config(cfg); merge_from_file; merge_from_list`` (scripts/train_VOC.py:26-33)."""
    cfg = get_cfg()
    add_config(cfg)
    if yaml_path:
        cfg.merge_from_file(yaml_path)
    cfg.merge_from_list(list(overrides))
    return cfg
