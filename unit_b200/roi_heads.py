"""RoI heads registered in ``ROI_HEADS_REGISTRY`` under the reference's names
(modeling/roi_heads/roi_heads.py:28,134,488,593,646,824): same ``from_config`` keys, same ``forward`` keywords,
same outputs, same parameter names -- so configs/VOC and configs/COCO YAMLs select them unchanged.

The per-image Python loops of the reference (label_and_sample_proposals, fast_rcnn_inference) and the ATen
sequences of get_similarity_matrices / predictor.forward are replaced by batched CUDA kernels (see layers.py,
predictors.py).  The res5 box head stays stock PyTorch (out of scope).
"""
from __future__ import annotations

import inspect
import logging
from typing import Dict, List, Optional

import torch
from torch import nn

from . import layers, ops
from .layers import Matcher, ROIPooler, label_and_sample, select_foreground_proposals
from .predictors import FusedSimilarity, _freeze, build_fastrcnn_head
from .registry import (ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY, ROI_MASK_HEAD_REGISTRY, configurable)
from .structures import Boxes, Instances, ShapeSpec

# modeling/roi_heads/roi_heads.py:191 -- COCO-80 name -> row of the glove_mean embedding table
_COCO = ['person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light',
         'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow',
         'elephant', 'bear', 'zebra', 'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee',
         'skis', 'snowboard', 'sports ball', 'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard',
         'tennis racket', 'bottle', 'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple',
         'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair', 'couch',
         'potted plant', 'bed', 'dining table', 'toilet', 'tv', 'laptop', 'mouse', 'remote', 'keyboard', 'cell phone',
         'microwave', 'oven', 'toaster', 'sink', 'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear',
         'hair drier', 'toothbrush']
_COCO_ID = {n: i for i, n in enumerate(_COCO)}
_VOC_TO_COCO = {'aeroplane': 'airplane', 'diningtable': 'dining table', 'motorbike': 'motorcycle',
                'pottedplant': 'potted plant', 'sofa': 'couch', 'tvmonitor': 'tv'}
VOC_CLASSES = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable",
               "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor"]
COCO_CLASSES_VOC_SPELLING = [{v: k for k, v in _VOC_TO_COCO.items()}.get(n, n) for n in _COCO]


def thing_classes_for(dataset_name: str) -> List[str]:
    """``MetadataCatalog.get(name).thing_classes`` for the reference's datasets (data/datasets/*/base_training.py);
    the real catalog is consulted first when Detectron2 is importable."""
    try:
        from detectron2.data import MetadataCatalog  # type: ignore

        tc = getattr(MetadataCatalog.get(dataset_name), "thing_classes", None)
        if tc:
            return list(tc)
    except Exception:
        pass
    return VOC_CLASSES if (dataset_name.startswith("voc") or "pascal" in dataset_name) else COCO_CLASSES_VOC_SPELLING


def build_box_head(cfg, input_shape):
    return ROI_BOX_HEAD_REGISTRY.get(cfg.MODEL.ROI_BOX_HEAD.NAME)(cfg, input_shape)


def build_mask_head(cfg, input_shape):
    return ROI_MASK_HEAD_REGISTRY.get(cfg.MODEL.ROI_MASK_HEAD.NAME)(cfg, input_shape)


class ROIHeads(nn.Module):
    """[D2] ROIHeads: sampling configuration + ``label_and_sample_proposals``."""

    @configurable
    def __init__(self, *, num_classes, batch_size_per_image, positive_fraction, proposal_matcher,
                 proposal_append_gt=True):
        super().__init__()
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.num_classes = num_classes
        self.proposal_matcher = proposal_matcher
        self.proposal_append_gt = proposal_append_gt
        self.sampling_generator: Optional[torch.Generator] = None  # host RNG of the two randperm draws per image

    @classmethod
    def from_config(cls, cfg):
        rh = cfg.MODEL.ROI_HEADS
        return {
            "batch_size_per_image": rh.BATCH_SIZE_PER_IMAGE,
            "positive_fraction": rh.POSITIVE_FRACTION,
            "num_classes": rh.NUM_CLASSES,
            "proposal_append_gt": rh.PROPOSAL_APPEND_GT,
            "proposal_matcher": Matcher(rh.IOU_THRESHOLDS, rh.IOU_LABELS, allow_low_quality_matches=False,
                                        return_vals=False),
        }

    @torch.no_grad()
    def label_and_sample_proposals(self, proposals: List[Instances], targets: List[Instances]) -> List[Instances]:
        if self.proposal_append_gt:
            proposals = layers.add_ground_truth_to_proposals([t.gt_boxes for t in targets], proposals)
        out, _, _ = label_and_sample(
            proposals, targets, num_classes=self.num_classes, batch_size_per_image=self.batch_size_per_image,
            positive_fraction=self.positive_fraction, thresholds=self.proposal_matcher.user_thresholds,
            labels=self.proposal_matcher.labels, sample=True, generator=self.sampling_generator)
        return out


class StandardROIHeads(ROIHeads):
    """[D2] StandardROIHeads; UniT overrides ``_init_box_head`` / ``_init_mask_head``."""

    @configurable
    def __init__(self, *, box_in_features, box_pooler, box_head, box_predictor, mask_in_features=None,
                 mask_pooler=None, mask_head=None, keypoint_in_features=None, keypoint_pooler=None,
                 keypoint_head=None, train_on_pred_boxes=False, **kwargs):
        super().__init__(**kwargs)
        self.in_features = self.box_in_features = box_in_features
        self.box_pooler = box_pooler
        self.box_head = box_head
        self.box_predictor = box_predictor
        self.mask_on = mask_in_features is not None
        if self.mask_on:
            self.mask_in_features = mask_in_features
            self.mask_pooler = mask_pooler
            self.mask_head = mask_head
        self.keypoint_on = keypoint_in_features is not None
        if self.keypoint_on:
            raise NotImplementedError("keypoint heads are not part of UniT")
        self.train_on_pred_boxes = train_on_pred_boxes

    @classmethod
    def from_config(cls, cfg, input_shape):
        ret = super().from_config(cfg)
        ret["train_on_pred_boxes"] = cfg.MODEL.ROI_BOX_HEAD.TRAIN_ON_PRED_BOXES
        if inspect.ismethod(cls._init_box_head):
            ret.update(cls._init_box_head(cfg, input_shape))
        if inspect.ismethod(cls._init_mask_head):
            ret.update(cls._init_mask_head(cfg, input_shape))
        return ret

    @classmethod
    def _init_box_head(cls, cfg, input_shape):
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        res = cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        in_channels = [input_shape[f].channels for f in in_features]
        assert len(set(in_channels)) == 1, in_channels
        box_pooler = ROIPooler(output_size=res, scales=tuple(1.0 / input_shape[k].stride for k in in_features),
                               sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                               pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE)
        box_head = build_box_head(cfg, ShapeSpec(channels=in_channels[0], height=res, width=res))
        return {"box_in_features": in_features, "box_pooler": box_pooler, "box_head": box_head,
                "box_predictor": None}

    @classmethod
    def _init_mask_head(cls, cfg, input_shape):
        return {}

    def _forward_mask(self, features, instances):
        return {} if self.training else instances

    def _forward_keypoint(self, features, instances):
        return {} if self.training else instances

    def forward_with_given_boxes(self, features, instances):
        assert not self.training
        assert instances[0].has("pred_boxes") and instances[0].has("pred_classes")
        return self._forward_mask(features, instances)


class WSROIHead(StandardROIHeads):
    """roi_heads.py:134-486 without the meta-attention branch (``visual_attention_head`` is only used by the
    meta-learning ``WSROIHead`` proper, which no shipped YAML selects)."""

    @configurable
    def __init__(self, *, box_in_features, box_pooler, box_head, box_predictor, mask_in_features=None,
                 mask_pooler=None, mask_head=None, keypoint_in_features=None, keypoint_pooler=None,
                 keypoint_head=None, weak_box_head=None, visual_attention_head=None, train_on_pred_boxes=False,
                 freeze_layers=(), **kwargs):
        self._base_classes_id = list(kwargs.pop("base_classes_id"))
        self._novel_classes_id = list(kwargs.pop("novel_classes_id"))
        self.train_dataset_name = kwargs.pop("train_dataset_name")
        self.weak_divisor = kwargs.pop("weak_divisor")
        self.terms = dict(kwargs.pop("terms"))
        self.load_proposals = kwargs.pop("load_proposals")
        self.visual_threshold = kwargs.pop("visual_threshold")
        self.similarity_combination = kwargs.pop("similarity_combination")
        self.topk = kwargs.pop("topk")
        super().__init__(box_in_features=box_in_features, box_pooler=box_pooler, box_head=box_head,
                         box_predictor=box_predictor, mask_in_features=mask_in_features, mask_pooler=mask_pooler,
                         mask_head=mask_head, keypoint_in_features=keypoint_in_features,
                         keypoint_pooler=keypoint_pooler, keypoint_head=keypoint_head,
                         train_on_pred_boxes=train_on_pred_boxes, **kwargs)
        flat = [y for x in self.terms.values() for y in x]
        self.compute_similarity = {"lingual": "lingual" in flat, "visual": "visual" in flat}
        if visual_attention_head is not None:
            raise NotImplementedError("the meta-attention WSROIHead is out of scope; use the *NoMeta heads")
        self.weak_box_head = weak_box_head
        self._class_mappings()
        _freeze(self, freeze_layers)
        self._spec_cache: Dict[str, ops.TransferSpec] = {}

    @classmethod
    def from_config(cls, cfg, input_shape):
        ret = super().from_config(cfg, input_shape)
        rh = cfg.MODEL.ROI_HEADS
        ret["freeze_layers"] = cfg.MODEL.FREEZE_LAYERS.ROI_HEADS
        ret["base_classes_id"] = cfg.DATASETS.FEWSHOT.BASE_CLASSES_ID
        ret["novel_classes_id"] = cfg.DATASETS.FEWSHOT.NOVEL_CLASSES_ID
        ret["train_dataset_name"] = cfg.DATASETS.TRAIN[0]
        ret["weak_divisor"] = rh.WEAK_CLASSIFIER_PROPOSAL_DIVISOR
        ret["terms"] = {"cls": rh.FINETUNE_TERMS.CLASSIFIER, "bbox": rh.FINETUNE_TERMS.BBOX}
        if cfg.MODEL.MASK_ON:
            ret["terms"]["seg"] = rh.FINETUNE_TERMS.MASK
        ret["load_proposals"] = cfg.MODEL.LOAD_PROPOSALS
        ret["visual_threshold"] = rh.VISUAL_ATTENTION_HEAD.VISUAL_SIMILARITY_THRESHOLD
        ret["similarity_combination"] = rh.VISUAL_ATTENTION_HEAD.SIMILARITY_COMBINATION
        ret["topk"] = rh.VISUAL_ATTENTION_HEAD.TOPK
        return ret

    @classmethod
    def _init_box_head(cls, cfg, input_shape):
        ret = super()._init_box_head(cfg, input_shape)
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        res = cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        in_channels = input_shape[in_features[0]].channels
        ret["box_predictor"] = build_fastrcnn_head(cfg, ret["box_head"].output_shape)
        if cfg.MODEL.ROI_HEADS.MULTI_BOX_HEAD:
            ret["weak_box_head"] = build_box_head(cfg, ShapeSpec(channels=in_channels, height=res, width=res))
        return ret

    def _class_mappings(self) -> None:
        """roi_heads.py:190-216."""
        names = thing_classes_for(self.train_dataset_name)
        idx = [_COCO_ID[_VOC_TO_COCO.get(n, n)] for n in names]
        self._coco_indexer_tensor = torch.tensor(idx).long()
        self._base_classes_tensor = torch.tensor(self._base_classes_id).long()
        self._novel_classes_tensor = torch.tensor(self._novel_classes_id).long()

    def move_mappings_to_gpu(self) -> None:
        if not self._coco_indexer_tensor.is_cuda:
            device = next(self.box_predictor.parameters()).device
            self._coco_indexer_tensor = self._coco_indexer_tensor.to(device)
            self._base_classes_tensor = self._base_classes_tensor.to(device)
            self._novel_classes_tensor = self._novel_classes_tensor.to(device)

    # -- similarity ------------------------------------------------------------------------------------------
    @staticmethod
    def _term_k(tl, tag: str) -> int:
        """roi_heads.py:274,285,296,307: k of the first term CONTAINING ``tag`` (the reference matches by substring, so
        ``WTopK-k`` also switches the ``TopK`` branch on)."""
        return int([t for t in tl if tag in t][0].split("-")[1])

    def _weight_terms(self, tl, device) -> torch.Tensor:
        """TopK / WTopK / LSDA (roi_heads.py:273-305): class-level [Nn,B] terms from the mean OICR predictor weights,
        unweighted sum (every term has the same 1/len(terms) weight).  A handful of tiny ATen ops, once per model."""
        cw = self.box_predictor.weak_detector_head.mean_oicr_weight()[0].detach().to(device)
        base_w = cw.index_select(0, self._base_classes_tensor)
        novel_w = cw.index_select(0, self._novel_classes_tensor)
        Nn, B = novel_w.shape[0], base_w.shape[0]
        out = torch.zeros(Nn, B, device=device)
        if any("TopK" in t for t in tl):
            _, idx = torch.topk(novel_w @ base_w.t(), self._term_k(tl, "TopK"), dim=-1)
            t = torch.zeros(Nn, B, device=device).scatter(1, idx, 1.0)
            out = out + t / t.sum(-1, keepdim=True)
        if any("WTopK" in t for t in tl):
            top, idx = torch.topk(novel_w @ base_w.t(), self._term_k(tl, "WTopK"), dim=-1)
            t = torch.zeros(Nn, B, device=device).scatter(1, idx, top)
            out = out + t / t.sum(-1, keepdim=True)
        if any("LSDA" in t for t in tl):
            dist = torch.norm(novel_w.unsqueeze(1) - base_w.unsqueeze(0), dim=-1)
            _, idx = torch.topk(dist, self._term_k(tl, "LSDA"), dim=-1, largest=False)
            t = torch.zeros(Nn, B, device=device).scatter(1, idx, 1.0)
            out = out + t / t.sum(-1, keepdim=True)
        return out

    def _transfer_spec(self, device) -> ops.TransferSpec:
        """Class-level part of ``get_similarity_matrices`` (roi_heads.py:266-334), reduced once per model/device:
        static[h] = sum of the class-level terms x their 1/len(terms) weight, wv[h] = weight of 'visual',
        ``spec.wk[h]`` = (weight, k) of a 'VisualK-k' term (added per RoI in ``get_similarity_matrices``)."""
        weight_terms = any(("TopK" in t) or ("LSDA" in t) for tl in self.terms.values() for t in tl)
        key = str(device)
        if weight_terms:  # TopK / WTopK / LSDA read the OICR weights: rebuild when they change (base training)
            key = (key,) + tuple(l.weight._version for l in self.box_predictor.weak_detector_head.oicr_predictors)
        if key in self._spec_cache:
            return self._spec_cache[key]
        self.move_mappings_to_gpu()
        Nn, B = len(self._novel_classes_id), len(self._base_classes_id)
        soft = None
        if self.compute_similarity["lingual"]:
            _, soft = ops.lingual_similarity(self.box_predictor.embeddings.weight, self._coco_indexer_tensor,
                                             self._base_classes_tensor, self._novel_classes_tensor)
        static, wv, norm, wk = {}, {}, {}, {}
        for head, tl in self.terms.items():
            tl = list(tl)
            if self.similarity_combination == "Sum":
                w = 1.0 / len(tl) if len(tl) else 0.0
                st = torch.zeros(Nn, B, device=device)
                if "lingual" in tl:
                    st = st + w * soft
                if any(("TopK" in t) or ("LSDA" in t) for t in tl):
                    st = st + w * self._weight_terms(tl, device)
                wv[head] = w if "visual" in tl else 0.0
                if any("VisualK" in t for t in tl):
                    if "visual" in tl:
                        raise ValueError("'VisualK-k' together with 'visual' makes the reference build a 4-D similarity "
                                         "(roi_heads.py:315-317) that its own transfer (fast_rcnn.py:407) rejects")
                    wk[head] = (w, self._term_k(tl, "VisualK"))
                if "Average" in tl:  # fill_(1.) overrides every other term (roi_heads.py:319-321)
                    st = torch.full((Nn, B), 1.0 / B, device=device)
                    wv[head] = 0.0
                    wk.pop(head, None)
                if len(tl) > 0 and "None" not in tl:
                    norm[head] = 1
                else:  # 0.0 * similarity (roi_heads.py:324-325)
                    st, wv[head], norm[head] = torch.zeros(Nn, B, device=device), 0.0, 0
                    wk.pop(head, None)
                static[head] = st
            else:
                # product mode starts from zeros (roi_heads.py:268,327-332): softmax(0) = uniform when any term
                static[head] = torch.full((Nn, B), (1.0 / B) if len(tl) > 0 else 0.0, device=device)
                wv[head], norm[head] = 0.0, 0
        spec = ops.TransferSpec(self.num_classes, self._base_classes_id, self._novel_classes_id, device, static, wv,
                                norm, self.visual_threshold)
        spec.wk = wk
        if len(self._spec_cache) > 8:
            self._spec_cache.clear()
        self._spec_cache[key] = spec
        return spec

    def _visualk_spec(self, spec: ops.TransferSpec, logits: torch.Tensor) -> ops.TransferSpec:
        """'VisualK-k' (roi_heads.py:306-315): per-RoI top-k of the renormalised base-class probabilities (softmax over
        the K foreground logits), added to the class-level terms as a per-RoI static [R,Nn,B] block."""
        static, per_roi = dict(spec.static), spec.static_per_roi
        order = {"cls": 0, "bbox": 1, "seg": 2}
        for head, (w, k) in spec.wk.items():
            cw = torch.softmax(logits.narrow(1, 0, self.num_classes), -1).index_select(1, self._base_classes_tensor)
            ws = cw / cw.sum(-1, keepdim=True).clamp(min=1e-9)
            top, idx = torch.topk(ws, k, dim=-1)
            t = torch.zeros_like(ws).scatter(1, idx, top)
            t = t / t.sum(-1, keepdim=True)
            static[head] = (static[head].unsqueeze(0) + w * t.unsqueeze(1)).contiguous()
            per_roi |= 1 << order[head]
        return spec.with_static(static, per_roi)

    def get_similarity_matrices(self, box_features: torch.Tensor, return_similarity: bool = False):
        """roi_heads.py:245-336.  Returns a :class:`FusedSimilarity`; ``.materialize()`` gives the explicit
        ``{'cls','bbox'[,'seg']: [R,Nn,B]}`` dict of the reference."""
        spec = self._transfer_spec(box_features.device)
        vis_logits = None
        if self.compute_similarity["visual"] or spec.wk:
            feats = box_features.mean(dim=[2, 3]) if box_features.dim() > 2 else box_features
            # The reference builds the similarity with autograd on (roi_heads.py:245-257, 618): with a trainable box
            # head (COCO-*-ft.yaml, VOC 10-shot split 2/3) the loss reaches box_features through it.  Frozen features
            # (the usual fine-tune setting, and inference) skip the graph.
            track = torch.is_grad_enabled() and feats.requires_grad
            defer = (not track and not spec.wk and not return_similarity and not torch.is_grad_enabled()
                     and getattr(self.box_predictor, "_can_pack", lambda t: False)(feats))
            if defer:  # inference: the predictor's packed GEMM produces these columns (predictors._packed_products)
                return FusedSimilarity(spec, None, tuple(self.terms.keys()), feats=feats,
                                       weak_head=self.box_predictor.weak_detector_head)
            with torch.set_grad_enabled(track):
                vis_logits = self.box_predictor.weak_detector_head.mean_logits(feats)
            with torch.no_grad():
                if spec.wk:
                    spec = self._visualk_spec(spec, vis_logits.detach())
        sim = FusedSimilarity(spec, vis_logits, tuple(self.terms.keys()))
        if return_similarity:
            raw, _ = ops.lingual_similarity(self.box_predictor.embeddings.weight, self._coco_indexer_tensor,
                                            self._base_classes_tensor, self._novel_classes_tensor)
            viz = None if vis_logits is None else vis_logits.index_select(1, self._base_classes_tensor)
            return sim, [raw, viz]
        return sim

    def box_losses(self, x: torch.Tensor, weak_branch: Optional[torch.Tensor], proposals: List[Instances]):
        """Similarity -> predictor -> losses for sampled proposals (the training half of roi_heads.py:618-624).  The
        shipped fine-tune setting (only cls_score_ft / bbox_pred_ft train, frozen box head) runs as ONE fused autograd
        node; everything else takes the modular path.  Returns (losses dict, similarity or None)."""
        pred = self.box_predictor
        spec = self._transfer_spec(x.device)
        if getattr(pred, "can_fuse_losses", None) is not None and pred.can_fuse_losses(x, weak_branch, spec):
            losses, _ = pred.forward_losses(x, weak_branch, spec, proposals)
            return losses, None
        # base training applies no transfer (roi_heads.py:519-521): the similarity is only built by the *FineTune heads
        similarity = self.get_similarity_matrices(x) if getattr(self, "ALWAYS_TRANSFER", False) else None
        predictions, _ = pred(x, supervised_branch_x_weak=weak_branch, novel_classes=self._novel_classes_tensor,
                              base_classes=self._base_classes_tensor, similarity=similarity)
        return pred.losses(predictions, proposals), similarity

    # -- shared forward pieces -----------------------------------------------------------------------------
    def _truncate_weak(self, weak_proposals):
        if self.load_proposals or weak_proposals is None:
            return weak_proposals
        n = self.batch_size_per_image // self.weak_divisor
        return [p[:n] for p in weak_proposals]

    def _weak_box_features(self, weak_features, weak_proposals):
        """Pooled + box-head features of the weakly labelled images (roi_heads.py:509-517)."""
        if weak_features is None:
            return None
        feats = [weak_features[f] for f in self.box_in_features]
        pooled = self.box_pooler(feats, [x.proposal_boxes for x in weak_proposals])
        head = self.weak_box_head if self.weak_box_head is not None else self.box_head
        x_weak = head(pooled)
        return x_weak.mean(dim=[2, 3]) if x_weak.dim() > 2 else x_weak

    def _box_features(self, features, proposals):
        feats = [features[f] for f in self.box_in_features]
        pooled = self.box_pooler(feats, [x.proposal_boxes for x in proposals])
        box_features = self.box_head(pooled)
        weak_branch = None
        if self.weak_box_head is not None:
            with torch.no_grad():
                weak_branch = self.weak_box_head(pooled)
                if weak_branch.dim() > 2:
                    weak_branch = weak_branch.mean(dim=[2, 3])
        return pooled, box_features, weak_branch


@ROI_HEADS_REGISTRY.register()
class WSROIHeadNoMeta(WSROIHead):
    """roi_heads.py:487-591."""

    ALWAYS_TRANSFER = False

    def _forward_box(self, features, proposals, weak_features=None, weak_proposals=None, weak_targets=None, tta=False,
                     return_similarity=False, train_only_weak=False, return_proposals=False):
        x_weak = self._weak_box_features(weak_features, weak_proposals)
        if train_only_weak:
            if self.ALWAYS_TRANSFER or not self.training:
                raise ValueError("train_only_weak needs a training head without base->novel transfer "
                                 "(the reference dereferences box_features=None there, roi_heads.py:245-250)")
            box_features, x, weak_branch = None, None, None
        else:
            _, box_features, weak_branch = self._box_features(features, proposals)
            x = box_features.mean(dim=[2, 3]) if box_features.dim() > 2 else box_features
        similarity, sim_values = None, None
        if (self.training and self.ALWAYS_TRANSFER and x is not None and x_weak is None and not self.train_on_pred_boxes
                and not return_similarity):
            losses, similarity = self.box_losses(x, weak_branch, proposals)  # fused when the setting allows it
            return losses, box_features, similarity
        if self.ALWAYS_TRANSFER or not self.training:
            if return_similarity:
                similarity, sim_values = self.get_similarity_matrices(box_features, return_similarity=True)
            else:
                similarity = self.get_similarity_matrices(box_features)
        predictions, weak_predictions = self.box_predictor(x, supervised_branch_x_weak=weak_branch,
                                                           novel_classes=self._novel_classes_tensor,
                                                           base_classes=self._base_classes_tensor, x_weak=x_weak,
                                                           similarity=similarity)
        if self.training:
            losses = self.box_predictor.losses(predictions, proposals, weak_predictions=weak_predictions,
                                               weak_proposals=weak_proposals, weak_targets=weak_targets,
                                               train_only_weak=train_only_weak)
            if self.train_on_pred_boxes:  # roi_heads.py:533-539: the next stage trains on the predicted boxes
                with torch.no_grad():
                    pred_boxes = self.box_predictor.predict_boxes_for_gt_classes(predictions, proposals)
                    for per_image, boxes_i in zip(proposals, pred_boxes):
                        per_image.proposal_boxes = Boxes(boxes_i)
            return losses, box_features, similarity
        pred_instances, filter_inds = self.box_predictor.inference(predictions, proposals, tta=tta)
        if return_similarity and not tta:
            for i, inst in enumerate(pred_instances):
                inst._lingual_similarity = sim_values[0]
                inst._visual_similarity = sim_values[1][filter_inds[i]] if sim_values[1] is not None else None
        all_proposals = [proposals, predictions] if return_proposals else None
        return pred_instances, all_proposals, similarity, filter_inds

    def forward(self, images, features, proposals, targets=None, weak_images=None, weak_features=None,
                weak_proposals=None, weak_targets=None, tta=False, return_similarity=False, train_only_weak=False,
                return_proposals=False):
        """See roi_heads.py:553-591."""
        del images
        self.move_mappings_to_gpu()
        if self.training and not train_only_weak:
            assert targets
            proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        weak_proposals = self._truncate_weak(weak_proposals)
        if self.training:
            losses, box_features, similarity = self._forward_box(features, proposals, weak_features=weak_features,
                                                                 weak_proposals=weak_proposals,
                                                                 weak_targets=weak_targets,
                                                                 train_only_weak=train_only_weak)
            losses.update(self._forward_mask_train(features, proposals, box_features, similarity))
            return proposals, losses
        pred_instances, all_proposals, similarity, filter_inds = self._forward_box(
            features, proposals, tta=tta, return_similarity=return_similarity, return_proposals=return_proposals)
        if not tta:
            pred_instances = self._forward_mask_test(features, pred_instances, similarity, filter_inds)
        return pred_instances, all_proposals

    def _forward_mask_train(self, features, proposals, box_features, similarity):
        return {}

    def _forward_mask_test(self, features, instances, similarity, filter_inds):
        return instances


@ROI_HEADS_REGISTRY.register()
class WSROIHeadFineTune(WSROIHeadNoMeta):
    """roi_heads.py:593-644: similarity + transfer in training too."""

    ALWAYS_TRANSFER = True


@ROI_HEADS_REGISTRY.register()
class WSROIHeadNoMetaWithMask(WSROIHeadNoMeta):
    """roi_heads.py:646-822: mask features = box_head(box_pooler(pred_boxes)) when ROI_MASK_HEAD.POOLER_TYPE is
    "None" (the shipped segm YAMLs), mask logits transferred with the per-detection 'seg' similarity."""

    @classmethod
    def _init_mask_head(cls, cfg, input_shape):
        if not cfg.MODEL.MASK_ON:
            return {}
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        res = cfg.MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION
        pooler_type = cfg.MODEL.ROI_MASK_HEAD.POOLER_TYPE
        if pooler_type == "None":
            pooler_type = None
        in_channels = [input_shape[f].channels for f in in_features][0]
        ret = {"mask_in_features": in_features}
        ret["mask_pooler"] = ROIPooler(output_size=res, scales=tuple(1.0 / input_shape[k].stride for k in in_features),
                                       sampling_ratio=cfg.MODEL.ROI_MASK_HEAD.POOLER_SAMPLING_RATIO,
                                       pooler_type=pooler_type) if pooler_type else None
        if pooler_type:
            shape = ShapeSpec(channels=in_channels, width=res, height=res)
        else:
            shape = build_box_head(cfg, ShapeSpec(channels=in_channels, height=res, width=res)).output_shape
        ret["mask_head"] = build_mask_head(cfg, shape)
        return ret

    def _forward_mask_train(self, features, proposals, box_features, similarity):
        if not self.mask_on:
            return {}
        raise NotImplementedError("mask_rcnn_loss (gt mask rasterisation) is training-only and out of scope "
                                  "(SURVEY.md section 8a row a11)")

    def _forward_mask_test(self, features, instances, similarity, filter_inds):
        """roi_heads.py:691-709 / 777-782 / 880-885: second ROIAlign on pred_boxes, res5, mask head with the
        similarity rows of the kept RoIs."""
        if not self.mask_on:
            return instances
        assert len(instances) == 1, "mask inference runs one image at a time (roi_heads.py:773,882)"
        if self.mask_pooler is not None:
            feats = [features[f] for f in self.mask_in_features]
            x = self.mask_pooler(feats, [i.pred_boxes for i in instances])
        else:
            feats = [features[f] for f in self.mask_in_features]
            x = self.box_head(self.box_pooler(feats, [i.pred_boxes for i in instances]))
        s_seg = None
        if similarity is not None and "seg" in self.terms:
            s = similarity.materialize()["seg"] if isinstance(similarity, FusedSimilarity) else similarity["seg"]
            s_seg = s[filter_inds[0]] if s.dim() > 2 else s
        spec = self._transfer_spec(x.device)
        return self.mask_head(x, instances, similarity=None if s_seg is None else {"seg": s_seg}, spec=spec)


@ROI_HEADS_REGISTRY.register()
class WSROIHeadWithMaskFineTune(WSROIHeadNoMetaWithMask):
    """roi_heads.py:824-952."""

    ALWAYS_TRANSFER = True


@ROI_HEADS_REGISTRY.register()
class WeakDetectorHead(StandardROIHeads):
    """roi_heads.py:28-132: pool -> box head -> weak predictor (MIL + OICR losses in training)."""

    @configurable
    def __init__(self, *, box_in_features, box_pooler, box_head, box_predictor, freeze_layers=(), **kwargs):
        for k in ("base_classes_id", "novel_classes_id", "train_dataset_name", "weak_divisor", "terms"):
            setattr(self, "_" + k, kwargs.pop(k, None))
        for k in ("mask_in_features", "mask_pooler", "mask_head", "keypoint_in_features", "keypoint_pooler",
                  "keypoint_head", "weak_box_head", "visual_attention_head"):
            kwargs.pop(k, None)
        super().__init__(box_in_features=box_in_features, box_pooler=box_pooler, box_head=box_head,
                         box_predictor=box_predictor, **kwargs)
        _freeze(self, freeze_layers)

    @classmethod
    def from_config(cls, cfg, input_shape):
        ret = super().from_config(cfg, input_shape)
        ret["freeze_layers"] = cfg.MODEL.FREEZE_LAYERS.ROI_HEADS
        ret["base_classes_id"] = cfg.DATASETS.FEWSHOT.BASE_CLASSES_ID
        ret["novel_classes_id"] = cfg.DATASETS.FEWSHOT.NOVEL_CLASSES_ID
        ret["train_dataset_name"] = cfg.DATASETS.TRAIN[0]
        ret["weak_divisor"] = cfg.MODEL.ROI_HEADS.WEAK_CLASSIFIER_PROPOSAL_DIVISOR
        ret["terms"] = {"cls": cfg.MODEL.ROI_HEADS.FINETUNE_TERMS.CLASSIFIER,
                        "bbox": cfg.MODEL.ROI_HEADS.FINETUNE_TERMS.BBOX}
        return ret

    @classmethod
    def _init_box_head(cls, cfg, input_shape):
        ret = super()._init_box_head(cfg, input_shape)
        ret["box_predictor"] = build_fastrcnn_head(cfg, ret["box_head"].output_shape)
        return ret

    def forward(self, images, features, proposals, targets=None, tta=False, **unused):
        del images
        feats = [features[f] for f in self.box_in_features]
        box_features = self.box_head(self.box_pooler(feats, [x.proposal_boxes for x in proposals]))
        predictions, _ = self.box_predictor(box_features)
        if self.training:  # roi_heads.py:91-93,116-118: ``targets`` are the image-level class ids of each image
            return proposals, dict(self.box_predictor.losses(predictions, proposals, targets))
        pred_instances, _ = self.box_predictor.inference(predictions, proposals, tta=tta)
        return pred_instances, {}


def build_roi_heads(cfg, input_shape):
    """[D2] build_roi_heads: ``ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)``."""
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)
