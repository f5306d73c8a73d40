"""Box predictors behind ``FAST_RCNN_REGISTRY`` / ``WEAK_DETECTOR_FAST_RCNN_REGISTRY`` with the reference's class
names, constructor arguments, ``forward`` keywords and state_dict keys
(modeling/roi_heads/fast_rcnn.py:287-585, weak_detector_fast_rcnn.py:38-187, 270-351).

What differs from the reference is the execution: the five-to-seven ``Linear`` layers of a forward are one packed
GEMM, and everything after it (visual similarity, combination, base->novel transfer, weak scores, fine-tune terms)
is ONE fused CUDA kernel (csrc/transfer.cu) instead of ~30 ATen launches.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn
from torch.nn import functional as F

from . import layers, ops
from .layers import Box2BoxTransform, Matcher, fast_rcnn_inference
from .registry import FAST_RCNN_REGISTRY, WEAK_DETECTOR_FAST_RCNN_REGISTRY, configurable
from .structures import Boxes, Instances, ShapeSpec

logger = logging.getLogger("unit_b200")


class LossDict(dict):
    """The losses dict of the reference (name -> scalar tensor) that may also carry ``total``: the sum of its values
    computed by the same launch that produced them (fused fine-tune node), so a trainer need not add them again."""

    total: Optional[torch.Tensor] = None


def _freeze(module: nn.Module, layers_to_freeze: Sequence[str]) -> None:
    """Freeze by first dotted component of the parameter name (fast_rcnn.py:353-358, roi_heads.py:166-171)."""
    for name, param in module.named_parameters():
        if any(layer == name.split(".")[0] for layer in layers_to_freeze):
            param.requires_grad = False


def _linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, precision: str) -> torch.Tensor:
    """Predictor GEMM: "tf32" = hand-written tcgen05 kernel (north_star tolerance 1e-2), "fp32" = cuBLAS fp32."""
    if precision == "tf32" and x.is_cuda and x.dtype == torch.float32 and x.shape[1] % 4 == 0:
        return ops.linear_tf32(x, w, b)
    return F.linear(x, w, b)


def _input_size(input_shape) -> int:
    if isinstance(input_shape, int):
        input_shape = ShapeSpec(channels=input_shape)
    return input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)


class FusedSimilarity:
    """Lazy similarity: carries the mean OICR logits of the box-head features; the fused kernel turns them into
    ``S[R,Nn,B]`` on the fly.  ``materialize`` returns the explicit dict the reference's
    ``get_similarity_matrices`` returns (roi_heads.py:245-336)."""

    def __init__(self, spec: ops.TransferSpec, vis_logits: Optional[torch.Tensor], heads: Sequence[str],
                 feats: Optional[torch.Tensor] = None, weak_head=None):
        self.spec = spec
        self.vis_logits = vis_logits
        self.heads = tuple(heads)
        # deferred form: the visual logits are mean-OICR(feats); the predictor computes them as extra columns of its
        # packed GEMM when it is handed the same ``feats`` (predictors._packed_products), else ``ensure`` does
        self.feats = feats
        self.weak_head = weak_head
        self._cache: Optional[Dict[str, torch.Tensor]] = None

    def ensure_vis_logits(self) -> Optional[torch.Tensor]:
        if self.vis_logits is None and self.feats is not None:
            with torch.no_grad():
                self.vis_logits = self.weak_head.mean_logits(self.feats)
        return self.vis_logits

    def materialize(self) -> Dict[str, torch.Tensor]:
        self.ensure_vis_logits()
        if self._cache is None:
            R = self.vis_logits.shape[0] if self.vis_logits is not None else 1
            K = self.spec.K
            dev = self.spec.class_kind.device
            zs = torch.zeros((R, K + 1), device=dev)
            zb = torch.zeros((R, 4 * K), device=dev)
            _, _, sims = ops.similarity_transfer_forward(self.spec, self.vis_logits, zs, zb, want_similarity=self.heads)
            self._cache = {h: sims[h] for h in self.heads}
        return self._cache


@WEAK_DETECTOR_FAST_RCNN_REGISTRY.register()
class WeakDetectorOutputsBase(nn.Module):
    """OICR weak detector head (weak_detector_fast_rcnn.py:38-187).  In scope: ``evaluation`` (the mean of its three
    refinement classifiers is both the weak score and the source of the visual similarity), ``predict_*`` /
    ``inference``, ``label_and_sample_proposals`` and the MIL + OICR training ``losses`` (SURVEY.md section 8f rank 3;
    the PCL variant with its proposal graph / KMeans and the regression branches stay out of scope and raise)."""

    @configurable
    def __init__(self, input_shape, *, box2box_transform, num_classes, cls_agnostic_bbox_reg=False, oicr_iter=3,
                 freeze_layers=(), detector_temp=1.0, classifier_temp=1.0, regression_branch=False,
                 proposal_matcher=None, test_score_thresh=0.0, test_nms_thresh=0.5, test_topk_per_image=100,
                 oicr_regression_branch=False, base_classes=None, novel_classes=None, fg_threshold=0.5,
                 bg_threshold=0.1, mil_multiplier=4.0, weak_detector_type="OICR", **unused):
        super().__init__()
        if regression_branch or oicr_regression_branch or oicr_iter <= 0:
            raise NotImplementedError("only the shipped OICR configuration (OICR_ITER>0, no regression branches) "
                                      "is implemented (configs/default_config.py:40-50)")
        self.num_classes = num_classes
        self.oicr_iter = oicr_iter
        self.fg_threshold, self.bg_threshold = fg_threshold, bg_threshold
        self.mil_multiplier = mil_multiplier
        self.weak_detector_type = weak_detector_type
        self.box_dim = len(box2box_transform.weights)
        self.num_bbox_reg_classes = 1 if cls_agnostic_bbox_reg else num_classes
        self.detector_temp, self.classifier_temp = detector_temp, classifier_temp
        self.regression_branch = regression_branch
        self.oicr_regression_branch = oicr_regression_branch
        self.proposal_matcher = proposal_matcher
        self.box2box_transform = box2box_transform
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.base_classes = torch.tensor(list(base_classes or [])).long()
        self.novel_classes = torch.tensor(list(novel_classes or [])).long()
        self.input_size = _input_size(input_shape)
        self.classifier_stream = nn.Linear(self.input_size, num_classes)
        self.detection_stream = nn.Linear(self.input_size, num_classes)
        self.oicr_predictors = nn.ModuleList([nn.Linear(self.input_size, num_classes + 1) for _ in range(oicr_iter)])
        for l in [self.classifier_stream, self.detection_stream, *self.oicr_predictors]:
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0.0)
        _freeze(self, freeze_layers)

    @classmethod
    def from_config(cls, cfg, input_shape):
        wd = cfg.MODEL.ROI_HEADS.FAST_RCNN.WEAK_DETECTOR
        return {
            "input_shape": input_shape,
            "box2box_transform": Box2BoxTransform(weights=cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
            "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "cls_agnostic_bbox_reg": cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG,
            "oicr_iter": wd.OICR_ITER,
            "fg_threshold": wd.FG_THRESHOLD,
            "bg_threshold": wd.BG_THRESHOLD,
            "mil_multiplier": wd.MIL_MULTIPLIER,
            "weak_detector_type": wd.TYPE,
            "freeze_layers": cfg.MODEL.FREEZE_LAYERS.FAST_RCNN,
            "detector_temp": wd.DETECTOR_TEMP,
            "classifier_temp": wd.CLASSIFIER_TEMP,
            "regression_branch": wd.REGRESSION_BRANCH,
            "proposal_matcher": Matcher(cfg.MODEL.ROI_HEADS.IOU_THRESHOLDS, cfg.MODEL.ROI_HEADS.IOU_LABELS,
                                        allow_low_quality_matches=False),
            "test_score_thresh": cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST,
            "test_nms_thresh": cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST,
            "test_topk_per_image": cfg.TEST.DETECTIONS_PER_IMAGE,
            "oicr_regression_branch": wd.OICR_REGRESSION_BRANCH,
            "base_classes": cfg.DATASETS.FEWSHOT.BASE_CLASSES_ID,
            "novel_classes": cfg.DATASETS.FEWSHOT.NOVEL_CLASSES_ID,
        }

    # -- packed weights ---------------------------------------------------------------------------------
    def mean_oicr_weight(self):
        """mean_k(W_k x + b_k) == (mean_k W_k) x + mean_k b_k: the 3 refinement classifiers collapse to one."""
        params = [q for p in self.oicr_predictors for q in (p.weight, p.bias)]
        frozen = not any(q.requires_grad for q in params)
        key = (tuple(q._version for q in params), params[0].device,
               0 if ops._is_fake(params[0]) else params[0].data_ptr())
        cached = self.__dict__.get("_mean_cache")
        if frozen and cached is not None and cached[0] == key:
            return cached[1], cached[2]
        w = torch.stack([p.weight for p in self.oicr_predictors]).mean(0)
        b = torch.stack([p.bias for p in self.oicr_predictors]).mean(0)
        if frozen:
            self.__dict__["_mean_cache"] = (key, w.detach(), b.detach())
        return w, b

    gemm_precision = "tf32"

    def mean_logits(self, x: torch.Tensor) -> torch.Tensor:
        w, b = self.mean_oicr_weight()
        return _linear(x, w, b, self.gemm_precision)

    def evaluation(self, x_weak: torch.Tensor):
        """weak_detector_fast_rcnn.py:167-187: ([list of OICR logits], zeros box deltas), None."""
        cls_output = [p(x_weak) for p in self.oicr_predictors]
        bbox_output = x_weak.new_zeros((x_weak.size(0), self.num_bbox_reg_classes * self.box_dim))
        return [cls_output, bbox_output], None

    def forward(self, x_weak: torch.Tensor):
        if self.training:
            classifier_stream = self.classifier_stream(x_weak) / self.classifier_temp
            detection_stream = self.detection_stream(x_weak) / self.detector_temp
            oicr_scores = [p(x_weak) for p in self.oicr_predictors]
            return [classifier_stream, detection_stream, oicr_scores, [], None, None], None
        return self.evaluation(x_weak)

    def image_label_vector(self, weak_targets, dev) -> torch.Tensor:
        """[n_img, K] with 1 at every class present in the image (weak_detector_fast_rcnn.py:203,213-216); its
        non-zero columns in ascending order are the reference's ``torch.unique(gt_class)``."""
        gt_vector = torch.zeros((len(weak_targets), self.num_classes), dtype=torch.float32, device=dev)
        lens = [int(t.numel()) for t in weak_targets]
        if sum(lens):
            rows = torch.tensor([i for i, n in enumerate(lens) for _ in range(n)], dtype=torch.int64).to(dev)
            cols = layers.cat([t.reshape(-1).to(dev).long() for t in weak_targets])
            gt_vector[rows, cols] = 1.0
        return gt_vector

    def oicr_supervision(self, weak_predictions, weak_proposals, weak_targets):
        """The (labels, cls_weights) pair ``compute_loss_inputs`` hands to every refinement classifier
        (weak_detector_fast_rcnn.py:218-228, 384-396), plus the image-level MIL loss they start from."""
        cls_stream, det_stream, oicr_scores = weak_predictions[0], weak_predictions[1], weak_predictions[2]
        dev = cls_stream.device
        counts = [len(p) for p in weak_proposals]
        offsets = ops.offsets_from_counts(counts, dev)
        boxes = layers.cat([p.proposal_boxes.tensor for p in weak_proposals])
        gt_vector = self.image_label_vector(weak_targets, dev)
        loss_im, probs, _ = ops.mil_loss(cls_stream, det_stream, offsets, gt_vector, self.mil_multiplier,
                                         max_rows=max(counts) if counts else 0)
        supervision = []
        with torch.no_grad():
            for idx in range(len(oicr_scores)):
                if idx > 0:
                    probs, _ = ops.softmax_decode(oicr_scores[idx - 1].detach(), None, None, want_boxes=False)
                labels, weights, _, _ = ops.oicr_targets(probs, boxes, offsets, gt_vector,
                                                         self.proposal_matcher.user_thresholds,
                                                         self.proposal_matcher.labels, self.bg_threshold)
                supervision.append((labels, weights))
        return loss_im, supervision

    def losses(self, weak_predictions, weak_proposals, weak_targets):
        """weak_detector_fast_rcnn.py:189-245 for TYPE == "OICR": ``loss_im_cls`` (MIL) and ``loss_oicr_{1..n}``.
        ``weak_targets``: one tensor of image-level class ids per image."""
        if self.weak_detector_type != "OICR":
            raise NotImplementedError("the PCL weak detector (proposal graph + KMeans, weak_detector_fast_rcnn.py:"
                                      "410-519) is out of scope; WEAK_DETECTOR.TYPE is 'OICR' in every shipped YAML")
        loss_im, supervision = self.oicr_supervision(weak_predictions, weak_proposals, weak_targets)
        final_losses = {"loss_im_cls": loss_im}
        for idx, (labels, weights) in enumerate(supervision):
            final_losses["loss_oicr_{}".format(idx + 1)] = ops.weighted_ce_loss(weak_predictions[2][idx], labels,
                                                                                weights)
        return final_losses

    # -- inference (weak_detector_fast_rcnn.py:270-306) ---------------------------------------------------
    def predict_boxes(self, predictions, proposals):
        _, proposal_deltas = predictions
        num = [len(p) for p in proposals]
        boxes = layers.cat([p.proposal_boxes.tensor for p in proposals])
        return self.box2box_transform.apply_deltas(proposal_deltas, boxes).split(num)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        scores = torch.sum(torch.softmax(torch.stack(scores, 0), -1), 0)
        return scores.split([len(p) for p in proposals], dim=0)

    def inference(self, predictions, proposals, tta=False):
        scores = self.predict_probs(predictions, proposals)
        if tta:
            return [torch.cat(scores, 0), predictions[1]], None
        boxes = self.predict_boxes(predictions, proposals)
        return fast_rcnn_inference(boxes, scores, [x.image_size for x in proposals], self.test_score_thresh,
                                   self.test_nms_thresh, self.test_topk_per_image)

    def label_and_sample_proposals(self, proposals, targets, return_match_vals=False):
        """weak_detector_fast_rcnn.py:320-351: fused IoU + UniT Matcher, every proposal kept."""
        res, matched, vals = layers.label_and_sample(
            proposals, targets, num_classes=self.num_classes, batch_size_per_image=0, positive_fraction=0.0,
            thresholds=self.proposal_matcher.user_thresholds, labels=self.proposal_matcher.labels, sample=False,
            want_vals=True)
        if return_match_vals:
            return res, matched, vals
        return res, matched


@FAST_RCNN_REGISTRY.register()
class SupervisedDetectorOutputsBase(nn.Module):
    """fast_rcnn.py:292-468.  ``forward(x, novel_classes, base_classes, supervised_branch_x_weak=None, x_weak=None,
    similarity=None) -> ([scores, bbox], weak_branch_return)``."""

    KIND = "Base"
    gemm_precision = "tf32"  # "fp32" routes the predictor GEMMs through cuBLAS fp32 (strict 1e-5 parity checks)

    @configurable
    def __init__(self, input_shape, *, box2box_transform, num_classes, test_score_thresh=0.0, test_nms_thresh=0.5,
                 test_topk_per_image=100, cls_agnostic_bbox_reg=False, smooth_l1_beta=0.0,
                 box_reg_loss_type="smooth_l1", loss_weight=1.0, weak_detector_head=None, regression_branch=False,
                 terms=None, freeze_layers=(), embedding_path="", embeddings=None):
        super().__init__()
        if cls_agnostic_bbox_reg:
            raise NotImplementedError("class-agnostic box regression is not used by any reference YAML")
        self.num_classes = num_classes
        self.terms = dict(terms or {})
        self.box_dim = len(box2box_transform.weights)
        self.num_bbox_reg_classes = num_classes
        self.box2box_transform = box2box_transform
        self.smooth_l1_beta = smooth_l1_beta
        self.box_reg_loss_type = box_reg_loss_type
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.loss_weight = loss_weight
        self.weak_detector_head = weak_detector_head
        self.regression_branch = regression_branch
        self.input_size = _input_size(input_shape)
        self.cls_score_delta = nn.Linear(self.input_size, num_classes + 1)
        self.bbox_pred_delta = nn.Linear(self.input_size, num_classes * self.box_dim)
        nn.init.constant_(self.cls_score_delta.weight, 0.0)
        nn.init.normal_(self.bbox_pred_delta.weight, std=0.001)
        for l in (self.cls_score_delta, self.bbox_pred_delta):
            nn.init.constant_(l.bias, 0.0)
        if embeddings is None:
            embeddings = torch.load(embedding_path, weights_only=False)["embeddings"]
        self.embeddings = nn.Embedding.from_pretrained(embeddings.float(), freeze=True)
        self._extra_init()
        _freeze(self, freeze_layers)

    def _extra_init(self) -> None:
        pass

    @classmethod
    def from_config(cls, cfg, input_shape):
        rh = cfg.MODEL.ROI_HEADS
        return {
            "input_shape": input_shape,
            "box2box_transform": Box2BoxTransform(weights=cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
            "num_classes": rh.NUM_CLASSES,
            "cls_agnostic_bbox_reg": cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG,
            "smooth_l1_beta": cfg.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA,
            "test_score_thresh": rh.SCORE_THRESH_TEST,
            "test_nms_thresh": rh.NMS_THRESH_TEST,
            "test_topk_per_image": cfg.TEST.DETECTIONS_PER_IMAGE,
            "box_reg_loss_type": cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE,
            "loss_weight": {"loss_box_reg": cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT},
            "weak_detector_head": WEAK_DETECTOR_FAST_RCNN_REGISTRY.get(rh.FAST_RCNN.WEAK_DETECTOR.NAME)(cfg,
                                                                                                        input_shape),
            "regression_branch": rh.FAST_RCNN.WEAK_DETECTOR.REGRESSION_BRANCH,
            "terms": {"cls": rh.FINETUNE_TERMS.CLASSIFIER, "bbox": rh.FINETUNE_TERMS.BBOX,
                      "seg": rh.FINETUNE_TERMS.MASK},
            "embedding_path": rh.EMBEDDING_PATH,
            "freeze_layers": cfg.MODEL.FREEZE_LAYERS.FAST_RCNN,
        }

    # -- similarity --------------------------------------------------------------------------------------
    def get_similarity(self, base_classes, novel_classes, indexer):
        """fast_rcnn.py:376-382: raw lingual similarity [Nn,B]."""
        raw, _ = ops.lingual_similarity(self.embeddings.weight, indexer, base_classes, novel_classes)
        return raw

    # -- forward -------------------------------------------------------------------------------------------
    def _ft_layers(self):
        return None, None

    def _linears(self, x: torch.Tensor):
        """One packed GEMM for every Linear that reads ``x``: [delta | bbox | (ft cls | ft bbox)].  Returns column
        views (delta, bbox, packed fine-tune block or None); the transfer kernel reads them in place (row strides)."""
        K1, K4 = self.num_classes + 1, self.num_classes * self.box_dim
        base = [self.cls_score_delta.weight, self.bbox_pred_delta.weight, self.cls_score_delta.bias,
                self.bbox_pred_delta.bias]
        ft_c, ft_b = self._ft_layers()
        frozen = not any(q.requires_grad for q in base)
        if frozen and ft_c is not None:
            # fine-tuning: the frozen [delta | bbox] block is packed once; only the small trainable block is re-packed
            key = (tuple(q._version for q in base), base[0].device, 0 if ops._is_fake(base[0]) else base[0].data_ptr())
            cached = self.__dict__.get("_pack_cache")
            if cached is None or cached[0] != key:
                cached = (key, torch.cat(base[:2], 0).detach(), torch.cat(base[2:], 0).detach())
                self.__dict__["_pack_cache"] = cached
            y = _linear(x, cached[1], cached[2], self.gemm_precision)
            yf = _linear(x, torch.cat([ft_c.weight, ft_b.weight], 0), torch.cat([ft_c.bias, ft_b.bias], 0),
                         self.gemm_precision)
            return y[:, :K1], y[:, K1:], yf
        ws, bs = base[:2], base[2:]
        if ft_c is not None:
            ws = ws + [ft_c.weight, ft_b.weight]
            bs = bs + [ft_c.bias, ft_b.bias]
        y = _linear(x, torch.cat(ws, 0), torch.cat(bs, 0), self.gemm_precision)
        return y[:, :K1], y[:, K1:K1 + K4], (y[:, K1 + K4:] if ft_c is not None else None)

    # -- packed weights (one GEMM for every Linear of the predictor) ---------------------------------------------
    class _Pack:
        """W [rows, D] / b [rows]: rows [0, K1) cls_score_delta, [K1, o_ft) bbox_pred_delta, [o_ft, o_vis) the
        fine-tune layers (when the predictor has them), [o_vis, o_vis + K1) the mean of the OICR refinement classifiers."""
        __slots__ = ("W", "b", "o_ft", "o_vis", "key")

    def _pack(self) -> "SupervisedDetectorOutputsBase._Pack":
        """The persistent packed weight matrix.  Frozen blocks are copied in when their version changes; the TRAINABLE
        fine-tune parameters are re-pointed at their rows (``param.data`` becomes a view of the pack), so the optimizer
        updates the pack in place and no per-step concatenation is needed.  ``.to()`` / external ``.data`` swaps are
        detected by pointer and the pack is rebuilt."""
        K1, K4 = self.num_classes + 1, self.num_classes * self.box_dim
        ft_c, ft_b = self._ft_layers()
        ow, ob = self.weak_detector_head.mean_oicr_weight()
        frozen = [self.cls_score_delta.weight, self.cls_score_delta.bias, self.bbox_pred_delta.weight,
                  self.bbox_pred_delta.bias] + [q for p_ in self.weak_detector_head.oicr_predictors
                                                for q in (p_.weight, p_.bias)]
        dev = self.cls_score_delta.weight.device
        o_ft = K1 + K4
        o_vis = o_ft + (K1 + K4 if ft_c is not None else 0)
        rows = o_vis + K1
        fake = ops._is_fake(frozen[0])  # FakeTensor tracing: no data pointers, nothing to re-point
        key = (tuple(q._version for q in frozen), () if fake else tuple(q.data_ptr() for q in frozen), str(dev), rows)
        pk = self.__dict__.get("_wpack")
        if pk is None or pk.key != key:
            old = pk
            pk = SupervisedDetectorOutputsBase._Pack()
            D = self.input_size
            pk.W = torch.zeros((rows, D), dtype=torch.float32, device=dev)
            pk.b = torch.zeros((rows,), dtype=torch.float32, device=dev)
            pk.o_ft, pk.o_vis, pk.key = o_ft, o_vis, key
            with torch.no_grad():
                pk.W[:K1].copy_(self.cls_score_delta.weight)
                pk.W[K1:o_ft].copy_(self.bbox_pred_delta.weight)
                pk.b[:K1].copy_(self.cls_score_delta.bias)
                pk.b[K1:o_ft].copy_(self.bbox_pred_delta.bias)
                pk.W[o_vis:].copy_(ow)
                pk.b[o_vis:].copy_(ob)
            self.__dict__["_wpack"] = pk
            del old
        if ft_c is not None and not fake:
            es = pk.W.element_size()
            want = ((ft_c.weight, pk.W[o_ft:o_ft + K1]), (ft_b.weight, pk.W[o_ft + K1:o_vis]),
                    (ft_c.bias, pk.b[o_ft:o_ft + K1]), (ft_b.bias, pk.b[o_ft + K1:o_vis]))
            for prm, view in want:
                if prm.data_ptr() != view.data_ptr():
                    with torch.no_grad():
                        view.copy_(prm.data)
                    prm.data = view
            del es
        return pk

    def _can_pack(self, x: torch.Tensor) -> bool:
        return (self.gemm_precision == "tf32" and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
                and x.shape[1] % 4 == 0 and self.weak_detector_head.oicr_iter > 0)

    def _packed_products(self, x, x_weak_branch, similarity):
        """Every Linear of an inference-mode forward as ONE grouped tcgen05 launch (+ one reduce): returns
        (delta, proposal_deltas, ft_packed or None, weak_scores); fills ``similarity.vis_logits`` when it was deferred
        on this same ``x``."""
        pk = self._pack()
        K1 = self.num_classes + 1
        xw = x if x_weak_branch is None else x_weak_branch
        y1, y2 = ops.predictor_gemm2(x, pk.W, pk.b, xw, pk.W[pk.o_vis:pk.o_vis + K1], pk.b[pk.o_vis:pk.o_vis + K1])
        if isinstance(similarity, FusedSimilarity) and similarity.vis_logits is None and similarity.feats is x:
            similarity.vis_logits = y1[:, pk.o_vis:pk.o_vis + K1]
        ft = y1[:, pk.o_ft:pk.o_vis] if pk.o_vis > pk.o_ft else None
        return y1[:, :K1], y1[:, K1:pk.o_ft], ft, y2[:, :K1]

    def _transfer_mode(self, similarity):
        """(do_transfer, novel_neg_inf, detach) for this predictor kind (fast_rcnn.py:401,427-428)."""
        return (similarity is not None and not self.training), self.training, False

    def forward(self, x, novel_classes, base_classes, supervised_branch_x_weak=None, x_weak=None, similarity=None):
        if x is None:  # train_only_weak (fast_rcnn.py:393-397): zero supervised predictions, weak head only
            scores = x_weak.new_zeros((x_weak.size(0), self.num_classes + 1))
            bbox = x_weak.new_zeros((x_weak.size(0), self.num_classes * self.box_dim))
            if self.training:
                scores = scores.index_fill(1, novel_classes, float("-inf"))
            return [scores, bbox], self.weak_detector_head(x_weak)[0]
        if not torch.is_grad_enabled() and self._can_pack(x):  # inference: one grouped GEMM for all five products
            delta, pd, ft_packed, weak_scores = self._packed_products(x, supervised_branch_x_weak, similarity)
        else:
            delta, pd, ft_packed = self._linears(x)
            with torch.no_grad():
                weak_scores = self.weak_detector_head.mean_logits(x if supervised_branch_x_weak is None
                                                                  else supervised_branch_x_weak)
        do_transfer, neg_inf, detach = self._transfer_mode(similarity)
        spec, vis_logits = self._resolve_similarity(similarity, base_classes, novel_classes, x.device)
        scores, bbox = ops.similarity_transfer(spec, vis_logits, delta, pd, weak_scores, None, None, do_transfer,
                                               neg_inf, detach, ft_packed=ft_packed)
        weak_branch_return = None
        if x_weak is not None:
            weak_branch_return, _ = self.weak_detector_head(x_weak)
        return [scores, bbox], weak_branch_return

    def _resolve_similarity(self, similarity, base_classes, novel_classes, dev):
        if isinstance(similarity, FusedSimilarity):
            return similarity.spec, similarity.ensure_vis_logits()
        base, novel = base_classes.tolist(), novel_classes.tolist()
        if similarity is None:
            return self._plain_spec(base, novel, dev), None
        # explicit matrices, exactly what the reference passes: [Nn,B] or [R,Nn,B] per head
        static, per_roi = {}, 0
        for i, h in enumerate(("cls", "bbox")):
            static[h] = similarity[h]
            if similarity[h].dim() > 2:
                per_roi |= 1 << i
        return ops.TransferSpec(self.num_classes, base, novel, dev, static, static_per_roi=per_roi), None

    def _plain_spec(self, base, novel, dev):
        key = (tuple(base), tuple(novel), str(dev))
        cache = self.__dict__.setdefault("_spec_cache", {})
        if key not in cache:
            cache[key] = ops.TransferSpec(self.num_classes, base, novel, dev)
        return cache[key]

    # -- losses / inference --------------------------------------------------------------------------------
    def losses(self, predictions, proposals, weak_predictions=None, weak_proposals=None, weak_targets=None,
               train_only_weak=False):
        """fast_rcnn.py:435-453: [D2] FastRCNNOutputs.losses (softmax CE + smooth-L1 on get_deltas)."""
        final_losses = {}
        if weak_predictions is not None:
            final_losses.update(self.weak_detector_head.losses(weak_predictions, weak_proposals, weak_targets))
        if train_only_weak:
            return final_losses
        scores, proposal_deltas = predictions
        if self.box_reg_loss_type != "smooth_l1":
            raise NotImplementedError("only the smooth_l1 box loss is used by the reference YAMLs")
        gt_classes = layers.cat([p.gt_classes for p in proposals])
        prop = layers.cat([p.proposal_boxes.tensor for p in proposals])
        gt_boxes = layers.cat([p.gt_boxes.tensor for p in proposals])
        loss_cls, loss_box = ops.fastrcnn_loss(scores, proposal_deltas, prop, gt_boxes, gt_classes,
                                               self.box2box_transform.weights, self.smooth_l1_beta)
        final_losses.update({"loss_cls": loss_cls, "loss_box_reg": loss_box})
        return final_losses

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        probs, _ = ops.softmax_decode(scores, None, None, want_boxes=False)
        return probs.split([len(p) for p in proposals], dim=0)

    def predict_boxes(self, predictions, proposals):
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        boxes = layers.cat([p.proposal_boxes.tensor for p in proposals])
        return self.box2box_transform.apply_deltas(proposal_deltas, boxes).split([len(p) for p in proposals])

    def predict_boxes_for_gt_classes(self, predictions, proposals):
        """[D2] FastRCNNOutputLayers.predict_boxes_for_gt_classes (roi_heads.py:535-539, TRAIN_ON_PRED_BOXES)."""
        from .outputs import predict_boxes_for_gt_classes

        return predict_boxes_for_gt_classes(self.box2box_transform, predictions, proposals)

    def inference(self, predictions, proposals, tta=False):
        """fast_rcnn.py:455-468: softmax + decode in one launch, then batched filter + NMS."""
        scores, proposal_deltas = predictions
        n = [len(p) for p in proposals]
        if tta:
            probs, _ = ops.softmax_decode(scores, None, None, want_boxes=False)
            return [probs, predictions[1]], None
        dets = self.inference_device(predictions, proposals)
        return layers.instances_from_detections(dets, [x.image_size for x in proposals])

    def inference_device(self, predictions, proposals):
        """Device half of ``inference`` (softmax + decode, filter, NMS, top-k): nothing is read back, so it can be
        captured in a CUDA graph; ``layers.instances_from_detections`` finishes on the host."""
        scores, proposal_deltas = predictions
        n = [len(p) for p in proposals]
        boxes_in = layers.cat([p.proposal_boxes.tensor for p in proposals])
        probs, boxes = ops.softmax_decode(scores, proposal_deltas, boxes_in, self.box2box_transform.weights,
                                          self.box2box_transform.scale_clamp)
        return layers.fast_rcnn_inference_device(boxes.split(n), probs.split(n), [x.image_size for x in proposals],
                                                 self.test_score_thresh, self.test_nms_thresh,
                                                 self.test_topk_per_image)


    def inference_tta(self, tta_predictions, proposals):
        """Test-time-augmentation aggregation of the meta-architecture (rcnn.py:495-527): the per-augmentation
        ``[probs, deltas]`` returned by ``inference(..., tta=True)`` are reduced -- class probabilities SUMMED, box deltas
        AVERAGED -- then decoded on the un-augmented proposals and sent through ``fast_rcnn_inference``."""
        scores = torch.stack([p[0] for p in tta_predictions]).sum(0)
        deltas = torch.stack([p[1] for p in tta_predictions]).mean(0)
        n = [len(p) for p in proposals]
        boxes = self.predict_boxes([scores, deltas], proposals)
        return fast_rcnn_inference(boxes, scores.split(n), [x.image_size for x in proposals],
                                   self.test_score_thresh, self.test_nms_thresh, self.test_topk_per_image)

@FAST_RCNN_REGISTRY.register()
class SupervisedDetectorOutputsFineTune(SupervisedDetectorOutputsBase):
    """fast_rcnn.py:470-533: adds zero-initialised ``cls_score_ft`` / ``bbox_pred_ft``; transfer applied in
    training too."""

    KIND = "FineTune"

    def _extra_init(self) -> None:
        self.cls_score_ft = nn.Linear(self.input_size, self.num_classes + 1)
        self.bbox_pred_ft = nn.Linear(self.input_size, self.num_bbox_reg_classes * self.box_dim)
        for l in (self.cls_score_ft, self.bbox_pred_ft):
            nn.init.constant_(l.weight, 0.0)
            nn.init.constant_(l.bias, 0.0)

    def _ft_layers(self):
        return self.cls_score_ft, self.bbox_pred_ft

    def _transfer_mode(self, similarity):
        return similarity is not None, False, False

    def can_fuse_losses(self, x: torch.Tensor, x_weak_branch: Optional[torch.Tensor], spec: ops.TransferSpec) -> bool:
        """True when the whole fine-tune step can run as the fused node ``ops.ft_step_losses``: the shipped fine-tune
        setting -- only cls_score_ft / bbox_pred_ft train, the box-head features carry no gradient."""
        trainable = [n for n, p_ in self.named_parameters() if p_.requires_grad]
        only_ft = all(n.split(".")[0] in ("cls_score_ft", "bbox_pred_ft") for n in trainable) and len(trainable) == 4
        return (self.training and torch.is_grad_enabled() and self._can_pack(x) and only_ft and not x.requires_grad
                and (x_weak_branch is None or not x_weak_branch.requires_grad) and not getattr(spec, "wk", None)
                and not spec.static_per_roi and self.box_reg_loss_type == "smooth_l1")

    def forward_losses(self, x, x_weak_branch, spec: ops.TransferSpec, proposals):
        """forward (fast_rcnn.py:484-533) + losses (:435-453) for sampled ``proposals`` as ONE autograd node.
        Returns ({'loss_cls', 'loss_box_reg'} as a LossDict carrying .total, [scores, bbox]) -- predictions detached."""
        pk = self._pack()
        gt_classes = layers.cat([p_.gt_classes for p_ in proposals])
        prop = layers.cat([p_.proposal_boxes.tensor for p_ in proposals])
        gt_boxes = layers.cat([p_.gt_boxes.tensor for p_ in proposals])
        xw = x if x_weak_branch is None else x_weak_branch
        loss_cls, loss_box, scores, bbox, total = ops.ft_step_losses(
            self.cls_score_ft.weight, self.cls_score_ft.bias, self.bbox_pred_ft.weight, self.bbox_pred_ft.bias, x, xw, pk,
            spec, prop, gt_boxes, gt_classes, self.box2box_transform.weights, self.smooth_l1_beta)
        losses = LossDict(loss_cls=loss_cls, loss_box_reg=loss_box)
        losses.total = total  # the sum the trainer would form, from the same launch
        return losses, [scores, bbox]


@FAST_RCNN_REGISTRY.register()
class SupervisedDetectorOutputsWeakFineTune(SupervisedDetectorOutputsBase):
    """fast_rcnn.py:535-585: transfer always applied, transferred part detached from the graph (:566,575)."""

    KIND = "WeakFineTune"

    def _transfer_mode(self, similarity):
        return similarity is not None, False, True


@FAST_RCNN_REGISTRY.register()
class WeakDetectorOutputsBaseWrapper(WeakDetectorOutputsBase):
    """fast_rcnn.py:287-290."""


def build_fastrcnn_head(cfg, input_shape):
    """fast_rcnn.py:587-589."""
    return FAST_RCNN_REGISTRY.get(cfg.MODEL.ROI_HEADS.FAST_RCNN.NAME)(cfg, input_shape)
