"""Oracle restatement of the Detectron2 functional glue on the RoI path.  TEST INFRASTRUCTURE.

Each function names the Detectron2 symbol it restates and the reference call site that uses it
(SURVEY.md section 8a / Appendix A).  ROIAlign and NMS are NOT restated: they call the compiled torchvision
0.26 CPU ops (``torch.ops.torchvision.roi_align`` / ``nms``), the kernels Detectron2 >= 0.4 dispatches to.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch
import torchvision
from torch.nn import functional as F

from .structures import Boxes, Instances


# --------------------------------------------------------------------------- layers helpers
def cat(tensors: List[torch.Tensor], dim: int = 0) -> torch.Tensor:
    """detectron2.layers.cat"""
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x: torch.Tensor):
    """detectron2.layers.nonzero_tuple (used by modeling/matcher.py:115, fast_rcnn.py:61)"""
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """detectron2.layers.batched_nms (>=0.4): torchvision.ops.boxes.batched_nms on float boxes.

    torchvision/ops/boxes.py:51-119: coordinate trick when numel <= 4000 (CPU) else per-class loop.
    """
    assert boxes.shape[-1] == 4
    return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


# --------------------------------------------------------------------------- Matcher (D2, 2 outputs)
class Matcher:
    """detectron2.modeling.matcher.Matcher -- same algorithm as the reference's modeling/matcher.py:54-119
    minus the third return value."""

    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False):
        thresholds = thresholds[:]
        assert thresholds[0] > 0
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def _match(self, m: torch.Tensor):
        assert m.dim() == 2
        if m.numel() == 0:
            n = m.size(1)
            return (
                m.new_full((n,), 0, dtype=torch.int64),
                m.new_full((n,), self.labels[0], dtype=torch.int8),
                m.new_full((n,), 0, dtype=torch.float32),
            )
        assert torch.all(m >= 0)
        matched_vals, matches = m.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for (l, low, high) in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            low_high = (matched_vals >= low) & (matched_vals < high)
            match_labels[low_high] = l
        if self.allow_low_quality_matches:
            highest_quality_foreach_gt, _ = m.max(dim=1)
            _, pred_inds = nonzero_tuple(m == highest_quality_foreach_gt[:, None])
            match_labels[pred_inds] = 1
        return matches, match_labels, matched_vals

    def __call__(self, match_quality_matrix: torch.Tensor):
        matches, labels, _ = self._match(match_quality_matrix)
        return matches, labels


class MatcherWithVals(Matcher):
    """The reference's own variant (modeling/matcher.py:54-98): also returns ``matched_vals``."""

    def __call__(self, match_quality_matrix: torch.Tensor):
        return self._match(match_quality_matrix)


# --------------------------------------------------------------------------- sampling
def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     generator: Optional[torch.Generator] = None):
    """detectron2.modeling.sampling.subsample_labels (called via ROIHeads._sample_proposals and at
    roi_heads.py:415).  ``generator`` makes the two randperm draws reproducible (SURVEY.md section 7 hard part 1e)."""
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = int(num_samples * positive_fraction)
    num_pos = min(positive.numel(), num_pos)
    num_neg = num_samples - num_pos
    num_neg = min(negative.numel(), num_neg)
    perm1 = torch.randperm(positive.numel(), device=positive.device, generator=generator)[:num_pos]
    perm2 = torch.randperm(negative.numel(), device=negative.device, generator=generator)[:num_neg]
    return positive[perm1], negative[perm2]


def add_ground_truth_to_proposals(gt_boxes: List[Boxes], proposals: List[Instances]) -> List[Instances]:
    """detectron2.modeling.proposal_generator.proposal_utils.add_ground_truth_to_proposals: GT appended AFTER
    the RPN proposals with objectness logit log((1-1e-10)/(1-(1-1e-10)))."""
    assert gt_boxes is not None and len(proposals) == len(gt_boxes)
    if len(proposals) == 0:
        return proposals
    out = []
    for gt_boxes_i, proposals_i in zip(gt_boxes, proposals):
        device = proposals_i.objectness_logits.device
        gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
        gt_logits = gt_logit_value * torch.ones(len(gt_boxes_i), device=device)
        gt_proposal = Instances(proposals_i.image_size)
        gt_proposal.proposal_boxes = gt_boxes_i
        gt_proposal.objectness_logits = gt_logits
        out.append(Instances.cat([proposals_i, gt_proposal]))
    return out


# --------------------------------------------------------------------------- box regression
_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    """detectron2.modeling.box_regression.Box2BoxTransform (constructed at fast_rcnn.py:335,
    weak_detector_fast_rcnn.py:119)."""

    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes: torch.Tensor, target_boxes: torch.Tensor) -> torch.Tensor:
        src_widths = src_boxes[:, 2] - src_boxes[:, 0]
        src_heights = src_boxes[:, 3] - src_boxes[:, 1]
        src_ctr_x = src_boxes[:, 0] + 0.5 * src_widths
        src_ctr_y = src_boxes[:, 1] + 0.5 * src_heights
        target_widths = target_boxes[:, 2] - target_boxes[:, 0]
        target_heights = target_boxes[:, 3] - target_boxes[:, 1]
        target_ctr_x = target_boxes[:, 0] + 0.5 * target_widths
        target_ctr_y = target_boxes[:, 1] + 0.5 * target_heights
        wx, wy, ww, wh = self.weights
        dx = wx * (target_ctr_x - src_ctr_x) / src_widths
        dy = wy * (target_ctr_y - src_ctr_y) / src_heights
        dw = ww * torch.log(target_widths / src_widths)
        dh = wh * torch.log(target_heights / src_heights)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_widths > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = deltas[:, 2::4] / ww
        dh = deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        x1 = pred_ctr_x - 0.5 * pred_w
        y1 = pred_ctr_y - 0.5 * pred_h
        x2 = pred_ctr_x + 0.5 * pred_w
        y2 = pred_ctr_y + 0.5 * pred_h
        pred_boxes = torch.stack((x1, y1, x2, y2), dim=-1)
        return pred_boxes.reshape(deltas.shape)


# --------------------------------------------------------------------------- ROIPooler
class ROIPooler(torch.nn.Module):
    """detectron2.modeling.poolers.ROIPooler, single- or multi-level; C4 uses one level.

    "ROIAlignV2" -> roi_align(aligned=True); "ROIAlign" -> aligned=False; "ROIPool" -> roi_pool.
    Constructed by StandardROIHeads._init_box_head (reached via roi_heads.py:220) and roi_heads.py:673-678.
    """

    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2
        self.output_size = output_size
        self.scales = tuple(scales)
        self.sampling_ratio = sampling_ratio
        self.pooler_type = pooler_type
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level = int(min_level)
        self.max_level = int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1
        self.canonical_level = canonical_level
        self.canonical_box_size = canonical_box_size

    def _pool(self, x: torch.Tensor, rois: torch.Tensor, scale: float) -> torch.Tensor:
        if self.pooler_type == "ROIAlignV2":
            return torch.ops.torchvision.roi_align(x, rois.to(x.dtype), scale, self.output_size[0],
                                                   self.output_size[1], self.sampling_ratio, True)
        if self.pooler_type == "ROIAlign":
            return torch.ops.torchvision.roi_align(x, rois.to(x.dtype), scale, self.output_size[0],
                                                   self.output_size[1], self.sampling_ratio, False)
        if self.pooler_type == "ROIPool":
            return torchvision.ops.roi_pool(x, rois.to(x.dtype), self.output_size, scale)
        raise ValueError("Unknown pooler type: {}".format(self.pooler_type))

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]) -> torch.Tensor:
        num_level_assignments = len(self.scales)
        assert isinstance(x, list) and isinstance(box_lists, list)
        assert len(x) == num_level_assignments
        assert len(box_lists) == x[0].size(0)
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        pooler_fmt_boxes = cat(
            [
                torch.cat((torch.full((len(b), 1), i, dtype=b.tensor.dtype, device=b.tensor.device), b.tensor), dim=1)
                for i, b in enumerate(box_lists)
            ],
            dim=0,
        )
        if num_level_assignments == 1:
            return self._pool(x[0], pooler_fmt_boxes, self.scales[0])
        box_sizes = torch.sqrt(cat([b.area() for b in box_lists]))
        level_assignments = torch.floor(self.canonical_level + torch.log2(box_sizes / self.canonical_box_size + 1e-8))
        level_assignments = torch.clamp(level_assignments, min=self.min_level, max=self.max_level).to(torch.int64)
        level_assignments = level_assignments - self.min_level
        num_boxes = pooler_fmt_boxes.size(0)
        output = torch.zeros((num_boxes, x[0].shape[1]) + self.output_size, dtype=x[0].dtype, device=x[0].device)
        for level, scale in enumerate(self.scales):
            inds = nonzero_tuple(level_assignments == level)[0]
            output.index_put_((inds,), self._pool(x[level], pooler_fmt_boxes[inds], scale))
        return output


# --------------------------------------------------------------------------- inference
def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh: float, nms_thresh: float,
                                     topk_per_image: int):
    """detectron2.modeling.roi_heads.fast_rcnn.fast_rcnn_inference_single_image."""
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    boxes = Boxes(boxes.reshape(-1, 4))
    boxes.clip(image_shape)
    boxes = boxes.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


def fast_rcnn_inference(boxes, scores, image_shapes, score_thresh, nms_thresh, topk_per_image):
    """detectron2 fast_rcnn_inference (called at fast_rcnn.py:461-468, weak_detector_fast_rcnn.py:299-306,
    meta_arch/rcnn.py:526)."""
    result_per_image = [
        fast_rcnn_inference_single_image(b, s, shape, score_thresh, nms_thresh, topk_per_image)
        for s, b, shape in zip(scores, boxes, image_shapes)
    ]
    return [x[0] for x in result_per_image], [x[1] for x in result_per_image]


def select_foreground_proposals(proposals: List[Instances], bg_label: int):
    """detectron2.modeling.roi_heads.roi_heads.select_foreground_proposals (roi_heads.py:344,696,892)."""
    assert isinstance(proposals, (list, tuple))
    assert isinstance(proposals[0], Instances)
    assert proposals[0].has("gt_classes")
    fg_proposals, fg_selection_masks = [], []
    for proposals_per_image in proposals:
        gt_classes = proposals_per_image.gt_classes
        fg_selection_mask = (gt_classes != -1) & (gt_classes != bg_label)
        fg_idxs = fg_selection_mask.nonzero().squeeze(1)
        fg_proposals.append(proposals_per_image[fg_idxs])
        fg_selection_masks.append(fg_selection_mask)
    return fg_proposals, fg_selection_masks


# --------------------------------------------------------------------------- masks
def mask_rcnn_inference(pred_mask_logits: torch.Tensor, pred_instances: List[Instances]) -> None:
    """detectron2.modeling.roi_heads.mask_head.mask_rcnn_inference (mask_head.py:36,92)."""
    cls_agnostic_mask = pred_mask_logits.size(1) == 1
    if cls_agnostic_mask:
        mask_probs_pred = pred_mask_logits.sigmoid()
    else:
        num_masks = pred_mask_logits.shape[0]
        class_pred = cat([i.pred_classes for i in pred_instances])
        indices = torch.arange(num_masks, device=class_pred.device)
        mask_probs_pred = pred_mask_logits[indices, class_pred][:, None].sigmoid()
    num_boxes_per_image = [len(i) for i in pred_instances]
    mask_probs_pred = mask_probs_pred.split(num_boxes_per_image, dim=0)
    for prob, instances in zip(mask_probs_pred, pred_instances):
        instances.pred_masks = prob


def mask_rcnn_loss(pred_mask_logits: torch.Tensor, instances: List[Instances], vis_period: int = 0):
    """detectron2 mask_rcnn_loss needs gt_masks.crop_and_resize (polygon/bitmask rasterisation): training-only and
    outside SURVEY.md section 8; not restated."""
    raise NotImplementedError("mask_rcnn_loss is outside the scoped hot path (SURVEY.md section 8a row a11)")


BYTES_PER_FLOAT = 4
GPU_MEM_LIMIT = 1024 ** 3


def _do_paste_mask(masks: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int, skip_empty: bool = True):
    device = masks.device
    if skip_empty:
        x0_int, y0_int = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(dtype=torch.int32)
        x1_int = torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(dtype=torch.int32)
        y1_int = torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(dtype=torch.int32)
    else:
        x0_int, y0_int = 0, 0
        x1_int, y1_int = img_w, img_h
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    N = masks.shape[0]
    img_y = torch.arange(y0_int, y1_int, device=device, dtype=torch.float32) + 0.5
    img_x = torch.arange(x0_int, x1_int, device=device, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(N, img_y.size(1), img_x.size(1))
    gy = img_y[:, :, None].expand(N, img_y.size(1), img_x.size(1))
    grid = torch.stack([gx, gy], dim=3)
    if not masks.dtype.is_floating_point:
        masks = masks.float()
    img_masks = F.grid_sample(masks, grid.to(masks.dtype), align_corners=False)
    if skip_empty:
        return img_masks[:, 0], (slice(y0_int, y1_int), slice(x0_int, x1_int))
    return img_masks[:, 0], ()


def paste_masks_in_image(masks: torch.Tensor, boxes, image_shape: Tuple[int, int], threshold: float = 0.5,
                         skip_empty: Optional[bool] = None) -> torch.Tensor:
    """detectron2.layers.mask_ops.paste_masks_in_image (reached through detector_postprocess,
    meta_arch/rcnn.py:423).  CPU path: one mask at a time inside its integer window (skip_empty=True)."""
    assert masks.shape[-1] == masks.shape[-2], "Only square mask predictions are supported"
    N = len(masks)
    if N == 0:
        return masks.new_empty((0,) + tuple(image_shape), dtype=torch.uint8)
    if not isinstance(boxes, torch.Tensor):
        boxes = boxes.tensor
    device = boxes.device
    assert len(boxes) == N, boxes.shape
    img_h, img_w = image_shape
    if device.type == "cpu":
        num_chunks = N
    else:
        num_chunks = int(np.ceil(N * int(img_h) * int(img_w) * BYTES_PER_FLOAT / GPU_MEM_LIMIT))
    if skip_empty is None:
        skip_empty = device.type == "cpu"
    chunks = torch.chunk(torch.arange(N, device=device), num_chunks)
    img_masks = torch.zeros(N, img_h, img_w, device=device, dtype=torch.bool if threshold >= 0 else torch.uint8)
    for inds in chunks:
        masks_chunk, spatial_inds = _do_paste_mask(masks[inds, None, :, :], boxes[inds], img_h, img_w,
                                                   skip_empty=skip_empty)
        if threshold >= 0:
            masks_chunk = (masks_chunk >= threshold).to(dtype=torch.bool)
        else:
            masks_chunk = (masks_chunk * 255).to(dtype=torch.uint8)
        img_masks[(inds,) + spatial_inds] = masks_chunk
    return img_masks


def detector_postprocess(results: Instances, output_height: int, output_width: int, mask_threshold: float = 0.5):
    """detectron2.modeling.postprocessing.detector_postprocess (meta_arch/rcnn.py:423)."""
    new_size = (output_height, output_width)
    scale_x, scale_y = (output_width / results.image_size[1], output_height / results.image_size[0])
    results = Instances(new_size, **results.get_fields())
    if results.has("pred_boxes"):
        output_boxes = results.pred_boxes
    elif results.has("proposal_boxes"):
        output_boxes = results.proposal_boxes
    else:
        output_boxes = None
    assert output_boxes is not None, "Predictions must contain boxes!"
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    results = results[output_boxes.nonempty()]
    if results.has("pred_masks"):
        results.pred_masks = paste_masks_in_image(
            results.pred_masks[:, 0, :, :], results.pred_boxes, results.image_size, threshold=mask_threshold
        )
    return results


# --------------------------------------------------------------------------- fvcore losses
def smooth_l1_loss(input: torch.Tensor, target: torch.Tensor, beta: float, reduction: str = "none") -> torch.Tensor:
    """fvcore.nn.smooth_l1_loss (imported at fast_rcnn.py:20)."""
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        cond = n < beta
        loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def giou_loss(boxes1: torch.Tensor, boxes2: torch.Tensor, reduction: str = "none", eps: float = 1e-7) -> torch.Tensor:
    """fvcore.nn.giou_loss."""
    x1, y1, x2, y2 = boxes1.unbind(dim=-1)
    x1g, y1g, x2g, y2g = boxes2.unbind(dim=-1)
    assert (x2 >= x1).all() and (y2 >= y1).all()
    xkis1, ykis1 = torch.max(x1, x1g), torch.max(y1, y1g)
    xkis2, ykis2 = torch.min(x2, x2g), torch.min(y2, y2g)
    intsctk = torch.zeros_like(x1)
    mask = (ykis2 > ykis1) & (xkis2 > xkis1)
    intsctk[mask] = (xkis2[mask] - xkis1[mask]) * (ykis2[mask] - ykis1[mask])
    unionk = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - intsctk
    iouk = intsctk / (unionk + eps)
    xc1, yc1 = torch.min(x1, x1g), torch.min(y1, y1g)
    xc2, yc2 = torch.max(x2, x2g), torch.max(y2, y2g)
    area_c = (xc2 - xc1) * (yc2 - yc1)
    miouk = iouk - ((area_c - unionk) / (area_c + eps))
    loss = 1 - miouk
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss
