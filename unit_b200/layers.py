"""Operator-level drop-ins behind the reference's interfaces: ``ROIPooler``, ``Matcher``, ``subsample_labels``,
``label_and_sample``, ``Box2BoxTransform``, ``batched_nms``, ``fast_rcnn_inference``, ``mask_rcnn_inference``,
``paste_masks_in_image``, ``detector_postprocess``.

Same names, argument meaning and return structures as the Detectron2 / UniT objects they replace (reference call
sites cited per object); every tensor operation is a hand-written CUDA kernel reached through unit_b200.ops.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .registry import get_event_storage
from .structures import Boxes, Instances

_SCALE_CLAMP = math.log(1000.0 / 16)


def _adjacent_rows(tensors: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """The tensors as ONE view when they are consecutive row blocks of the same buffer (the per-image slices of a
    flat kernel output, e.g. what ``sample_from_draw`` hands out), else None."""
    first = tensors[0]
    if first.requires_grad or not first.is_contiguous() or first.dim() == 0 or ops._is_fake(first):
        return None
    base, rest = first.untyped_storage().data_ptr(), first.shape[1:]
    nxt, rows = first.storage_offset(), 0
    for t in tensors:
        if (t.requires_grad or t.dtype != first.dtype or t.shape[1:] != rest or not t.is_contiguous()
                or t.untyped_storage().data_ptr() != base or t.storage_offset() != nxt):
            return None
        nxt += t.numel()
        rows += t.shape[0]
    # strides computed from the shape: ``is_contiguous`` ignores the stride of a size-1 dim, so a 1-row slice of a
    # wider parent may carry a row stride that is wrong for the combined view
    strides, acc = [], 1
    for d in reversed((rows,) + tuple(rest)):
        strides.append(acc)
        acc *= max(int(d), 1)
    return first.as_strided((rows,) + tuple(rest), tuple(reversed(strides)), first.storage_offset())


def cat(tensors: List[torch.Tensor], dim: int = 0) -> torch.Tensor:
    """``torch.cat`` for kernel INPUTS: the result may alias the sources (a single tensor is returned as is, adjacent
    row blocks of one buffer come back as one view), so callers must treat it as read-only.  Everything in this
    package that mutates boxes in place (Boxes.clip / scale) works on tensors it owns, never on this result."""
    if len(tensors) == 1:
        return tensors[0]
    if dim == 0:
        view = _adjacent_rows(tensors)
        if view is not None:
            return view
    return torch.cat(tensors, dim)


def nonzero_tuple(x: torch.Tensor):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


# ------------------------------------------------------------------------------------------------- ROIPooler
class ROIPooler(torch.nn.Module):
    """[D2] ``ROIPooler`` (built by StandardROIHeads._init_box_head, reached via roi_heads.py:220, and at
    roi_heads.py:673-678).  ``forward(x: List[Tensor NCHW], box_lists: List[Boxes]) -> Tensor[sum R_i, C, P, P]``.

    "ROIAlignV2" -> aligned=True, "ROIAlign" -> aligned=False.  RoIs are concatenated in image order, which is what
    lets the slab-resident ROIAlign kernel walk (image, channel slab) units.
    """

    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2
        if pooler_type not in ("ROIAlign", "ROIAlignV2"):
            raise ValueError(f"unit_b200.ROIPooler supports ROIAlign / ROIAlignV2, got {pooler_type!r}")
        self.output_size = tuple(output_size)
        self.scales = tuple(scales)
        self.sampling_ratio = int(sampling_ratio)
        self.pooler_type = pooler_type
        self.aligned = pooler_type == "ROIAlignV2"
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1
        self.canonical_level = canonical_level
        self.canonical_box_size = canonical_box_size

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]) -> torch.Tensor:
        assert isinstance(x, list) and isinstance(box_lists, list)
        assert len(x) == len(self.scales)
        assert len(box_lists) == x[0].size(0), f"{len(box_lists)} box lists for batch {x[0].size(0)}"
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        rois = ops.boxes_to_rois(cat([b.tensor for b in box_lists]),
                                 ops.offsets_from_counts([len(b) for b in box_lists], x[0].device))
        if len(self.scales) == 1:
            return ops.roi_align(x[0], rois, self.output_size, self.scales[0], self.sampling_ratio, self.aligned,
                                 rois_sorted=True)
        # multi-level (FPN) assignment: [D2] assign_boxes_to_levels
        box_sizes = torch.sqrt(cat([b.area() for b in box_lists]))
        lv = torch.floor(self.canonical_level + torch.log2(box_sizes / self.canonical_box_size + 1e-8))
        lv = torch.clamp(lv, min=self.min_level, max=self.max_level).to(torch.int64) - self.min_level
        out = torch.zeros((rois.size(0), x[0].shape[1]) + self.output_size, dtype=x[0].dtype, device=x[0].device)
        for level, scale in enumerate(self.scales):
            inds = nonzero_tuple(lv == level)[0]
            if inds.numel():
                out.index_put_((inds,), ops.roi_align(x[level], rois[inds], self.output_size, scale,
                                                      self.sampling_ratio, self.aligned, rois_sorted=True))
        return out


# ------------------------------------------------------------------------------------------------- Matcher
class Matcher:
    """``Matcher`` of modeling/matcher.py:6-119 (UniT: 3 outputs incl. ``matched_vals``); constructed with
    ``return_vals=False`` it is the 2-output [D2] ``Matcher`` used by ROIHeads.label_and_sample_proposals."""

    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False,
                 return_vals: bool = True):
        thresholds = list(thresholds)
        assert thresholds[0] > 0
        assert all(low <= high for low, high in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in (-1, 0, 1) for l in labels)
        assert len(labels) == len(thresholds) + 1
        self.user_thresholds = thresholds
        self.thresholds = [-float("inf")] + thresholds + [float("inf")]
        self.labels = list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches
        self.return_vals = return_vals

    def __call__(self, match_quality_matrix: torch.Tensor):
        assert match_quality_matrix.dim() == 2
        out = ops.matcher(match_quality_matrix, self.user_thresholds, self.labels, self.allow_low_quality_matches,
                          want_vals=True)
        return out if self.return_vals else out[:2]


# ------------------------------------------------------------------------------------------------- sampling
def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     generator: Optional[torch.Generator] = None):
    """[D2] ``subsample_labels`` (roi_heads.py:415): the two ``randperm`` draws come from a HOST generator (the
    global CPU generator when ``generator`` is None) so that the sampled indices are reproducible bit for bit
    against the CPU reference; compaction and gathering stay on the GPU."""
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = torch.randperm(positive.numel(), generator=generator)[:num_pos].to(labels.device)
    perm2 = torch.randperm(negative.numel(), generator=generator)[:num_neg].to(labels.device)
    return positive[perm1], negative[perm2]


def _plain_rpn_fields(p: Instances) -> bool:
    """Proposals as the RPN emits them: fp32 CUDA proposal_boxes + objectness_logits and nothing else."""
    f = p.get_fields()
    if set(f) != {"proposal_boxes", "objectness_logits"}:
        return False
    b, l = f["proposal_boxes"].tensor, f["objectness_logits"]
    return b.is_cuda and b.dtype == torch.float32 and l.dtype == torch.float32 and l.dim() == 1 and not ops._is_fake(b)


def add_ground_truth_to_proposals(gt_boxes: List[Boxes], proposals: List[Instances]) -> List[Instances]:
    """[D2] proposal_utils.add_ground_truth_to_proposals: GT appended AFTER the RPN proposals."""
    assert len(proposals) == len(gt_boxes)
    out = []
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    if proposals and all(_plain_rpn_fields(p) for p in proposals) and all(g.tensor.dtype == torch.float32 for g in gt_boxes):
        # one launch for the whole batch; every image's fields are views of ONE buffer, so the labelling's cat is free
        boxes, logits = ops.append_gt([p.proposal_boxes.tensor for p in proposals],
                                      [p.objectness_logits for p in proposals], [g.tensor for g in gt_boxes],
                                      gt_logit_value)
        off = 0
        for gt_i, prop_i in zip(gt_boxes, proposals):
            n = len(prop_i) + len(gt_i)
            out.append(Instances(prop_i.image_size, proposal_boxes=Boxes(boxes[off:off + n]),
                                 objectness_logits=logits[off:off + n]))
            off += n
        return out
    for gt_i, prop_i in zip(gt_boxes, proposals):
        gt_prop = Instances(prop_i.image_size)
        gt_prop.proposal_boxes = gt_i
        gt_prop.objectness_logits = prop_i.objectness_logits.new_full((len(gt_i),), gt_logit_value)
        out.append(Instances.cat([prop_i, gt_prop]))
    return out


class LabelMatch:
    """Device-side result of the label phase (fused IoU+match, label+compaction) for a batch of images."""

    __slots__ = ("prop_counts", "gt_counts", "prop_boxes", "gt_boxes", "gt_classes", "po", "go", "matches", "mlabels",
                 "vals", "prop_classes", "pos_idx", "neg_idx", "counts", "prop_logits")


def label_match(proposals: List[Instances], targets: List[Instances], *, num_classes: int,
                thresholds: Sequence[float], labels: Sequence[int]) -> LabelMatch:
    """Phase 1 of ``label_and_sample``: two launches, no host synchronisation (capturable in a CUDA graph)."""
    dev = proposals[0].proposal_boxes.device
    lm = LabelMatch()
    lm.prop_counts = [len(p) for p in proposals]
    lm.gt_counts = [len(t) for t in targets]
    lm.prop_boxes = cat([p.proposal_boxes.tensor for p in proposals])
    # the one pass-through field RPN proposals carry, concatenated like the boxes (a view when the images already share
    # a buffer, as after add_ground_truth_to_proposals): gathered by the sampling launch itself
    lm.prop_logits = (cat([p.objectness_logits for p in proposals])
                      if all(p.has("objectness_logits") and p.objectness_logits.dtype == torch.float32
                             and p.objectness_logits.dim() == 1 for p in proposals) else None)
    lm.gt_boxes = cat([t.gt_boxes.tensor for t in targets])
    lm.gt_classes = cat([t.gt_classes for t in targets]).to(torch.int64)
    lm.po = ops.offsets_from_counts(lm.prop_counts, dev)
    lm.go = ops.offsets_from_counts(lm.gt_counts, dev)
    lm.matches, lm.mlabels, lm.vals = ops.iou_match(lm.gt_boxes, lm.go, lm.prop_boxes, lm.po, thresholds, labels,
                                                    want_vals=True)
    lm.prop_classes, lm.pos_idx, lm.neg_idx, lm.counts = ops.label_proposals(lm.matches, lm.mlabels, lm.gt_classes,
                                                                             lm.go, lm.po, num_classes)
    return lm


class SampleDraw:
    """Host-side result of the draw phase: the reference's ``randperm(#pos)``, ``randperm(#neg)`` per image, packed as
    [perm_pos | perm_neg | ppo pno pso nso] (int64).  With ``capacity`` the two permutation regions have that fixed
    length, so the device copy of the buffer has a static layout (CUDA-graph replay)."""

    __slots__ = ("host", "pso", "nso", "pos_len", "neg_len", "n_img")

    @property
    def sizes(self) -> List[int]:
        return [(self.pso[i + 1] - self.pso[i]) + (self.nso[i + 1] - self.nso[i]) for i in range(self.n_img)]


def draw_permutations(counts_h: Sequence[Sequence[int]], batch_size_per_image: int, positive_fraction: float,
                      generator: Optional[torch.Generator] = None, capacity: Optional[int] = None,
                      out: Optional[torch.Tensor] = None) -> SampleDraw:
    """Phase 2 ([D2] ``subsample_labels``, sampling.py): same draws, same order (pos then neg, image by image)."""
    n_img = len(counts_h)
    max_pos = int(batch_size_per_image * positive_fraction)
    perm_pos, perm_neg = [], []
    ppo, pno, pso, nso = [0], [0], [0], [0]
    for i in range(n_img):
        n_pos_all, n_neg_all = counts_h[i]
        num_pos = min(n_pos_all, max_pos)
        num_neg = min(n_neg_all, batch_size_per_image - num_pos)
        perm_pos.append(torch.randperm(n_pos_all, generator=generator))
        perm_neg.append(torch.randperm(n_neg_all, generator=generator))
        ppo.append(ppo[-1] + n_pos_all)
        pno.append(pno[-1] + n_neg_all)
        pso.append(pso[-1] + num_pos)
        nso.append(nso[-1] + num_neg)
    d = SampleDraw()
    d.pso, d.nso, d.n_img = pso, nso, n_img
    # the four int32 offset arrays travel in the tail of the int64 buffer, stored AS int32 (the device side views the
    # tail as int32: no conversion kernel in the step)
    offs = torch.tensor(ppo + pno + pso + nso, dtype=torch.int32)
    if capacity is None:
        d.pos_len, d.neg_len = ppo[-1], pno[-1]
        tail = torch.zeros(offs.numel(), dtype=torch.int64)
        tail.view(torch.int32)[:offs.numel()] = offs
        d.host = torch.cat(perm_pos + perm_neg + [tail])
    else:
        assert ppo[-1] <= capacity and pno[-1] <= capacity
        d.pos_len = d.neg_len = capacity
        d.host = out if out is not None else torch.empty(2 * capacity + offs.numel(), dtype=torch.int64)
        torch.cat(perm_pos, out=d.host[:ppo[-1]])
        torch.cat(perm_neg, out=d.host[capacity:capacity + pno[-1]])
        d.host[2 * capacity:].view(torch.int32)[:offs.numel()] = offs
    return d


_ROW_OFF_CACHE: dict = {}


def sample_from_draw(lm: LabelMatch, draw: SampleDraw, devbuf: torch.Tensor, proposals: List[Instances],
                     targets: List[Instances], want_vals: bool = False):
    """Phase 3: one gather launch + the Instances the reference's ``label_and_sample_proposals`` returns."""
    n_img = draw.n_img
    pso, nso = draw.pso, draw.nso
    S = pso[-1] + nso[-1]
    d_perm_pos = devbuf[:draw.pos_len]
    d_perm_neg = devbuf[draw.pos_len:draw.pos_len + draw.neg_len]
    offs = devbuf[draw.pos_len + draw.neg_len:].view(torch.int32)
    k = n_img + 1
    res_g = ops.sample_gather(
        lm.pos_idx, lm.neg_idx, d_perm_pos, offs[:k], d_perm_neg, offs[k:2 * k], offs[2 * k:3 * k], offs[3 * k:4 * k],
        lm.po, lm.go, S, lm.prop_boxes, lm.prop_classes, lm.matches, lm.gt_boxes, lm.prop_logits)
    sampled, s_boxes, s_classes, s_matched, s_gt = res_g[:5]
    s_logits = res_g[5] if lm.prop_logits is not None else None
    out, matched_list, vals_list = [], [], []
    off = 0
    num_fg, num_bg = [], []
    # pass-through proposal fields (objectness_logits ...): for more than two images ONE gather per field over the
    # concatenated field (global row = sampled row + the image's proposal offset) instead of one launch per image
    extra = [n for n in proposals[0].get_fields()
             if n != "proposal_boxes" and not (n == "objectness_logits" and s_logits is not None)] if proposals else []
    batched = {}
    if n_img > 2 and extra:
        per_img = [(pso[i + 1] - pso[i]) + (nso[i + 1] - nso[i]) for i in range(n_img)]
        starts, acc = [], 0
        for c in lm.prop_counts:
            starts.append(acc)
            acc += c
        key = (tuple(per_img), tuple(lm.prop_counts), str(sampled.device))
        row_off = _ROW_OFF_CACHE.get(key)
        if row_off is None:
            if len(_ROW_OFF_CACHE) > 64:
                _ROW_OFF_CACHE.clear()
            row_off = torch.repeat_interleave(torch.tensor(starts, dtype=torch.int64),
                                              torch.tensor(per_img, dtype=torch.int64)).to(sampled.device)
            if not ops._is_fake(row_off):
                _ROW_OFF_CACHE[key] = row_off
        gidx = sampled + row_off
        for name in extra:
            batched[name] = cat([p.get(name) for p in proposals])[gidx]
    for i, (p, t) in enumerate(zip(proposals, targets)):
        n = (pso[i + 1] - pso[i]) + (nso[i + 1] - nso[i])
        sl = slice(off, off + n)
        idx = sampled[sl]
        res = Instances(p.image_size)
        res.proposal_boxes = Boxes(s_boxes[sl])
        for name, value in p.get_fields().items():
            if name == "objectness_logits" and s_logits is not None:
                res.set(name, s_logits[sl])
            elif name != "proposal_boxes":
                res.set(name, batched[name][sl] if name in batched else value[idx])
        res.gt_classes = s_classes[sl]
        if lm.gt_counts[i] > 0:
            for name, value in t.get_fields().items():
                if name.startswith("gt_") and not res.has(name):
                    res.set(name, Boxes(s_gt[sl]) if name == "gt_boxes" else value[s_matched[sl]])
        else:
            res.gt_boxes = Boxes(s_gt[sl])
        out.append(res)
        matched_list.append(s_matched[sl])
        vals_list.append(lm.vals[po_slice(lm.prop_counts, i)][idx] if want_vals else None)
        num_fg.append(pso[i + 1] - pso[i])
        num_bg.append(nso[i + 1] - nso[i])
        off += n
    storage = get_event_storage()
    storage.put_scalar("roi_head/num_fg_samples", sum(num_fg) / max(len(num_fg), 1))
    storage.put_scalar("roi_head/num_bg_samples", sum(num_bg) / max(len(num_bg), 1))
    return out, matched_list, (vals_list if want_vals else None)


def label_and_sample(proposals: List[Instances], targets: List[Instances], *, num_classes: int,
                     batch_size_per_image: int, positive_fraction: float, thresholds: Sequence[float],
                     labels: Sequence[int], sample: bool = True, generator: Optional[torch.Generator] = None,
                     want_vals: bool = False):
    """Batched core of [D2] ``ROIHeads.label_and_sample_proposals`` (roi_heads.py:459,563,794,925) and of
    weak_detector_fast_rcnn.py:320-351 (``sample=False``: keep every proposal).

    All images go through three launches (fused IoU+match, label+compaction, gather) and ONE device->host read
    (the fg/bg counts the host needs to draw the reference's randperm(#pos), randperm(#neg)).
    Returns (proposals_with_gt, matched_idxs per image, matched_vals per image or None).
    """
    lm = label_match(proposals, targets, num_classes=num_classes, thresholds=thresholds, labels=labels)
    if not sample:
        out, matched_list, vals_list = [], [], []
        off = 0
        for i, (p, t) in enumerate(zip(proposals, targets)):
            sl = slice(off, off + lm.prop_counts[i])
            res = Instances(p.image_size, **p.get_fields())
            res.gt_classes = lm.prop_classes[sl]
            m = lm.matches[sl]
            if lm.gt_counts[i] > 0:
                for name, value in t.get_fields().items():
                    if name.startswith("gt_") and not res.has(name):
                        res.set(name, value[m])
            else:
                res.gt_boxes = Boxes(lm.prop_boxes.new_zeros((lm.prop_counts[i], 4)))
            out.append(res)
            matched_list.append(m)
            vals_list.append(lm.vals[sl])
            off += lm.prop_counts[i]
        return out, matched_list, (vals_list if want_vals else None)

    counts_h = lm.counts.cpu().tolist()  # the one host sync (the reference syncs per image on nonzero())
    draw = draw_permutations(counts_h, batch_size_per_image, positive_fraction, generator)
    devbuf = draw.host.to(lm.prop_boxes.device, non_blocking=True)
    return sample_from_draw(lm, draw, devbuf, proposals, targets, want_vals)


def po_slice(counts: Sequence[int], i: int) -> slice:
    start = sum(counts[:i])
    return slice(start, start + counts[i])


# ------------------------------------------------------------------------------------------------- boxes
class Box2BoxTransform:
    """[D2] ``Box2BoxTransform`` (fast_rcnn.py:335, weak_detector_fast_rcnn.py:119)."""

    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _SCALE_CLAMP):
        self.weights = tuple(float(w) for w in weights)
        self.scale_clamp = scale_clamp

    def apply_deltas(self, deltas: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        _, out = ops.softmax_decode(None, deltas, boxes, self.weights, self.scale_clamp, want_probs=False)
        return out

    def get_deltas(self, src_boxes: torch.Tensor, target_boxes: torch.Tensor) -> torch.Tensor:
        return ops.box_get_deltas(src_boxes, target_boxes, self.weights)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """[D2] layers.batched_nms == torchvision.ops.batched_nms on CUDA tensors (coordinate trick iff numel <= 100000)."""
    return ops.batched_nms(boxes, scores, idxs, iou_threshold, nms_mode=ops.NMS_TV_CUDA_RULE)


# ------------------------------------------------------------------------------------------------- inference
def fast_rcnn_inference_device(boxes: Sequence[torch.Tensor], scores: Sequence[torch.Tensor],
                                image_shapes: Sequence[Tuple[int, int]], score_thresh: float, nms_thresh: float,
                                topk_per_image: int, nms_mode: int = ops.NMS_TV_CUDA_RULE):
    """Device half of ``fast_rcnn_inference``: filter + class-wise NMS + top-k for all images in two launches, nothing
    read back (capturable).  Returns padded (det_boxes [n,topk,4], det_scores, det_classes, det_roi, det_counts)."""
    dev = boxes[0].device
    counts = [b.shape[0] for b in boxes]
    off = ops.offsets_from_counts(counts, dev)
    hw = ops.f32_const([[float(h), float(w)] for (h, w) in image_shapes], dev)
    db, ds, dc, dr, cnt, _ = ops.detect(cat(list(boxes)), cat(list(scores)), off, hw, score_thresh, nms_thresh,
                                        topk_per_image, nms_mode)
    return db, ds, dc, dr, cnt


def instances_from_detections(dets, image_shapes: Sequence[Tuple[int, int]]):
    """Host half: ONE read of the per-image detection counts, then views of the padded device outputs."""
    db, ds, dc, dr, cnt = dets
    cnt_h = cnt.cpu().tolist()
    results, kept = [], []
    for i, shape in enumerate(image_shapes):
        n = cnt_h[i]
        inst = Instances(tuple(shape))
        inst.pred_boxes = Boxes(db[i, :n])
        inst.scores = ds[i, :n]
        inst.pred_classes = dc[i, :n]
        results.append(inst)
        kept.append(dr[i, :n])
    return results, kept


def fast_rcnn_inference(boxes: Sequence[torch.Tensor], scores: Sequence[torch.Tensor],
                        image_shapes: Sequence[Tuple[int, int]], score_thresh: float, nms_thresh: float,
                        topk_per_image: int, nms_mode: int = ops.NMS_TV_CUDA_RULE):
    """[D2] ``fast_rcnn_inference`` (fast_rcnn.py:461-468, weak_detector_fast_rcnn.py:299-306, rcnn.py:526):
    all images in two launches and one host read of the detection counts.
    Returns (List[Instances(pred_boxes, scores, pred_classes)], List[kept RoI index])."""
    if len(boxes) == 0:
        return [], []
    dets = fast_rcnn_inference_device(boxes, scores, image_shapes, score_thresh, nms_thresh, topk_per_image, nms_mode)
    return instances_from_detections(dets, image_shapes)


def mask_rcnn_inference(pred_mask_logits: torch.Tensor, pred_instances: List[Instances]) -> None:
    """[D2] ``mask_rcnn_inference`` for logits that already went through the transfer (class select + sigmoid)."""
    if pred_mask_logits.size(1) == 1:
        probs = pred_mask_logits.sigmoid()
    else:
        cls = cat([i.pred_classes for i in pred_instances])
        probs = pred_mask_logits[torch.arange(pred_mask_logits.shape[0], device=cls.device), cls][:, None].sigmoid()
    for prob, inst in zip(probs.split([len(i) for i in pred_instances], dim=0), pred_instances):
        inst.pred_masks = prob


def paste_masks_in_image(masks: torch.Tensor, boxes, image_shape: Tuple[int, int], threshold: float = 0.5):
    """[D2] ``paste_masks_in_image`` -> bool [D,H,W] (one write-bound kernel, no 1 GB chunking)."""
    if not isinstance(boxes, torch.Tensor):
        boxes = boxes.tensor
    if len(masks) == 0:
        return masks.new_empty((0,) + tuple(image_shape), dtype=torch.uint8)
    return ops.mask_paste(masks, boxes, image_shape, threshold)


def detector_postprocess(results: Instances, output_height: int, output_width: int, mask_threshold: float = 0.5):
    """[D2] ``detector_postprocess`` (meta_arch/rcnn.py:423): rescale, clip, drop empty, paste masks."""
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances((output_height, output_width), **results.get_fields())
    output_boxes = results.pred_boxes if results.has("pred_boxes") else results.proposal_boxes
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    results = results[output_boxes.nonempty()]
    if results.has("pred_masks"):
        results.pred_masks = paste_masks_in_image(results.pred_masks[:, 0, :, :], results.pred_boxes,
                                                  results.image_size, threshold=mask_threshold)
    return results


def select_foreground_proposals(proposals: List[Instances], bg_label: int):
    """[D2] ``select_foreground_proposals`` (roi_heads.py:344,696,892)."""
    fg, masks = [], []
    for p in proposals:
        gt_classes = p.gt_classes
        m = (gt_classes != -1) & (gt_classes != bg_label)
        fg.append(p[m.nonzero().squeeze(1)])
        masks.append(m)
    return fg, masks
