"""Oracle restatement of the UniT side of the RoI stage (functional, CPU, fp32).  TEST INFRASTRUCTURE.

Every function cites the reference lines it follows.  It is pinned against the reference's own files executed
verbatim (oracle.shim) by tests/test_oracle.py and against the committed fixtures in tests/golden/.
Weights are passed as a flat dict keyed by the reference's state_dict names (SURVEY.md section 5 "Checkpoint"):
  cls_score_delta.{weight,bias}  bbox_pred_delta.{weight,bias}  cls_score_ft.*  bbox_pred_ft.*
  weak_detector_head.oicr_predictors.{0,1,2}.{weight,bias}  embeddings.weight
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch.nn import functional as F

from .d2.ops import Box2BoxTransform, MatcherWithVals, fast_rcnn_inference, mask_rcnn_inference
from .d2.structures import Boxes, Instances

# modeling/roi_heads/roi_heads.py:191 -- COCO-80 name -> row of glove_mean
COCO_NAME_TO_ID = {n: i for i, n in enumerate([
    'person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light',
    'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant',
    'bear', 'zebra', 'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard',
    'sports ball', 'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket', 'bottle',
    'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli',
    'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair', 'couch', 'potted plant', 'bed', 'dining table', 'toilet',
    'tv', 'laptop', 'mouse', 'remote', 'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink',
    'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear', 'hair drier', 'toothbrush'])}
_RENAME = {'aeroplane': 'airplane', 'diningtable': 'dining table', 'motorbike': 'motorcycle',
           'pottedplant': 'potted plant', 'sofa': 'couch', 'tvmonitor': 'tv'}


def coco_indexer(thing_classes: Sequence[str]) -> torch.Tensor:
    """roi_heads.py:190-216 ``_class_mappings``: dataset class order -> glove_mean row."""
    return torch.tensor([COCO_NAME_TO_ID[_RENAME.get(n, n)] for n in thing_classes], dtype=torch.long)


def lingual_similarity(embeddings: torch.Tensor, indexer: torch.Tensor, base: torch.Tensor, novel: torch.Tensor):
    """fast_rcnn.py:376-382 ``get_similarity``: E[idx][novel] @ E[idx][base]^T  -> [N,B] (raw, pre-softmax)."""
    label_embeddings = embeddings[indexer]
    return torch.mm(label_embeddings.index_select(0, novel), label_embeddings.index_select(0, base).t())


def oicr_mean_logits(x: torch.Tensor, w: Dict[str, torch.Tensor], oicr_iter: int = 3) -> torch.Tensor:
    """weak_detector_fast_rcnn.py:172-175 + the mean at roi_heads.py:252 / fast_rcnn.py:366."""
    outs = [F.linear(x, w[f"weak_detector_head.oicr_predictors.{k}.weight"],
                     w[f"weak_detector_head.oicr_predictors.{k}.bias"]) for k in range(oicr_iter)]
    return torch.mean(torch.stack(outs, 0), 0)


def visual_similarity(probs: torch.Tensor, base: torch.Tensor, threshold: float) -> torch.Tensor:
    """roi_heads.py:255-257: softmax over all K+1, take base columns, renormalise (clamp 1e-9), zero below thr."""
    v = torch.softmax(probs, -1).index_select(1, base)
    v = v / v.sum(-1, keepdim=True).clamp(min=1e-9)
    v = v.clone()
    v[v < threshold] = 0
    return v


def _first_k(terms: List[str], tag: str) -> int:
    """roi_heads.py:274,285,296,307: ``int([x for x in terms if tag in x][0].split("-")[1])`` (substring match)."""
    return int([x for x in terms if tag in x][0].split("-")[1])


def similarity_matrices(lingual: Optional[torch.Tensor], visual: Optional[torch.Tensor],
                        terms: Dict[str, List[str]], n_novel: int, n_base: int, combination: str = "Sum",
                        class_weights: Optional[torch.Tensor] = None, mean_logits: Optional[torch.Tensor] = None,
                        base: Optional[torch.Tensor] = None, novel: Optional[torch.Tensor] = None,
                        num_classes: Optional[int] = None):
    """roi_heads.py:266-336, every term.  ``class_weights`` = mean of the OICR predictor weights [K+1,D] (TopK / WTopK
    / LSDA, :275,286,297), ``mean_logits`` = mean OICR logits of the RoIs [R,K+1] (VisualK, :308).  The reference tests
    ``'TopK' in x`` by SUBSTRING, so a ``WTopK-k`` term also switches the ``TopK`` branch on; reproduced."""
    dev = (visual if visual is not None else (lingual if lingual is not None else class_weights)).device
    similarity = {}
    for head_type, tl in terms.items():
        s = torch.zeros(n_novel, n_base, device=dev)
        if combination == "Sum":
            weight = 1.0 / len(tl) if len(tl) else 0.0
            if "lingual" in tl:
                s = s + weight * torch.softmax(lingual, dim=-1)
            if any("TopK" in x for x in tl):
                k = _first_k(tl, "TopK")
                ws = torch.mm(class_weights.index_select(0, novel), class_weights.index_select(0, base).t())
                _, idx = torch.topk(ws, k, dim=-1)
                t = torch.zeros(n_novel, n_base, device=dev).scatter(1, idx, 1.0)
                s = s + weight * (t / torch.sum(t, dim=-1, keepdim=True))
            if any("WTopK" in x for x in tl):
                k = _first_k(tl, "WTopK")
                ws = torch.mm(class_weights.index_select(0, novel), class_weights.index_select(0, base).t())
                top, idx = torch.topk(ws, k, dim=-1)
                t = torch.zeros(n_novel, n_base, device=dev).scatter(1, idx, top)
                s = s + weight * (t / torch.sum(t, dim=-1, keepdim=True))
            if any("LSDA" in x for x in tl):
                k = _first_k(tl, "LSDA")
                ws = torch.norm(class_weights.index_select(0, novel).unsqueeze(1) -
                                class_weights.index_select(0, base).unsqueeze(0), dim=-1)
                _, idx = torch.topk(ws, k, dim=-1, largest=False)
                t = torch.zeros(n_novel, n_base, device=dev).scatter(1, idx, 1.0)
                s = s + weight * (t / torch.sum(t, dim=-1, keepdim=True))
            if any("VisualK" in x for x in tl):
                k = _first_k(tl, "VisualK")
                cw = torch.softmax(mean_logits.narrow(1, 0, num_classes), -1).index_select(1, base)
                ws = cw / torch.sum(cw, -1, keepdim=True).clamp(min=1e-9)
                top, idx = torch.topk(ws, k, dim=-1)
                t = torch.zeros(ws.size(0), n_base, device=dev).scatter(1, idx, top)
                t = t / torch.sum(t, dim=-1, keepdim=True)
                s = s.unsqueeze(0) + weight * t.unsqueeze(1)
            if "visual" in tl:
                s = s.unsqueeze(0) + weight * visual.unsqueeze(1)  # 4-D after VisualK, as in the reference
            if "Average" in tl:
                s = s.fill_(1.0)
                s = s / torch.sum(s, dim=-1, keepdim=True)
            if len(tl) > 0 and ("None" not in tl):
                s = s / torch.sum(s, dim=-1, keepdim=True).clamp(min=1e-9)
            else:
                s = 0.0 * s
        else:
            if "lingual" in tl:
                s = s * lingual
            if "visual" in tl:
                s = s.unsqueeze(0) * visual.unsqueeze(1)
            if len(tl) > 0:
                s = torch.softmax(s, -1)
        similarity[head_type] = s
    return similarity


def transfer(delta_scores: torch.Tensor, proposal_deltas: torch.Tensor, similarity: Dict[str, torch.Tensor],
             base: torch.Tensor, novel: torch.Tensor, num_classes: int, box_dim: int = 4):
    """fast_rcnn.py:403-423 (== :503-523): novel logits += S_cls . base logits; novel box deltas := S_bbox . base."""
    transfered = torch.zeros_like(delta_scores)
    base_scores = delta_scores.index_select(1, base)
    if similarity["cls"].dim() > 2:
        t = torch.bmm(similarity["cls"], base_scores.unsqueeze(2)).squeeze(2)
    else:
        t = torch.mm(base_scores, similarity["cls"].t())
    transfered = transfered.index_copy(1, novel, t)
    delta_scores = delta_scores + transfered

    pd = proposal_deltas.view(-1, num_classes, box_dim)
    base_pd = pd.index_select(1, base)
    out = torch.zeros_like(pd)
    if similarity["bbox"].dim() > 2:
        tb = torch.bmm(similarity["bbox"], base_pd)
    else:
        tb = torch.matmul(base_pd.transpose(1, 2), similarity["bbox"].t()).transpose(1, 2)
    out = out.index_copy(1, novel, tb)
    out = out.index_copy(1, base, base_pd)
    return delta_scores, out.view(-1, num_classes * box_dim)


def predictor_forward(x: torch.Tensor, x_weak_branch: Optional[torch.Tensor], w: Dict[str, torch.Tensor],
                      similarity: Optional[Dict[str, torch.Tensor]], base: torch.Tensor, novel: torch.Tensor,
                      num_classes: int, kind: str = "Base", training: bool = False):
    """SupervisedDetectorOutputs{Base,FineTune,WeakFineTune}.forward (fast_rcnn.py:384-433, 484-533, 542-585)
    with ``x_weak=None`` (no weak-image branch; that branch is training-only MIL/OICR, out of scope).

    x: box_head features [R,D]; x_weak_branch: weak_box_head features (``supervised_branch_x_weak``) or None.
    Returns (scores [R,K+1], bbox [R,4K]).
    """
    delta_scores = F.linear(x, w["cls_score_delta.weight"], w["cls_score_delta.bias"])
    proposal_deltas = F.linear(x, w["bbox_pred_delta.weight"], w["bbox_pred_delta.bias"])
    weak_scores = oicr_mean_logits(x if x_weak_branch is None else x_weak_branch, w)
    do_transfer = similarity is not None and (kind != "Base" or not training)
    if do_transfer:
        delta_scores, proposal_deltas = transfer(delta_scores, proposal_deltas, similarity, base, novel, num_classes)
    scores = delta_scores + weak_scores          # get_cls_logits, fast_rcnn.py:360-368 (oicr_iter > 0 branch)
    bbox = proposal_deltas + 0.0                 # get_cls_bbox with the all-zero weak box output (:181)
    if kind == "FineTune":
        scores = scores + F.linear(x, w["cls_score_ft.weight"], w["cls_score_ft.bias"])
        bbox = bbox + F.linear(x, w["bbox_pred_ft.weight"], w["bbox_pred_ft.bias"])
    if kind == "Base" and training:
        scores = scores.index_fill(1, novel, -float("inf"))   # fast_rcnn.py:427-428
    return scores, bbox


def mask_transfer(logits: torch.Tensor, similarity_seg: torch.Tensor, base: torch.Tensor, novel: torch.Tensor,
                  x_delta: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mask_head.py:18-31 / 74-88: novel mask logits := S_seg . base mask logits (base copied through)."""
    if logits.numel() > 0:
        mask_base = logits.index_select(1, base)
        flat = mask_base.view(*mask_base.size()[:2], -1)
        if similarity_seg.dim() > 2:
            comb = torch.bmm(similarity_seg, flat)
        else:
            comb = torch.matmul(flat.transpose(1, 2), similarity_seg.t()).transpose(1, 2)
        mask_novel = comb.view(mask_base.size(0), -1, *mask_base.size()[2:])
        final = torch.zeros_like(logits)
        final = final.index_copy(1, novel, mask_novel)
        final = final.index_copy(1, base, mask_base)
        logits = final
    if x_delta is not None:
        logits = logits + x_delta
    return logits


def box_inference(scores: torch.Tensor, bbox: torch.Tensor, proposal_boxes: List[torch.Tensor],
                  image_sizes: List[tuple], score_thresh: float = 0.05, nms_thresh: float = 0.5, topk: int = 100,
                  weights=(10.0, 10.0, 5.0, 5.0)):
    """``box_predictor.inference`` (fast_rcnn.py:455-468): softmax, apply_deltas, fast_rcnn_inference."""
    n = [len(b) for b in proposal_boxes]
    probs = F.softmax(scores, dim=-1).split(n, dim=0)
    boxes = Box2BoxTransform(weights).apply_deltas(bbox, torch.cat(proposal_boxes, 0)).split(n)
    return fast_rcnn_inference(boxes, probs, image_sizes, score_thresh, nms_thresh, topk)


UniTMatcher = MatcherWithVals  # modeling/matcher.py:54-98


# ------------------------------------------------------------------------------------ weak-image training losses
def mil_scores(cls_logits: torch.Tensor, det_logits: torch.Tensor, counts: Sequence[int]):
    """weak_detector_fast_rcnn.py:202-210: per image, softmax over classes times softmax over the image's proposals;
    returns (x [R,K], class_vectors [n_img,K])."""
    xs, vecs = [], []
    for c, d in zip(cls_logits.split(list(counts)), det_logits.split(list(counts))):
        x = torch.softmax(c, -1) * torch.softmax(d, 0)
        xs.append(x)
        vecs.append(x.sum(0))
    return torch.cat(xs, 0), torch.stack(vecs)


def mil_loss(class_vectors: torch.Tensor, image_classes: Sequence[torch.Tensor], multiplier: float,
             eps: float = 1e-6) -> torch.Tensor:
    """weak_detector_fast_rcnn.py:211-216,247-250: BCE between the clamped class vector and the multi-hot labels."""
    gt = torch.zeros_like(class_vectors)
    for i, c in enumerate(image_classes):
        gt[i, torch.unique(c)] = 1.0
    return F.binary_cross_entropy(class_vectors.clamp(eps, 1 - eps), gt) * multiplier


def oicr_targets(probs: torch.Tensor, boxes: Sequence[torch.Tensor], image_classes: Sequence[torch.Tensor],
                 thresholds: Sequence[float], match_labels: Sequence[int], bg_threshold: float, num_classes: int):
    """compute_loss_inputs (weak_detector_fast_rcnn.py:384-396) = get_proposal_clusters (:353-382) followed by
    label_and_sample_proposals (:320-351).  Returns (labels i64 [R], cls_weights [R], picked: per image the index of
    the proposal chosen for every unique class)."""
    matcher = UniTMatcher(list(thresholds), list(match_labels), allow_low_quality_matches=False)
    labels, weights, picked = [], [], []
    start = 0
    for b, cls in zip(boxes, image_classes):
        p = probs[start:start + len(b)].clone()
        start += len(b)
        uniq = torch.unique(cls)
        idxs, scores = [], []
        for c in uniq.tolist():  # :358-365 -- best proposal of the class, then its whole row is zeroed
            s, i = p[:, c].max(dim=0)
            idxs.append(int(i))
            scores.append(s)
            p[int(i), :] = 0.0
        picked.append(torch.tensor(idxs, dtype=torch.int64))
        pseudo = b[idxs]
        iou = pairwise_iou_tensor(pseudo, b)
        matched_idxs, matched_labels, matched_vals = matcher(iou)
        gt = uniq[matched_idxs].clone()  # :308-318
        gt[matched_labels == 0] = num_classes
        gt[matched_labels == -1] = -1
        w = torch.stack(scores)[matched_idxs].clone()  # :392-396
        if bg_threshold > 0.0:
            w[matched_vals < bg_threshold] = 0.0
        labels.append(gt)
        weights.append(w)
    return torch.cat(labels), torch.cat(weights), picked


def pairwise_iou_tensor(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    from .d2.structures import pairwise_iou
    return pairwise_iou(Boxes(b1), Boxes(b2))


def weak_losses(cls_logits: torch.Tensor, det_logits: torch.Tensor, oicr_scores: Sequence[torch.Tensor],
                boxes: Sequence[torch.Tensor], image_classes: Sequence[torch.Tensor], thresholds=(0.5,),
                match_labels=(0, 1), bg_threshold: float = 0.1, multiplier: float = 1.0):
    """WeakDetectorOutputsBase.losses for TYPE == "OICR" without regression branches
    (weak_detector_fast_rcnn.py:189-228).  cls_logits / det_logits are already divided by their temperatures."""
    counts = [len(b) for b in boxes]
    K = cls_logits.shape[1]
    x, vecs = mil_scores(cls_logits, det_logits, counts)
    out = {"loss_im_cls": mil_loss(vecs, image_classes, multiplier)}
    sup = []
    probs = x.detach()
    for idx, score in enumerate(oicr_scores):
        if idx > 0:
            probs = torch.softmax(oicr_scores[idx - 1].detach(), dim=-1)
        labels, weights, picked = oicr_targets(probs, boxes, image_classes, thresholds, match_labels, bg_threshold, K)
        sup.append((labels, weights, picked))
        if score.numel() > 0:  # :220-227
            out["loss_oicr_{}".format(idx + 1)] = (F.cross_entropy(score, labels, reduction="none") * weights).mean()
        else:
            out["loss_oicr_{}".format(idx + 1)] = 0.0 * score.sum()
    return out, sup
