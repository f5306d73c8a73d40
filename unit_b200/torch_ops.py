"""``torch.library`` registration of the RoI-stage kernels: ``torch.ops.unit_b200.*``.

The north_star asks for "a thin C-ABI torch custom-op layer": every op below is a ``torch.library.custom_op`` whose real
implementation enqueues the sm_100a kernels of libunit_b200.so through the C ABI (ctypes, ``unit_b200.ops``), with
  * ``register_fake``      -- shapes / dtypes only, so the heads trace under FakeTensor / torch.export / AOT autograd
                              without touching a GPU (tests/test_host.py::test_heads_trace_under_fake_tensor), and
  * ``register_autograd``  -- backward formulas expressed with the backward ops of the same library.
``unit_b200.ops`` routes its public entry points through these ops, so the registered heads (roi_heads.py,
predictors.py, layers.py) are built from them.  There is still no CPU kernel: calling an op with CPU tensors raises.

Ops (all under the ``unit_b200::`` namespace)
  boxes_to_rois, roi_align, roi_align_backward            [D2] ROIPooler -> [TV] roi_align / _roi_align_backward
  iou_match                                               [D2] pairwise_iou + modeling/matcher.py
  predictor_linear, predictor_wgrad, predictor_gemm2      fast_rcnn.py:386-392, 486-489 (tcgen05 TF32)
  similarity_transfer, similarity_transfer_bwd, similarity_transfer_bwd_vis
                                                          roi_heads.py:245-336 + fast_rcnn.py:503-528
  softmax_decode, detect                                  fast_rcnn.py:455-468 -> [D2] fast_rcnn_inference
  mask_paste                                              [D2] paste_masks_in_image
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd

from . import ops as _k  # the ctypes-level implementations (``*_impl``)

_NS = "unit_b200"


def _empty(like: Tensor, shape, dtype=None) -> Tensor:
    return like.new_empty(tuple(int(s) for s in shape), dtype=dtype or like.dtype)


# ------------------------------------------------------------------------------------------------- ROIAlign
@custom_op(f"{_NS}::boxes_to_rois", mutates_args=())
def boxes_to_rois(boxes: Tensor, offsets: Tensor) -> Tensor:
    return _k._boxes_to_rois_impl(boxes, offsets)


@boxes_to_rois.register_fake
def _(boxes, offsets):
    return _empty(boxes, (boxes.shape[0], 5), torch.float32)


@custom_op(f"{_NS}::roi_align", mutates_args=())
def roi_align(feat: Tensor, rois: Tensor, pooled_h: int, pooled_w: int, spatial_scale: float, sampling_ratio: int,
              aligned: bool, rois_sorted: bool) -> Tensor:
    return _k._roi_align_forward_impl(feat, rois, (pooled_h, pooled_w), spatial_scale, sampling_ratio, aligned,
                                      rois_sorted)


@roi_align.register_fake
def _(feat, rois, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned, rois_sorted):
    return _empty(feat, (rois.shape[0], feat.shape[1], pooled_h, pooled_w))


@custom_op(f"{_NS}::roi_align_backward", mutates_args=())
def roi_align_backward(grad_out: Tensor, rois: Tensor, n: int, c: int, h: int, w: int, spatial_scale: float,
                       sampling_ratio: int, aligned: bool, rois_sorted: bool) -> Tensor:
    return _k._roi_align_backward_impl(grad_out, rois, (n, c, h, w), spatial_scale, sampling_ratio, aligned, rois_sorted)


@roi_align_backward.register_fake
def _(grad_out, rois, n, c, h, w, spatial_scale, sampling_ratio, aligned, rois_sorted):
    return _empty(grad_out, (n, c, h, w))


def _roi_align_setup(ctx, inputs, output):
    feat, rois, ph, pw, scale, sr, aligned, srt = inputs
    ctx.save_for_backward(rois)
    ctx.cfg = (tuple(feat.shape), scale, sr, aligned, srt)


def _roi_align_bwd(ctx, grad_out):
    (rois,) = ctx.saved_tensors
    (n, c, h, w), scale, sr, aligned, srt = ctx.cfg
    g = torch.ops.unit_b200.roi_align_backward(grad_out.contiguous(), rois, n, c, h, w, scale, sr, aligned, srt)
    return g, None, None, None, None, None, None, None


register_autograd(f"{_NS}::roi_align", _roi_align_bwd, setup_context=_roi_align_setup)


# ------------------------------------------------------------------------------------------------- IoU + Matcher
@custom_op(f"{_NS}::iou_match", mutates_args=())
def iou_match(gt_boxes: Tensor, gt_offsets: Tensor, prop_boxes: Tensor, prop_offsets: Tensor,
              thresholds: Sequence[float], labels: Sequence[int]) -> Tuple[Tensor, Tensor, Tensor]:
    return _k._iou_match_impl(gt_boxes, gt_offsets, prop_boxes, prop_offsets, list(thresholds), list(labels), True)


@iou_match.register_fake
def _(gt_boxes, gt_offsets, prop_boxes, prop_offsets, thresholds, labels):
    p = prop_boxes.shape[0]
    return (_empty(prop_boxes, (p,), torch.int64), _empty(prop_boxes, (p,), torch.int8),
            _empty(prop_boxes, (p,), torch.float32))


# ------------------------------------------------------------------------------------------------- predictor GEMMs
@custom_op(f"{_NS}::predictor_linear", mutates_args=())
def predictor_linear(x: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tensor:
    return _k._predictor_gemm_forward_impl(x, w, bias)


@predictor_linear.register_fake
def _(x, w, bias):
    return _empty(x, (x.shape[0], w.shape[0]), torch.float32)


@custom_op(f"{_NS}::predictor_wgrad", mutates_args=())
def predictor_wgrad(gy: Tensor, x: Tensor) -> Tuple[Tensor, Tensor]:
    """(gy^T x, column sums of gy) on the tcgen05 tensor cores."""
    n = gy.shape[1]
    ld = (n + 127) // 128 * 128
    gp = gy.new_zeros((gy.shape[0], ld))
    gp[:, :n] = gy
    gw = gy.new_empty((n, x.shape[1]))
    gb = gy.new_empty((n,))
    _k.predictor_wgrad(gp, x, n, [0, n], [gw], [gb], [None], accumulate=False)
    return gw, gb


@predictor_wgrad.register_fake
def _(gy, x):
    return _empty(gy, (gy.shape[1], x.shape[1])), _empty(gy, (gy.shape[1],))


def _linear_setup(ctx, inputs, output):
    x, w, bias = inputs
    ctx.save_for_backward(x, w)
    ctx.has_bias = bias is not None


def _linear_bwd(ctx, gy):
    x, w = ctx.saved_tensors
    gx = gw = gb = None
    if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
        gw, gb = torch.ops.unit_b200.predictor_wgrad(gy.contiguous(), x)
        if not ctx.has_bias:
            gb = None
    if ctx.needs_input_grad[0]:
        # d/dx = gy . W is a [R,N] x [N,K] product with N <= a few hundred: a plain library GEMM (not on the fine-tune
        # path, where the box-head features carry no gradient)
        gx = gy @ w
    return gx, gw, gb


register_autograd(f"{_NS}::predictor_linear", _linear_bwd, setup_context=_linear_setup)


@custom_op(f"{_NS}::predictor_gemm2", mutates_args=())
def predictor_gemm2(x1: Tensor, w1: Tensor, b1: Optional[Tensor], x2: Tensor, w2: Tensor,
                    b2: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    y1, y2 = _k._predictor_gemm2_impl(x1, w1, b1, x2, w2, b2)
    return y1, y2


@predictor_gemm2.register_fake
def _(x1, w1, b1, x2, w2, b2):
    pad = lambda n: (n + 31) // 32 * 32
    return (_empty(x1, (x1.shape[0], pad(w1.shape[0])), torch.float32),
            _empty(x1, (x1.shape[0], pad(w2.shape[0])), torch.float32))


# ------------------------------------------------------------------------------------------------- transfer
@custom_op(f"{_NS}::lingual_similarity", mutates_args=())
def lingual_similarity(emb: Tensor, indexer: Tensor, base: Tensor, novel: Tensor) -> Tuple[Tensor, Tensor]:
    return _k._lingual_similarity_impl(emb, indexer, base, novel)


@lingual_similarity.register_fake
def _(emb, indexer, base, novel):
    shape = (novel.numel(), base.numel())
    return _empty(emb, shape, torch.float32), _empty(emb, shape, torch.float32)


@custom_op(f"{_NS}::similarity_transfer", mutates_args=())
def similarity_transfer(vis_logits: Optional[Tensor], static_cls: Optional[Tensor], static_bbox: Optional[Tensor],
                        static_seg: Optional[Tensor], base: Tensor, novel: Tensor, class_kind: Tensor,
                        delta_scores: Tensor, proposal_deltas: Tensor, weak_scores: Optional[Tensor],
                        ft_scores: Optional[Tensor], ft_deltas: Optional[Tensor], vis_threshold: float,
                        wv: Sequence[float], norm: Sequence[int], do_transfer: bool, novel_neg_inf: bool,
                        static_per_roi: int, want: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Fused visual similarity + combination + base->novel transfer.  ``want`` bit h: materialise S of head h
    (cls, bbox, seg); heads not wanted come back as empty tensors."""
    spec = _k.TransferSpec.from_tensors(int(class_kind.numel()), base, novel, class_kind,
                                        {"cls": static_cls, "bbox": static_bbox, "seg": static_seg}, list(wv),
                                        list(norm), vis_threshold, static_per_roi)
    heads = tuple(h for i, h in enumerate(("cls", "bbox", "seg")) if (want >> i) & 1)
    s, b, sims = _k._similarity_transfer_forward_impl(spec, vis_logits, delta_scores, proposal_deltas, weak_scores,
                                                      ft_scores, ft_deltas, do_transfer, novel_neg_inf, heads)
    pick = lambda t: t if t is not None else s.new_empty((0,))  # a fresh tensor each: outputs must not alias
    return s, b, pick(sims["cls"]), pick(sims["bbox"]), pick(sims["seg"])


@similarity_transfer.register_fake
def _(vis_logits, static_cls, static_bbox, static_seg, base, novel, class_kind, delta_scores, proposal_deltas,
      weak_scores, ft_scores, ft_deltas, vis_threshold, wv, norm, do_transfer, novel_neg_inf, static_per_roi, want):
    r = delta_scores.shape[0]
    f32 = torch.float32
    sim = lambda i: _empty(delta_scores, (r, novel.numel(), base.numel()) if (want >> i) & 1 else (0,), f32)
    return (_empty(delta_scores, tuple(delta_scores.shape), f32), _empty(delta_scores, tuple(proposal_deltas.shape), f32),
            sim(0), sim(1), sim(2))


# ------------------------------------------------------------------------------------------------- decode / detect
@custom_op(f"{_NS}::softmax_decode", mutates_args=())
def softmax_decode(scores: Optional[Tensor], deltas: Optional[Tensor], proposals: Optional[Tensor],
                   weights: Sequence[float], scale_clamp: float) -> Tuple[Tensor, Tensor]:
    """softmax(scores) and/or Box2BoxTransform.apply_deltas(deltas, proposals); an absent half is an empty tensor."""
    any_t = scores if scores is not None else deltas
    probs, boxes = _k._softmax_decode_impl(scores, deltas, proposals, tuple(weights), scale_clamp, scores is not None,
                                           deltas is not None)
    e = any_t.new_empty((0,), dtype=torch.float32)
    return probs if probs is not None else e, boxes if boxes is not None else e


@softmax_decode.register_fake
def _(scores, deltas, proposals, weights, scale_clamp):
    any_t = scores if scores is not None else deltas
    f32 = torch.float32
    return (_empty(any_t, tuple(scores.shape) if scores is not None else (0,), f32),
            _empty(any_t, tuple(deltas.shape) if deltas is not None else (0,), f32))


@custom_op(f"{_NS}::detect", mutates_args=())
def detect(boxes: Tensor, probs: Tensor, roi_offsets: Tensor, image_hw: Tensor, score_thresh: float,
           nms_thresh: float, topk: int, nms_mode: int) -> List[Tensor]:
    """fast_rcnn_inference for a batch -> [det_boxes, det_scores, det_classes, det_roi, det_counts, cand_boxes,
    cand_scores, cand_roi, cand_cls, cand_counts]."""
    db, ds, dc, dr, cnt, cands = _k._detect_impl(boxes, probs, roi_offsets, image_hw, score_thresh, nms_thresh, topk,
                                                 nms_mode)
    return [db, ds, dc, dr, cnt, *cands]


@detect.register_fake
def _(boxes, probs, roi_offsets, image_hw, score_thresh, nms_thresh, topk, nms_mode):
    n_img = roi_offsets.numel() - 1
    r, k1 = probs.shape
    cap = max(r * (k1 - 1), 1)
    tk = topk if topk >= 0 else cap
    f32, i32, i64 = torch.float32, torch.int32, torch.int64
    return [_empty(boxes, (n_img, tk, 4), f32), _empty(boxes, (n_img, tk), f32), _empty(boxes, (n_img, tk), i64),
            _empty(boxes, (n_img, tk), i64), _empty(boxes, (max(n_img, 1),), i32), _empty(boxes, (cap, 4), f32),
            _empty(boxes, (cap,), f32), _empty(boxes, (cap,), i32), _empty(boxes, (cap,), i32),
            _empty(boxes, (max(n_img, 1),), i32)]


# ------------------------------------------------------------------------------------------------- masks
@custom_op(f"{_NS}::mask_paste", mutates_args=())
def mask_paste(masks: Tensor, boxes: Tensor, img_h: int, img_w: int, threshold: float) -> Tensor:
    return _k._mask_paste_impl(masks, boxes, (img_h, img_w), threshold)


@mask_paste.register_fake
def _(masks, boxes, img_h, img_w, threshold):
    return _empty(masks, (masks.shape[0], img_h, img_w), torch.bool)


OP_NAMES = ("lingual_similarity", "boxes_to_rois", "roi_align", "roi_align_backward", "iou_match", "predictor_linear", "predictor_wgrad",
            "predictor_gemm2", "similarity_transfer", "softmax_decode", "detect", "mask_paste")
