"""Tensor-level entry points of the RoI stage: every function enqueues hand-written sm_100a kernels from
libunit_b200.so on the current CUDA stream through the C ABI (include/unit_b200.h).

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); there is no CPU path and no fallback:
a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import TransferParams, check, lib

_F32 = torch.float32
NOVEL_TAG = 1000000


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    """All tensors on ONE CUDA device, and that device must be the current one: the library launches on the current
    device's current stream (one process per GPU), so a tensor of another device would silently run elsewhere."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("unit_b200 ops run on CUDA tensors only (there is no CPU path)")
        if dev is not None and t.device != dev:
            raise RuntimeError(f"unit_b200 op called with tensors on {dev} and {t.device}")
        dev = t.device
    if dev is None:
        raise RuntimeError("unit_b200 op called without tensors")
    if dev.index is not None and dev.index != torch.cuda.current_device():
        raise RuntimeError(f"unit_b200 op called with tensors on {dev} while the current device is "
                           f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device({dev.index})")
    return dev


def _c(t: torch.Tensor, dtype=None) -> torch.Tensor:
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


def _is_fake(t: torch.Tensor) -> bool:
    from torch._subclasses.fake_tensor import FakeTensor
    return isinstance(t, FakeTensor)


_WS = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


_I32_CACHE = {}


def _i32(values: Sequence[int], dev: torch.device) -> torch.Tensor:
    """Small read-only int32 device array (offsets, class index sets).  Cached by value: repeated shapes cost no
    host->device copy, and none is issued while a CUDA graph is being captured after a warm-up pass."""
    key = (tuple(int(v) for v in values), str(dev))
    t = _I32_CACHE.get(key)
    if t is None:
        if len(_I32_CACHE) >= 512:
            _I32_CACHE.clear()
        t = torch.tensor(list(key[0]), dtype=torch.int32, device=dev)
        if not _is_fake(t):  # tracing under FakeTensorMode must not poison the cache of real constants
            _I32_CACHE[key] = t
    return t


_F32_CACHE = {}


def f32_const(values: Sequence[Sequence[float]], dev: torch.device) -> torch.Tensor:
    """Small read-only fp32 device array (image sizes ...), cached by value like ``_i32``."""
    key = (tuple(tuple(float(x) for x in row) for row in values), str(dev))
    t = _F32_CACHE.get(key)
    if t is None:
        if len(_F32_CACHE) >= 512:
            _F32_CACHE.clear()
        t = torch.tensor([list(r) for r in key[0]], dtype=torch.float32, device=dev)
        if not _is_fake(t):
            _F32_CACHE[key] = t
    return t


_ONES_CACHE = {}


def _ones(n: int, dev: torch.device) -> torch.Tensor:
    key = (int(n), str(dev))
    t = _ONES_CACHE.get(key)
    if t is None:
        if len(_ONES_CACHE) >= 64:
            _ONES_CACHE.clear()
        t = torch.ones(int(n), dtype=torch.float32, device=dev)
        if not _is_fake(t):
            _ONES_CACHE[key] = t
    return t


def offsets_from_counts(counts: Sequence[int], dev: torch.device) -> torch.Tensor:
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    return _i32(off, dev)


# ------------------------------------------------------------------------------------------------- ROIAlign
def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.UNIT_F32
    if t.dtype == torch.bfloat16:
        return _lib.UNIT_BF16
    raise RuntimeError(f"roi_align: unsupported dtype {t.dtype} (f32 and bf16 only)")


def _boxes_to_rois_impl(boxes: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """[D2] convert_boxes_to_pooler_format in one launch: [R,4] boxes in image order + int32 offsets -> [R,5] rois."""
    dev = _need_cuda(boxes, offsets)
    boxes = _c(boxes, _F32)
    R = boxes.shape[0]
    rois = torch.empty((R, 5), dtype=_F32, device=dev)
    check(lib().unit_boxes_to_rois(_ptr(boxes), _ptr(offsets), offsets.numel() - 1, R, _ptr(rois), _stream()),
          "unit_boxes_to_rois")
    return rois


def _roi_align_forward_impl(feat: torch.Tensor, rois: torch.Tensor, output_size: Tuple[int, int], spatial_scale: float,
                      sampling_ratio: int, aligned: bool, rois_sorted: bool) -> torch.Tensor:
    dev = _need_cuda(feat, rois)
    feat = _c(feat)
    rois = _c(rois, _F32)
    N, C, H, W = feat.shape
    R = rois.shape[0]
    ph, pw = output_size
    out = torch.empty((R, C, ph, pw), dtype=feat.dtype, device=dev)
    if R == 0:
        return out
    ws_bytes = lib().unit_roi_align_workspace_bytes(N, C, H, W, R, _lib.UNIT_F32)
    ws = _workspace(dev, ws_bytes)
    check(lib().unit_roi_align_fwd(_ptr(feat), _ptr(rois), _ptr(out), N, C, H, W, R, ph, pw, float(spatial_scale),
                                   int(sampling_ratio), int(bool(aligned)), _dtype_code(feat), int(bool(rois_sorted)),
                                   _ptr(ws), ws.numel(), _stream()), "unit_roi_align_fwd")
    return out


def _roi_align_backward_impl(grad_out: torch.Tensor, rois: torch.Tensor, input_shape: Sequence[int], spatial_scale: float,
                       sampling_ratio: int, aligned: bool, rois_sorted: bool) -> torch.Tensor:
    dev = _need_cuda(grad_out, rois)
    grad_out = _c(grad_out)
    rois = _c(rois, _F32)
    N, C, H, W = [int(v) for v in input_shape]
    R, _, ph, pw = grad_out.shape
    grad_in = torch.empty((N, C, H, W), dtype=grad_out.dtype, device=dev)
    ws_bytes = lib().unit_roi_align_workspace_bytes(N, C, H, W, R, _dtype_code(grad_out))
    ws = _workspace(dev, ws_bytes)
    check(lib().unit_roi_align_bwd(_ptr(grad_out), _ptr(rois), _ptr(grad_in), N, C, H, W, R, ph, pw,
                                   float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                                   _dtype_code(grad_out), int(bool(rois_sorted)), _ptr(ws), ws.numel(), _stream()),
          "unit_roi_align_bwd")
    return grad_in


def boxes_to_rois(boxes: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """[D2] convert_boxes_to_pooler_format in one launch (torch.ops.unit_b200.boxes_to_rois)."""
    return torch.ops.unit_b200.boxes_to_rois(boxes, offsets)


def roi_align_forward(feat: torch.Tensor, rois: torch.Tensor, output_size: Tuple[int, int], spatial_scale: float,
                      sampling_ratio: int, aligned: bool, rois_sorted: bool) -> torch.Tensor:
    return torch.ops.unit_b200.roi_align(feat, rois, int(output_size[0]), int(output_size[1]), float(spatial_scale),
                                         int(sampling_ratio), bool(aligned), bool(rois_sorted))


def roi_align_backward(grad_out: torch.Tensor, rois: torch.Tensor, input_shape: Sequence[int], spatial_scale: float,
                       sampling_ratio: int, aligned: bool, rois_sorted: bool) -> torch.Tensor:
    n, c, h, w = [int(v) for v in input_shape]
    return torch.ops.unit_b200.roi_align_backward(grad_out, rois, n, c, h, w, float(spatial_scale), int(sampling_ratio),
                                                  bool(aligned), bool(rois_sorted))


def roi_align(feat: torch.Tensor, rois: torch.Tensor, output_size, spatial_scale: float = 1.0,
              sampling_ratio: int = 0, aligned: bool = True, rois_sorted: bool = False) -> torch.Tensor:
    """[TV] torchvision.ops.roi_align / [D2] ROIAlign drop-in (rois: [R,5] batch_idx,x1,y1,x2,y2); differentiable
    w.r.t. ``feat`` (torch.ops.unit_b200.roi_align, backward = torch.ops.unit_b200.roi_align_backward)."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    return roi_align_forward(feat, rois, tuple(output_size), spatial_scale, sampling_ratio, aligned, rois_sorted)


# ------------------------------------------------------------------------------------------------- IoU / matcher
def pairwise_iou(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """[D2] pairwise_iou on raw [G,4] / [P,4] tensors -> [G,P]."""
    dev = _need_cuda(boxes1, boxes2)
    b1, b2 = _c(boxes1, _F32), _c(boxes2, _F32)
    G, Pn = b1.shape[0], b2.shape[0]
    out = torch.empty((G, Pn), dtype=_F32, device=dev)
    if G and Pn:
        check(lib().unit_pairwise_iou(_ptr(b1), _ptr(b2), _ptr(out), G, Pn, _stream()), "unit_pairwise_iou")
    return out


def _thr_arrays(thresholds: Sequence[float], labels: Sequence[int]):
    T = len(thresholds)
    thr = (ctypes.c_float * T)(*[float(t) for t in thresholds])
    lab = (ctypes.c_int * (T + 1))(*[int(l) for l in labels])
    return thr, lab, T


def matcher(iou: torch.Tensor, thresholds: Sequence[float], labels: Sequence[int],
            allow_low_quality_matches: bool = False, want_vals: bool = True):
    """modeling/matcher.py Matcher.__call__ on a [G,P] quality matrix -> (matches i64, labels i8[, vals f32])."""
    dev = _need_cuda(iou)
    iou = _c(iou, _F32)
    G, Pn = iou.shape
    matches = torch.empty((Pn,), dtype=torch.int64, device=dev)
    mlabels = torch.empty((Pn,), dtype=torch.int8, device=dev)
    vals = torch.empty((Pn,), dtype=_F32, device=dev) if want_vals else None
    thr, lab, T = _thr_arrays(thresholds, labels)
    ws = _workspace(dev, max(G, 1) * 4)
    check(lib().unit_matcher(_ptr(iou), G, Pn, thr, lab, T, int(bool(allow_low_quality_matches)), _ptr(matches),
                             _ptr(mlabels), _ptr(vals), _ptr(ws), ws.numel(), _stream()), "unit_matcher")
    return (matches, mlabels, vals) if want_vals else (matches, mlabels)


def _iou_match_impl(gt_boxes: torch.Tensor, gt_offsets: torch.Tensor, prop_boxes: torch.Tensor, prop_offsets: torch.Tensor,
              thresholds: Sequence[float], labels: Sequence[int], want_vals: bool = True):
    """Fused pairwise_iou + Matcher for all images (offsets: int32 [n_img+1] on device)."""
    dev = _need_cuda(prop_boxes, prop_offsets, gt_offsets)
    gt_boxes, prop_boxes = _c(gt_boxes, _F32), _c(prop_boxes, _F32)
    n_img = prop_offsets.numel() - 1
    Pt = prop_boxes.shape[0]
    matches = torch.empty((Pt,), dtype=torch.int64, device=dev)
    mlabels = torch.empty((Pt,), dtype=torch.int8, device=dev)
    vals = torch.empty((Pt,), dtype=_F32, device=dev) if want_vals else None
    thr, lab, T = _thr_arrays(thresholds, labels)
    check(lib().unit_iou_match(_ptr(gt_boxes), _ptr(gt_offsets), _ptr(prop_boxes), _ptr(prop_offsets), n_img, Pt, thr,
                               lab, T, _ptr(matches), _ptr(mlabels), _ptr(vals), _stream()), "unit_iou_match")
    return matches, mlabels, vals


def iou_match(gt_boxes: torch.Tensor, gt_offsets: torch.Tensor, prop_boxes: torch.Tensor, prop_offsets: torch.Tensor,
              thresholds: Sequence[float], labels: Sequence[int], want_vals: bool = True):
    """Fused pairwise_iou + Matcher for all images (torch.ops.unit_b200.iou_match)."""
    m, l, v = torch.ops.unit_b200.iou_match(gt_boxes, gt_offsets, prop_boxes, prop_offsets,
                                            [float(t) for t in thresholds], [int(x) for x in labels])
    return m, l, (v if want_vals else None)


def label_proposals(matches, mlabels, gt_classes, gt_offsets, prop_offsets, num_classes: int):
    """-> prop_classes i64 [P], pos_idx i64 [P], neg_idx i64 [P] (per-image compacted), counts i32 [n_img,2]."""
    dev = _need_cuda(matches)
    n_img = prop_offsets.numel() - 1
    Pt = matches.numel()
    prop_classes = torch.empty((Pt,), dtype=torch.int64, device=dev)
    pos_idx = torch.empty((Pt,), dtype=torch.int64, device=dev)
    neg_idx = torch.empty((Pt,), dtype=torch.int64, device=dev)
    counts = torch.zeros((n_img, 2), dtype=torch.int32, device=dev)
    gt_classes = _c(gt_classes, torch.int64)
    check(lib().unit_label_proposals(_ptr(matches), _ptr(mlabels), _ptr(gt_classes), _ptr(gt_offsets),
                                     _ptr(prop_offsets), n_img, Pt, int(num_classes), _ptr(prop_classes),
                                     _ptr(pos_idx), _ptr(neg_idx), _ptr(counts), _stream()), "unit_label_proposals")
    return prop_classes, pos_idx, neg_idx, counts


def append_gt(prop_boxes: Sequence[torch.Tensor], prop_logits: Sequence[torch.Tensor],
              gt_boxes: Sequence[torch.Tensor], gt_logit: float):
    """[D2] add_ground_truth_to_proposals for all images in one launch: -> (boxes [T,4], logits [T]) with image i's
    rows = its proposals followed by its GT boxes, images back to back (T = sum(P_i + G_i)).  fp32 CUDA inputs."""
    dev = _need_cuda(*prop_boxes, *prop_logits, *gt_boxes)
    n = len(prop_boxes)
    pb = [_c(t, _F32) for t in prop_boxes]
    pl = [_c(t, _F32) for t in prop_logits]
    gb = [_c(t, _F32) for t in gt_boxes]
    pc = [int(t.shape[0]) for t in pb]
    gc = [int(t.shape[0]) for t in gb]
    total = sum(pc) + sum(gc)
    out_boxes = torch.empty((total, 4), dtype=_F32, device=dev)
    out_logits = torch.empty((total,), dtype=_F32, device=dev)
    if n and total:
        vp = ctypes.c_void_p * n
        ci = ctypes.c_int * n
        check(lib().unit_append_gt(vp(*[t.data_ptr() for t in pb]), vp(*[t.data_ptr() for t in pl]),
                                   vp(*[t.data_ptr() for t in gb]), ci(*pc), ci(*gc), n, float(gt_logit),
                                   _ptr(out_boxes), _ptr(out_logits), _stream()), "unit_append_gt")
    return out_boxes, out_logits


def sample_gather(pos_idx, neg_idx, perm_pos, perm_pos_off, perm_neg, perm_neg_off, pos_sel_off, neg_sel_off,
                  prop_offsets, gt_offsets, S_total: int, prop_boxes, prop_classes, matches, gt_boxes,
                  prop_field: Optional[torch.Tensor] = None):
    """-> (sampled_idx, boxes, classes, matched, gt_boxes[, field]): ``prop_field`` is one fp32 pass-through field of
    the proposals (objectness_logits), concatenated like prop_boxes and gathered by the same launch."""
    dev = _need_cuda(pos_idx)
    n_img = prop_offsets.numel() - 1
    sampled = torch.empty((S_total,), dtype=torch.int64, device=dev)
    out_boxes = torch.empty((S_total, 4), dtype=_F32, device=dev)
    out_classes = torch.empty((S_total,), dtype=torch.int64, device=dev)
    out_matched = torch.empty((S_total,), dtype=torch.int64, device=dev)
    out_gt = torch.empty((S_total, 4), dtype=_F32, device=dev)
    out_field = None
    if prop_field is not None:
        prop_field = _c(prop_field, _F32)
        out_field = torch.empty((S_total,), dtype=_F32, device=dev)
    check(lib().unit_sample_gather(_ptr(pos_idx), _ptr(neg_idx), _ptr(perm_pos), _ptr(perm_pos_off), _ptr(perm_neg),
                                   _ptr(perm_neg_off), _ptr(pos_sel_off), _ptr(neg_sel_off), _ptr(prop_offsets),
                                   _ptr(gt_offsets), n_img, S_total, _ptr(_c(prop_boxes, _F32)), _ptr(prop_classes),
                                   _ptr(matches), _ptr(_c(gt_boxes, _F32)), _ptr(sampled), _ptr(out_boxes),
                                   _ptr(out_classes), _ptr(out_matched), _ptr(out_gt), _ptr(prop_field),
                                   _ptr(out_field), _stream()),
          "unit_sample_gather")
    if prop_field is not None:
        return sampled, out_boxes, out_classes, out_matched, out_gt, out_field
    return sampled, out_boxes, out_classes, out_matched, out_gt


# ------------------------------------------------------------------------------------------------- decode / NMS
SCALE_CLAMP = math.log(1000.0 / 16)


def _softmax_decode_impl(scores: Optional[torch.Tensor], deltas: Optional[torch.Tensor], proposals: Optional[torch.Tensor],
                   weights=(10.0, 10.0, 5.0, 5.0), scale_clamp: float = SCALE_CLAMP, want_probs=True, want_boxes=True):
    dev = _need_cuda(scores, deltas)
    R = (scores if scores is not None else deltas).shape[0]
    probs = boxes = None
    K1 = scores.shape[1] if scores is not None else 1
    KB = deltas.shape[1] // 4 if deltas is not None else 0
    if want_probs:
        scores = _c(scores, _F32)
        probs = torch.empty((R, K1), dtype=_F32, device=dev)
    if want_boxes:
        deltas = _c(deltas, _F32)
        proposals = _c(proposals, _F32)
        boxes = torch.empty((R, 4 * KB), dtype=_F32, device=dev)
    if R:
        check(lib().unit_softmax_decode(_ptr(scores) if want_probs else None, _ptr(deltas) if want_boxes else None,
                                        _ptr(proposals) if want_boxes else None, _ptr(probs), _ptr(boxes), R, K1, KB,
                                        float(weights[0]), float(weights[1]), float(weights[2]), float(weights[3]),
                                        float(scale_clamp), _stream()), "unit_softmax_decode")
    return probs, boxes


def softmax_decode(scores: Optional[torch.Tensor], deltas: Optional[torch.Tensor], proposals: Optional[torch.Tensor],
                   weights=(10.0, 10.0, 5.0, 5.0), scale_clamp: float = SCALE_CLAMP, want_probs=True, want_boxes=True):
    """softmax and / or Box2BoxTransform.apply_deltas in one launch (torch.ops.unit_b200.softmax_decode)."""
    probs, boxes = torch.ops.unit_b200.softmax_decode(scores if want_probs else None, deltas if want_boxes else None,
                                                      proposals if want_boxes else None, [float(w) for w in weights],
                                                      float(scale_clamp))
    return (probs if want_probs else None), (boxes if want_boxes else None)


def box_get_deltas(src: torch.Tensor, tgt: torch.Tensor, weights=(10.0, 10.0, 5.0, 5.0)) -> torch.Tensor:
    dev = _need_cuda(src, tgt)
    src, tgt = _c(src, _F32), _c(tgt, _F32)
    R = src.shape[0]
    out = torch.empty((R, 4), dtype=_F32, device=dev)
    if R:
        check(lib().unit_box_get_deltas(_ptr(src), _ptr(tgt), _ptr(out), R, float(weights[0]), float(weights[1]),
                                        float(weights[2]), float(weights[3]), _stream()), "unit_box_get_deltas")
    return out


class _FastRCNNLossFn(torch.autograd.Function):
    """[D2] FastRCNNOutputs.losses fused with its gradient: (loss_cls, loss_box_reg)."""

    @staticmethod
    def forward(ctx, scores, deltas, proposals, gt_boxes, gt_classes, weights, beta):
        dev = _need_cuda(scores, deltas)
        scores, deltas = _c(scores, _F32), _c(deltas, _F32)
        R, K1 = scores.shape
        losses = torch.empty((2,), dtype=_F32, device=dev)
        d_scores = torch.empty_like(scores)
        d_deltas = torch.empty_like(deltas)
        ws = _workspace(dev, max(R, 1) * 8)
        check(lib().unit_fastrcnn_loss(_ptr(scores), _ptr(deltas), _ptr(_c(proposals, _F32)), _ptr(_c(gt_boxes, _F32)),
                                       _ptr(_c(gt_classes, torch.int64)), R, K1 - 1, float(weights[0]),
                                       float(weights[1]), float(weights[2]), float(weights[3]), float(beta),
                                       _ptr(losses), _ptr(d_scores), _ptr(d_deltas), _ptr(ws), ws.numel(), _stream()),
              "unit_fastrcnn_loss")
        ctx.save_for_backward(d_scores, d_deltas)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_cls, g_box):
        d_scores, d_deltas = ctx.saved_tensors
        return d_scores * g_cls, d_deltas * g_box, None, None, None, None, None


def fastrcnn_loss(scores, deltas, proposals, gt_boxes, gt_classes, weights=(10.0, 10.0, 5.0, 5.0), beta=0.0):
    return _FastRCNNLossFn.apply(scores, deltas, proposals, gt_boxes, gt_classes, tuple(weights), float(beta))


class _RowLossesFn(torch.autograd.Function):
    """Per-RoI Fast R-CNN losses (FastRCNNOutputsReduction / NLL / Regression) with their gradients from one launch."""

    @staticmethod
    def forward(ctx, scores, deltas, proposals, gt_boxes, gt_classes, weights, beta, nll):
        dev = _need_cuda(scores)
        scores = _c(scores, _F32)
        R, K1 = scores.shape
        K = K1 - 1
        gt_classes = _c(gt_classes, torch.int64)
        row_ce = torch.empty((R,), dtype=_F32, device=dev)
        d_scores = torch.empty_like(scores)
        has_box = deltas is not None
        row_box = torch.zeros((R, 4), dtype=_F32, device=dev)
        d_box = torch.zeros((R, 4), dtype=_F32, device=dev)
        if has_box:
            deltas = _c(deltas, _F32)
        check(lib().unit_fastrcnn_row_losses(_ptr(scores), _ptr(deltas) if has_box else None,
                                             _ptr(_c(proposals, _F32)) if has_box else None,
                                             _ptr(_c(gt_boxes, _F32)) if has_box else None, _ptr(gt_classes), R, K,
                                             float(weights[0]), float(weights[1]), float(weights[2]), float(weights[3]),
                                             float(beta), int(bool(nll)), _ptr(row_ce), _ptr(row_box), _ptr(d_scores),
                                             _ptr(d_box), _stream()), "unit_fastrcnn_row_losses")
        ctx.save_for_backward(d_scores, d_box, gt_classes)
        ctx.K, ctx.has_box = K, has_box
        return row_ce, row_box

    @staticmethod
    def backward(ctx, g_ce, g_box):
        d_scores, d_box, gt_classes = ctx.saved_tensors
        g_scores = d_scores * g_ce[:, None]
        g_deltas = None
        if ctx.has_box and ctx.needs_input_grad[1]:
            K = ctx.K
            g_deltas = torch.zeros((d_scores.shape[0], 4 * K), dtype=_F32, device=d_scores.device)
            fg = ((gt_classes >= 0) & (gt_classes < K)).nonzero().squeeze(1)
            cols = 4 * gt_classes[fg][:, None] + torch.arange(4, device=fg.device)
            g_deltas[fg[:, None], cols] = (d_box * g_box)[fg]
        return g_scores, g_deltas, None, None, None, None, None, None


def fastrcnn_row_losses(scores, deltas, proposals, gt_boxes, gt_classes, weights=(10.0, 10.0, 5.0, 5.0), beta=0.0,
                        nll=False):
    """-> (row_ce [R], row_box [R,4]) differentiable w.r.t. scores / deltas."""
    return _RowLossesFn.apply(scores, deltas, proposals, gt_boxes, gt_classes, tuple(weights), float(beta), bool(nll))


# ---------------------------------------------------------------------------------- weak-image training losses
class _MILLossFn(torch.autograd.Function):
    """weak_detector_fast_rcnn.py:189-214 fused with its gradient: (loss_im_cls, mil_scores, class_vector)."""

    @staticmethod
    def forward(ctx, cls_logits, det_logits, img_offsets, gt_vector, multiplier, max_rows):
        dev = _need_cuda(cls_logits, det_logits, img_offsets, gt_vector)
        cls_logits, det_logits = _c(cls_logits, _F32), _c(det_logits, _F32)
        gt_vector = _c(gt_vector, _F32)
        R, K = cls_logits.shape
        n_img = img_offsets.numel() - 1
        mil = torch.empty((R, K), dtype=_F32, device=dev)
        class_vec = torch.empty((n_img, K), dtype=_F32, device=dev)
        loss = torch.empty((1,), dtype=_F32, device=dev)
        d_cls, d_det = torch.empty_like(mil), torch.empty_like(mil)
        max_rows = R if max_rows is None else min(int(max_rows), R)
        ws = _workspace(dev, lib().unit_mil_loss_workspace_bytes(n_img, max_rows, K))
        check(lib().unit_mil_loss(_ptr(cls_logits), _ptr(det_logits), _ptr(img_offsets), _ptr(gt_vector), n_img, R,
                                  max_rows, K, float(multiplier), _ptr(mil), _ptr(class_vec), _ptr(loss), _ptr(d_cls),
                                  _ptr(d_det), _ptr(ws), ws.numel(), _stream()), "unit_mil_loss")
        ctx.save_for_backward(d_cls, d_det)
        ctx.mark_non_differentiable(mil, class_vec)
        return loss[0], mil, class_vec

    @staticmethod
    def backward(ctx, g_loss, _g_mil, _g_vec):
        d_cls, d_det = ctx.saved_tensors
        return d_cls * g_loss, d_det * g_loss, None, None, None, None


def mil_loss(cls_logits, det_logits, img_offsets, gt_vector, multiplier: float = 1.0,
             max_rows: Optional[int] = None):
    """-> (loss_im_cls scalar, mil_scores [R,K] detached, class_vector [n_img,K] detached).  ``max_rows``: the largest
    per-image proposal count when the caller knows it on the host (sizes the grid; default: all rows)."""
    return _MILLossFn.apply(cls_logits, det_logits, img_offsets, gt_vector, float(multiplier), max_rows)


def oicr_targets(probs: torch.Tensor, prop_boxes: torch.Tensor, prop_offsets: torch.Tensor, gt_vector: torch.Tensor,
                 thresholds: Sequence[float], labels: Sequence[int], bg_threshold: float):
    """compute_loss_inputs (weak_detector_fast_rcnn.py:384-408) for every image in one launch.
    -> labels i64 [P], cls_weights f32 [P], pgt_index i64 [n_img,K] (-1 = class absent), pgt_scores [n_img,K]."""
    dev = _need_cuda(probs, prop_boxes, prop_offsets, gt_vector)
    probs, prop_boxes, gt_vector = _c(probs, _F32), _c(prop_boxes, _F32), _c(gt_vector, _F32)
    n_img, K = gt_vector.shape
    Pt, ld = probs.shape
    out_labels = torch.empty((Pt,), dtype=torch.int64, device=dev)
    weights = torch.empty((Pt,), dtype=_F32, device=dev)
    pgt_index = torch.empty((n_img, K), dtype=torch.int64, device=dev)
    pgt_scores = torch.empty((n_img, K), dtype=_F32, device=dev)
    thr, lab, T = _thr_arrays(thresholds, labels)
    check(lib().unit_oicr_targets(_ptr(probs), ld, _ptr(prop_boxes), _ptr(prop_offsets), _ptr(gt_vector), n_img, Pt, K,
                                  thr, lab, T, float(bg_threshold), _ptr(out_labels), _ptr(weights), _ptr(pgt_index),
                                  _ptr(pgt_scores), _stream()), "unit_oicr_targets")
    return out_labels, weights, pgt_index, pgt_scores


class _WeightedCEFn(torch.autograd.Function):
    """weighted_softmax_with_loss (weak_detector_fast_rcnn.py:220-227) fused with its gradient."""

    @staticmethod
    def forward(ctx, scores, labels, weights):
        dev = _need_cuda(scores, labels, weights)
        scores = _c(scores, _F32)
        R, K1 = scores.shape
        loss = torch.empty((1,), dtype=_F32, device=dev)
        d_scores = torch.empty_like(scores)
        ws = _workspace(dev, max(R, 1) * 4)
        check(lib().unit_weighted_ce_loss(_ptr(scores), _ptr(_c(labels, torch.int64)), _ptr(_c(weights, _F32)), R, K1,
                                          _ptr(loss), _ptr(d_scores), _ptr(ws), ws.numel(), _stream()),
              "unit_weighted_ce_loss")
        ctx.save_for_backward(d_scores)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_scores,) = ctx.saved_tensors
        return d_scores * g, None, None


def weighted_ce_loss(scores, labels, weights):
    return _WeightedCEFn.apply(scores, labels, weights)


NMS_CLASSWISE, NMS_COORD_TRICK, NMS_TV_CUDA_RULE, NMS_TV_CPU_RULE = 0, 1, 2, 3


def _detect_impl(boxes: torch.Tensor, probs: torch.Tensor, roi_offsets: torch.Tensor, image_hw: torch.Tensor,
           score_thresh: float, nms_thresh: float, topk: int, nms_mode: int = NMS_TV_CUDA_RULE):
    """fast_rcnn_inference for a batch: returns det_boxes [n,topk,4], det_scores [n,topk], det_classes i64,
    det_roi i64 (index into the image's finite rows), det_counts i32 [n] -- all on device, no sync."""
    dev = _need_cuda(boxes, probs)
    boxes, probs = _c(boxes, _F32), _c(probs, _F32)
    n_img = roi_offsets.numel() - 1
    R, K1 = probs.shape
    K = K1 - 1
    KB = boxes.shape[1] // 4
    cap = max(R * K, 1)
    cand_boxes = torch.empty((cap, 4), dtype=_F32, device=dev)
    cand_scores = torch.empty((cap,), dtype=_F32, device=dev)
    cand_roi = torch.empty((cap,), dtype=torch.int32, device=dev)
    cand_cls = torch.empty((cap,), dtype=torch.int32, device=dev)
    cand_counts = torch.empty((max(n_img, 1),), dtype=torch.int32, device=dev)
    ws_bytes = lib().unit_nms_workspace_bytes(n_img, R * K)
    ws = _workspace(dev, max(ws_bytes, R * 8 + 256))
    check(lib().unit_detect_filter(_ptr(boxes), _ptr(probs), _ptr(roi_offsets), _ptr(image_hw), n_img, R, K, KB,
                                   float(score_thresh), _ptr(cand_boxes), _ptr(cand_scores), _ptr(cand_roi),
                                   _ptr(cand_cls), _ptr(cand_counts), _ptr(ws), ws.numel(), _stream()),
          "unit_detect_filter")
    topk = int(topk) if topk >= 0 else cap
    det_boxes = torch.empty((n_img, topk, 4), dtype=_F32, device=dev)
    det_scores = torch.empty((n_img, topk), dtype=_F32, device=dev)
    det_classes = torch.empty((n_img, topk), dtype=torch.int64, device=dev)
    det_roi = torch.empty((n_img, topk), dtype=torch.int64, device=dev)
    det_counts = torch.empty((max(n_img, 1),), dtype=torch.int32, device=dev)
    check(lib().unit_detect_nms(_ptr(cand_boxes), _ptr(cand_scores), _ptr(cand_roi), _ptr(cand_cls),
                                _ptr(cand_counts), _ptr(roi_offsets), n_img, R, K, float(nms_thresh), int(nms_mode),
                                topk, _ptr(det_boxes), _ptr(det_scores), _ptr(det_classes), _ptr(det_roi),
                                _ptr(det_counts), _ptr(ws), ws.numel(), _stream()), "unit_detect_nms")
    return det_boxes, det_scores, det_classes, det_roi, det_counts, (cand_boxes, cand_scores, cand_roi, cand_cls,
                                                                      cand_counts)


def detect(boxes: torch.Tensor, probs: torch.Tensor, roi_offsets: torch.Tensor, image_hw: torch.Tensor,
           score_thresh: float, nms_thresh: float, topk: int, nms_mode: int = NMS_TV_CUDA_RULE):
    """fast_rcnn_inference for a batch (torch.ops.unit_b200.detect): det_boxes [n,topk,4], det_scores, det_classes,
    det_roi, det_counts, (candidates) -- all on device, no sync."""
    o = torch.ops.unit_b200.detect(boxes, probs, roi_offsets, image_hw, float(score_thresh), float(nms_thresh),
                                   int(topk), int(nms_mode))
    return o[0], o[1], o[2], o[3], o[4], tuple(o[5:])


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: Optional[torch.Tensor], iou_threshold: float,
                nms_mode: int = NMS_TV_CUDA_RULE, max_keep: int = -1) -> torch.Tensor:
    """[TV] torchvision.ops.batched_nms (idxs given) / nms (idxs None) drop-in: int64 indices, score-descending."""
    dev = _need_cuda(boxes, scores)
    boxes, scores = _c(boxes.float()), _c(scores, _F32)
    N = boxes.shape[0]
    if N == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    if idxs is not None:
        idxs = _c(idxs, torch.int64)
    keep = torch.empty((N,), dtype=torch.int64, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    ws_bytes = lib().unit_nms_workspace_bytes(1, N)
    ws = _workspace(dev, ws_bytes)
    check(lib().unit_batched_nms(_ptr(boxes), _ptr(scores), _ptr(idxs), N, float(iou_threshold), int(nms_mode),
                                 int(max_keep), _ptr(keep), _ptr(count), _ptr(ws), ws.numel(), _stream()),
          "unit_batched_nms")
    return keep[: int(count.item())]  # same host sync the reference's nms performs


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return batched_nms(boxes, scores, None, iou_threshold)


# ------------------------------------------------------------------------------------------------- transfer
def _lingual_similarity_impl(emb: torch.Tensor, indexer: torch.Tensor, base: torch.Tensor, novel: torch.Tensor):
    """fast_rcnn.py:376-382 -> (raw [Nn,B], softmax(raw, -1))."""
    dev = _need_cuda(emb)
    emb = _c(emb, _F32)
    indexer, base, novel = _c(indexer, torch.int64), _c(base, torch.int64), _c(novel, torch.int64)
    B, Nn = base.numel(), novel.numel()
    raw = torch.empty((Nn, B), dtype=_F32, device=dev)
    soft = torch.empty((Nn, B), dtype=_F32, device=dev)
    check(lib().unit_lingual_similarity(_ptr(emb), _ptr(indexer), _ptr(base), _ptr(novel), emb.shape[1], B, Nn,
                                        _ptr(raw), _ptr(soft), _stream()), "unit_lingual_similarity")
    return raw, soft


def lingual_similarity(emb: torch.Tensor, indexer: torch.Tensor, base: torch.Tensor, novel: torch.Tensor):
    """fast_rcnn.py:376-382 -> (raw [Nn,B], softmax(raw, -1))  (torch.ops.unit_b200.lingual_similarity)."""
    return torch.ops.unit_b200.lingual_similarity(emb, indexer, base, novel)


def make_class_kind(num_classes: int, base: Sequence[int], novel: Sequence[int], dev) -> torch.Tensor:
    kind = [-1] * num_classes
    for i, b in enumerate(base):
        kind[int(b)] = i
    for i, n in enumerate(novel):
        kind[int(n)] = NOVEL_TAG + i
    return _i32(kind, dev)


class TransferSpec:
    """Static (per model) part of the transfer: class index sets, class-level similarity terms, term weights."""

    def __init__(self, num_classes: int, base: Sequence[int], novel: Sequence[int], dev,
                 static: Optional[dict] = None, wv: Optional[dict] = None, norm: Optional[dict] = None,
                 vis_threshold: float = 0.0, static_per_roi: int = 0):
        self.K, self.B, self.Nn = int(num_classes), len(base), len(novel)
        self.base_i32 = _i32(base, dev)
        self.novel_i32 = _i32(novel, dev)
        self.class_kind = make_class_kind(num_classes, base, novel, dev)
        self.static = {k: (None if v is None else _c(v.to(dev), _F32)) for k, v in (static or {}).items()}
        self.wv = dict(wv or {})
        self.norm = dict(norm or {})
        self.vis_threshold = float(vis_threshold)
        self.static_per_roi = int(static_per_roi)
        self.wk = {}

    @classmethod
    def from_tensors(cls, num_classes: int, base_i32, novel_i32, class_kind, static: dict, wv: Sequence[float],
                     norm: Sequence[int], vis_threshold: float, static_per_roi: int) -> "TransferSpec":
        """Rebuild a spec from the flat arguments of torch.ops.unit_b200.similarity_transfer."""
        self = cls.__new__(cls)
        self.K, self.B, self.Nn = int(num_classes), int(base_i32.numel()), int(novel_i32.numel())
        self.base_i32, self.novel_i32, self.class_kind = base_i32, novel_i32, class_kind
        self.static = dict(static)
        self.wv = dict(zip(("cls", "bbox", "seg"), wv))
        self.norm = dict(zip(("cls", "bbox", "seg"), norm))
        self.vis_threshold, self.static_per_roi, self.wk = float(vis_threshold), int(static_per_roi), {}
        return self

    def with_static(self, static: dict, static_per_roi: int) -> "TransferSpec":
        """Same class sets and term weights with other static blocks (per-RoI terms: bit h of ``static_per_roi``)."""
        import copy
        other = copy.copy(self)
        other.static = {k: (None if v is None else _c(v, _F32)) for k, v in static.items()}
        other.static_per_roi = int(static_per_roi)
        return other

    def params(self, R: int, do_transfer: bool, novel_neg_inf: bool) -> TransferParams:
        g = lambda d, k, default: d.get(k, default)
        return TransferParams(R, self.K, self.B, self.Nn, self.vis_threshold,
                              float(g(self.wv, "cls", 0.0)), float(g(self.wv, "bbox", 0.0)),
                              float(g(self.wv, "seg", 0.0)),
                              int(g(self.norm, "cls", 0)), int(g(self.norm, "bbox", 0)), int(g(self.norm, "seg", 0)),
                              int(do_transfer), int(novel_neg_inf), self.static_per_roi)


def _similarity_transfer_forward_impl(spec: TransferSpec, vis_logits, delta_scores, proposal_deltas, weak_scores=None,
                                ft_scores=None, ft_deltas=None, do_transfer=True, novel_neg_inf=False,
                                want_similarity: Sequence[str] = ()):
    dev = _need_cuda(delta_scores, proposal_deltas)
    R = delta_scores.shape[0]

    def rows(t):  # column blocks of a packed GEMM output go in as (pointer, row stride): no copy
        if t is None:
            return None, 0
        if t.dtype == _F32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
            return t, t.stride(0)
        return _c(t, _F32), 0

    (delta_scores, ld_d), (proposal_deltas, ld_p) = rows(delta_scores), rows(proposal_deltas)
    (ft_scores, ld_fs), (ft_deltas, ld_fd) = rows(ft_scores), rows(ft_deltas)
    (vis_logits, ld_v), (weak_scores, ld_w) = rows(vis_logits), rows(weak_scores)
    out_scores = torch.empty(tuple(delta_scores.shape), dtype=_F32, device=dev)
    out_bbox = torch.empty(tuple(proposal_deltas.shape), dtype=_F32, device=dev)
    sims = {h: (torch.empty((R, spec.Nn, spec.B), dtype=_F32, device=dev) if h in want_similarity else None)
            for h in ("cls", "bbox", "seg")}
    if R:
        p = spec.params(R, do_transfer, novel_neg_inf)
        p.ld_delta_scores, p.ld_proposal_deltas, p.ld_ft_scores, p.ld_ft_deltas = ld_d, ld_p, ld_fs, ld_fd
        p.ld_vis_logits, p.ld_weak_scores = ld_v, ld_w
        check(lib().unit_similarity_transfer(ctypes.byref(p), _ptr(vis_logits), _ptr(spec.static.get("cls")),
                                             _ptr(spec.static.get("bbox")), _ptr(spec.static.get("seg")),
                                             _ptr(spec.base_i32), _ptr(spec.novel_i32), _ptr(spec.class_kind),
                                             _ptr(delta_scores), _ptr(proposal_deltas), _ptr(weak_scores),
                                             _ptr(ft_scores), _ptr(ft_deltas), _ptr(out_scores), _ptr(out_bbox),
                                             _ptr(sims["cls"]), _ptr(sims["bbox"]), _ptr(sims["seg"]), _stream()),
              "unit_similarity_transfer")
    return out_scores, out_bbox, sims


def similarity_transfer_forward(spec: TransferSpec, vis_logits, delta_scores, proposal_deltas, weak_scores=None,
                                ft_scores=None, ft_deltas=None, do_transfer=True, novel_neg_inf=False,
                                want_similarity: Sequence[str] = ()):
    """Fused similarity + transfer (torch.ops.unit_b200.similarity_transfer) -> (scores, bbox, {head: S or None})."""
    g = lambda d, k: d.get(k, 0)
    want = sum(1 << i for i, h in enumerate(("cls", "bbox", "seg")) if h in want_similarity)
    s_, b_, sc, sb, ss = torch.ops.unit_b200.similarity_transfer(
        vis_logits, spec.static.get("cls"), spec.static.get("bbox"), spec.static.get("seg"), spec.base_i32,
        spec.novel_i32, spec.class_kind, delta_scores, proposal_deltas, weak_scores, ft_scores, ft_deltas,
        float(spec.vis_threshold), [float(g(spec.wv, h)) for h in ("cls", "bbox", "seg")],
        [int(g(spec.norm, h)) for h in ("cls", "bbox", "seg")], bool(do_transfer), bool(novel_neg_inf),
        int(spec.static_per_roi), want)
    pick = lambda t, i: t if (want >> i) & 1 else None
    return s_, b_, {"cls": pick(sc, 0), "bbox": pick(sb, 1), "seg": pick(ss, 2)}


def similarity_transfer_backward(spec: TransferSpec, s_cls, s_bbox, g_scores, g_bbox, detach_transfer=False):
    dev = _need_cuda(g_scores, g_bbox)
    R = g_scores.shape[0]
    g_scores, g_bbox = _c(g_scores, _F32), _c(g_bbox, _F32)
    g_delta = torch.empty_like(g_scores)
    g_pd = torch.empty_like(g_bbox)
    if R:
        p = spec.params(R, True, False)
        check(lib().unit_similarity_transfer_bwd(ctypes.byref(p), _ptr(s_cls), _ptr(s_bbox), _ptr(spec.base_i32),
                                                 _ptr(spec.novel_i32), _ptr(spec.class_kind), _ptr(g_scores),
                                                 _ptr(g_bbox), int(bool(detach_transfer)), _ptr(g_delta), _ptr(g_pd),
                                                 _stream()), "unit_similarity_transfer_bwd")
    return g_delta, g_pd


def similarity_transfer_backward_vis(spec: TransferSpec, vis_logits, delta_scores, proposal_deltas, g_scores, g_bbox):
    """d loss / d vis_logits of the fused similarity + transfer (recomputes S in the kernel; nothing is saved)."""
    dev = _need_cuda(vis_logits, g_scores, g_bbox)
    R = g_scores.shape[0]
    g_vis = torch.empty((R, spec.K + 1), dtype=_F32, device=dev)
    if R:
        def rows(t):
            if t.dtype == _F32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
                return t, t.stride(0)
            return _c(t, _F32), 0

        (delta_scores, ld_d), (proposal_deltas, ld_p) = rows(delta_scores), rows(proposal_deltas)
        vis_logits, ld_v = rows(vis_logits)
        p = spec.params(R, True, False)
        p.ld_delta_scores, p.ld_proposal_deltas, p.ld_vis_logits = ld_d, ld_p, ld_v
        check(lib().unit_similarity_transfer_bwd_vis(ctypes.byref(p), _ptr(vis_logits),
                                                     _ptr(spec.static.get("cls")), _ptr(spec.static.get("bbox")),
                                                     _ptr(spec.base_i32), _ptr(spec.novel_i32), _ptr(delta_scores),
                                                     _ptr(proposal_deltas), _ptr(_c(g_scores, _F32)),
                                                     _ptr(_c(g_bbox, _F32)), _ptr(g_vis), _stream()),
              "unit_similarity_transfer_bwd_vis")
    return g_vis


class _TransferFn(torch.autograd.Function):
    """Autograd of the fused transfer w.r.t. (vis_logits, delta_scores, proposal_deltas, ft_scores, ft_deltas).

    The gradient w.r.t. ``vis_logits`` is what the reference gets from running get_similarity_matrices
    (roi_heads.py:245-257) with autograd on: with a trainable box head the fine-tune loss reaches box_features
    through the visual similarity.  The class-level terms (lingual, TopK ...) are built from frozen tensors
    (embeddings; ``.clone().detach()`` weights, roi_heads.py:275,286,297) and carry no gradient in the reference either.
    """

    @staticmethod
    def forward(ctx, spec, vis_logits, delta_scores, proposal_deltas, weak_scores, ft_scores, ft_deltas, do_transfer,
                novel_neg_inf, detach_transfer, ft_packed=None):
        need_grad = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        ctx.need_delta_grad = need_grad
        # detach_transfer (fast_rcnn.py:566,575) cuts the transferred part out of the graph: no similarity gradient
        ctx.need_vis_grad = bool(ctx.needs_input_grad[1] and do_transfer and vis_logits is not None
                                 and not detach_transfer)
        ctx.ft_is_packed = ft_packed is not None
        if ft_packed is not None:  # [R, (K+1) + 4K]: both fine-tune blocks of one GEMM output, one gradient tensor
            K1 = delta_scores.shape[1]
            ft_scores, ft_deltas = ft_packed[:, :K1], ft_packed[:, K1:]
        want = ("cls", "bbox") if (need_grad and do_transfer and not detach_transfer) else ()
        out_scores, out_bbox, sims = similarity_transfer_forward(spec, vis_logits, delta_scores, proposal_deltas,
                                                                 weak_scores, ft_scores, ft_deltas, do_transfer,
                                                                 novel_neg_inf, want)
        ctx.spec, ctx.do_transfer, ctx.detach, ctx.neg_inf = spec, do_transfer, detach_transfer, novel_neg_inf
        saved = [t for t in (sims["cls"], sims["bbox"]) if t is not None]
        ctx.n_sims = len(saved)
        if ctx.need_vis_grad:
            saved += [vis_logits, delta_scores, proposal_deltas]
        ctx.save_for_backward(*saved)
        ctx.has_ft = (ft_scores is not None, ft_deltas is not None)
        return out_scores, out_bbox

    @staticmethod
    def backward(ctx, g_scores, g_bbox):
        saved = ctx.saved_tensors
        g_scores, g_bbox = g_scores.contiguous(), g_bbox.contiguous()
        if ctx.neg_inf:
            kind = ctx.spec.class_kind
            g_scores = g_scores.clone()
            g_scores[:, :-1][:, kind >= NOVEL_TAG] = 0
        if not ctx.need_delta_grad:  # frozen delta layers (fine-tuning): nothing flows through the transfer
            g_delta = g_pd = None
        elif ctx.do_transfer:
            s_cls, s_bbox = (saved[0], saved[1]) if ctx.n_sims == 2 else (None, None)
            g_delta, g_pd = similarity_transfer_backward(ctx.spec, s_cls, s_bbox, g_scores, g_bbox,
                                                         detach_transfer=ctx.detach or s_cls is None)
        else:
            g_delta, g_pd = g_scores, g_bbox
        g_vis = None
        if ctx.need_vis_grad:
            vis_logits, delta_scores, proposal_deltas = saved[ctx.n_sims:]
            g_vis = similarity_transfer_backward_vis(ctx.spec, vis_logits, delta_scores, proposal_deltas, g_scores,
                                                     g_bbox)
        if ctx.ft_is_packed:
            return (None, g_vis, g_delta, g_pd, None, None, None, None, None, None,
                    torch.cat([g_scores, g_bbox], 1))
        return (None, g_vis, g_delta, g_pd, None, g_scores if ctx.has_ft[0] else None,
                g_bbox if ctx.has_ft[1] else None, None, None, None, None)


def similarity_transfer(spec: TransferSpec, vis_logits, delta_scores, proposal_deltas, weak_scores=None,
                        ft_scores=None, ft_deltas=None, do_transfer=True, novel_neg_inf=False,
                        detach_transfer=False, ft_packed=None):
    if vis_logits is not None and vis_logits.requires_grad and do_transfer and not detach_transfer:
        if spec.static_per_roi or getattr(spec, "wk", None):
            raise NotImplementedError("similarity_transfer: gradients through a per-RoI static similarity block "
                                      "('VisualK-k' terms) are not implemented; only the 'visual' term is")
    return _TransferFn.apply(spec, vis_logits, delta_scores, proposal_deltas, weak_scores, ft_scores, ft_deltas,
                             bool(do_transfer), bool(novel_neg_inf), bool(detach_transfer), ft_packed)


# ------------------------------------------------------------------------------------------------- predictor GEMM
def _predictor_gemm_forward_impl(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """y = x @ w.T + bias on the tcgen05 tensor cores (TF32 multiply, fp32 accumulate in TMEM)."""
    dev = _need_cuda(x, w)
    x, w = _c(x, _F32), _c(w, _F32)
    b = None if bias is None else _c(bias, _F32)
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=_F32, device=dev)
    if M == 0:
        return y
    ws_bytes = lib().unit_predictor_gemm_workspace_bytes(M, N, K)
    ws = _workspace(dev, ws_bytes)
    check(lib().unit_predictor_gemm(_ptr(x), _ptr(w), _ptr(b), _ptr(y), M, N, K, _ptr(ws), ws.numel(), _stream()),
          "unit_predictor_gemm")
    return y


def predictor_gemm_forward(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """y = x @ w.T + bias on the tcgen05 tensor cores (torch.ops.unit_b200.predictor_linear)."""
    return torch.ops.unit_b200.predictor_linear(x, w, bias)


def linear_tf32(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable predictor Linear: forward and weight / bias gradients on the hand-written tcgen05 kernels
    (torch.ops.unit_b200.predictor_linear / predictor_wgrad); only d/dx -- unused on the fine-tune path -- is a
    library GEMM."""
    return torch.ops.unit_b200.predictor_linear(x, w, bias)


def _predictor_gemm2_impl(x1: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], x2: Optional[torch.Tensor] = None,
                    w2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None):
    """Two products sharing M and K in ONE tcgen05 launch (+ one reduce launch):
    y1 = x1 @ w1.T + b1 and y2 = x2 @ w2.T + b2.  Output rows are padded to a multiple of 32 columns (zeros), so
    column blocks can be handed on as (pointer, row stride) views.  Returns (y1, y2 or None)."""
    dev = _need_cuda(x1, w1)
    x1, w1 = _c(x1, _F32), _c(w1, _F32)
    M, K = x1.shape
    N1 = w1.shape[0]
    ld1 = (N1 + 31) // 32 * 32
    y1 = torch.empty((M, ld1), dtype=_F32, device=dev)
    N2, ld2, y2 = 0, 0, None
    if x2 is not None:
        x2, w2 = _c(x2, _F32), _c(w2, _F32)
        N2 = w2.shape[0]
        ld2 = (N2 + 31) // 32 * 32
        y2 = torch.empty((M, ld2), dtype=_F32, device=dev)
    if M == 0:
        return y1, y2
    cb = lambda t: None if t is None else _c(t, _F32)
    b1, b2 = cb(b1), cb(b2)
    ws = _workspace(dev, lib().unit_predictor_gemm2_workspace_bytes(M, N1, N2, K))
    check(lib().unit_predictor_gemm2(_ptr(x1), _ptr(w1), _ptr(b1), _ptr(y1), N1, ld1, _ptr(x2), _ptr(w2), _ptr(b2),
                                     _ptr(y2), N2, ld2, M, K, _ptr(ws), ws.numel(), _stream()), "unit_predictor_gemm2")
    return y1, y2


def predictor_gemm2(x1: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], x2: Optional[torch.Tensor] = None,
                    w2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None):
    """Two products sharing M and K in one launch (torch.ops.unit_b200.predictor_gemm2); see the implementation."""
    if x2 is None:
        return _predictor_gemm2_impl(x1, w1, b1)
    return torch.ops.unit_b200.predictor_gemm2(x1, w1, b1, x2, w2, b2)


def predictor_wgrad(gy: torch.Tensor, x: torch.Tensor, n_cols: int, seg_rows: Sequence[int],
                    w_dst: Sequence[Optional[torch.Tensor]], b_dst: Sequence[Optional[torch.Tensor]],
                    scales: Sequence[Optional[torch.Tensor]], accumulate: bool) -> None:
    """dW = gy[:, :n_cols].T @ x and db = gy[:, :n_cols].sum(0) on the tcgen05 tensor cores, written (or accumulated)
    straight into the destination buffers per row segment -- normally the parameters' ``.grad`` views of the flat
    all-reduce bucket.  ``gy`` is [R, ld] with ld a multiple of 128 and zeros past ``n_cols``; ``scales[s]`` is a
    device scalar multiplying segment s (the upstream gradient of its loss)."""
    dev = _need_cuda(gy, x)
    assert gy.dtype == _F32 and gy.dim() == 2 and gy.stride(1) == 1 and gy.stride(0) % 128 == 0
    x = _c(x, _F32)
    R, K = x.shape
    ld = gy.stride(0)
    ws = _workspace(dev, lib().unit_predictor_wgrad_workspace_bytes(R, K))
    seg_rows = [int(v) for v in seg_rows]
    for n0 in range(0, n_cols, 128):  # the kernel handles <= 128 gradient rows per launch (one UMMA M tile)
        n1 = min(n_cols, n0 + 128)
        rows, wd, bd, sc = [0], [], [], []
        for s in range(len(seg_rows) - 1):
            a, b = max(seg_rows[s], n0), min(seg_rows[s + 1], n1)
            if a >= b:
                continue
            rows.append(b - n0)
            off = a - seg_rows[s]
            wd.append(None if w_dst[s] is None else w_dst[s][off:])
            bd.append(None if b_dst[s] is None else b_dst[s][off:])
            sc.append(scales[s])
        nseg = len(rows) - 1
        c_rows = (ctypes.c_int * (nseg + 1))(*rows)
        vp = lambda ts: (ctypes.c_void_p * nseg)(*[None if t is None else t.data_ptr() for t in ts])
        g = gy[:, n0:]
        check(lib().unit_predictor_wgrad(ctypes.c_void_p(g.data_ptr()), ld, _ptr(x), R, n1 - n0, K, nseg, c_rows, vp(wd),
                                         vp(bd), vp(sc), int(bool(accumulate)), _ptr(ws), ws.numel(), _stream()),
              "unit_predictor_wgrad")


def fastrcnn_loss_packed(scores, deltas, proposals, gt_boxes, gt_classes, weights=(10.0, 10.0, 5.0, 5.0), beta=0.0):
    """FastRCNNOutputs.losses with both gradients in one packed, zero-padded buffer:
    -> (losses [3] = (loss_cls, loss_box_reg, their sum), d_packed [R, ld] = [dL_cls/dscores | dL_box/ddeltas | 0],
    ld = multiple of 128)."""
    dev = _need_cuda(scores, deltas)
    scores, deltas = _c(scores, _F32), _c(deltas, _F32)
    R, K1 = scores.shape
    K = K1 - 1
    ld = (5 * K + 1 + 127) // 128 * 128
    losses = torch.empty((3,), dtype=_F32, device=dev)
    d_packed = torch.empty((R, ld), dtype=_F32, device=dev)
    ws = _workspace(dev, max(R, 1) * 8)
    check(lib().unit_fastrcnn_loss_packed(_ptr(scores), _ptr(deltas), _ptr(_c(proposals, _F32)), _ptr(_c(gt_boxes, _F32)),
                                          _ptr(_c(gt_classes, torch.int64)), R, K, float(weights[0]), float(weights[1]),
                                          float(weights[2]), float(weights[3]), float(beta), _ptr(losses),
                                          _ptr(d_packed), ld, _ptr(ws), ws.numel(), _stream()),
          "unit_fastrcnn_loss_packed")
    return losses, d_packed


class _FTStepLossFn(torch.autograd.Function):
    """The fine-tune predictor + losses as ONE autograd node (fast_rcnn.py:484-533 followed by FastRCNNOutputs.losses):
    packed tcgen05 GEMM [delta | bbox | ft | mean-OICR](x) and mean-OICR(x_weak) -> fused similarity + transfer ->
    fused CE + smooth-L1 (+ their gradients).  Backward is one tcgen05 weight-gradient GEMM that writes the gradients
    of cls_score_ft / bbox_pred_ft (scaled by the upstream loss gradients) directly into the parameters' existing
    ``.grad`` buffers -- the views of the flat all-reduce bucket -- or returns them when no buffer is bound.

    Preconditions (checked by the caller): only the four ft tensors require grad; x / x_weak do not."""

    @staticmethod
    def forward(ctx, cls_w, cls_b, box_w, box_b, x, xw, pack, spec, proposals, gt_boxes, gt_classes, weights, beta):
        K1 = spec.K + 1
        y1, y2 = predictor_gemm2(x, pack.W, pack.b, xw, pack.W[pack.o_vis:pack.o_vis + K1],
                                 pack.b[pack.o_vis:pack.o_vis + K1])
        delta, pd = y1[:, :K1], y1[:, K1:pack.o_ft]
        ft_s, ft_d = y1[:, pack.o_ft:pack.o_ft + K1], y1[:, pack.o_ft + K1:pack.o_vis]
        vis = y1[:, pack.o_vis:pack.o_vis + K1]
        scores, bbox, _ = similarity_transfer_forward(spec, vis, delta, pd, y2[:, :K1], ft_s, ft_d, True, False, ())
        losses, d_packed = fastrcnn_loss_packed(scores, bbox, proposals, gt_boxes, gt_classes, weights, beta)
        ctx.save_for_backward(x, d_packed)
        ctx.params = (cls_w, cls_b, box_w, box_b)
        ctx.K1 = K1
        ctx.n_cols = 5 * spec.K + 1
        total = losses[2]
        ctx.set_materialize_grads(False)  # no zero tensors for the unused gradients of scores / bbox / total
        ctx.mark_non_differentiable(scores, bbox, total)
        return losses[0], losses[1], scores, bbox, total

    @staticmethod
    def backward(ctx, g_cls, g_box, _gs, _gb, _gt):
        x, d_packed = ctx.saved_tensors
        cls_w, cls_b, box_w, box_b = ctx.params
        params = (cls_w, cls_b, box_w, box_b)
        need = ctx.needs_input_grad[:4]
        bound = all((not n) or (p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == _F32)
                    for p, n in zip(params, need))
        if bound:  # accumulate in place (what autograd's AccumulateGrad would do with a returned tensor)
            dst = [p.grad if n else None for p, n in zip(params, need)]
            ret = (None, None, None, None)
        else:
            dst = [torch.empty_like(p) if n else None for p, n in zip(params, need)]
            ret = tuple(dst)
        if g_cls is None and g_box is None:
            return (None,) * 13
        zero = None
        if g_cls is None or g_box is None:  # only one of the two losses was differentiated
            zero = torch.zeros((1,), dtype=_F32, device=x.device)
        sc = [zero if g is None else _c(g.reshape(1), _F32) for g in (g_cls, g_box)]
        # overwrite_bound_grads(): the caller guarantees the bound buffers hold nothing to keep (it would have zeroed
        # them), so the kernel writes instead of accumulating and the zero fill is not needed
        predictor_wgrad(d_packed, x, ctx.n_cols, [0, ctx.K1, ctx.n_cols], [dst[0], dst[2]], [dst[1], dst[3]], sc,
                        accumulate=bound and not _OVERWRITE_BOUND[0])
        return ret + (None,) * 9


_OVERWRITE_BOUND = [False]


class overwrite_bound_grads:
    """Context for ONE backward of the fused fine-tune node: gradients are WRITTEN into the bound ``.grad`` buffers
    (the flat all-reduce bucket) instead of accumulated, which makes the bucket's zero fill before the step
    unnecessary.  Only for callers that own the buffers and would have zeroed them (RoIStage)."""

    def __enter__(self):
        self._prev, _OVERWRITE_BOUND[0] = _OVERWRITE_BOUND[0], True
        return self

    def __exit__(self, *exc):
        _OVERWRITE_BOUND[0] = self._prev
        return False


def ft_step_losses(cls_w, cls_b, box_w, box_b, x, xw, pack, spec, proposals, gt_boxes, gt_classes,
                   weights=(10.0, 10.0, 5.0, 5.0), beta=0.0):
    """-> (loss_cls, loss_box_reg, scores [detached], bbox [detached], loss_cls + loss_box_reg [detached])."""
    return _FTStepLossFn.apply(cls_w, cls_b, box_w, box_b, x, xw, pack, spec, proposals, gt_boxes, gt_classes,
                               tuple(weights), float(beta))


# ------------------------------------------------------------------------------------------------- masks
def mask_transfer(logits: torch.Tensor, s_seg: Optional[torch.Tensor], spec: TransferSpec,
                  x_delta: Optional[torch.Tensor] = None, pred_classes: Optional[torch.Tensor] = None,
                  want_logits: bool = False, want_probs: bool = True):
    dev = _need_cuda(logits)
    logits = _c(logits, _F32)
    D, K, M1, M2 = logits.shape
    MM = M1 * M2
    out_logits = torch.empty_like(logits) if want_logits else None
    out_probs = torch.empty((D, 1, M1, M2), dtype=_F32, device=dev) if want_probs else None
    if D == 0:
        return out_logits, out_probs
    s = None if s_seg is None else _c(s_seg, _F32)
    xd = None if x_delta is None else _c(x_delta, _F32)
    pc = None if pred_classes is None else _c(pred_classes, torch.int64)
    check(lib().unit_mask_transfer(_ptr(logits), _ptr(s), int(s is not None and s.dim() == 2), _ptr(spec.base_i32),
                                   _ptr(spec.novel_i32), _ptr(spec.class_kind), _ptr(xd), _ptr(pc), _ptr(out_logits),
                                   _ptr(out_probs), D, K, spec.B, spec.Nn, MM, _stream()), "unit_mask_transfer")
    return out_logits, out_probs


def _mask_paste_impl(masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int], threshold: float = 0.5):
    """[D2] paste_masks_in_image: masks [D,M,M] -> bool [D,H,W]."""
    dev = _need_cuda(masks, boxes)
    masks, boxes = _c(masks, _F32), _c(boxes, _F32)
    D, M = masks.shape[0], masks.shape[-1]
    h, w = int(image_shape[0]), int(image_shape[1])
    out = torch.empty((D, h, w), dtype=torch.uint8, device=dev)
    if D:
        check(lib().unit_mask_paste(_ptr(masks), _ptr(boxes), D, M, h, w, float(threshold), _ptr(out), _stream()),
              "unit_mask_paste")
    return out.view(torch.bool)


def mask_paste(masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int], threshold: float = 0.5):
    """[D2] paste_masks_in_image (torch.ops.unit_b200.mask_paste): masks [D,M,M] -> bool [D,H,W]."""
    return torch.ops.unit_b200.mask_paste(masks, boxes, int(image_shape[0]), int(image_shape[1]), float(threshold))


from . import torch_ops as _torch_ops  # noqa: E402,F401  (registers torch.ops.unit_b200.* on top of the *_impl functions)
