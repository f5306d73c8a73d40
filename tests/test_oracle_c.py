"""CPU: the plain-C restatement (oracle/c) vs the torch/torchvision oracle -- two independent statements of the
integer-exact arithmetic must agree bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torchvision

from conftest import ROOT, assert_close_rms, random_boxes, seeded
from oracle.d2.ops import MatcherWithVals
from oracle.d2.structures import Boxes, pairwise_iou


@pytest.fixture(scope="module")
def clib():
    d = os.path.join(ROOT, "oracle", "c")
    subprocess.run(["make", "-s", "-C", d], check=True)
    return ctypes.CDLL(os.path.join(d, "liboracle.so"))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_c_iou_and_matcher_bit_exact(clib):
    g = seeded(3)
    gt = random_boxes(7, 800, 1333, g, 32.0)
    pr = random_boxes(500, 800, 1333, g, 16.0)
    pr[:40] = (pr[:40] / 16).round() * 16
    pr[40:60] = gt[torch.randint(0, 7, (20,), generator=g)]
    iou = np.zeros((7, 500), np.float32)
    clib.oracle_pairwise_iou(_p(gt.numpy()), 7, _p(pr.numpy()), 500, _p(iou))
    ref = pairwise_iou(Boxes(gt), Boxes(pr))
    assert torch.equal(torch.from_numpy(iou), ref)
    m, l, v = np.zeros(500, np.int64), np.zeros(500, np.int8), np.zeros(500, np.float32)
    thr, lab = np.array([0.5], np.float32), np.array([0, 1], np.int32)
    clib.oracle_matcher(_p(iou), 7, 500, _p(thr), _p(lab), 1, _p(m), _p(l), _p(v))
    rm, rl, rv = MatcherWithVals([0.5], [0, 1])(ref)
    assert torch.equal(torch.from_numpy(m), rm) and torch.equal(torch.from_numpy(l), rl)
    assert torch.equal(torch.from_numpy(v), rv)


def test_c_nms_bit_exact(clib):
    g = seeded(4)
    boxes = random_boxes(1500, 600, 600, g)
    boxes[:400] = (boxes[:400] / 32).round() * 32
    scores = torch.rand(1500, generator=g)
    scores[200:300] = 0.25
    keep = np.zeros(1500, np.int64)
    clib.oracle_nms.restype = ctypes.c_int
    n = clib.oracle_nms(_p(boxes.numpy()), _p(scores.numpy()), 1500, ctypes.c_float(0.5), _p(keep))
    ref = torchvision.ops.nms(boxes, scores, 0.5)
    assert n == ref.numel() and torch.equal(torch.from_numpy(keep[:n]), ref)


def test_c_roi_align_matches_torchvision(clib):
    g = seeded(5)
    feat = torch.randn(2, 3, 20, 31, generator=g)
    b = random_boxes(24, 320, 496, g, 8.0)
    rois = torch.cat([torch.randint(0, 2, (24, 1), generator=g).float(), b], 1)
    rois[0, 1:] = torch.tensor([-30.0, -20.0, 60.0, 70.0])
    out = np.zeros((24, 3, 14, 14), np.float32)
    clib.oracle_roi_align_fwd(_p(feat.numpy()), 2, 3, 20, 31, _p(rois.numpy()), 24, 14, 14, ctypes.c_float(1 / 16), 0, 1,
                              _p(out))
    ref = torch.ops.torchvision.roi_align(feat, rois, 1 / 16, 14, 14, 0, True)
    assert_close_rms(torch.from_numpy(out), ref, 1e-5, "C roi_align")


def test_c_oicr_targets_match_reference_fixture(clib):
    """OICR pseudo-labelling in plain C vs the reference run verbatim (tests/golden/weak_losses.pt) and vs the torch
    restatement on a larger seeded case with ties: labels and picked proposals bit-exact, weights exact copies."""
    from conftest import load_golden
    from oracle import unit_ref

    def run_c(probs, boxes, classes, bg, K):
        probs = np.ascontiguousarray(probs.numpy(), np.float32)
        labels, weights, picked = [], [], []
        start = 0
        for b, cls in zip(boxes, classes):
            R = len(b)
            uniq = np.ascontiguousarray(torch.unique(cls).numpy(), np.int64)
            lab, w, pk = np.zeros(R, np.int64), np.zeros(R, np.float32), np.zeros(len(uniq), np.int64)
            thr, ml = np.array([0.5], np.float32), np.array([0, 1], np.int32)
            p = np.ascontiguousarray(probs[start:start + R])
            clib.oracle_oicr_targets(_p(p), p.shape[1], _p(np.ascontiguousarray(b.numpy(), np.float32)), R, _p(uniq),
                                     len(uniq), _p(thr), _p(ml), 1, ctypes.c_float(bg), K, _p(lab), _p(w), _p(pk))
            start += R
            labels.append(torch.from_numpy(lab)); weights.append(torch.from_numpy(w)); picked.append(torch.from_numpy(pk))
        return torch.cat(labels), torch.cat(weights), picked

    gold = load_golden("weak_losses.pt")
    probs = gold["mil_scores"]
    for idx, sup in enumerate(gold["supervision"]):
        if idx > 0:
            probs = torch.softmax(gold["oicr_scores"][idx - 1], -1)
        lab, w, _ = run_c(probs, gold["proposal_boxes"], gold["targets"], gold["bg_threshold"], 20)
        assert torch.equal(lab, sup["labels"]) and torch.equal(w, sup["cls_weights"])
    g = seeded(41)
    boxes = []
    for n in (700, 33, 1200):
        b = random_boxes(n, 600, 800, g, 16.0)
        b[:n // 2] = b[torch.randint(0, 8, (n // 2,), generator=g)] + torch.randn(n // 2, 4, generator=g) * 6
        b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 2)
        boxes.append(b)
    classes = [torch.tensor([7, 3, 7, 60]), torch.tensor([0]), torch.tensor([79, 5, 12, 12, 41])]
    probs = torch.softmax(torch.randn(1933, 81, generator=g) * 3, -1)
    dup = torch.arange(0, 1932, 5)
    probs[dup + 1] = probs[dup]
    want_l, want_w, want_p = unit_ref.oicr_targets(probs, boxes, classes, [0.5], [0, 1], 0.1, 80)
    lab, w, pk = run_c(probs, boxes, classes, 0.1, 80)
    assert torch.equal(lab, want_l) and torch.equal(w, want_w)
    assert all(torch.equal(a, b) for a, b in zip(pk, want_p))
