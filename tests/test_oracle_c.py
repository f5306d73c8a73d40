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
