"""Weak-image training losses on the GPU: MIL image-level loss, OICR pseudo-labelling
and the weighted refinement loss vs (a) the reference run verbatim (tests/golden/weak_losses.pt) and (b) the CPU
oracle restatement on larger seeded inputs.  Labels, picked proposals and loss weights are bit-exact."""
import os
import sys

import pytest
import torch

from conftest import ROOT, assert_close_rms, load_golden, random_boxes, seeded

sys.path.insert(0, ROOT)
from oracle import unit_ref  # noqa: E402  (the checker, never the thing measured)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from unit_b200 import ops as _ops
    return _ops


def _weak_head(gold):
    from unit_b200.config import load_cfg
    from unit_b200.predictors import WeakDetectorOutputsBase
    from unit_b200.structures import ShapeSpec

    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_base.yaml"),
                   ["MODEL.ROI_HEADS.EMBEDDING_PATH", os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])
    wd = WeakDetectorOutputsBase(cfg, ShapeSpec(channels=gold["x"].shape[1]))
    wd.load_state_dict(gold["state"], strict=True)
    assert wd.bg_threshold == gold["bg_threshold"] and wd.mil_multiplier == gold["mil_multiplier"]
    return wd.cuda().train()


def _proposals(gold):
    from unit_b200.structures import Boxes, Instances
    return [Instances(gold["image_size"], proposal_boxes=Boxes(b.cuda())) for b in gold["proposal_boxes"]]


def test_weak_losses_match_reference_fixture():
    gold = load_golden("weak_losses.pt")
    wd = _weak_head(gold)
    x = gold["x"].cuda().requires_grad_(True)
    preds, _ = wd(x)
    assert_close_rms(preds[0].detach().cpu(), gold["cls_stream"], 1e-5, "classifier stream")
    assert_close_rms(preds[1].detach().cpu(), gold["det_stream"], 1e-5, "detection stream")
    props, targets = _proposals(gold), [t.cuda() for t in gold["targets"]]
    _, supervision = wd.oicr_supervision(preds, props, targets)
    for (labels, weights), g in zip(supervision, gold["supervision"]):
        assert torch.equal(labels.cpu(), g["labels"])
        assert torch.allclose(weights.cpu(), g["cls_weights"], rtol=2e-6, atol=0)
        assert torch.equal(weights.cpu() == 0, g["cls_weights"] == 0)
    losses = wd.losses(preds, props, targets)
    assert set(losses) == set(gold["losses"])
    for k, v in gold["losses"].items():
        assert abs(losses[k].item() - v.item()) <= 1e-5 * max(abs(v.item()), 1e-3), (k, losses[k].item(), v.item())
    sum(losses.values()).backward()
    assert_close_rms(x.grad.cpu(), gold["grad_x"], 2e-5, "d loss / d x_weak")


def test_weak_losses_from_the_fixture_streams(ops):
    """Same fixture, feeding the reference's own prediction tensors: isolates the loss kernels from the Linears."""
    gold = load_golden("weak_losses.pt")
    counts = [len(b) for b in gold["proposal_boxes"]]
    off = ops.offsets_from_counts(counts, torch.device("cuda"))
    gt = torch.zeros(len(counts), 20)
    for i, t in enumerate(gold["targets"]):
        gt[i, t] = 1
    cls_s = gold["cls_stream"].cuda().requires_grad_(True)
    det_s = gold["det_stream"].cuda().requires_grad_(True)
    loss, mil, vec = ops.mil_loss(cls_s, det_s, off, gt.cuda(), gold["mil_multiplier"])
    assert_close_rms(mil.cpu(), gold["mil_scores"], 1e-5, "mil scores")
    assert abs(loss.item() - gold["losses"]["loss_im_cls"].item()) <= 1e-5 * gold["losses"]["loss_im_cls"].item()
    loss.backward()
    assert_close_rms(cls_s.grad.cpu(), gold["grad_cls_stream"], 1e-5, "d loss_im_cls / d classifier stream")
    assert_close_rms(det_s.grad.cpu(), gold["grad_det_stream"], 1e-5, "d loss_im_cls / d detection stream")
    boxes = torch.cat(gold["proposal_boxes"]).cuda()
    probs = gold["mil_scores"].cuda()
    for idx, g in enumerate(gold["supervision"]):
        if idx > 0:
            probs = torch.softmax(gold["oicr_scores"][idx - 1], -1).cuda()
        labels, weights, _, _ = ops.oicr_targets(probs, boxes, off, gt.cuda(), [0.5], [0, 1], gold["bg_threshold"])
        assert torch.equal(labels.cpu(), g["labels"])
        assert torch.equal(weights.cpu(), g["cls_weights"])  # copies of the given probabilities: exact
        s = gold["oicr_scores"][idx].cuda().requires_grad_(True)
        l = ops.weighted_ce_loss(s, labels, weights)
        want = gold["losses"]["loss_oicr_{}".format(idx + 1)].item()
        assert abs(l.item() - want) <= 1e-5 * max(abs(want), 1e-3)
        l.backward()
        assert_close_rms(s.grad.cpu(), gold["grad_oicr_scores"][idx], 1e-5, f"d loss_oicr_{idx + 1}")


@pytest.mark.parametrize("K,ld,counts,seed", [(20, 20, (300, 0, 1500, 7), 1), (80, 81, (2000, 2000, 1, 640), 2),
                                              (80, 81, (4096,) * 8, 3)])
def test_oicr_targets_bit_exact_vs_oracle(ops, K, ld, counts, seed):
    g = seeded(400 + seed)
    boxes, classes = [], []
    for n in counts:
        b = random_boxes(max(n, 1), 600, 800, g, 16.0)[:n]
        if n > 8:  # clusters around a few boxes so that IoU >= 0.5 and IoU < 0.1 both occur
            k = n // 2
            b[:k] = b[torch.randint(0, 8, (k,), generator=g)] + torch.randn(k, 4, generator=g) * 6
            b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 2)
        boxes.append(b)
        classes.append(torch.randint(0, K, (int(torch.randint(1, 7, (1,), generator=g)),), generator=g))
    R = sum(counts)
    probs = torch.softmax(torch.randn(R, ld, generator=g) * 3, -1)
    dup = torch.arange(0, R - 1, 7)
    probs[dup + 1] = probs[dup]  # exact ties between neighbouring proposals: the first one must win
    if counts[0] > 2:
        probs[1, :] = 0.0
    gt = torch.zeros(len(counts), K)
    for i, c in enumerate(classes):
        gt[i, c] = 1
    # the reference cannot handle an image without proposals (torch.stack of nothing): skip those in the oracle
    keep = [i for i, n in enumerate(counts) if n > 0]
    want_l, want_w, picked = unit_ref.oicr_targets(
        torch.cat([probs[sum(counts[:i]):sum(counts[:i + 1])] for i in keep]), [boxes[i] for i in keep],
        [classes[i] for i in keep], [0.5], [0, 1], 0.1, K)
    off = ops.offsets_from_counts(counts, torch.device("cuda"))
    labels, weights, pidx, pscore = ops.oicr_targets(probs.cuda(), torch.cat(boxes).cuda(), off, gt.cuda(), [0.5],
                                                     [0, 1], 0.1)
    assert torch.equal(labels.cpu(), want_l)
    assert torch.equal(weights.cpu(), want_w)
    for j, i in enumerate(keep):
        uniq = torch.unique(classes[i])
        assert torch.equal(pidx[i].cpu()[uniq], picked[j])
        absent = torch.ones(K, dtype=torch.bool)
        absent[uniq] = False
        assert (pidx[i].cpu()[absent] == -1).all()
    assert (labels < K).any() and (labels == K).any()


def test_oicr_targets_three_way_matcher(ops):
    """IOU_THRESHOLDS [0.3, 0.7] / IOU_LABELS [0, -1, 1]: ignore labels come out as -1 and carry no loss."""
    g = seeded(77)
    n = 500
    b = random_boxes(n, 600, 800, g, 16.0)
    b[:250] = b[torch.randint(0, 6, (250,), generator=g)] + torch.randn(250, 4, generator=g) * 10
    b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 2)
    probs = torch.softmax(torch.randn(n, 21, generator=g) * 2, -1)
    cls = torch.tensor([4, 9, 9, 17])
    want_l, want_w, _ = unit_ref.oicr_targets(probs, [b], [cls], [0.3, 0.7], [0, -1, 1], 0.1, 20)
    gt = torch.zeros(1, 20)
    gt[0, cls] = 1
    off = ops.offsets_from_counts([n], torch.device("cuda"))
    labels, weights, _, _ = ops.oicr_targets(probs.cuda(), b.cuda(), off, gt.cuda(), [0.3, 0.7], [0, -1, 1], 0.1)
    assert torch.equal(labels.cpu(), want_l) and torch.equal(weights.cpu(), want_w)
    assert (labels == -1).any()
    s = torch.randn(n, 21, generator=g).cuda().requires_grad_(True)
    l = ops.weighted_ce_loss(s, labels, weights)
    l.backward()
    ok = want_l >= 0
    sc = s.detach().cpu().requires_grad_(True)
    ref = (torch.nn.functional.cross_entropy(sc[ok], want_l[ok], reduction="none") * want_w[ok]).sum() / n
    ref.backward()
    assert abs(l.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert_close_rms(s.grad.cpu(), sc.grad, 1e-5, "weighted CE grad with ignored rows")
    assert (s.grad.cpu()[~ok] == 0).all()


@pytest.mark.parametrize("K,counts", [(20, (1000, 3, 0, 2000)), (80, (4000, 4000))])
def test_mil_loss_matches_oracle(ops, K, counts):
    g = seeded(500 + K)
    R = sum(counts)
    cls_s = (torch.randn(R, K, generator=g) * 2).requires_grad_(True)
    det_s = (torch.randn(R, K, generator=g) * 2).requires_grad_(True)
    classes = [torch.randint(0, K, (3,), generator=g) for _ in counts]
    x, vecs = unit_ref.mil_scores(cls_s, det_s, counts)
    vecs.retain_grad()
    want = unit_ref.mil_loss(vecs, classes, 4.0)
    want.backward()
    # sum of |terms| behind each gradient element (both are differences of nearly equal products where p ~ v):
    #   d_det = g q (p - v),  d_cls = p (g q - sum_k g_k x_k)
    with torch.no_grad():
        n_rep = torch.tensor(counts)
        gk = vecs.grad.abs().repeat_interleave(n_rep, 0)
        vk = vecs.repeat_interleave(n_rep, 0)
        p = torch.softmax(cls_s, -1)
        q = x / p.clamp_min(1e-30)
        mag_det = 0.5 * gk * q * (p + vk)
        mag_cls = 0.5 * p * (gk * q + (gk * x).sum(1, keepdim=True))
    gt = torch.zeros(len(counts), K)
    for i, c in enumerate(classes):
        gt[i, c] = 1
    off = ops.offsets_from_counts(counts, torch.device("cuda"))
    c2 = cls_s.detach().cuda().requires_grad_(True)
    d2 = det_s.detach().cuda().requires_grad_(True)
    loss, mil, vec = ops.mil_loss(c2, d2, off, gt.cuda(), 4.0)
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    assert_close_rms(mil.cpu(), x.detach(), 1e-5, "mil scores")
    assert_close_rms(vec.cpu(), vecs.detach(), 1e-5, "class vectors")
    (loss * 0.5).backward()
    assert_close_rms(c2.grad.cpu(), 0.5 * cls_s.grad, 1e-5, "d mil / d classifier stream", magnitude=mag_cls)
    assert_close_rms(d2.grad.cpu(), 0.5 * det_s.grad, 1e-5, "d mil / d detection stream", magnitude=mag_det)


def test_weak_branch_through_the_roi_head():
    """WSROIHeadNoMeta.forward with weak images (roi_heads.py:496-552): the base losses gain loss_im_cls and
    loss_oicr_*; train_only_weak returns the weak losses alone."""
    from test_heads_gpu import _StandInBoxHead, _build
    from unit_b200 import d2compat  # noqa: F401
    from unit_b200.registry import ROI_BOX_HEAD_REGISTRY
    from unit_b200.structures import Boxes, Instances

    if "StandInBoxHead" not in ROI_BOX_HEAD_REGISTRY:
        ROI_BOX_HEAD_REGISTRY._do_register("StandInBoxHead", _StandInBoxHead)
    cfg, head = _build("voc_split1_base.yaml", 64, ROI_BOX_HEAD_REGISTRY)
    head = head.cuda().train()
    g = seeded(9)
    feats = {"res4": torch.randn(2, 64, 38, 50, generator=g).cuda()}
    wfeats = {"res4": torch.randn(2, 64, 38, 50, generator=g).cuda()}
    img = (600, 800)
    props, targets, wprops = [], [], []
    base = torch.tensor(list(cfg.DATASETS.FEWSHOT.BASE_CLASSES_ID))  # novel logits are -inf while training
    for i in range(2):
        gtb = random_boxes(3, 600, 800, g, 60.0)
        pb = torch.cat([gtb + torch.randn(3, 4, generator=g) * 4, random_boxes(200, 600, 800, g, 16.0)])
        props.append(Instances(img, proposal_boxes=Boxes(pb.cuda()), objectness_logits=torch.zeros(len(pb)).cuda()))
        targets.append(Instances(img, gt_boxes=Boxes(gtb.cuda()), gt_classes=base[torch.randint(0, len(base), (3,), generator=g)].cuda()))
        wb = random_boxes(300, 600, 800, g, 16.0)
        wprops.append(Instances(img, proposal_boxes=Boxes(wb.cuda()), objectness_logits=torch.zeros(300).cuda()))
    wtargets = [torch.tensor([2, 5]).cuda(), torch.tensor([11]).cuda()]
    _, losses = head(None, feats, props, targets, weak_images=torch.zeros(1).cuda(), weak_features=wfeats,
                     weak_proposals=wprops, weak_targets=wtargets)
    assert {"loss_cls", "loss_box_reg", "loss_im_cls", "loss_oicr_1", "loss_oicr_2", "loss_oicr_3"} <= set(losses)
    assert all(torch.isfinite(v) for v in losses.values())
    sum(losses.values()).backward()
    wh = head.box_predictor.weak_detector_head
    assert wh.classifier_stream.weight.grad.abs().sum() > 0 and wh.oicr_predictors[2].weight.grad.abs().sum() > 0
    _, only = head(None, None, None, None, weak_images=torch.zeros(1).cuda(), weak_features=wfeats,
                   weak_proposals=wprops, weak_targets=wtargets, train_only_weak=True)
    assert set(only) == {"loss_im_cls", "loss_oicr_1", "loss_oicr_2", "loss_oicr_3"}
    n_weak = cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE // cfg.MODEL.ROI_HEADS.WEAK_CLASSIFIER_PROPOSAL_DIVISOR
    assert n_weak > 0
    for k in only:  # same weak inputs -> same weak losses with or without the supervised branch
        assert abs(only[k].item() - losses[k].item()) <= 1e-6 * max(abs(losses[k].item()), 1e-3), k
