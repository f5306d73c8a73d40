"""Generate tests/golden/*.pt by executing the reference's own files VERBATIM (through oracle.shim).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The fixtures pin oracle.unit_ref (and through it the CUDA path) to what the authors' code computes:
  matcher.pt        modeling/matcher.py Matcher.__call__ on fixed + random IoU matrices
  lingual.pt        fast_rcnn.py get_similarity on the shipped data/embeddings/glove_mean (VOC split 1, COCO split)
  predictor_*.pt    roi_heads.py get_similarity_matrices + fast_rcnn.py predictor forward (+ inference)
  head_voc.pt       WSROIHeadNoMeta.forward end to end (pooler -> stand-in box head -> transfer -> NMS)
  mask_head.pt      mask_head.py MaskRCNNConvUpsampleHeadWithFineTune.forward (transfer + mask_rcnn_inference)
  weak_label.pt     weak_detector_fast_rcnn.py label_and_sample_proposals (pairwise_iou + UniT Matcher)
  weak_losses.pt    weak_detector_fast_rcnn.py WeakDetectorOutputsBase.forward + losses (MIL + 3 OICR refinements),
                    the per-iteration compute_loss_inputs outputs and the gradients of the summed loss
  outputs_variants.pt  fast_rcnn.py FastRCNNOutputsReduction / NLL / Regression (+ the weak detector's Regression) losses
                    and gradients, [D2] predict_boxes_for_gt_classes
  glove_mean.pt     the reference's only shipped data fixture, re-saved as a bare tensor
"""
from __future__ import annotations

import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import shim  # noqa: E402
from oracle.d2.structures import Boxes, Instances, ShapeSpec  # noqa: E402

EMB = os.path.join(shim.REFERENCE_ROOT, "data", "embeddings", "glove_mean")
D_FEAT = 64  # predictor input width used by the fixtures (arithmetic is independent of it)


def _seeded(seed):
    return torch.Generator().manual_seed(seed)


def randomize_(module, seed, scale=0.2):
    """Replace the (partly zero) inits by seeded N(0, scale^2) so softmax / transfer are not degenerate."""
    g = _seeded(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            if name.startswith("embeddings") or ".proj." in name or name.startswith("proj."):
                continue
            s = scale * (0.1 if "bbox" in name else 1.0)
            p.copy_(torch.randn(p.shape, generator=g) * s)


def boxes_in_image(n, h, w, g, min_size=8.0):
    cx = torch.rand(n, generator=g) * w
    cy = torch.rand(n, generator=g) * h
    bw = min_size + 0.6 * w * torch.rand(n, generator=g) ** 2
    bh = min_size + 0.6 * h * torch.rand(n, generator=g) ** 2
    b = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, w)
    b[:, 1::2] = b[:, 1::2].clamp(0, h)
    return b


def make_matcher(ns):
    M = ns.matcher.Matcher
    out = {}
    iou = torch.tensor([[0.1, 0.6, 0.5, 0.0], [0.3, 0.6, 0.2, 0.0]])
    out["kat_iou"] = iou
    out["kat_default"] = M([0.5], [0, 1])(iou)
    out["kat_empty"] = M([0.5], [0, 1])(torch.zeros(0, 4))
    m3 = M([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    out["kat_lowq"] = m3(iou)
    g = _seeded(7)
    r = torch.rand(6, 257, generator=g)
    r[:, 10] = 0.0
    r[2, 20] = r[4, 20] = 0.9  # tie -> first index
    r[:, 30] = 0.5             # exactly on the threshold -> positive
    out["rand_iou"] = r
    out["rand_default"] = M([0.5], [0, 1])(r)
    out["rand_lowq"] = M([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(r)
    return out


def make_lingual(ns):
    out = {}
    for tag, yaml_rel in (("voc", "VOC/VOC-RCNN-101-C4-split1.yaml"), ("coco", "COCO/COCO-RCNN-50-C4-split1.yaml")):
        cfg = shim.reference_cfg(yaml_rel, ["MODEL.ROI_HEADS.EMBEDDING_PATH", EMB])
        head = ns.roi_heads.WSROIHeadNoMeta(cfg, {"res4": ShapeSpec(channels=8, stride=16)})
        sim = head.box_predictor.get_similarity(base_classes=head._base_classes_tensor,
                                                novel_classes=head._novel_classes_tensor,
                                                indexer=head._coco_indexer_tensor)
        out[tag] = {"indexer": head._coco_indexer_tensor.clone(), "base": head._base_classes_tensor.clone(),
                    "novel": head._novel_classes_tensor.clone(), "lingual": sim.detach().clone()}
    return out


def _predictor_case(ns, yaml_rel, seed, R, overrides=()):
    cfg = shim.reference_cfg(yaml_rel, ["MODEL.ROI_HEADS.EMBEDDING_PATH", EMB] + list(overrides))
    shim._StandInBoxHead.OUT = D_FEAT
    try:
        head = ns.roi_heads.__dict__[cfg.MODEL.ROI_HEADS.NAME](cfg, {"res4": ShapeSpec(channels=8, stride=16)})
    finally:
        shim._StandInBoxHead.OUT = 2048
    randomize_(head.box_predictor, seed)
    g = _seeded(seed + 1)
    x = torch.relu(torch.randn(R, D_FEAT, generator=g))
    xw = torch.relu(torch.randn(R, D_FEAT, generator=g))
    return cfg, head, x, xw


def make_predictor(ns, tag, yaml_rel, seed, R, training, overrides=()):
    cfg, head, x, xw = _predictor_case(ns, yaml_rel, seed, R, overrides)
    head.train(training)
    pred = head.box_predictor
    head.move_mappings_to_gpu() if False else None
    with torch.no_grad():
        kind = type(pred).__name__
        sim = None
        if (not training) or kind != "SupervisedDetectorOutputsBase":
            sim = head.get_similarity_matrices(x)
        (scores, bbox), _ = pred(x, supervised_branch_x_weak=xw, novel_classes=head._novel_classes_tensor,
                                 base_classes=head._base_classes_tensor, x_weak=None, similarity=sim)
        out = {
            "kind": kind, "training": training, "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "threshold": cfg.MODEL.ROI_HEADS.VISUAL_ATTENTION_HEAD.VISUAL_SIMILARITY_THRESHOLD,
            "terms": {k: list(v) for k, v in head.terms.items()},
            "x": x, "x_weak_branch": xw,
            "weights": {k: v.detach().clone() for k, v in pred.state_dict().items()
                        if not k.startswith("weak_detector_head.classifier") and not k.startswith(
                            "weak_detector_head.detection") and not k.startswith("embeddings")},
            "indexer": head._coco_indexer_tensor.clone(), "base": head._base_classes_tensor.clone(),
            "novel": head._novel_classes_tensor.clone(),
            "similarity": None if sim is None else {k: v.clone() for k, v in sim.items()},
            "scores": scores.clone(), "bbox": bbox.clone(),
        }
        if not training:
            gb = _seeded(seed + 2)
            h, w = 800, 1333
            props = [Instances((h, w), proposal_boxes=Boxes(boxes_in_image(R, h, w, gb)),
                               objectness_logits=torch.zeros(R))]
            insts, kept = pred.inference([scores, bbox], props)
            out["proposal_boxes"] = props[0].proposal_boxes.tensor.clone()
            out["image_size"] = (h, w)
            out["det_boxes"] = insts[0].pred_boxes.tensor.clone()
            out["det_scores"] = insts[0].scores.clone()
            out["det_classes"] = insts[0].pred_classes.clone()
            out["det_roi_idx"] = kept[0].clone()
            tta_out, _ = pred.inference([scores, bbox], props, tta=True)
            out["tta_probs"] = tta_out[0].clone()
    return out


def make_head_voc(ns):
    cfg = shim.reference_cfg("VOC/VOC-RCNN-101-C4-split1.yaml", ["MODEL.ROI_HEADS.EMBEDDING_PATH", EMB])
    C, H, W = 16, 25, 42
    shim._StandInBoxHead.OUT = D_FEAT
    try:
        head = ns.roi_heads.WSROIHeadNoMeta(cfg, {"res4": ShapeSpec(channels=C, stride=16)})
    finally:
        shim._StandInBoxHead.OUT = 2048
    randomize_(head.box_predictor, 31)
    head.eval()
    g = _seeded(32)
    feats = torch.randn(2, C, H, W, generator=g)
    img = (H * 16, W * 16)
    props = [Instances(img, proposal_boxes=Boxes(boxes_in_image(48, img[0], img[1], g, 16.0)),
                       objectness_logits=torch.randn(48, generator=g)) for _ in range(2)]
    with torch.no_grad():
        pooled = head.box_pooler([feats], [p.proposal_boxes for p in props])
        insts, _ = head(None, {"res4": feats}, props)
    return {
        "features": feats, "image_size": img,
        "proposal_boxes": [p.proposal_boxes.tensor.clone() for p in props],
        "pooled_sum_per_roi": pooled.sum(dim=(1, 2, 3)), "pooled_first": pooled[:2].clone(),
        "state_dict": {k: v.detach().clone() for k, v in head.state_dict().items()
                       if "classifier_stream" not in k and "detection_stream" not in k and "embeddings" not in k},
        "det_boxes": [i.pred_boxes.tensor.clone() for i in insts],
        "det_scores": [i.scores.clone() for i in insts],
        "det_classes": [i.pred_classes.clone() for i in insts],
    }


def make_mask_head(ns):
    cfg = shim.reference_cfg("COCO/COCO-RCNN-50-C4-split1-segm-ft.yaml", ["MODEL.ROI_HEADS.EMBEDDING_PATH", EMB])
    mh = ns.mask_head.MaskRCNNConvUpsampleHeadWithFineTune(cfg, ShapeSpec(channels=32, height=7, width=7))
    g = _seeded(41)
    with torch.no_grad():
        for _, p in sorted(mh.named_parameters()):
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    mh.eval()
    D, K = 4, 80
    base = torch.tensor(cfg.DATASETS.FEWSHOT.BASE_CLASSES_ID)
    novel = torch.tensor(cfg.DATASETS.FEWSHOT.NOVEL_CLASSES_ID)
    x = torch.randn(D, 32, 7, 7, generator=g)
    s = torch.rand(D, len(novel), len(base), generator=g)
    s = s / s.sum(-1, keepdim=True)
    classes = torch.tensor([0, 7, 62, 79])
    inst = [Instances((480, 640), pred_classes=classes,
                      pred_boxes=Boxes(boxes_in_image(D, 480, 640, g, 12.0)))]
    with torch.no_grad():
        fixed, delta = mh.layers(x)
        mh(x, inst, similarity={"seg": s}, base_classes=base, novel_classes=novel)
    from oracle.d2.ops import paste_masks_in_image

    pasted = paste_masks_in_image(inst[0].pred_masks[:, 0], inst[0].pred_boxes, (480, 640), 0.5)
    return {"logits_fixed": fixed, "logits_delta": delta, "similarity_seg": s, "base": base, "novel": novel,
            "pred_classes": classes, "pred_boxes": inst[0].pred_boxes.tensor.clone(),
            "pred_masks": inst[0].pred_masks.clone(), "image_size": (480, 640),
            "pasted_sum": pasted.sum(dim=(1, 2)).to(torch.int64), "pasted_packed": torch.from_numpy(
                __import__("numpy").packbits(pasted.numpy(), axis=-1))}


def make_weak_label(ns):
    cfg = shim.reference_cfg("VOC/VOC-RCNN-101-C4-split1.yaml", ["MODEL.ROI_HEADS.EMBEDDING_PATH", EMB])
    wd = ns.weak.WeakDetectorOutputsBase(cfg, ShapeSpec(channels=8))
    g = _seeded(51)
    img = (600, 800)
    out = {"image_size": img, "cases": []}
    for n_gt in (3, 0, 5):
        gt = boxes_in_image(n_gt, img[0], img[1], g, 40.0)
        props = boxes_in_image(64, img[0], img[1], g, 16.0)
        if n_gt:
            jit = gt[torch.randint(0, n_gt, (16,), generator=g)] + torch.randn(16, 4, generator=g) * 6
            props[:16] = jit
            props[16] = gt[0]
        gcls = torch.randint(0, 20, (n_gt,), generator=g)
        p = [Instances(img, proposal_boxes=Boxes(props), objectness_logits=torch.zeros(64))]
        t = [Instances(img, gt_boxes=Boxes(gt), gt_classes=gcls)]
        res, assign, vals = wd.label_and_sample_proposals(p, t, return_match_vals=True)
        out["cases"].append({"gt_boxes": gt, "gt_classes": gcls, "proposal_boxes": props,
                             "out_gt_classes": res[0].gt_classes.clone(), "assign": assign[0].clone(),
                             "vals": vals[0].clone(), "out_gt_boxes": res[0].gt_boxes.tensor.clone()})
    return out


def make_weak_losses(ns):
    cfg = shim.reference_cfg("VOC/VOC-RCNN-101-C4-split1.yaml", ["MODEL.ROI_HEADS.EMBEDDING_PATH", EMB])
    wd = ns.weak.WeakDetectorOutputsBase(cfg, ShapeSpec(channels=D_FEAT))
    randomize_(wd, 61, scale=0.35)
    wd.train()
    g = _seeded(62)
    img = (600, 800)
    counts = (96, 130, 64)
    targets = [torch.tensor([3, 7, 3, 11]), torch.tensor([5]), torch.tensor([19, 0, 8, 8, 14])]
    props = []
    for n in counts:  # clusters of jittered copies so that many proposals overlap the picked boxes
        seeds = boxes_in_image(8, img[0], img[1], g, 60.0)
        b = boxes_in_image(n, img[0], img[1], g, 16.0)
        k = n // 2
        b[:k] = seeds[torch.randint(0, 8, (k,), generator=g)] + torch.randn(k, 4, generator=g) * 8
        b[:, 0::2] = b[:, 0::2].clamp(0, img[1])
        b[:, 1::2] = b[:, 1::2].clamp(0, img[0])
        b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 4)
        props.append(b[torch.randperm(n, generator=g)])
    x = torch.randn(sum(counts), D_FEAT, generator=g).requires_grad_(True)
    P = [Instances(img, proposal_boxes=Boxes(b), objectness_logits=torch.zeros(len(b))) for b in props]
    preds, _ = wd(x)
    cls_s, det_s, oicr_scores = preds[0], preds[1], preds[2]
    losses = wd.losses(preds, P, targets)
    total = sum(losses.values())
    leaves = [cls_s, det_s, *oicr_scores]
    grads = torch.autograd.grad(total, leaves + [x], retain_graph=True)
    # the supervision every refinement classifier received (weak_detector_fast_rcnn.py:218-228)
    import numpy as np
    indices = np.insert(np.cumsum(counts), 0, 0)
    uniq = [torch.unique(t) for t in targets]
    with torch.no_grad():
        mil = torch.cat([torch.softmax(c, -1) * torch.softmax(d, 0)
                         for c, d in zip(cls_s.split(list(counts)), det_s.split(list(counts)))], 0)
        sup = []
        for idx in range(len(oicr_scores)):
            probs = mil if idx == 0 else torch.softmax(oicr_scores[idx - 1].detach(), -1)
            li = wd.compute_loss_inputs(P, probs.clone(), uniq, None, indices)
            sup.append({"labels": li["labels"].clone(), "cls_weights": li["cls_weights"].clone()})
    return {"state": {k: v.detach().clone() for k, v in wd.state_dict().items()}, "x": x.detach().clone(),
            "image_size": img, "proposal_boxes": props, "targets": targets,
            "bg_threshold": cfg.MODEL.ROI_HEADS.FAST_RCNN.WEAK_DETECTOR.BG_THRESHOLD,
            "mil_multiplier": cfg.MODEL.ROI_HEADS.FAST_RCNN.WEAK_DETECTOR.MIL_MULTIPLIER,
            "cls_stream": cls_s.detach().clone(), "det_stream": det_s.detach().clone(),
            "oicr_scores": [o.detach().clone() for o in oicr_scores], "mil_scores": mil,
            "losses": {k: v.detach().clone() for k, v in losses.items()}, "supervision": sup,
            "grad_cls_stream": grads[0], "grad_det_stream": grads[1], "grad_oicr_scores": list(grads[2:5]),
            "grad_x": grads[5]}


def make_outputs_variants(ns):
    """fast_rcnn.py:24-130 FastRCNNOutputsReduction / NLL / Regression and weak_detector_fast_rcnn.py:23-37, run
    verbatim on seeded predictions (losses + gradients), plus [D2] predict_boxes_for_gt_classes (restated D2 glue)."""
    import types as _types

    from oracle.d2.modeling import FastRCNNOutputLayers
    from oracle.d2.ops import Box2BoxTransform

    g = _seeded(71)
    K, img = 20, (600, 800)
    counts = (70, 58)
    b2b = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
    props = []
    for n in counts:
        pb = boxes_in_image(n, img[0], img[1], g, 16.0)
        gt = pb + torch.randn(n, 4, generator=g) * 6
        gt[:, 2:] = torch.maximum(gt[:, 2:], gt[:, :2] + 4)
        cls = torch.randint(0, K + 1, (n,), generator=g)
        props.append(Instances(img, proposal_boxes=Boxes(pb), gt_boxes=Boxes(gt), gt_classes=cls))
    R = sum(counts)
    scores = (2.0 * torch.randn(R, K + 1, generator=g)).requires_grad_(True)
    deltas = (0.5 * torch.randn(R, 4 * K, generator=g)).requires_grad_(True)
    weights = torch.rand(R, generator=g)
    out = {"image_size": img, "proposal_boxes": [p.proposal_boxes.tensor for p in props],
           "gt_boxes": [p.gt_boxes.tensor for p in props], "gt_classes": [p.gt_classes for p in props],
           "scores": scores.detach().clone(), "deltas": deltas.detach().clone(), "weights": weights, "cases": {}}

    def run(tag, obj, reduce):
        losses = obj.losses()
        total = sum(reduce(v) for v in losses.values())
        gs, gd = torch.autograd.grad(total, [scores, deltas], allow_unused=True)
        out["cases"][tag] = {"losses": {k: v.detach().clone() for k, v in losses.items()}, "grad_scores": gs,
                             "grad_deltas": gd}

    F = ns.fast_rcnn
    for beta in (0.0, 0.4):
        run(f"reduction_beta{beta}", F.FastRCNNOutputsReduction(b2b, scores, deltas, props, beta, "smooth_l1"),
            lambda v: (v * torch.linspace(0.5, 1.5, v.numel()).view(v.shape)).sum())
        run(f"regression_beta{beta}", F.FastRCNNOutputsRegression(b2b, scores, deltas, props, weights, beta, "smooth_l1"),
            lambda v: v)
        run(f"weak_regression_beta{beta}",
            ns.weak.FastRCNNOutputsRegression(b2b, scores, deltas, props, weights, beta, "smooth_l1"), lambda v: v)
    logp = torch.log_softmax(scores, -1)
    nll = F.FastRCNNOutputsNLL(b2b, logp, deltas, props, 0.0, "smooth_l1")
    losses = nll.losses()
    gs, gd = torch.autograd.grad(sum(losses.values()), [scores, deltas])
    out["cases"]["nll"] = {"losses": {k: v.detach().clone() for k, v in losses.items()}, "grad_scores": gs,
                           "grad_deltas": gd}
    holder = _types.SimpleNamespace(box2box_transform=b2b)
    with torch.no_grad():
        pb = FastRCNNOutputLayers.predict_boxes_for_gt_classes(holder, (scores.detach(), deltas.detach()), props)
    out["pred_boxes_for_gt_classes"] = [b.clone() for b in pb]
    return out


def main():
    assert shim.reference_available(), "needs /root/reference"
    ns = shim.load_reference()
    torch.manual_seed(0)
    emb = torch.load(EMB)["embeddings"]
    sha = hashlib.sha256(open(EMB, "rb").read()).hexdigest()
    fixtures = {
        "glove_mean.pt": {"embeddings": emb.clone(), "source_sha256": sha},
        "matcher.pt": make_matcher(ns),
        "lingual.pt": make_lingual(ns),
        "predictor_voc_base_eval.pt": make_predictor(ns, "voc", "VOC/VOC-RCNN-101-C4-split1.yaml", 11, 24, False),
        "predictor_voc_base_train.pt": make_predictor(ns, "voc", "VOC/VOC-RCNN-101-C4-split1.yaml", 12, 8, True),
        "predictor_voc_ft_train.pt": make_predictor(ns, "voc", "VOC/FT/10_shot/VOC-RCNN-101-C4-split1-ft.yaml", 13,
                                                    24, True),
        "predictor_voc_ft_eval.pt": make_predictor(ns, "voc", "VOC/FT/10_shot/VOC-RCNN-101-C4-split1-ft.yaml", 14,
                                                   24, False),
        "predictor_coco_ft_eval.pt": make_predictor(ns, "coco", "COCO/COCO-RCNN-50-C4-split1-ft.yaml", 15, 8, False),
        # non-default similarity terms (roi_heads.py:273-315), reachable through FINETUNE_TERMS overrides
        "predictor_voc_ft_terms.pt": make_predictor(
            ns, "voc", "VOC/FT/10_shot/VOC-RCNN-101-C4-split1-ft.yaml", 16, 24, False,
            ["MODEL.ROI_HEADS.FINETUNE_TERMS.CLASSIFIER", ["lingual", "WTopK-3", "visual"],
             "MODEL.ROI_HEADS.FINETUNE_TERMS.BBOX", ["LSDA-2", "VisualK-4"]]),
        "head_voc.pt": make_head_voc(ns),
        "mask_head.pt": make_mask_head(ns),
        "weak_label.pt": make_weak_label(ns),
        "weak_losses.pt": make_weak_losses(ns),
        "outputs_variants.pt": make_outputs_variants(ns),
    }
    only = set(sys.argv[1:])  # optional: regenerate just the named fixtures
    for name, obj in fixtures.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, name)
        torch.save(obj, path)
        print(f"{name:32s} {os.path.getsize(path) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
