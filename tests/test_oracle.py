"""CPU: the oracle against the committed golden fixtures (made by executing the reference's files verbatim) and,
in the build container, against the reference itself."""
import pytest
import torch

from conftest import assert_close_rms, load_golden
from oracle import shim, unit_ref
from oracle.d2.ops import MatcherWithVals
from oracle.d2.structures import Boxes, pairwise_iou


def test_glove_fixture_known_answers():
    g = load_golden("glove_mean.pt")
    e = g["embeddings"]
    assert g["source_sha256"] == "4a1437f4c6e7c4bc5be6112aeb42ce906f864d0fcc15a4f9e6c902c75b5f53ee"
    assert e.shape == (80, 300) and e.dtype == torch.float32
    assert abs(e.sum().item() - 73.055890) < 1e-3
    assert torch.allclose(e[0, :4], torch.tensor([-0.28545, 0.18613, -0.36656, -0.028399]), atol=1e-5)


def test_lingual_similarity_kat():
    """VOC split 1 lingual similarity known answers from the shipped glove_mean."""
    emb = load_golden("glove_mean.pt")["embeddings"]
    gold = load_golden("lingual.pt")["voc"]
    assert gold["indexer"].tolist() == [4, 1, 14, 8, 39, 5, 2, 15, 56, 19, 60, 16, 17, 3, 0, 58, 18, 57, 6, 62]
    idx = unit_ref.coco_indexer(shim.VOC_CLASSES)
    assert torch.equal(idx, gold["indexer"])
    L = unit_ref.lingual_similarity(emb, idx, gold["base"], gold["novel"])
    assert torch.allclose(L, gold["lingual"], rtol=1e-6, atol=1e-5)
    assert abs(L.min().item() - 3.904293) < 1e-3 and abs(L.max().item() - 34.102333) < 1e-3
    soft = torch.softmax(L, -1)
    base = gold["base"].tolist()
    names = shim.VOC_CLASSES
    best = [names[base[i]] for i in soft.argmax(1).tolist()]
    assert best == ["cat", "train", "sheep", "bicycle", "chair"]  # bird, bus, cow, motorbike, sofa
    assert abs(soft[0].max().item() - 0.687040) < 1e-4
    coco = load_golden("lingual.pt")["coco"]
    assert torch.equal(unit_ref.coco_indexer(shim.COCO_CLASSES), coco["indexer"])


def test_matcher_golden():
    gold = load_golden("matcher.pt")
    for key, args in (("kat_default", ([0.5], [0, 1], False)), ("kat_lowq", ([0.3, 0.7], [0, -1, 1], True))):
        out = MatcherWithVals(args[0], args[1], allow_low_quality_matches=args[2])(gold["kat_iou"])
        for a, b in zip(out, gold[key]):
            assert torch.equal(a, b)
    assert gold["kat_default"][0].tolist() == [1, 0, 0, 0] and gold["kat_default"][1].tolist() == [0, 1, 1, 0]
    assert gold["kat_lowq"][1].tolist() == [-1, 1, -1, 0]
    out = MatcherWithVals([0.5], [0, 1])(torch.zeros(0, 4))
    for a, b in zip(out, gold["kat_empty"]):
        assert torch.equal(a, b)
    for key, args in (("rand_default", ([0.5], [0, 1], False)), ("rand_lowq", ([0.3, 0.7], [0, -1, 1], True))):
        out = MatcherWithVals(args[0], args[1], allow_low_quality_matches=args[2])(gold["rand_iou"])
        for a, b in zip(out, gold[key]):
            assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["predictor_voc_base_eval.pt", "predictor_voc_base_train.pt",
                                  "predictor_voc_ft_train.pt", "predictor_voc_ft_eval.pt",
                                  "predictor_coco_ft_eval.pt"])
def test_unit_ref_against_reference_fixture(name):
    gold = load_golden(name)
    emb = load_golden("glove_mean.pt")["embeddings"]
    w = dict(gold["weights"])
    base, novel = gold["base"], gold["novel"]
    K = gold["num_classes"]
    kind = {"SupervisedDetectorOutputsBase": "Base", "SupervisedDetectorOutputsFineTune": "FineTune"}[gold["kind"]]
    sim = None
    if gold["similarity"] is not None:
        L = unit_ref.lingual_similarity(emb, gold["indexer"], base, novel)
        V = unit_ref.visual_similarity(unit_ref.oicr_mean_logits(gold["x"], w), base, gold["threshold"])
        sim = unit_ref.similarity_matrices(L, V, gold["terms"], len(novel), len(base))
        for h in ("cls", "bbox"):
            assert torch.allclose(sim[h], gold["similarity"][h], rtol=1e-5, atol=1e-6)
    scores, bbox = unit_ref.predictor_forward(gold["x"], gold["x_weak_branch"], w, sim, base, novel, K, kind=kind,
                                              training=gold["training"])
    assert torch.allclose(scores, gold["scores"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(bbox, gold["bbox"], rtol=1e-5, atol=1e-6)
    if not gold["training"]:
        insts, kept = unit_ref.box_inference(scores, bbox, [gold["proposal_boxes"]], [tuple(gold["image_size"])])
        assert torch.equal(insts[0].pred_classes, gold["det_classes"])
        assert torch.equal(kept[0], gold["det_roi_idx"])
        assert torch.allclose(insts[0].scores, gold["det_scores"], rtol=1e-5, atol=1e-7)


def test_unit_ref_nondefault_similarity_terms_fixture():
    """TopK / WTopK / LSDA / VisualK (roi_heads.py:273-315) against the reference run verbatim with
    FINETUNE_TERMS.CLASSIFIER = [lingual, WTopK-3, visual], BBOX = [LSDA-2, VisualK-4]."""
    gold = load_golden("predictor_voc_ft_terms.pt")
    emb = load_golden("glove_mean.pt")["embeddings"]
    w = dict(gold["weights"])
    base, novel, K = gold["base"], gold["novel"], gold["num_classes"]
    assert gold["terms"]["cls"] == ["lingual", "WTopK-3", "visual"] and gold["terms"]["bbox"] == ["LSDA-2", "VisualK-4"]
    logits = unit_ref.oicr_mean_logits(gold["x"], w)
    L = unit_ref.lingual_similarity(emb, gold["indexer"], base, novel)
    V = unit_ref.visual_similarity(logits, base, gold["threshold"])
    cw = torch.stack([w[f"weak_detector_head.oicr_predictors.{k}.weight"] for k in range(3)]).mean(0)
    sim = unit_ref.similarity_matrices(L, V, gold["terms"], len(novel), len(base), class_weights=cw,
                                       mean_logits=logits, base=base, novel=novel, num_classes=K)
    for h in ("cls", "bbox"):
        assert sim[h].shape == gold["similarity"][h].shape
        assert torch.allclose(sim[h], gold["similarity"][h], rtol=1e-5, atol=1e-6)
    scores, bbox = unit_ref.predictor_forward(gold["x"], gold["x_weak_branch"], w, sim, base, novel, K,
                                              kind="FineTune", training=False)
    assert torch.allclose(scores, gold["scores"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(bbox, gold["bbox"], rtol=1e-5, atol=1e-6)


def test_mask_transfer_fixture():
    gold = load_golden("mask_head.pt")
    full = unit_ref.mask_transfer(gold["logits_fixed"], gold["similarity_seg"], gold["base"], gold["novel"],
                                  gold["logits_delta"])
    D = full.shape[0]
    probs = full[torch.arange(D), gold["pred_classes"]][:, None].sigmoid()
    assert torch.allclose(probs, gold["pred_masks"], rtol=1e-5, atol=1e-6)


def test_weak_label_fixture():
    gold = load_golden("weak_label.pt")
    m = MatcherWithVals([0.5], [0, 1])
    for case in gold["cases"]:
        matches, labels, vals = m(pairwise_iou(Boxes(case["gt_boxes"]), Boxes(case["proposal_boxes"])))
        if len(case["gt_boxes"]):
            cls = case["gt_classes"][matches]
            cls[labels == 0] = 20
        else:
            cls = torch.zeros_like(matches) + 20
        assert torch.equal(cls, case["out_gt_classes"])
        assert torch.equal(matches, case["assign"])
        assert torch.equal(vals, case["vals"])


@pytest.mark.skipif(not shim.reference_available(), reason="reference checkout not present (GPU box)")
def test_reference_head_verbatim_matches_fixture():
    """Re-run the reference's own WSROIHeadNoMeta here and compare with the committed fixture (guards the fixtures
    against drift of torch / torchvision versions)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden

    ns = shim.load_reference()
    fresh = make_golden.make_head_voc(ns)
    gold = load_golden("head_voc.pt")
    for i in range(2):
        assert torch.equal(fresh["det_classes"][i], gold["det_classes"][i])
        assert torch.allclose(fresh["det_scores"][i], gold["det_scores"][i], rtol=1e-5, atol=1e-7)
    fresh = make_golden.make_matcher(ns)
    assert torch.equal(fresh["rand_default"][0], load_golden("matcher.pt")["rand_default"][0])


def test_unit_ref_weak_losses_fixture():
    """MIL + OICR losses (weak_detector_fast_rcnn.py:189-228, 353-408) against the reference run verbatim: labels
    bit-exact, loss weights / losses / gradients to fp32 round-off."""
    gold = load_golden("weak_losses.pt")
    cls_s = gold["cls_stream"].clone().requires_grad_(True)
    det_s = gold["det_stream"].clone().requires_grad_(True)
    oicr = [o.clone().requires_grad_(True) for o in gold["oicr_scores"]]
    losses, sup = unit_ref.weak_losses(cls_s, det_s, oicr, gold["proposal_boxes"], gold["targets"],
                                       bg_threshold=gold["bg_threshold"], multiplier=gold["mil_multiplier"])
    assert set(losses) == set(gold["losses"])
    for k, v in gold["losses"].items():
        assert torch.allclose(losses[k], v, rtol=1e-6, atol=1e-8), k
    for (labels, weights, _), g in zip(sup, gold["supervision"]):
        assert torch.equal(labels, g["labels"])
        assert torch.allclose(weights, g["cls_weights"], rtol=1e-6, atol=0)
        assert (labels < 20).any() and (labels == 20).any() and (weights == 0).any()
    grads = torch.autograd.grad(sum(losses.values()), [cls_s, det_s, *oicr])
    for got, want in zip(grads, [gold["grad_cls_stream"], gold["grad_det_stream"], *gold["grad_oicr_scores"]]):
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-9)
