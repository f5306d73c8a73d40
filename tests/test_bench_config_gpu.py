"""Oracle parity ON THE BENCHMARKED CONFIGURATION.

bench.py times BASELINE.json configs[1] with the shipped defaults: 2 x [1024,50,84] features, 2000 proposals + GT per
image, 512 sampled RoIs per image, predictor GEMMs on the tcgen05 TF32 kernel.  These tests run exactly that and check
it against the CPU oracle (torchvision CPU ROIAlign + restated Detectron2 / UniT glue):
  * sampled RoIs / classes / matched GT: bit-exact (same host generator);
  * ROIAlign forward / backward at C=1024, 50x84, 2x512: fp32 rel 1e-5, bf16 I/O one rounding (2^-8);
  * loss, scores, box deltas, cls_score_ft / bbox_pred_ft gradients through similarity -> transfer -> loss with the
    TF32 GEMM: rel 1e-2 (north_star: "bf16/tf32 transfer rel 1e-2");
  * inference (eager and CUDA-graph replay) with the TF32 GEMM: scores / deltas rel 1e-2, detections equal up to
    candidates whose score sits within the TF32 error of the 0.05 threshold or of an NMS decision.
"""
import pytest
import torch
import torchvision  # noqa: F401  (registers torch.ops.torchvision.*: the CPU reference kernels used below)

from conftest import assert_close_rms, random_boxes, seeded

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bench():
    import bench as b

    return b


@pytest.fixture(scope="module")
def roi_case():
    """bf16-representable inputs so the fp32 and the bf16-I/O runs share ONE CPU reference."""
    g = seeded(2024)
    feat = torch.randn(2, 1024, 50, 84, generator=g).bfloat16().float()
    rois = torch.cat([torch.cat([torch.full((512, 1), float(i)), random_boxes(512, 800, 1333, g, 16.0)], 1)
                      for i in range(2)])
    gout = torch.randn(1024, 1024, 14, 14, generator=g).bfloat16().float()
    torch.set_num_threads(max(torch.get_num_threads(), 1))
    ref_fwd = torch.ops.torchvision.roi_align(feat, rois, 1 / 16, 14, 14, 0, True)
    ref_bwd = torch.ops.torchvision._roi_align_backward(gout, rois, 1 / 16, 14, 14, 2, 1024, 50, 84, 0, True)
    return feat, rois, gout, ref_fwd, ref_bwd


def test_roi_align_forward_full_size_fp32(roi_case):
    from unit_b200 import ops

    feat, rois, _, ref, _ = roi_case
    out = ops.roi_align_forward(feat.cuda(), rois.cuda(), (14, 14), 1 / 16, 0, True, True)
    assert_close_rms(out.cpu(), ref, 1e-5, "roi_align fwd, C=1024 50x84 2x512, fp32")


def test_roi_align_forward_full_size_bf16(roi_case):
    from unit_b200 import ops

    feat, rois, _, ref, _ = roi_case
    out = ops.roi_align_forward(feat.bfloat16().cuda(), rois.cuda(), (14, 14), 1 / 16, 0, True, True)
    assert out.dtype == torch.bfloat16
    err = (out.float().cpu() - ref).abs()  # fp32 accumulation, one bf16 rounding of the result
    assert (err <= 2 ** -8 * ref.abs() + 1e-6).all(), err.max()


def test_roi_align_backward_full_size_fp32(roi_case):
    from unit_b200 import ops

    feat, rois, gout, _, ref = roi_case
    got = ops.roi_align_backward(gout.cuda(), rois.cuda(), feat.shape, 1 / 16, 0, True, True)
    # every cell sums hundreds of RoI contributions in a different order than the CPU loop: bound by sum |terms|
    mag = torch.ops.torchvision._roi_align_backward(gout.abs(), rois, 1 / 16, 14, 14, 2, 1024, 50, 84, 0, True)
    assert_close_rms(got.cpu(), ref, 1e-5, "roi_align bwd, C=1024 50x84 2x512, fp32", magnitude=mag)
    roi_case_mag.append(mag)


roi_case_mag = []


def test_roi_align_backward_full_size_bf16(roi_case):
    from unit_b200 import ops

    feat, rois, gout, _, ref = roi_case
    got = ops.roi_align_backward(gout.bfloat16().cuda(), rois.cuda(), feat.shape, 1 / 16, 0, True, True)
    assert got.dtype == torch.bfloat16
    mag = roi_case_mag[0] if roi_case_mag else torch.ops.torchvision._roi_align_backward(
        gout.abs(), rois, 1 / 16, 14, 14, 2, 1024, 50, 84, 0, True)
    err = (got.float().cpu() - ref).abs()
    assert (err <= 2 ** -8 * ref.abs() + 1e-5 * mag + 1e-30).all(), err.max()


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def test_train_step_bench_config_tf32_vs_oracle(bench):
    """RoIStage.train_step exactly as bench.py runs it (default gemm_precision == "tf32") vs the CPU oracle
    (reference: modeling/roi_heads/roi_heads.py:595-644 -> fast_rcnn.py:484-533 -> FastRCNNOutputs.losses)."""
    from unit_b200.distributed import FlatGradBucket
    from unit_b200.stage import RoIStage
    from unit_b200.structures import Boxes, Instances

    dev = torch.device("cuda")
    head = bench.build_head(dev)
    pred = head.box_predictor
    assert pred.gemm_precision == "tf32" and pred.weak_detector_head.gemm_precision == "tf32"
    head.sampling_generator = seeded(1000)
    bucket = FlatGradBucket([p for p in head.parameters() if p.requires_grad])
    w, meta, x, xw, gp = bench.cpu_workload()
    x_dev, xw_dev = x.to(dev), xw.to(dev)
    stage = RoIStage(head, lambda pooled: (x_dev, xw_dev), bucket)
    host = bench.make_inputs(2000)
    feats, props, gts, gcls = host
    d_props = [Instances(bench.IMG_HW, proposal_boxes=Boxes(p.to(dev)), objectness_logits=torch.zeros(len(p), device=dev))
               for p in props]
    d_tgts = [Instances(bench.IMG_HW, gt_boxes=Boxes(t.to(dev)), gt_classes=c.to(dev)) for t, c in zip(gts, gcls)]
    gp_dev = gp.to(dev)

    # ---- sampling: bit-exact against the oracle with the same host generator
    sampled = head.label_and_sample_proposals(d_props, d_tgts)
    _, _, _, _, ref_pooled, ref_gfeat = bench.cpu_reference_step(host, w, meta, seeded(1000), x, xw, gp)
    ref = bench.cpu_reference_step.last
    for i, s in enumerate(sampled):
        assert torch.equal(s.proposal_boxes.tensor.cpu(), ref["sampled_boxes"][i]), f"sampled RoIs, image {i}"
        assert torch.equal(s.gt_classes.cpu(), ref["sampled_classes"][i]), f"sampled classes, image {i}"
        assert torch.equal(s.gt_boxes.tensor.cpu(), ref["sampled_gt"][i]), f"matched GT, image {i}"

    # ---- the step itself (fresh generator with the same seed -> the same sample)
    head.sampling_generator = seeded(1000)
    loss, grad_feat = stage.train_step(feats.to(dev), d_props, d_tgts, grad_pooled_fn=lambda pooled: gp_dev)
    assert _rel(loss, ref["loss"]) <= 1e-2, (loss.item(), ref["loss"].item())
    for name, param in (("cls_score_ft.weight", pred.cls_score_ft.weight), ("cls_score_ft.bias", pred.cls_score_ft.bias),
                        ("bbox_pred_ft.weight", pred.bbox_pred_ft.weight), ("bbox_pred_ft.bias", pred.bbox_pred_ft.bias)):
        assert _rel(param.grad, ref["grads"][name]) <= 1e-2, name
    mag = torch.ops.torchvision._roi_align_backward(gp.abs(), ref["rois"], 1 / 16, 14, 14, 2, 1024, 50, 84, 0, True)
    assert_close_rms(grad_feat.cpu(), ref_gfeat, 1e-5, "dL/dfeatures at bench shapes", magnitude=mag)

    # ---- forward pieces on the sampled RoIs: pooled (fp32 rel 1e-5), transferred scores / deltas (tf32 rel 1e-2)
    from unit_b200 import ops

    pooled = ops.roi_align_forward(feats.to(dev), ref["rois"].to(dev), (14, 14), 1 / 16, 0, True, True)
    assert_close_rms(pooled.cpu(), ref_pooled, 1e-5, "pooled at bench shapes")
    with torch.no_grad():
        sim = head.get_similarity_matrices(x_dev)
        (scores, bbox), _ = pred(x_dev, supervised_branch_x_weak=xw_dev, novel_classes=head._novel_classes_tensor,
                                 base_classes=head._base_classes_tensor, similarity=sim)
    assert _rel(scores, ref["scores"]) <= 1e-2
    assert _rel(bbox, ref["bbox"]) <= 1e-2

    # ---- and the graphed step bench.py actually times equals the eager one
    head.sampling_generator = seeded(1000)
    bucket.zero_()
    lg, gg = stage.train_step_graphed(feats.to(dev), d_props, d_tgts, grad_pooled_fn=lambda pooled: gp_dev)
    assert _rel(lg, ref["loss"]) <= 1e-2
    assert _rel(pred.cls_score_ft.weight.grad, ref["grads"]["cls_score_ft.weight"]) <= 1e-2
    assert_close_rms(gg.cpu(), ref_gfeat, 1e-5, "dL/dfeatures, graphed", magnitude=mag)


@pytest.mark.parametrize("graphed", [False, True])
def test_inference_bench_config_tf32_vs_oracle(bench, graphed):
    """RoIStage.infer / infer_graphed with the default TF32 predictor GEMM, 2 images x 512 proposals, 2048-d features
    (reference: roi_heads.py:487-591 inference branch -> fast_rcnn.py:455-468 -> fast_rcnn_inference)."""
    from oracle import unit_ref
    from unit_b200.stage import RoIStage
    from unit_b200.structures import Boxes, Instances

    dev = torch.device("cuda")
    head = bench.build_head(dev).eval()
    w, meta, x, xw, _ = bench.cpu_workload()
    x_dev, xw_dev = x.to(dev), xw.to(dev)
    stage = RoIStage(head, lambda pooled: (x_dev, xw_dev))
    feats, props, _, _ = bench.make_inputs(3000)
    boxes = [p[:512] for p in props]
    d_props = [Instances(bench.IMG_HW, proposal_boxes=Boxes(b.to(dev)), objectness_logits=torch.zeros(512, device=dev))
               for b in boxes]
    fn = stage.infer_graphed if graphed else stage.infer
    f_dev = feats.to(dev)
    insts, kept = fn(f_dev, d_props)
    if graphed:  # second call = a replay
        insts, kept = fn(f_dev, d_props)
    base, novel, idx = meta
    L = unit_ref.lingual_similarity(w["embeddings.weight"], idx, base, novel)
    V = unit_ref.visual_similarity(unit_ref.oicr_mean_logits(x, w), base, head.visual_threshold)
    sim = unit_ref.similarity_matrices(L, V, {k: list(v) for k, v in head.terms.items()}, len(novel), len(base))
    scores, bbox = unit_ref.predictor_forward(x, xw, w, sim, base, novel, bench.K_CLASSES, kind="FineTune", training=False)
    ref_inst, ref_kept = unit_ref.box_inference(scores, bbox, boxes, [bench.IMG_HW] * 2)
    for i in range(2):
        got = {(int(r), int(c)): (float(s), b) for r, c, s, b in zip(kept[i].cpu(), insts[i].pred_classes.cpu(),
                                                                    insts[i].scores.cpu(), insts[i].pred_boxes.tensor.cpu())}
        want = {(int(r), int(c)): (float(s), b) for r, c, s, b in zip(ref_kept[i], ref_inst[i].pred_classes,
                                                                     ref_inst[i].scores, ref_inst[i].pred_boxes.tensor)}
        common = set(got) & set(want)
        assert len(common) >= 0.95 * max(len(want), 1), (len(got), len(want), len(common))
        for key in common:
            assert abs(got[key][0] - want[key][0]) <= 1e-2 * max(want[key][0], 1e-3)
            assert (got[key][1] - want[key][1]).abs().max() <= 1e-2 * max(bench.IMG_HW)
