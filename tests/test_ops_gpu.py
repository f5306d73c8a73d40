"""Op-boundary parity on the GPU: every C-ABI kernel vs the CPU oracle on identical seeded tensors.

Bars (BASELINE.json north_star): bit-exact for matcher labels / sampled indices / NMS keep lists (torch.equal);
ROIAlign / decode / transfer within |a-b| <= 1e-5 * max(|ref|, rms(ref)).
"""
import math

import pytest
import torch
import torchvision

from conftest import assert_close_rms, load_golden, random_boxes, seeded

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from unit_b200 import ops as o

    return o


def _rois(n_img, per_img, h, w, g, min_size=16.0):
    boxes = [random_boxes(per_img, h, w, g, min_size) for _ in range(n_img)]
    rois = torch.cat([torch.cat([torch.full((per_img, 1), float(i)), b], 1) for i, b in enumerate(boxes)], 0)
    return rois


# ----------------------------------------------------------------------------------------------- ROIAlign
@pytest.mark.parametrize("shape", [(2, 64, 50, 84, 96), (1, 16, 25, 42, 40), (3, 8, 13, 17, 33)])
def test_roi_align_forward_slab_matches_torchvision(ops, shape):
    n, c, h, w, per = shape
    g = seeded(100 + c)
    feat = torch.randn(n, c, h, w, generator=g)
    rois = _rois(n, per, h * 16, w * 16, g)
    # RoIs that leave the map, degenerate and inverted boxes (SURVEY.md section 4 tier 3)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 100.0, 90.0])
    rois[1, 1:] = torch.tensor([w * 16 - 60.0, h * 16 - 50.0, w * 16 + 80.0, h * 16 + 70.0])
    rois[2, 1:] = torch.tensor([50.0, 60.0, 50.0, 200.0])
    rois[3, 1:] = torch.tensor([300.0, 300.0, 200.0, 100.0])
    rois[4, 1:] = torch.tensor([-500.0, -500.0, -300.0, -300.0])
    rois[5, 1:] = torch.tensor([0.0, 0.0, w * 16.0, h * 16.0])
    ref = torch.ops.torchvision.roi_align(feat, rois, 1.0 / 16, 14, 14, 0, True)
    out = ops.roi_align(feat.cuda(), rois.cuda(), 14, 1.0 / 16, 0, True, rois_sorted=True)
    torch.cuda.synchronize()
    assert_close_rms(out.cpu(), ref, 1e-5, "roi_align fwd (slab)")


def test_roi_align_forward_generic_paths(ops):
    g = seeded(7)
    feat = torch.randn(2, 6, 20, 31, generator=g)
    rois = _rois(2, 20, 320, 496, g)
    perm = torch.randperm(rois.shape[0], generator=g)
    rois = rois[perm]  # unsorted batch indices -> generic kernel
    for (ps, sr, aligned) in [(7, 0, True), (14, 2, False), (14, 0, True), (5, 3, True)]:
        ref = torch.ops.torchvision.roi_align(feat, rois, 1.0 / 16, ps, ps, sr, aligned)
        out = ops.roi_align(feat.cuda(), rois.cuda(), ps, 1.0 / 16, sr, aligned, rois_sorted=False)
        assert_close_rms(out.cpu(), ref, 1e-5, f"roi_align generic ps={ps} sr={sr} aligned={aligned}")


def test_roi_align_large_grid_and_sampling_ratio_in_slab(ops):
    # grid > MAXG (direct in-kernel path) and explicit sampling ratio inside the slab kernel
    g = seeded(9)
    feat = torch.randn(1, 8, 120, 100, generator=g)  # does not fit the slab -> generic, but keep the case
    rois = torch.tensor([[0, 0.0, 0.0, 1600.0, 1900.0], [0, 10.0, 10.0, 1500.0, 300.0]])
    ref = torch.ops.torchvision.roi_align(feat, rois, 1.0 / 16, 14, 14, 0, True)
    out = ops.roi_align(feat.cuda(), rois.cuda(), 14, 1.0 / 16, 0, True, rois_sorted=True)
    assert_close_rms(out.cpu(), ref, 1e-5, "roi_align big map")
    feat = torch.randn(1, 8, 40, 100, generator=g)  # fits; first roi has gw = 8 > MAXG
    rois = torch.tensor([[0, 0.0, 0.0, 1600.0, 600.0], [0, 10.0, 10.0, 150.0, 300.0], [0, 5.0, 5.0, 900.0, 630.0]])
    for sr in (0, 2, 7):
        ref = torch.ops.torchvision.roi_align(feat, rois, 1.0 / 16, 14, 14, sr, True)
        out = ops.roi_align(feat.cuda(), rois.cuda(), 14, 1.0 / 16, sr, True, rois_sorted=True)
        assert_close_rms(out.cpu(), ref, 1e-5, f"roi_align slab sr={sr}")


def test_roi_align_bf16_io(ops):
    g = seeded(11)
    feat = torch.randn(2, 16, 50, 84, generator=g).bfloat16()
    rois = _rois(2, 64, 800, 1333, g)
    ref = torch.ops.torchvision.roi_align(feat.float(), rois, 1.0 / 16, 14, 14, 0, True)
    out = ops.roi_align(feat.cuda(), rois.cuda(), 14, 1.0 / 16, 0, True, rois_sorted=True)
    assert out.dtype == torch.bfloat16
    # fp32 accumulation, one bf16 rounding of the output: 2^-8 relative
    err = (out.float().cpu() - ref).abs()
    assert (err <= 2 ** -8 * ref.abs() + 1e-6).all(), err.max()


@pytest.mark.parametrize("shape", [(2, 64, 50, 84, 96, 0), (1, 128, 20, 30, 40, 2), (3, 192, 7, 9, 10, 0)])
def test_roi_align_backward_channel_lane_bf16(ops, shape):
    """bf16 grad_out / grad_feat through the channel-lane backward (pair-view TMA tiles): fp32 accumulation, ONE bf16
    rounding of the result."""
    n, c, h, w, per, sr = shape
    g = seeded(350 + c + h)
    rois = _rois(n, per, h * 16, w * 16, g)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 100.0, 90.0])
    rois[1, 1:] = torch.tensor([0.0, 0.0, w * 16.0, h * 16.0])                 # whole image -> direct path
    rois[2, 1:] = torch.tensor([33.0, 47.0, 33.0, 47.0])                       # zero-size
    rois[3, 1:] = torch.tensor([-500.0, -500.0, -100.0, -100.0])               # entirely outside
    rois = rois[torch.randperm(rois.shape[0], generator=g)]
    gout = torch.randn(rois.shape[0], c, 14, 14, generator=g).bfloat16()
    ref = torch.ops.torchvision._roi_align_backward(gout.float(), rois, 1.0 / 16, 14, 14, n, c, h, w, sr, True)
    mag = torch.ops.torchvision._roi_align_backward(gout.float().abs(), rois, 1.0 / 16, 14, 14, n, c, h, w, sr, True)
    got = ops.roi_align_backward(gout.cuda(), rois.cuda(), (n, c, h, w), 1.0 / 16, sr, True, False)
    assert got.dtype == torch.bfloat16
    err = (got.float().cpu() - ref).abs()
    assert (err <= 2 ** -8 * ref.abs() + 1e-5 * mag + 1e-30).all(), err.max()
    assert ref.abs().max() > 0


@pytest.mark.parametrize("shape", [(2, 32, 50, 84, 64), (1, 8, 13, 17, 21)])
def test_roi_align_backward_matches_torchvision(ops, shape):
    n, c, h, w, per = shape
    g = seeded(200 + c)
    rois = _rois(n, per, h * 16, w * 16, g)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 100.0, 90.0])
    gout = torch.randn(rois.shape[0], c, 14, 14, generator=g)
    ref = torch.ops.torchvision._roi_align_backward(gout, rois, 1.0 / 16, 14, 14, n, c, h, w, 0, True)
    feat = torch.zeros(n, c, h, w, device="cuda", requires_grad=True)
    out = ops.roi_align(feat, rois.cuda(), 14, 1.0 / 16, 0, True, rois_sorted=True)
    out.backward(gout.cuda())
    assert_close_rms(feat.grad.cpu(), ref, 1e-5, "roi_align bwd (slab)")
    # generic (unsorted) backward
    perm = torch.randperm(rois.shape[0], generator=g)
    feat2 = torch.zeros(n, c, h, w, device="cuda", requires_grad=True)
    out2 = ops.roi_align(feat2, rois[perm].cuda(), 14, 1.0 / 16, 0, True, rois_sorted=False)
    out2.backward(gout[perm].cuda())
    assert_close_rms(feat2.grad.cpu(), ref, 1e-5, "roi_align bwd (generic)")


@pytest.mark.parametrize("shape", [(2, 64, 50, 84, 96, 0), (1, 128, 20, 30, 40, 0), (2, 64, 50, 84, 48, 2),
                                   (1, 64, 120, 200, 24, 0), (3, 192, 7, 9, 10, 0)])
def test_roi_align_backward_channel_lane(ops, shape):
    """C % 64 == 0 takes the channel-lane backward (roi_align_bwd_cl.cu): adaptive and fixed sampling grids, RoIs
    hanging over every image border, degenerate and whole-image RoIs (grid > 6 -> direct path), unsorted RoIs."""
    n, c, h, w, per, sr = shape
    g = seeded(300 + c + h)
    rois = _rois(n, per, h * 16, w * 16, g)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 100.0, 90.0])
    rois[1, 1:] = torch.tensor([0.0, 0.0, w * 16.0, h * 16.0])                 # whole image
    rois[2, 1:] = torch.tensor([w * 16.0 - 20, h * 16.0 - 20, w * 16.0 + 60, h * 16.0 + 50])
    rois[3, 1:] = torch.tensor([33.0, 47.0, 33.0, 47.0])                       # zero-size
    rois[4, 1:] = torch.tensor([-500.0, -500.0, -100.0, -100.0])               # entirely outside
    rois[5, 1:] = torch.tensor([5.0, 5.0, 9.0, w * 4.0])                        # thin
    rois = rois[torch.randperm(rois.shape[0], generator=g)]
    gout = torch.randn(rois.shape[0], c, 14, 14, generator=g)
    ref = torch.ops.torchvision._roi_align_backward(gout, rois, 1.0 / 16, 14, 14, n, c, h, w, sr, True)
    mag = torch.ops.torchvision._roi_align_backward(gout.abs(), rois, 1.0 / 16, 14, 14, n, c, h, w, sr, True)
    got = ops.roi_align_backward(gout.cuda(), rois.cuda(), (n, c, h, w), 1.0 / 16, sr, True, False)
    assert_close_rms(got.cpu(), ref, 1e-5, "roi_align bwd (channel-lane)", magnitude=mag)


@pytest.mark.parametrize("shape", [(2, 64, 50, 84, 96), (1, 128, 20, 30, 40), (3, 8, 7, 9, 10), (2, 16, 64, 96, 50)])
def test_roi_align_backward_sorted_edge_rois(ops, shape):
    """Sorted RoIs through the autograd-facing entry point: RoIs hanging over every border, zero-size, outside, thin,
    whole-image, and one unclipped giant RoI (sampling grid 27 > 6 -> the direct path of the channel-lane kernel;
    C = 8 / 16 -> the column-owner kernel)."""
    n, c, h, w, per = shape
    g = seeded(400 + c + h)
    rois = _rois(n, per, h * 16, w * 16, g)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 100.0, 90.0])
    rois[1, 1:] = torch.tensor([0.0, 0.0, w * 16.0, h * 16.0])                 # whole image
    rois[2, 1:] = torch.tensor([w * 16.0 - 20, h * 16.0 - 20, w * 16.0 + 60, h * 16.0 + 50])
    rois[3, 1:] = torch.tensor([33.0, 47.0, 33.0, 47.0])                       # zero-size
    rois[4, 1:] = torch.tensor([-500.0, -500.0, -100.0, -100.0])               # entirely outside
    rois[5, 1:] = torch.tensor([5.0, 5.0, 9.0, w * 4.0])                        # thin
    rois[6, 1:] = torch.tensor([-2000.0, -1500.0, 4000.0, 3000.0])             # unclipped giant: fixup path
    gout = torch.randn(rois.shape[0], c, 14, 14, generator=g)
    ref = torch.ops.torchvision._roi_align_backward(gout, rois, 1.0 / 16, 14, 14, n, c, h, w, 0, True)
    mag = torch.ops.torchvision._roi_align_backward(gout.abs(), rois, 1.0 / 16, 14, 14, n, c, h, w, 0, True)
    got = ops.roi_align_backward(gout.cuda(), rois.cuda(), (n, c, h, w), 1.0 / 16, 0, True, True)
    assert_close_rms(got.cpu(), ref, 1e-5, "roi_align bwd (sorted, edge RoIs)", magnitude=mag)


def test_roi_align_linearity_full_size(ops):
    """Size-independent property at BASELINE.json's full size: ROIAlign is linear in the feature map, and
    <roi_align(f), g> == <f, roi_align_bwd(g)> (adjointness of forward and backward)."""
    g = seeded(5)
    feat = torch.randn(2, 1024, 50, 84, generator=g).cuda()
    feat2 = torch.randn(2, 1024, 50, 84, generator=g).cuda()
    rois = _rois(2, 512, 800, 1333, g).cuda()
    a = ops.roi_align(feat, rois, 14, 1 / 16, 0, True, True)
    b = ops.roi_align(feat2, rois, 14, 1 / 16, 0, True, True)
    ab = ops.roi_align(feat + 2 * feat2, rois, 14, 1 / 16, 0, True, True)
    assert_close_rms(ab.cpu(), (a + 2 * b).cpu(), 2e-5, "linearity")
    gout = torch.randn(a.shape, generator=g).cuda()
    gin = ops.roi_align_backward(gout, rois, feat.shape, 1 / 16, 0, True, True)
    lhs = (a.double() * gout.double()).sum().item()
    rhs = (feat.double() * gin.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


# ----------------------------------------------------------------------------------------------- IoU / matcher
def test_pairwise_iou_bit_exact(ops):
    from oracle.d2.structures import Boxes, pairwise_iou

    g = seeded(21)
    gt = random_boxes(9, 800, 1333, g, 32.0)
    pr = random_boxes(2000, 800, 1333, g, 16.0)
    pr[:50] = gt[torch.randint(0, 9, (50,), generator=g)]                      # IoU == 1
    pr[50:100] = (pr[50:100] / 16).round() * 16                                  # coarse grid -> many exact ties
    gt[0] = torch.tensor([0.0, 0.0, 64.0, 64.0])
    pr[100] = torch.tensor([32.0, 0.0, 96.0, 64.0])                              # IoU exactly 1/3
    pr[101] = torch.tensor([64.0, 64.0, 128.0, 128.0])                           # touching -> 0
    pr[102] = torch.tensor([10.0, 10.0, 10.0, 40.0])                             # zero area
    ref = pairwise_iou(Boxes(gt), Boxes(pr))
    out = ops.pairwise_iou(gt.cuda(), pr.cuda())
    assert torch.equal(out.cpu(), ref)
    assert ops.pairwise_iou(gt[:0].cuda(), pr.cuda()).shape == (0, 2000)


def test_matcher_golden_and_random(ops):
    gold = load_golden("matcher.pt")
    m, l, v = ops.matcher(gold["kat_iou"].cuda(), [0.5], [0, 1])
    for got, want in zip((m, l, v), gold["kat_default"]):
        assert torch.equal(got.cpu(), want)
    m, l, v = ops.matcher(torch.zeros(0, 4).cuda(), [0.5], [0, 1])
    for got, want in zip((m, l, v), gold["kat_empty"]):
        assert torch.equal(got.cpu(), want)
    m, l, v = ops.matcher(gold["kat_iou"].cuda(), [0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    for got, want in zip((m, l, v), gold["kat_lowq"]):
        assert torch.equal(got.cpu(), want)
    m, l, v = ops.matcher(gold["rand_iou"].cuda(), [0.5], [0, 1])
    for got, want in zip((m, l, v), gold["rand_default"]):
        assert torch.equal(got.cpu(), want)
    m, l, v = ops.matcher(gold["rand_iou"].cuda(), [0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    for got, want in zip((m, l, v), gold["rand_lowq"]):
        assert torch.equal(got.cpu(), want)


def test_iou_match_fused_batched_bit_exact(ops):
    from oracle.d2.ops import MatcherWithVals
    from oracle.d2.structures import Boxes, pairwise_iou

    g = seeded(31)
    gts, props = [], []
    for n_gt in (5, 0, 1, 8):
        gt = random_boxes(n_gt, 800, 1333, g, 32.0)
        pr = random_boxes(1000 + n_gt, 800, 1333, g, 16.0)
        if n_gt:
            k = 250
            pr[:k] = gt[torch.randint(0, n_gt, (k,), generator=g)] * (1 + 0.1 * (torch.rand(k, 4, generator=g) - 0.5))
            pr[-n_gt:] = gt
        gts.append(gt)
        props.append(pr)
    dev = torch.device("cuda")
    go = ops.offsets_from_counts([len(x) for x in gts], dev)
    po = ops.offsets_from_counts([len(x) for x in props], dev)
    m, l, v = ops.iou_match(torch.cat(gts).cuda(), go, torch.cat(props).cuda(), po, [0.5], [0, 1])
    ref = MatcherWithVals([0.5], [0, 1])
    off = 0
    for gt, pr in zip(gts, props):
        rm, rl, rv = ref(pairwise_iou(Boxes(gt), Boxes(pr)))
        sl = slice(off, off + len(pr))
        assert torch.equal(m[sl].cpu(), rm) and torch.equal(l[sl].cpu(), rl) and torch.equal(v[sl].cpu(), rv)
        off += len(pr)


# ----------------------------------------------------------------------------------------------- decode
def test_softmax_decode(ops):
    from oracle.d2.ops import Box2BoxTransform

    g = seeded(41)
    R, K = 777, 20
    scores = torch.randn(R, K + 1, generator=g) * 3
    deltas = torch.randn(R, 4 * K, generator=g)
    deltas[:, 2::4] *= 6  # exercise the scale clamp
    props = random_boxes(R, 800, 1333, g)
    probs, boxes = ops.softmax_decode(scores.cuda(), deltas.cuda(), props.cuda())
    assert_close_rms(probs.cpu(), torch.softmax(scores, -1), 1e-5, "softmax")
    assert_close_rms(boxes.cpu(), Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).apply_deltas(deltas, props), 1e-5, "decode")
    tgt = random_boxes(R, 800, 1333, g)
    d = ops.box_get_deltas(props.cuda(), tgt.cuda())
    assert_close_rms(d.cpu(), Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).get_deltas(props, tgt), 1e-5, "get_deltas")


# ----------------------------------------------------------------------------------------------- NMS
def _tie_free_scores(n, g, lo=0.05):
    s = lo + (1 - lo) * torch.rand(n, generator=g)
    return (s + torch.arange(n) * 2.0 ** -24).float()


@pytest.mark.parametrize("n,k", [(1, 3), (33, 1), (900, 20), (3000, 20), (6000, 80)])
def test_batched_nms_bit_exact_both_regimes(ops, n, k):
    g = seeded(50 + n)
    boxes = random_boxes(n, 800, 1333, g)
    boxes[: n // 4] = (boxes[: n // 4] / 32).round() * 32  # many IoU ties incl. exactly 0.5
    scores = _tie_free_scores(n, g)
    idxs = torch.randint(0, k, (n,), generator=g)
    ref = torchvision.ops.batched_nms(boxes, scores, idxs, 0.5)  # CPU rule: coordinate trick iff 4n <= 4000
    out = ops.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5, nms_mode=ops.NMS_TV_CPU_RULE)
    assert torch.equal(out.cpu(), ref)
    # both explicit formulations against their torchvision counterparts
    ref_v = torchvision.ops.boxes._batched_nms_vanilla(boxes, scores, idxs, 0.5)
    ref_c = torchvision.ops.boxes._batched_nms_coordinate_trick(boxes, scores, idxs, 0.5)
    assert torch.equal(ops.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5, ops.NMS_CLASSWISE).cpu(), ref_v)
    assert torch.equal(ops.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5, ops.NMS_COORD_TRICK).cpu(), ref_c)


def test_plain_nms_and_ties(ops):
    g = seeded(61)
    boxes = random_boxes(500, 400, 400, g)
    scores = torch.rand(500, generator=g)
    scores[100:200] = 0.5  # equal scores: stable order (lower index first)
    ref = torchvision.ops.nms(boxes, scores, 0.5)
    assert torch.equal(ops.nms(boxes.cuda(), scores.cuda(), 0.5).cpu(), ref)
    # negative coordinates force the literal all-pairs coordinate trick
    b2 = boxes - 200
    idxs = torch.randint(0, 5, (500,), generator=g)
    s2 = _tie_free_scores(500, g)
    ref = torchvision.ops.boxes._batched_nms_coordinate_trick(b2, s2, idxs, 0.5)
    assert torch.equal(ops.batched_nms(b2.cuda(), s2.cuda(), idxs.cuda(), 0.5, ops.NMS_COORD_TRICK).cpu(), ref)
    assert ops.nms(boxes[:0].cuda(), scores[:0].cuda(), 0.5).numel() == 0


@pytest.mark.parametrize("K,R,thresh", [(20, 512, 0.05), (80, 1000, 0.05), (80, 1000, 0.003), (3, 700, 0.01),
                                          (2, 6000, 0.01)])
def test_fast_rcnn_inference_batched_bit_exact(ops, K, R, thresh):
    """Filter + class-wise NMS + top-k vs [D2] fast_rcnn_inference.  The grouped multi-CTA NMS takes the first two
    cases; thresh 0.003 leaves > 16384 candidates per image (segmented single-CTA kernel); K = 2 with 6000 proposals puts > 4096 candidates of one image in
    one class group (also the segmented kernel); K = 3 leaves most groups empty."""
    from oracle.d2.ops import fast_rcnn_inference

    g = seeded(70 + K)
    n_img = 3
    sizes = [(800, 1333), (600, 900), (750, 1000)]
    boxes, probs = [], []
    for (h, w) in sizes:
        base = random_boxes(R, h, w, g)
        b = base.repeat_interleave(K, 0).view(R, K, 4) + torch.randn(R, K, 4, generator=g) * 8
        b = torch.cat([torch.minimum(b[..., :2], b[..., 2:]), torch.maximum(b[..., :2], b[..., 2:])], -1)
        p = torch.softmax(torch.randn(R, K + 1, generator=g) * 2.5, -1)
        boxes.append(b.reshape(R, 4 * K).contiguous())
        probs.append(p)
    boxes[1][5, 3] = float("nan")       # non-finite rows are dropped and the returned roi index shifts
    probs[1][9, 2] = float("inf")
    ref_inst, ref_idx = fast_rcnn_inference(boxes, probs, sizes, thresh, 0.5, 100)
    dev = torch.device("cuda")
    off = ops.offsets_from_counts([R] * n_img, dev)
    hw = torch.tensor(sizes, dtype=torch.float32, device=dev)
    for mode in (ops.NMS_TV_CPU_RULE,):
        db, ds, dc, dr, cnt, _ = ops.detect(torch.cat(boxes).cuda(), torch.cat(probs).cuda(), off, hw, thresh, 0.5,
                                            100, nms_mode=mode)
        cnt = cnt.cpu().tolist()
        for i in range(n_img):
            n = cnt[i]
            assert n == len(ref_inst[i])
            assert torch.equal(dc[i, :n].cpu(), ref_inst[i].pred_classes)
            assert torch.equal(dr[i, :n].cpu(), ref_idx[i])
            assert torch.equal(ds[i, :n].cpu(), ref_inst[i].scores)
            assert torch.equal(db[i, :n].cpu(), ref_inst[i].pred_boxes.tensor)


# ----------------------------------------------------------------------------------------------- sampling
def test_label_and_sample_bit_exact(ops):
    from oracle.d2.ops import MatcherWithVals, subsample_labels
    from oracle.d2.structures import Boxes, pairwise_iou

    g = seeded(81)
    K = 20
    gts, gcls, props = [], [], []
    for n_gt in (4, 0, 7):
        gt = random_boxes(n_gt, 800, 1333, g, 32.0)
        pr = random_boxes(1000, 800, 1333, g, 16.0)
        if n_gt:
            pr[:300] = gt[torch.randint(0, n_gt, (300,), generator=g)] * (1 + 0.08 * (torch.rand(300, 4, generator=g) - .5))
        gts.append(gt)
        gcls.append(torch.randint(0, K, (n_gt,), generator=g))
        props.append(torch.cat([pr, gt]))
    dev = torch.device("cuda")
    go = ops.offsets_from_counts([len(x) for x in gts], dev)
    po = ops.offsets_from_counts([len(x) for x in props], dev)
    m, l, _ = ops.iou_match(torch.cat(gts).cuda(), go, torch.cat(props).cuda(), po, [0.5], [0, 1])
    pc, pos, neg, counts = ops.label_proposals(m, l, torch.cat(gcls).cuda(), go, po, K)
    counts_h = counts.cpu()
    gen_ref, gen = seeded(5), seeded(5)
    perm_pos, perm_neg, ppo, pno, pso, nso = [], [], [0], [0], [0], [0]
    for i in range(3):
        npos_all, nneg_all = int(counts_h[i, 0]), int(counts_h[i, 1])
        num_pos = min(npos_all, int(512 * 0.25))
        num_neg = min(nneg_all, 512 - num_pos)
        perm_pos.append(torch.randperm(npos_all, generator=gen))
        perm_neg.append(torch.randperm(nneg_all, generator=gen))
        ppo.append(ppo[-1] + npos_all)
        pno.append(pno[-1] + nneg_all)
        pso.append(pso[-1] + num_pos)
        nso.append(nso[-1] + num_neg)
    S = pso[-1] + nso[-1]
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    sampled, ob, oc, om, ogt = ops.sample_gather(pos, neg, torch.cat(perm_pos).cuda(), i32(ppo),
                                                 torch.cat(perm_neg).cuda(), i32(pno), i32(pso), i32(nso), po, go, S,
                                                 torch.cat(props).cuda(), pc, m, torch.cat(gts).cuda())
    # oracle: [D2] ROIHeads._sample_proposals with the same generator stream
    ref = MatcherWithVals([0.5], [0, 1])
    out_off = 0
    for i in range(3):
        rm, rl, _ = ref(pairwise_iou(Boxes(gts[i]), Boxes(props[i])))
        if len(gts[i]):
            cls = gcls[i][rm]
            cls[rl == 0] = K
            cls[rl == -1] = -1
        else:
            cls = torch.zeros_like(rm) + K
        p_idx, n_idx = subsample_labels(cls, 512, 0.25, K, generator=gen_ref)
        s_ref = torch.cat([p_idx, n_idx])
        n = len(s_ref)
        sl = slice(out_off, out_off + n)
        assert torch.equal(sampled[sl].cpu(), s_ref)
        assert torch.equal(oc[sl].cpu(), cls[s_ref])
        assert torch.equal(ob[sl].cpu(), props[i][s_ref])
        if len(gts[i]):
            assert torch.equal(ogt[sl].cpu(), gts[i][rm[s_ref]])
        else:
            assert torch.equal(ogt[sl].cpu(), torch.zeros(n, 4))
        out_off += n
    assert out_off == S


def test_label_and_sample_full_batch_64_images(ops):
    """BASELINE.json configs[2] size: 64 images x 1000 proposals (+GT) in ONE batched call of the public
    ``layers.label_and_sample`` vs the per-image [D2] sequence of the oracle with the same generator stream."""
    from oracle.d2.ops import MatcherWithVals, subsample_labels
    from oracle.d2.structures import Boxes as RBoxes, pairwise_iou
    from unit_b200 import layers
    from unit_b200.structures import Boxes, Instances

    g = seeded(640)
    K, n_img = 20, 64
    props, tgts, cpu = [], [], []
    for i in range(n_img):
        n_gt = int(torch.randint(0, 9, (1,), generator=g))
        gt = random_boxes(n_gt, 800, 1333, g, 32.0)
        pr = random_boxes(1000, 800, 1333, g, 16.0)
        if n_gt:
            k = 250
            src = gt[torch.randint(0, n_gt, (k,), generator=g)]
            wh = torch.cat([src[:, 2:] - src[:, :2]] * 2, 1)
            pr[:k] = (src + 0.2 * (torch.rand(k, 4, generator=g) - 0.5) * wh).clamp(min=0)
        gc = torch.randint(0, K, (n_gt,), generator=g)
        allp = torch.cat([pr, gt])  # [D2] add_ground_truth_to_proposals: GT appended after the proposals
        cpu.append((allp, gt, gc))
        props.append(Instances((800, 1333), proposal_boxes=Boxes(allp.cuda()),
                               objectness_logits=torch.zeros(len(allp), device="cuda")))
        tgts.append(Instances((800, 1333), gt_boxes=Boxes(gt.cuda()), gt_classes=gc.cuda()))
    out, _, _ = layers.label_and_sample(props, tgts, num_classes=K, batch_size_per_image=512, positive_fraction=0.25,
                                        thresholds=[0.5], labels=[0, 1], generator=seeded(7))
    gen = seeded(7)
    ref = MatcherWithVals([0.5], [0, 1])
    for i, (allp, gt, gc) in enumerate(cpu):
        rm, rl, _ = ref(pairwise_iou(RBoxes(gt), RBoxes(allp)))
        if len(gt):
            cls = gc[rm]
            cls[rl == 0] = K
            cls[rl == -1] = -1
        else:
            cls = torch.zeros_like(rm) + K
        p_idx, n_idx = subsample_labels(cls, 512, 0.25, K, generator=gen)
        s_ref = torch.cat([p_idx, n_idx])
        assert torch.equal(out[i].proposal_boxes.tensor.cpu(), allp[s_ref]), i
        assert torch.equal(out[i].gt_classes.cpu(), cls[s_ref]), i
        if len(gt):
            assert torch.equal(out[i].gt_boxes.tensor.cpu(), gt[rm[s_ref]]), i


# ----------------------------------------------------------------------------------------------- transfer
def _spec_from_golden(ops, gold, lingual_soft, dev):
    terms = gold["terms"]
    K = gold["num_classes"]
    base, novel = gold["base"].tolist(), gold["novel"].tolist()
    static, wv, norm = {}, {}, {}
    for head, tl in terms.items():
        w = 1.0 / len(tl) if len(tl) else 0.0
        static[head] = (w * lingual_soft) if "lingual" in tl else None
        wv[head] = w if "visual" in tl else 0.0
        norm[head] = 1 if (len(tl) > 0 and "None" not in tl) else 0
    return ops.TransferSpec(K, base, novel, dev, static, wv, norm, gold["threshold"])


@pytest.mark.parametrize("name", ["predictor_voc_base_eval.pt", "predictor_voc_ft_train.pt",
                                  "predictor_voc_ft_eval.pt", "predictor_coco_ft_eval.pt",
                                  "predictor_voc_base_train.pt"])
def test_transfer_against_reference_golden(ops, name):
    import torch.nn.functional as F

    gold = load_golden(name)
    emb = load_golden("glove_mean.pt")["embeddings"]
    dev = torch.device("cuda")
    raw, soft = ops.lingual_similarity(emb.cuda(), gold["indexer"].cuda(), gold["base"].cuda(), gold["novel"].cuda())
    lg = load_golden("lingual.pt")["voc" if gold["num_classes"] == 20 else "coco"]
    assert_close_rms(raw.cpu(), lg["lingual"], 1e-5, "lingual raw")
    spec = _spec_from_golden(ops, gold, soft, dev)
    w = {k: v.cuda() for k, v in gold["weights"].items()}
    x, xw = gold["x"].cuda(), gold["x_weak_branch"].cuda()
    oicr = lambda t: torch.stack([F.linear(t, w[f"weak_detector_head.oicr_predictors.{i}.weight"],
                                           w[f"weak_detector_head.oicr_predictors.{i}.bias"]) for i in range(3)]).mean(0)
    delta = F.linear(x, w["cls_score_delta.weight"], w["cls_score_delta.bias"])
    pd = F.linear(x, w["bbox_pred_delta.weight"], w["bbox_pred_delta.bias"])
    ft = "cls_score_ft.weight" in w
    fts = F.linear(x, w["cls_score_ft.weight"], w["cls_score_ft.bias"]) if ft else None
    ftd = F.linear(x, w["bbox_pred_ft.weight"], w["bbox_pred_ft.bias"]) if ft else None
    do_transfer = gold["similarity"] is not None
    neg_inf = gold["kind"] == "SupervisedDetectorOutputsBase" and gold["training"]
    scores, bbox, sims = ops.similarity_transfer_forward(spec, oicr(x), delta, pd, oicr(xw), fts, ftd, do_transfer,
                                                         neg_inf, ("cls", "bbox") if do_transfer else ())
    assert_close_rms(scores.cpu(), gold["scores"], 1e-5, "scores")
    assert_close_rms(bbox.cpu(), gold["bbox"], 1e-5, "bbox")
    if do_transfer:
        # S rows are probability vectors (scale 1).  softmax(L) amplifies the fp32 rounding of the 300-d embedding
        # dot products (|L| up to 34) to ~3e-5 relative IN THE ORACLE ITSELF, so the bar here is 1e-5 of the row scale;
        # the north_star outputs (scores, bbox) above are held to 1e-5 relative.
        for h in ("cls", "bbox"):
            err = (sims[h].cpu() - gold["similarity"][h]).abs().max().item()
            assert err <= 1e-5, f"S_{h}: max abs err {err:.3e}"


def test_transfer_backward_matches_autograd_of_oracle(ops):
    from oracle import unit_ref

    g = seeded(91)
    R, K = 40, 20
    base = [0, 1, 3, 4, 6, 7, 8, 10, 11, 12, 14, 15, 16, 18, 19]
    novel = [2, 5, 9, 13, 17]
    dev = torch.device("cuda")
    ling = torch.softmax(torch.randn(5, 15, generator=g), -1)
    vis_logits = torch.randn(R, K + 1, generator=g)
    spec = ops.TransferSpec(K, base, novel, dev, {"cls": 0.5 * ling, "bbox": 0.5 * ling}, {"cls": 0.5, "bbox": 0.5},
                            {"cls": 1, "bbox": 1}, 0.02)
    delta = torch.randn(R, K + 1, generator=g, requires_grad=True)
    pd = torch.randn(R, 4 * K, generator=g, requires_grad=True)
    gs, gb = torch.randn(R, K + 1, generator=g), torch.randn(R, 4 * K, generator=g)
    v = unit_ref.visual_similarity(vis_logits, torch.tensor(base), 0.02)
    sim = unit_ref.similarity_matrices(ling.log(), v, {"cls": ["lingual", "visual"], "bbox": ["lingual", "visual"]}, 5, 15)
    s_ref, b_ref = unit_ref.transfer(delta, pd, sim, torch.tensor(base), torch.tensor(novel), K)
    (s_ref * gs).sum().backward(retain_graph=True)
    (b_ref * gb).sum().backward()
    d2 = delta.detach().cuda().requires_grad_(True)
    p2 = pd.detach().cuda().requires_grad_(True)
    s, b = ops.similarity_transfer(spec, vis_logits.cuda(), d2, p2)
    assert_close_rms(s.detach().cpu(), s_ref.detach(), 1e-5, "fwd scores")
    ((s * gs.cuda()).sum() + (b * gb.cuda()).sum()).backward()
    assert_close_rms(d2.grad.cpu(), delta.grad, 1e-5, "grad delta_scores")
    assert_close_rms(p2.grad.cpu(), pd.grad, 1e-5, "grad proposal_deltas")


def test_transfer_reads_packed_gemm_outputs_in_place(ops):
    """Column blocks of packed GEMM outputs go in as (pointer, row stride); the fine-tune block as ONE packed tensor
    with one packed gradient.  Same numbers, bit for bit, as the dense / separate-tensor call."""
    g = seeded(92)
    R, K = 50, 20
    K1, K4 = K + 1, 4 * K
    base = [0, 1, 3, 4, 6, 7, 8, 10, 11, 12, 14, 15, 16, 18, 19]
    novel = [2, 5, 9, 13, 17]
    dev = torch.device("cuda")
    ling = torch.softmax(torch.randn(5, 15, generator=g), -1)
    spec = ops.TransferSpec(K, base, novel, dev, {"cls": 0.5 * ling, "bbox": 0.5 * ling}, {"cls": 0.5, "bbox": 0.5},
                            {"cls": 1, "bbox": 1}, 0.02)
    vis = torch.randn(R, K1, generator=g).cuda()
    y = torch.randn(R, K1 + K4 + 3, generator=g).cuda()       # [delta | bbox | padding]
    yf = torch.randn(R, K1 + K4, generator=g).cuda()
    ws = torch.randn(R, K1, generator=g).cuda()
    gs, gb = torch.randn(R, K1, generator=g).cuda(), torch.randn(R, K4, generator=g).cuda()
    for training in (False, True):
        d1 = y[:, :K1].clone().requires_grad_(True)
        p1 = y[:, K1:K1 + K4].clone().requires_grad_(True)
        f1 = yf.clone().requires_grad_(True)
        s1, b1 = ops.similarity_transfer(spec, vis, d1, p1, ws, f1[:, :K1], f1[:, K1:], True, training, False)
        (s1.nan_to_num(neginf=0.0) * gs).sum().backward(retain_graph=True)
        (b1 * gb).sum().backward()
        y2 = y.clone().requires_grad_(True)
        f2 = yf.clone().requires_grad_(True)
        s2, b2 = ops.similarity_transfer(spec, vis, y2[:, :K1], y2[:, K1:K1 + K4], ws, None, None, True, training,
                                         False, ft_packed=f2)
        ((s2.nan_to_num(neginf=0.0) * gs).sum() + (b2 * gb).sum()).backward()
        assert torch.equal(s1, s2) and torch.equal(b1, b2)
        assert torch.equal(f1.grad, f2.grad)
        assert torch.equal(d1.grad, y2.grad[:, :K1]) and torch.equal(p1.grad, y2.grad[:, K1:K1 + K4])
        assert (y2.grad[:, K1 + K4:] == 0).all()
    # frozen delta layers: no gradient is computed for them, the packed fine-tune gradient is unchanged
    f3 = yf.clone().requires_grad_(True)
    s3, b3 = ops.similarity_transfer(spec, vis, y[:, :K1], y[:, K1:K1 + K4], ws, None, None, True, True, False,
                                     ft_packed=f3)
    ((s3.nan_to_num(neginf=0.0) * gs).sum() + (b3 * gb).sum()).backward()
    assert torch.equal(f3.grad, f2.grad)


def test_append_gt_and_field_gather(ops):
    """[D2] add_ground_truth_to_proposals as one launch (unit_append_gt) and the objectness_logits gathered by the
    sampling launch: same tensors as the torch formulation, per-image results are views of one buffer."""
    import math

    from unit_b200 import layers
    from unit_b200.structures import Boxes, Instances

    g = seeded(41)
    props, gts, tgts = [], [], []
    for n, k in ((37, 3), (0, 2), (50, 0), (64, 5)):
        pb = random_boxes(n, 480, 640, g, 8.0)
        gb = random_boxes(k, 480, 640, g, 24.0)
        props.append(Instances((480, 640), proposal_boxes=Boxes(pb.cuda()),
                               objectness_logits=torch.randn(n, generator=g).cuda()))
        gts.append(Boxes(gb.cuda()))
        tgts.append(Instances((480, 640), gt_boxes=Boxes(gb.cuda()),
                              gt_classes=torch.randint(0, 20, (k,), generator=g).cuda()))
    out = layers.add_ground_truth_to_proposals(gts, props)
    logit = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    for o, p, gt in zip(out, props, gts):
        assert torch.equal(o.proposal_boxes.tensor, torch.cat([p.proposal_boxes.tensor, gt.tensor]))
        want = torch.cat([p.objectness_logits, torch.full((len(gt),), logit, device="cuda")])
        assert torch.equal(o.objectness_logits, want)
    allb = layers.cat([o.proposal_boxes.tensor for o in out])
    assert allb.data_ptr() == out[0].proposal_boxes.tensor.data_ptr() and allb.shape[0] == sum(len(o) for o in out)
    # sampled proposals carry the logits of the rows they were drawn from
    gen = torch.Generator().manual_seed(5)
    sampled, _, _ = layers.label_and_sample(out, tgts, num_classes=20, batch_size_per_image=32, positive_fraction=0.25,
                                            thresholds=[0.5], labels=[0, 1], generator=gen)
    for s_, o in zip(sampled, out):
        if len(s_) == 0:
            continue
        eq = (s_.proposal_boxes.tensor[:, None, :] == o.proposal_boxes.tensor[None, :, :]).all(-1)
        assert eq.any(1).all()
        src = eq.float().argmax(1)
        assert torch.equal(s_.objectness_logits, o.objectness_logits[src])


# ----------------------------------------------------------------------------------------------- masks
def test_mask_transfer_and_paste_golden(ops):
    gold = load_golden("mask_head.pt")
    dev = torch.device("cuda")
    spec = ops.TransferSpec(80, gold["base"].tolist(), gold["novel"].tolist(), dev)
    full, probs = ops.mask_transfer(gold["logits_fixed"].cuda(), gold["similarity_seg"].cuda(), spec,
                                    gold["logits_delta"].cuda(), gold["pred_classes"].cuda(), want_logits=True)
    assert_close_rms(probs.cpu(), gold["pred_masks"], 1e-5, "mask probs")
    from oracle import unit_ref

    ref_full = unit_ref.mask_transfer(gold["logits_fixed"], gold["similarity_seg"], gold["base"], gold["novel"],
                                      gold["logits_delta"])
    assert_close_rms(full.cpu(), ref_full, 1e-5, "mask logits")
    pasted = ops.mask_paste(gold["pred_masks"][:, 0].cuda(), gold["pred_boxes"].cuda(), gold["image_size"], 0.5)
    import numpy as np

    want = torch.from_numpy(np.unpackbits(gold["pasted_packed"].numpy(), axis=-1)[..., : gold["image_size"][1]]).bool()
    mism = (pasted.cpu() != want).sum().item()
    assert mism <= 1e-5 * want.numel(), f"{mism} mismatching mask pixels"
    assert torch.equal(pasted.cpu().sum(dim=(1, 2)), gold["pasted_sum"]) or mism > 0


def test_mask_paste_full_size(ops):
    from oracle.d2.ops import paste_masks_in_image

    g = seeded(95)
    D, M, H, W = 20, 28, 800, 1333
    masks = torch.rand(D, M, M, generator=g)
    boxes = random_boxes(D, H, W, g, 24.0)
    want = paste_masks_in_image(masks, boxes, (H, W), 0.5)
    got = ops.mask_paste(masks.cuda(), boxes.cuda(), (H, W), 0.5).cpu()
    mism = (got != want).sum().item()
    assert mism <= 1e-5 * want.numel(), f"{mism} mismatching pixels of {want.numel()}"


def _paste_cases():
    g = seeded(96)
    cases = []
    for (D, M, H, W) in ((9, 28, 61, 83), (6, 7, 29, 37), (5, 14, 40, 64)):
        masks = torch.rand(D, M, M, generator=g)
        boxes = random_boxes(D, H, W, g, 3.0)
        boxes[0] = torch.tensor([-7.5, -3.25, W + 9.0, H + 4.5])     # larger than the canvas
        boxes[1] = torch.tensor([W - 2.5, H - 1.75, W + 6.0, H + 8.0])  # mostly outside
        boxes[2] = torch.tensor([5.2, 6.1, 5.9, 6.6])                 # smaller than a pixel
        boxes[3] = torch.tensor([10.0, 4.0, 10.0, 9.0])               # empty (zero width): irregular path
        boxes[4] = torch.tensor([0.0, 0.0, float(W), float(H)])       # exactly the canvas
        cases.append((masks, boxes, H, W))
    return cases


def test_mask_paste_edge_cases(ops):
    """Separable single-pass paste (mask_paste_rows_kernel): odd widths (unaligned heads / tails of every strip), boxes
    outside / larger than / smaller than a pixel, an empty box, several mask sizes -- against the oracle's
    paste_masks_in_image, and byte-for-byte against the library's flat per-pixel kernel (UNIT_PASTE_FLAT=1, run in a
    second process because the switch is read once)."""
    import os
    import subprocess
    import sys
    import tempfile

    from oracle.d2.ops import paste_masks_in_image

    cases = _paste_cases()
    got = [ops.mask_paste(m.cuda(), b.cuda(), (H, W), 0.5).cpu() for m, b, H, W in cases]
    for (m, b, H, W), o in zip(cases, got):
        keep = [i for i in range(b.shape[0]) if i != 3]  # the oracle divides by the zero width
        want = paste_masks_in_image(m[keep], b[keep], (H, W), 0.5)
        mism = (o[keep] != want).sum().item()
        assert mism <= 1e-4 * want.numel(), f"{mism} mismatching pixels of {want.numel()} at {(H, W)}"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "flat.pt")
        code = ("import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r); from unit_b200 import ops; "
                "from test_ops_gpu import _paste_cases; "
                "torch.save([ops.mask_paste(m.cuda(), b.cuda(), (H, W), 0.5).cpu() for m, b, H, W in _paste_cases()], %r)"
                % (root, os.path.join(root, "tests"), path))
        subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, UNIT_PASTE_FLAT="1"), timeout=300)
        flat = torch.load(path)
    for o, f in zip(got, flat):
        assert torch.equal(o, f), "separable paste differs from the per-pixel kernel"


def test_fused_fastrcnn_loss_and_grads(ops):
    """[D2] FastRCNNOutputs.losses (CE mean + smooth-L1 on get_deltas / R) and its autograd, fused in one kernel."""
    import torch.nn.functional as F
    from oracle.d2.ops import Box2BoxTransform, smooth_l1_loss

    g = seeded(97)
    for K, beta in ((20, 0.0), (80, 0.5)):
        R = 333
        scores = (torch.randn(R, K + 1, generator=g) * 2).requires_grad_(True)
        deltas = torch.randn(R, 4 * K, generator=g).requires_grad_(True)
        props = random_boxes(R, 800, 1333, g, 16.0)
        gts = random_boxes(R, 800, 1333, g, 24.0)
        cls = torch.randint(0, K + 1, (R,), generator=g)
        loss_cls = F.cross_entropy(scores, cls)
        fg = ((cls >= 0) & (cls < K)).nonzero().squeeze(1)
        cols = 4 * cls[fg][:, None] + torch.arange(4)
        tgt = Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).get_deltas(props, gts)[fg]
        loss_box = smooth_l1_loss(deltas[fg[:, None], cols], tgt, beta, reduction="sum") / R
        (2.0 * loss_cls + 3.0 * loss_box).backward()
        s2 = scores.detach().cuda().requires_grad_(True)
        d2 = deltas.detach().cuda().requires_grad_(True)
        lc, lb = ops.fastrcnn_loss(s2, d2, props.cuda(), gts.cuda(), cls.cuda(), beta=beta)
        (2.0 * lc + 3.0 * lb).backward()
        assert_close_rms(lc.detach().cpu(), loss_cls.detach(), 1e-5, "loss_cls")
        assert_close_rms(lb.detach().cpu(), loss_box.detach(), 1e-5, "loss_box")
        assert_close_rms(s2.grad.cpu(), scores.grad, 1e-5, "d scores")
        assert_close_rms(d2.grad.cpu(), deltas.grad, 1e-5, "d deltas")


# ----------------------------------------------------------------------------------------------- torch.library layer
def test_opcheck_custom_ops(ops):
    """torch.library.opcheck on the registered ops: schema, fake implementation vs real outputs (shapes / dtypes /
    devices), autograd registration and AOT dispatch -- for roi_align, iou_match, similarity_transfer, detect,
    mask_paste and the predictor linear."""
    from torch.library import opcheck

    g = seeded(77)
    dev = "cuda"
    feat = torch.randn(2, 64, 25, 42, generator=g).to(dev).requires_grad_(True)
    rois = _rois(2, 24, 400, 672, g).to(dev)
    opcheck(torch.ops.unit_b200.roi_align, (feat, rois, 14, 14, 1 / 16, 0, True, True),
            test_utils=("test_schema", "test_faketensor", "test_autograd_registration", "test_aot_dispatch_static"))
    gout = torch.randn(48, 64, 14, 14, generator=g).to(dev)
    opcheck(torch.ops.unit_b200.roi_align_backward, (gout, rois, 2, 64, 25, 42, 1 / 16, 0, True, True),
            test_utils=("test_schema", "test_faketensor"))
    gt = random_boxes(5, 400, 672, g, 32.0).to(dev)
    pb = random_boxes(300, 400, 672, g, 16.0).to(dev)
    go, po = ops.offsets_from_counts([5], torch.device(dev)), ops.offsets_from_counts([300], torch.device(dev))
    opcheck(torch.ops.unit_b200.iou_match, (gt, go, pb, po, [0.5], [0, 1]), test_utils=("test_schema", "test_faketensor"))
    # similarity + transfer (VOC: 15 base + 5 novel)
    K, B, Nn, R = 20, 15, 5, 64
    base, novel = list(range(15)), list(range(15, 20))
    spec = ops.TransferSpec(K, base, novel, torch.device(dev),
                            static={"cls": torch.softmax(torch.randn(Nn, B, generator=g), -1).to(dev),
                                    "bbox": torch.softmax(torch.randn(Nn, B, generator=g), -1).to(dev)},
                            wv={"cls": 0.5, "bbox": 0.5}, norm={"cls": 1, "bbox": 1}, vis_threshold=0.02)
    vis = torch.randn(R, K + 1, generator=g).to(dev)
    ds = torch.randn(R, K + 1, generator=g).to(dev)
    pd = torch.randn(R, 4 * K, generator=g).to(dev)
    opcheck(torch.ops.unit_b200.similarity_transfer,
            (vis, spec.static["cls"], spec.static["bbox"], None, spec.base_i32, spec.novel_i32, spec.class_kind, ds, pd,
             None, None, None, 0.02, [0.5, 0.5, 0.0], [1, 1, 0], True, False, 0, 3),
            test_utils=("test_schema", "test_faketensor"))
    probs = torch.softmax(torch.randn(300, K + 1, generator=g) * 2, -1).to(dev)
    boxes = pb.repeat_interleave(K, 0).view(300, K * 4).contiguous()
    hw = torch.tensor([[400.0, 672.0]], device=dev)
    opcheck(torch.ops.unit_b200.detect, (boxes, probs, po, hw, 0.05, 0.5, 100, ops.NMS_TV_CUDA_RULE),
            test_utils=("test_schema", "test_faketensor"))
    masks = torch.rand(6, 28, 28, generator=g).to(dev)
    opcheck(torch.ops.unit_b200.mask_paste, (masks, pb[:6].contiguous(), 400, 672, 0.5),
            test_utils=("test_schema", "test_faketensor"))
    x = torch.randn(256, 128, generator=g).to(dev)
    w = (torch.randn(40, 128, generator=g) * 0.1).to(dev).requires_grad_(True)
    b = torch.zeros(40, device=dev, requires_grad=True)
    opcheck(torch.ops.unit_b200.predictor_linear, (x, w, b),
            test_utils=("test_schema", "test_faketensor", "test_autograd_registration", "test_aot_dispatch_static"))
