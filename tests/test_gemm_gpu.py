"""The tcgen05 predictor GEMM (TF32 multiply, fp32 accumulate) vs an fp64 reference: north_star bar "tf32 rel 1e-2"."""
import pytest
import torch

from conftest import assert_close_rms, seeded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(1024, 202, 2048), (128, 16, 32), (300, 101, 64), (1000, 482, 2048), (7, 21, 256),
                                   (1024, 21, 2048)])
def test_predictor_gemm_tf32(M, N, K):
    from unit_b200 import ops

    g = seeded(M + N + K)
    x = torch.relu(torch.randn(M, K, generator=g))
    w = torch.randn(N, K, generator=g) * 0.05
    b = torch.randn(N, generator=g)
    ref = (x.double() @ w.double().t() + b.double())
    y = ops.predictor_gemm_forward(x.cuda(), w.cuda(), b.cuda())
    torch.cuda.synchronize()
    scale = (x.double().abs() @ w.double().abs().t()).clamp(min=1e-6)   # |x|.|w|: the natural error scale of a dot
    err = ((y.double().cpu() - ref).abs() / scale).max().item()
    assert err < 2e-3, f"relative-to-|x||w| error {err:.3e}"          # TF32 has 10 mantissa bits: ~5e-4 expected
    rel = ((y.double().cpu() - ref).norm() / ref.norm()).item()
    assert rel < 1e-2, rel
    # determinism (fixed split-K summation order)
    y2 = ops.predictor_gemm_forward(x.cuda(), w.cuda(), b.cuda())
    assert torch.equal(y, y2)
    y3 = ops.predictor_gemm_forward(x.cuda(), w.cuda(), None)
    assert torch.allclose(y3 + b.cuda(), y, rtol=1e-6, atol=1e-6)


def test_linear_tf32_autograd():
    from unit_b200 import ops

    g = seeded(3)
    x = torch.randn(256, 128, generator=g).cuda()
    w = (torch.randn(40, 128, generator=g) * 0.1).cuda().requires_grad_(True)
    b = torch.zeros(40).cuda().requires_grad_(True)
    gy = torch.randn(256, 40, generator=g).cuda()
    y = ops.linear_tf32(x, w, b)
    y.backward(gy)
    # weight gradient: TF32 multiply / fp32 accumulate like the forward (north_star: tf32 transfer rel 1e-2)
    assert_close_rms(w.grad.cpu(), (gy.t() @ x).cpu(), 1e-2, "d/dW of linear_tf32")
    assert torch.allclose(b.grad, gy.sum(0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M,N1,N2,K", [(1024, 223, 21, 2048), (300, 101, 21, 256), (1000, 883, 81, 2048), (64, 8, 0, 64)])
def test_predictor_gemm2_grouped(M, N1, N2, K):
    """Two products in one launch, padded output rows (zeros past N): vs fp64."""
    from unit_b200 import ops

    g = seeded(M + N1 + N2)
    x1 = torch.relu(torch.randn(M, K, generator=g))
    w1 = torch.randn(N1, K, generator=g) * 0.05
    b1 = torch.randn(N1, generator=g)
    args = [x1.cuda(), w1.cuda(), b1.cuda()]
    if N2:
        x2 = torch.relu(torch.randn(M, K, generator=g))
        w2 = torch.randn(N2, K, generator=g) * 0.05
        b2 = torch.randn(N2, generator=g)
        args += [x2.cuda(), w2.cuda(), b2.cuda()]
    y1, y2 = ops.predictor_gemm2(*args)
    assert y1.shape == (M, (N1 + 31) // 32 * 32) and (y1[:, N1:] == 0).all()
    ref1 = x1.double() @ w1.double().t() + b1.double()
    assert ((y1[:, :N1].double().cpu() - ref1).norm() / ref1.norm()).item() < 1e-2
    scale = (x1.double().abs() @ w1.double().abs().t()).clamp(min=1e-6)
    assert ((y1[:, :N1].double().cpu() - ref1).abs() / scale).max().item() < 2e-3
    if N2:
        assert (y2[:, N2:] == 0).all()
        ref2 = x2.double() @ w2.double().t() + b2.double()
        assert ((y2[:, :N2].double().cpu() - ref2).norm() / ref2.norm()).item() < 1e-2
    else:
        assert y2 is None


@pytest.mark.parametrize("R,N,K,accumulate", [(1024, 101, 2048, False), (1000, 101, 2048, True), (77, 21, 64, False),
                                              (512, 401, 256, True)])
def test_predictor_wgrad_tcgen05(R, N, K, accumulate):
    """dW = gy^T x and db = colsum(gy) from MN-major tcgen05 operands, per-segment scaling, overwrite / accumulate,
    N > 128 (COCO: 401 gradient rows -> four launches) vs fp64."""
    from unit_b200 import ops

    g = seeded(R + N + K)
    ld = (N + 127) // 128 * 128
    gy = torch.zeros(R, ld)
    gy[:, :N] = torch.randn(R, N, generator=g) * 0.01
    x = torch.relu(torch.randn(R, K, generator=g))
    n_a = N // 5 + 1  # first segment: the class rows, second: the box rows
    s_a, s_b = torch.tensor([0.7]), torch.tensor([-1.3])
    base = torch.randn(N, K, generator=g) if accumulate else torch.zeros(N, K)
    base_b = torch.randn(N, generator=g) if accumulate else torch.zeros(N)
    wa, wb = base[:n_a].clone().cuda(), base[n_a:].clone().cuda()
    ba, bb = base_b[:n_a].clone().cuda(), base_b[n_a:].clone().cuda()
    ops.predictor_wgrad(gy.cuda(), x.cuda(), N, [0, n_a, N], [wa, wb], [ba, bb], [s_a.cuda(), s_b.cuda()], accumulate)
    torch.cuda.synchronize()
    ref = gy[:, :N].double().t() @ x.double()
    ref[:n_a] *= 0.7
    ref[n_a:] *= -1.3
    ref_b = gy[:, :N].double().sum(0)
    ref_b[:n_a] *= 0.7
    ref_b[n_a:] *= -1.3
    got = torch.cat([wa, wb]).double().cpu() - base.double()
    got_b = torch.cat([ba, bb]).double().cpu() - base_b.double()
    scale = (gy[:, :N].double().abs().t() @ x.double().abs()).clamp(min=1e-9)
    assert ((got - ref).abs() / scale).max().item() < 2e-3 + (1e-5 if accumulate else 0)
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2
    assert torch.allclose(got_b, ref_b, rtol=1e-4, atol=1e-5)


def test_tensor_map_entry_from_a_fresh_thread():
    """A thread whose FIRST CUDA action is a call into the library (an autograd worker running the fused node's
    backward with no gradient tensors to materialise): cuTensorMapEncodeTiled needs a driver context current on that
    thread, which ensure_driver_context() binds."""
    import threading

    from unit_b200 import ops

    g = seeded(5)
    R, K, N = 256, 512, 101
    x = torch.randn(R, K, generator=g).cuda()
    gy = torch.zeros(R, 128).cuda()
    gy[:, :N] = torch.randn(R, N, generator=g).cuda()
    w, b, sc = torch.zeros(N, K).cuda(), torch.zeros(N).cuda(), torch.ones(1).cuda()
    ops.predictor_wgrad(gy, x, N, [0, N], [w], [b], [sc], accumulate=False)  # main thread; sizes the workspace
    torch.cuda.synchronize()
    want = w.clone()
    w.zero_()
    torch.cuda.synchronize()
    err = []

    def run():
        try:
            ops.predictor_wgrad(gy, x, N, [0, N], [w], [b], [sc], accumulate=False)
        except Exception as e:  # noqa: BLE001
            err.append(e)

    t = threading.Thread(target=run)
    t.start()
    t.join()
    torch.cuda.synchronize()
    assert not err, err
    assert torch.equal(w, want)


def test_fused_ft_step_matches_modular():
    """ops.ft_step_losses (grouped GEMM -> transfer -> packed loss; backward = tcgen05 wgrad into bound .grad buffers)
    vs the modular predictor.forward + losses + autograd at the same TF32 precision."""
    import os
    from conftest import ROOT
    from unit_b200 import d2compat, ops  # noqa: F401
    from unit_b200.config import load_cfg
    from unit_b200.registry import ROI_BOX_HEAD_REGISTRY
    from unit_b200.roi_heads import build_roi_heads
    from unit_b200.structures import Boxes, Instances, ShapeSpec

    class _Feat(torch.nn.Module):
        def __init__(self, cfg, input_shape):
            super().__init__()

        @property
        def output_shape(self):
            return ShapeSpec(channels=256, height=1, width=1)

    if "FeatOnly256" not in ROI_BOX_HEAD_REGISTRY:
        ROI_BOX_HEAD_REGISTRY._do_register("FeatOnly256", _Feat)
    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_ft.yaml"),
                   ["MODEL.ROI_BOX_HEAD.NAME", "FeatOnly256", "MODEL.ROI_HEADS.EMBEDDING_PATH",
                    os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])

    def make():
        head = build_roi_heads(cfg, {"res4": ShapeSpec(channels=16, stride=16)})
        g = seeded(11)
        with torch.no_grad():
            for name, p in sorted(head.named_parameters()):
                if "embeddings" not in name:
                    p.copy_(torch.randn(p.shape, generator=g) * (0.02 if "bbox" in name else 0.2))
        return head.cuda().train()

    g = seeded(12)
    R, K = 600, 20
    x = torch.relu(torch.randn(R, 256, generator=g)).cuda()
    xw = torch.relu(torch.randn(R, 256, generator=g)).cuda()
    from conftest import random_boxes

    props = []
    for i in range(2):
        pb = random_boxes(R // 2, 800, 1333, g, 16.0)
        gt = pb + torch.randn(R // 2, 4, generator=g) * 3
        inst = Instances((800, 1333), proposal_boxes=Boxes(pb.cuda()), gt_boxes=Boxes(gt.cuda()),
                         gt_classes=torch.randint(0, K + 1, (R // 2,), generator=g).cuda())
        props.append(inst)
    fused, modular = make(), make()
    assert fused.box_predictor.can_fuse_losses(x, xw, fused._transfer_spec(x.device))
    # bound .grad buffers (what FlatGradBucket installs): the fused backward accumulates into them in place
    from unit_b200.distributed import FlatGradBucket

    bucket = FlatGradBucket([p for p in fused.parameters() if p.requires_grad])
    lf, _ = fused.box_losses(x, xw, props)
    assert abs(lf.total.item() - (lf["loss_cls"].item() + lf["loss_box_reg"].item())) < 1e-6  # third output of the launch
    (lf["loss_cls"] + 2.0 * lf["loss_box_reg"]).backward()
    modular.box_predictor.can_fuse_losses = lambda *a, **k: False
    lm, _ = modular.box_losses(x, xw, props)
    (lm["loss_cls"] + 2.0 * lm["loss_box_reg"]).backward()
    for k in ("loss_cls", "loss_box_reg"):
        assert abs(lf[k].item() - lm[k].item()) <= 1e-3 * max(abs(lm[k].item()), 1e-3), k
    pf, pm = fused.box_predictor, modular.box_predictor
    assert pf.cls_score_ft.weight.grad.data_ptr() == bucket.flat.data_ptr()
    for name in ("cls_score_ft", "bbox_pred_ft"):
        for part in ("weight", "bias"):
            a, b = getattr(getattr(pf, name), part).grad, getattr(getattr(pm, name), part).grad
            assert ((a - b).norm() / b.norm().clamp(min=1e-12)).item() < 1e-2, (name, part)
    # overwrite_bound_grads: the bucket holds garbage, the backward writes instead of accumulating (RoIStage's path)
    want = bucket.flat.clone()
    bucket.flat.fill_(123.0)
    lf4, _ = fused.box_losses(x, xw, props)
    one = torch.ones((), device="cuda")
    with ops.overwrite_bound_grads():
        torch.autograd.backward([lf4["loss_cls"], lf4["loss_box_reg"]], [one, 2.0 * one])
    assert torch.allclose(bucket.flat, want, rtol=1e-5, atol=1e-7)
    # no bound buffers: gradients come back through autograd as usual
    fused2 = make()
    lf2, _ = fused2.box_losses(x, xw, props)
    (lf2["loss_cls"] + 2.0 * lf2["loss_box_reg"]).backward()
    assert torch.allclose(fused2.box_predictor.cls_score_ft.weight.grad, pf.cls_score_ft.weight.grad, rtol=1e-5, atol=1e-8)
    # the parameters live in the packed matrix: an optimizer step is seen by the next forward
    with torch.no_grad():
        pf.cls_score_ft.weight.add_(0.5)
    lf3, _ = fused.box_losses(x, xw, props)
    assert abs(lf3["loss_cls"].item() - lf["loss_cls"].item()) > 1e-4
