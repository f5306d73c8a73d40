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
