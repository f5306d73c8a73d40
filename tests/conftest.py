import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def seeded(seed):
    return torch.Generator().manual_seed(seed)


def random_boxes(n, h, w, g, min_size=8.0, frac=0.6):
    """Synthetic RPN-like boxes: centre uniform, side = min + frac*dim*u^2, clipped."""
    cx = torch.rand(n, generator=g) * w
    cy = torch.rand(n, generator=g) * h
    bw = min_size + frac * w * torch.rand(n, generator=g) ** 2
    bh = min_size + frac * h * torch.rand(n, generator=g) ** 2
    b = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, w)
    b[:, 1::2] = b[:, 1::2].clamp(0, h)
    return b


def assert_close_rms(actual, expected, rtol=1e-5, what="", magnitude=None):
    """|a-b| <= rtol * max(|ref|, rms(ref))  (SURVEY.md section 7 hard part 5: 'fp32 rel 1e-5' with an RMS floor).

    `magnitude` (optional, same shape): sum of |terms| behind each element of a scatter/accumulate result.  fp32
    summation in a different order can only be expected to agree to rtol * sum|terms| (not rtol * |sum|) where the
    terms cancel, so when given it replaces |ref| in the bound."""
    expected = expected.double()
    actual = actual.double().to(expected.device)
    rms = expected.pow(2).mean().sqrt().item() if expected.numel() else 0.0
    scale = expected.abs() if magnitude is None else torch.maximum(expected.abs(), magnitude.double().to(expected.device))
    tol = rtol * torch.clamp(scale, min=max(rms, 1e-30))
    err = (actual - expected).abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} outside tol; max err {err.max().item():.3e}, rms {rms:.3e}"
