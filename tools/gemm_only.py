"""Run only the packed predictor GEMM a few times (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from unit_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.randn(1024, 2048, generator=g).cuda()
w = torch.randn(202, 2048, generator=g).cuda()
b = torch.randn(202, generator=g).cuda()
for _ in range(3):
    y = ops.predictor_gemm_forward(x, w, b)
torch.cuda.synchronize()
