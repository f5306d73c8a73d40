"""Run only the predictor GEMMs a few times (for ncu): the grouped forward (packed [delta | bbox | ft | mean-OICR] rows on x
+ mean-OICR rows on x_weak, VOC fine-tune shapes) and the MN-major weight-gradient GEMM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from unit_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(1024, 2048, generator=g)).cuda()
xw = torch.relu(torch.randn(1024, 2048, generator=g)).cuda()
w = (torch.randn(223, 2048, generator=g) * 0.05).cuda()
b = torch.randn(223, generator=g).cuda()
gy = torch.zeros(1024, 128)
gy[:, :101] = torch.randn(1024, 101, generator=g) * 0.01
gy = gy.cuda()
gw_c, gw_b = torch.zeros(21, 2048).cuda(), torch.zeros(80, 2048).cuda()
gb_c, gb_b = torch.zeros(21).cuda(), torch.zeros(80).cuda()
for _ in range(3):
    y1, y2 = ops.predictor_gemm2(x, w, b, xw, w[202:], b[202:])
    ops.predictor_wgrad(gy, x, 101, [0, 21, 101], [gw_c, gw_b], [gb_c, gb_b], [None, None], accumulate=False)
torch.cuda.synchronize()
