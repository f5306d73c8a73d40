"""Per-launch times of the inference tail (softmax+decode, filter, NMS) at COCO size; CUDA events, L2 flushed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from unit_b200 import ops, _lib

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(31)
out = {}
for name, n_img, per, K in (("voc", 2, 512, 20), ("coco", 2, 1000, 80), ("coco16", 16, 1000, 80)):
    R = n_img * per
    scores = torch.softmax(4.0 * torch.randn(R, K + 1, generator=g), -1).to(dev)
    deltas = (0.2 * torch.randn(R, 4 * K, generator=g)).to(dev)
    pb = torch.cat([bench._boxes(per, 800, 1333, g) for _ in range(n_img)]).to(dev)
    off = ops.offsets_from_counts([per] * n_img, dev)
    hw = torch.tensor([[800.0, 1333.0]] * n_img, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _, boxes = ops.softmax_decode(None, deltas, pb, want_probs=False)
    res = ops.detect(boxes, scores, off, hw, 0.05, 0.5, 100)
    n0 = _lib.launch_count()
    ops.detect(boxes, scores, off, hw, 0.05, 0.5, 100)
    t_dec = bench.time_kernel(lambda: ops.softmax_decode(None, deltas, pb, want_probs=False), 10, flush)
    t_det = bench.time_kernel(lambda: ops.detect(boxes, scores, off, hw, 0.05, 0.5, 100), 10, flush)
    out[name] = {"decode_ms": t_dec, "filter_nms_ms": t_det, "candidates": int(res[5][4].sum().item()),
                 "launches_detect": _lib.launch_count() - n0 - 0}
print(json.dumps(out))
