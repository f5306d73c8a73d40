"""Micro-benchmark: mask transfer and mask paste at the COCO size (100 detections, 80 classes, 28x28 -> 800x1333)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from unit_b200 import ops
from conftest import random_boxes, seeded


def ev(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    g = seeded(3)
    dev = torch.device('cuda')
    D, K = 100, 80
    spec = ops.TransferSpec(K, list(range(60)), list(range(60, 80)), dev)
    logits = torch.randn(D, K, 28, 28, generator=g).cuda()
    sim = torch.softmax(torch.randn(D, 20, 60, generator=g), -1).cuda()
    cls = torch.randint(0, K, (D,), generator=g).cuda()
    boxes = random_boxes(D, 800, 1333, g, 24.0).cuda()
    _, probs = ops.mask_transfer(logits, sim, spec, None, cls)
    m = probs[:, 0].contiguous()
    out = {"mask_transfer_ms": ev(lambda: ops.mask_transfer(logits, sim, spec, None, cls)),
           "mask_paste_ms": ev(lambda: ops.mask_paste(m, boxes, (800, 1333), 0.5)),
           "paste_bytes_written": D * 800 * 1333}
    out["mask_paste_GBs"] = out["paste_bytes_written"] / out["mask_paste_ms"] / 1e6
    # device time without the host's per-call cost: the C-ABI call into 4 rotating canvases (4 x 107 MB > L2), 8 calls
    # captured in one CUDA graph, replayed
    from unit_b200 import _lib
    from unit_b200.ops import _ptr, _stream, check
    canv = [torch.empty(D, 800, 1333, dtype=torch.uint8, device=dev) for _ in range(4)]
    side = torch.cuda.Stream()
    def call(k):
        check(_lib.lib().unit_mask_paste(_ptr(m), _ptr(boxes), D, 28, 800, 1333, 0.5, _ptr(canv[k % 4]), _stream()), "unit_mask_paste")
    with torch.cuda.stream(side):
        call(0)
        side.synchronize()
        g8 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g8, stream=side):
            for k in range(8):
                call(k)
    out["mask_paste_device_ms"] = ev(g8.replay) / 8
    out["mask_paste_device_GBs"] = out["paste_bytes_written"] / out["mask_paste_device_ms"] / 1e6
    assert torch.equal(canv[0].bool(), ops.mask_paste(m, boxes, (800, 1333), 0.5))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
