"""Probe of the mask paste kernel: device time (graph replay, 4 rotating canvases) for normal / tiny boxes and a plain memset."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from unit_b200 import _lib, ops
from unit_b200.ops import _ptr, _stream, check
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
    D = 100
    m = torch.rand(D, 28, 28, generator=g).cuda()
    boxes = random_boxes(D, 800, 1333, g, 24.0).cuda()
    tiny = boxes.clone(); tiny[:, 2] = tiny[:, 0] + 2; tiny[:, 3] = tiny[:, 1] + 2
    canv = [torch.empty(D, 800, 1333, dtype=torch.uint8, device=dev) for _ in range(4)]
    side = torch.cuda.Stream()
    out = {}

    def graph_of(fn):
        with torch.cuda.stream(side):
            fn(0); side.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                for k in range(8):
                    fn(k)
        return gr

    def paste(bx):
        return lambda k: check(_lib.lib().unit_mask_paste(_ptr(m), _ptr(bx), D, 28, 800, 1333, 0.5, _ptr(canv[k % 4]), _stream()), "paste")

    out["paste_ms"] = ev(graph_of(paste(boxes)).replay) / 8
    out["paste_tiny_boxes_ms"] = ev(graph_of(paste(tiny)).replay) / 8
    out["memset_ms"] = ev(graph_of(lambda k: canv[k % 4].zero_()).replay) / 8
    area = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).sum().item()
    out["window_pixels"] = area
    if "--once" in sys.argv:
        paste(boxes)(0); torch.cuda.synchronize()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
