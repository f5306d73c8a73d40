"""Micro-benchmark: weak-image losses (MIL + 3 OICR refinements, forward + gradients) per call, CUDA-event timed,
next to the CPU oracle restatement on the same inputs.  python tools/micro_weak.py [n_img per_img K]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from unit_b200 import ops
from unit_b200._lib import launch_count
from conftest import random_boxes, seeded
from oracle import unit_ref


def main():
    n_img, per, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 2000, 20)
    g = seeded(1)
    counts = [per] * n_img
    R = per * n_img
    boxes = [random_boxes(per, 600, 800, g, 16.0) for _ in counts]
    cls_s, det_s = torch.randn(R, K, generator=g), torch.randn(R, K, generator=g)
    oicr = [torch.randn(R, K + 1, generator=g) for _ in range(3)]
    classes = [torch.randint(0, K, (3,), generator=g) for _ in counts]
    gt = torch.zeros(n_img, K)
    for i, c in enumerate(classes):
        gt[i, c] = 1
    dev = torch.device('cuda')
    off = ops.offsets_from_counts(counts, dev)
    d = dict(c=cls_s.cuda(), d=det_s.cuda(), o=[o.cuda() for o in oicr], b=torch.cat(boxes).cuda(), gt=gt.cuda())

    def step():
        c, dd = d['c'].requires_grad_(True), d['d'].requires_grad_(True)
        loss, probs, _ = ops.mil_loss(c, dd, off, d['gt'], 1.0, max_rows=per)
        total = loss
        for i in range(3):
            if i:
                probs, _ = ops.softmax_decode(d['o'][i - 1], None, None, want_boxes=False)
            lab, w, _, _ = ops.oicr_targets(probs, d['b'], off, d['gt'], [0.5], [0, 1], 0.1)
            total = total + ops.weighted_ce_loss(d['o'][i].requires_grad_(True), lab, w)
        return total

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    l0 = launch_count()
    ts = []
    for _ in range(30):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); step(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    launches = (launch_count() - l0) / 30
    ts.sort()
    t0 = time.perf_counter()
    for _ in range(3):
        unit_ref.weak_losses(cls_s, det_s, oicr, boxes, classes)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    print(json.dumps({"n_img": n_img, "proposals_per_image": per, "K": K, "gpu_ms_per_call": ts[len(ts) // 2],
                      "kernel_launches": launches, "cpu_oracle_ms_per_call": cpu_ms,
                      "cpu_threads": torch.get_num_threads()}))


if __name__ == '__main__':
    main()
