#!/usr/bin/env python
"""bench.py -- RoI-stage images/sec (BASELINE.json metric) for the B200-native UniT RoI stage.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], per GPU; weak scaling -- every rank gets the same amount of work):
VOC R101-C4 split-1 10-shot FINE-TUNE RoI-head train step on 2 synthetic 800x1333 images:
  res4 features [2,1024,50,84]; 2000 RPN proposals + G GT boxes per image -> fused IoU+Matcher -> fg/bg sampling of
  512 RoIs/image -> ROIAlign forward [1024,1024,14,14] -> (res5 box head: OUT OF SCOPE, replaced by fixed synthetic
  [1024,2048] box features and a fixed synthetic dL/dpooled) -> packed predictor GEMM -> fused similarity + base->novel
  transfer -> CE + smooth-L1 -> backward to cls_score_ft / bbox_pred_ft -> ROIAlign backward -> (N>1) one NCCL
  all-reduce of the flat 0.83 MB gradient bucket.
One JSON line on stdout (rank 0).  `value` = images/sec with inputs resident in HBM; `e2e` = the same step called with
HOST (pinned) buffers: H2D of features/proposals/GT and D2H of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_IMG = 2           # images per GPU per step
P_RPN = 2000        # RPN proposals per image in training ([D2] POST_NMS_TOPK_TRAIN)
BATCH = 512         # sampled RoIs per image
C, H, W = 1024, 50, 84
IMG_HW = (800, 1333)
K_CLASSES = 20
FEAT_DIM = 2048
N_SETS = 4          # rotating input sets: 4 x 34.4 MB features (+ 822 MB of ROIAlign output per step) > 126 MB L2


def _seeded(seed):
    return torch.Generator().manual_seed(seed)


def _boxes(n, h, w, g, min_size=16.0):
    cx = torch.rand(n, generator=g) * w
    cy = torch.rand(n, generator=g) * h
    bw = min_size + 0.6 * w * torch.rand(n, generator=g) ** 2
    bh = min_size + 0.6 * h * torch.rand(n, generator=g) ** 2
    b = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, w)
    b[:, 1::2] = b[:, 1::2].clamp(0, h)
    return b


def make_inputs(seed, dtype=torch.float32):
    """Synthetic inputs of SURVEY.md section 8d (host tensors)."""
    g = _seeded(seed)
    feats = torch.randn(N_IMG, C, H, W, generator=g).to(dtype)
    props, gts, gcls = [], [], []
    for _ in range(N_IMG):
        n_gt = int(torch.randint(1, 9, (1,), generator=g))
        gt = _boxes(n_gt, IMG_HW[0], IMG_HW[1], g, 32.0)
        pb = _boxes(P_RPN, IMG_HW[0], IMG_HW[1], g, 16.0)
        k = P_RPN // 4  # 25 % of the proposals are GT boxes jittered by <= 10 % so that positives exist
        src = gt[torch.randint(0, n_gt, (k,), generator=g)]
        wh = torch.cat([src[:, 2:] - src[:, :2]] * 2, 1)
        pb[:k] = src + 0.2 * (torch.rand(k, 4, generator=g) - 0.5) * wh
        pb[:, 0::2] = pb[:, 0::2].clamp(0, IMG_HW[1])
        pb[:, 1::2] = pb[:, 1::2].clamp(0, IMG_HW[0])
        props.append(pb)
        gts.append(gt)
        gcls.append(torch.randint(0, K_CLASSES, (n_gt,), generator=g))
    return feats, props, gts, gcls


def build_head(device):
    from unit_b200 import d2compat  # noqa: F401
    from unit_b200.config import load_cfg
    from unit_b200.registry import ROI_BOX_HEAD_REGISTRY
    from unit_b200.roi_heads import build_roi_heads
    from unit_b200.structures import ShapeSpec

    class OutOfScopeBoxHead(torch.nn.Module):
        """Placeholder for Res5BoxHead (stock cuDNN convs, out of scope): contributes no parameters or work."""

        def __init__(self, cfg, input_shape):
            super().__init__()

        @property
        def output_shape(self):
            return ShapeSpec(channels=FEAT_DIM, height=1, width=1)

    if "OutOfScopeBoxHead" not in ROI_BOX_HEAD_REGISTRY:
        ROI_BOX_HEAD_REGISTRY._do_register("OutOfScopeBoxHead", OutOfScopeBoxHead)
    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_ft.yaml"),
                   ["MODEL.ROI_BOX_HEAD.NAME", "OutOfScopeBoxHead", "MODEL.ROI_HEADS.EMBEDDING_PATH",
                    os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])
    head = build_roi_heads(cfg, {"res4": ShapeSpec(channels=C, stride=16)})
    g = _seeded(4242)
    with torch.no_grad():  # the reference's inits x20 so the softmax is not flat; FT weights N(0, 0.01^2) not zeros
        for name, p in sorted(head.named_parameters()):
            if "embeddings" in name:
                continue
            std = 0.001 if ("bbox_pred_delta" in name and name.endswith("weight")) else 0.01
            if name.endswith("bias"):
                p.zero_()
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std * (1.0 if "_ft" in name else 20.0))
    return head.to(device).train()


class Workload:
    def __init__(self, device, dtype, rank, use_graph=True, prefetch=True, upload_bf16=True):
        from unit_b200.distributed import FlatGradBucket
        from unit_b200.stage import RoIStage
        from unit_b200.structures import Boxes, Instances

        self.device, self.dtype, self.use_graph = device, dtype, use_graph
        self.prefetch = bool(prefetch and use_graph and N_SETS > 1)  # labelling of step i+1 overlaps step i
        self._grad_pooled_fn = lambda pooled: self.grad_pooled
        self.Boxes, self.Instances = Boxes, Instances
        self.head = build_head(device)
        self.head.sampling_generator = _seeded(1000 + rank)
        self.bucket = FlatGradBucket([p for p in self.head.parameters() if p.requires_grad])
        g = _seeded(77 + rank)
        R = N_IMG * BATCH
        self.x = torch.relu(torch.randn(R, FEAT_DIM, generator=g)).to(device)
        self.xw = torch.relu(torch.randn(R, FEAT_DIM, generator=g)).to(device)
        self.grad_pooled = torch.randn(R, C, 14, 14, generator=g).to(dtype).to(device)
        self.stage = RoIStage(self.head, lambda pooled: (self.x, self.xw), self.bucket)
        self.host_sets = [make_inputs(2000 + 10 * rank + s, dtype) for s in range(N_SETS)]
        # e2e arm: the res4 features leave the host as bf16 (the dtype BASELINE.json configs[1] names; half the bytes of
        # fp32) and are widened to the step's dtype on the device, inside the timed region.  The synthetic features are
        # rounded to bf16 once, so the resident and the host-fed arm compute on identical values.
        self.upload_bf16 = bool(upload_bf16) and dtype == torch.float32
        if self.upload_bf16:
            self.host_sets = [(f.bfloat16().float(), pr, gt, gc) for (f, pr, gt, gc) in self.host_sets]
        self.pinned = [((f.bfloat16() if self.upload_bf16 else f).pin_memory(), [p.pin_memory() for p in pr],
                        [t.pin_memory() for t in gt], [c.pin_memory() for c in gc]) for (f, pr, gt, gc) in self.host_sets]
        self.dev_sets = [self._to_device(s, False) for s in self.host_sets]
        self._stage_bf16 = ([torch.empty((N_IMG, C, H, W), dtype=torch.bfloat16, device=device) for _ in range(N_SETS)]
                            if self.upload_bf16 else None)
        self._copy_stream = torch.cuda.Stream(device=device)
        self._copy_done = [torch.cuda.Event() for _ in range(N_SETS)]
        self._set_free = [torch.cuda.Event() for _ in range(N_SETS)]
        for e in self._set_free:
            e.record(torch.cuda.current_stream(device))
        self._loss_pinned = [torch.zeros(1).pin_memory() for _ in range(N_SETS)]
        self._loss_ready = [torch.cuda.Event() for _ in range(N_SETS)]
        self._e2e_n, self._issued, self._loss_pending = 0, -1, None

    def _to_device(self, s, non_blocking):
        f, pr, gt, gc = s
        dev = self.device
        feats = f.to(dev, non_blocking=non_blocking)
        props = [self.Instances(IMG_HW, proposal_boxes=self.Boxes(p.to(dev, non_blocking=non_blocking)),
                                objectness_logits=torch.zeros(len(p), device=dev)) for p in pr]
        tgts = [self.Instances(IMG_HW, gt_boxes=self.Boxes(t.to(dev, non_blocking=non_blocking)),
                               gt_classes=c.to(dev, non_blocking=non_blocking)) for t, c in zip(gt, gc)]
        return feats, props, tgts

    def _run(self, feats, props, tgts):
        fn = self.stage.train_step_graphed if self.use_graph else self.stage.train_step
        return fn(feats, props, tgts, grad_pooled_fn=self._grad_pooled_fn)

    def step(self, i):
        out = self._run(*self.dev_sets[i % N_SETS])
        if self.prefetch:
            self.stage.prefetch_labels(*self.dev_sets[(i + 1) % N_SETS])
        return out

    def _h2d(self, k):
        """Pinned host -> the (reused) device buffers of input set k, on the current stream."""
        f, pr, gt, gc = self.pinned[k]
        feats, props, tgts = self.dev_sets[k]
        if self.upload_bf16:
            self._stage_bf16[k].copy_(f, non_blocking=True)   # 17.2 MB over PCIe
            feats.copy_(self._stage_bf16[k])                  # widened on the device (same stream)
        else:
            feats.copy_(f, non_blocking=True)
        for p, src in zip(props, pr):
            p.proposal_boxes.tensor.copy_(src, non_blocking=True)
        for t, b, c in zip(tgts, gt, gc):
            t.gt_boxes.tensor.copy_(b, non_blocking=True)
            t.gt_classes.copy_(c, non_blocking=True)

    E2E_DEPTH = 2  # input sets in flight ahead of the step that computes (N_SETS = 4 buffers rotate)

    def _issue_inputs(self, j):
        """Host -> device copy of step j's inputs on the copy stream, then (graphs only) its labelling on the label
        stream -- both overlap the steps that are computing."""
        k = j % N_SETS
        self._copy_stream.wait_event(self._set_free[k])  # the last step that read set k has finished
        with torch.cuda.stream(self._copy_stream):
            self._h2d(k)
            self._copy_done[k].record(self._copy_stream)
        if self.prefetch:
            self.stage.prefetch_labels(*self.dev_sets[k], after=self._copy_done[k])
        self._issued = j

    def step_e2e(self, _i=None):
        """Same step from HOST buffers: every call copies one step's features / proposals / GT from pinned memory
        (34.5 MB) and reads one step's loss back.  As a data loader with a prefetch depth of two would, the copy of
        step j+2 is issued while step j computes, and the loss of step j is read (pinned, asynchronous copy) once
        step j+1 has been enqueued; ``e2e_finish`` reads the last one, inside the timed region."""
        j = self._e2e_n
        self._e2e_n += 1
        main = torch.cuda.current_stream(self.device)
        while self._issued < j + self.E2E_DEPTH:
            self._issue_inputs(self._issued + 1)
        k = j % N_SETS
        main.wait_event(self._copy_done[k])
        loss, _ = self._run(*self.dev_sets[k])
        self._set_free[k].record(main)
        self._loss_pinned[k].copy_(loss.detach().reshape(1), non_blocking=True)
        self._loss_ready[k].record(main)
        prev, self._loss_pending = self._loss_pending, k
        return self._read_loss(prev)

    def _read_loss(self, k):
        if k is None:
            return None
        self._loss_ready[k].synchronize()
        return float(self._loss_pinned[k].item())  # the device -> host read of a step's result

    def e2e_finish(self):
        prev, self._loss_pending = self._loss_pending, None
        return self._read_loss(prev)

    def h2d_bytes(self):
        f, pr, gt, gc = self.pinned[0]
        return int(f.numel() * f.element_size() + sum(p.numel() * 4 for p in pr) + sum(t.numel() * 4 for t in gt) +
                   sum(c.numel() * 8 for c in gc))


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [x.strip() for x in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as closest to its GPU, BEFORE any pinned host buffer is allocated: with
    one rank per GPU the 34.5 MB/step host->device copies then come from the GPU's own NUMA node (what a launcher
    with --cpu-bind does).  Returns the CPU count bound to, or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES may renumber the devices: go through the PCI address
            pr = torch.cuda.get_device_properties(index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(
                "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def time_kernel(fn, iters, flush):
    """Average device time (ms) of one launch on the current stream, L2 flushed between launches."""
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


# ------------------------------------------------------------------------------------------------- CPU reference
def cpu_reference_step(host_set, weights, head_meta, gen, x, xw, grad_pooled, roi_sample=None):
    """The reference's CPU path for the same step, restated in oracle/ (torchvision CPU ROIAlign kernels +
    Detectron2 glue + UniT transfer).  ``roi_sample`` bounds the ROIAlign part to that many RoIs per image."""
    from oracle import unit_ref
    from oracle.d2.ops import Box2BoxTransform, MatcherWithVals, subsample_labels
    from oracle.d2.structures import Boxes, pairwise_iou
    import torch.nn.functional as F

    feats, props, gts, gcls = host_set
    t0 = time.perf_counter()
    matcher = MatcherWithVals([0.5], [0, 1])
    s_boxes, s_cls, s_gt = [], [], []
    for pb, gt, gc in zip(props, gts, gcls):
        allp = torch.cat([pb, gt])
        m, l, _ = matcher(pairwise_iou(Boxes(gt), Boxes(allp)))
        cls = gc[m]
        cls[l == 0] = K_CLASSES
        pos, neg = subsample_labels(cls, BATCH, 0.25, K_CLASSES, generator=gen)
        idx = torch.cat([pos, neg])
        s_boxes.append(allp[idx])
        s_cls.append(cls[idx])
        s_gt.append(gt[m[idx]])
    t_label = time.perf_counter() - t0
    t0 = time.perf_counter()
    n_roi = BATCH if roi_sample is None else roi_sample
    rois = torch.cat([torch.cat([torch.full((n_roi, 1), float(i)), b[:n_roi]], 1) for i, b in enumerate(s_boxes)])
    f32 = feats.float()
    pooled = torch.ops.torchvision.roi_align(f32, rois, 1 / 16, 14, 14, 0, True)
    gfeat = torch.ops.torchvision._roi_align_backward(grad_pooled[: rois.shape[0]].float(), rois, 1 / 16, 14, 14,
                                                      N_IMG, C, H, W, 0, True)
    t_roi = time.perf_counter() - t0
    t0 = time.perf_counter()
    base, novel, idx = head_meta
    w = dict(weights)
    for k in ("cls_score_ft.weight", "cls_score_ft.bias", "bbox_pred_ft.weight", "bbox_pred_ft.bias"):
        w[k] = w[k].clone().requires_grad_(True)
    L = unit_ref.lingual_similarity(w["embeddings.weight"], idx, base, novel)
    V = unit_ref.visual_similarity(unit_ref.oicr_mean_logits(x, w), base, 0.02)
    sim = unit_ref.similarity_matrices(L, V, {"cls": ["lingual", "visual"], "bbox": ["lingual", "visual"]}, 5, 15)
    scores, bbox = unit_ref.predictor_forward(x, xw, w, sim, base, novel, K_CLASSES, kind="FineTune", training=True)
    gt_classes = torch.cat(s_cls)
    loss_cls = F.cross_entropy(scores, gt_classes)
    fg = ((gt_classes >= 0) & (gt_classes < K_CLASSES)).nonzero().squeeze(1)
    cols = 4 * gt_classes[fg][:, None] + torch.arange(4)
    tgt = Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).get_deltas(torch.cat(s_boxes), torch.cat(s_gt))[fg]
    loss = loss_cls + (bbox[fg[:, None], cols] - tgt).abs().sum() / gt_classes.numel()
    loss.backward()
    t_pred = time.perf_counter() - t0
    cpu_reference_step.last = {  # everything the parity test at bench shapes compares (tests/test_bench_config_gpu.py)
        "sampled_boxes": s_boxes, "sampled_classes": s_cls, "sampled_gt": s_gt, "rois": rois, "scores": scores.detach(),
        "bbox": bbox.detach(), "loss_cls": loss_cls.detach(), "loss": loss.detach(),
        "grads": {k: w[k].grad for k in ("cls_score_ft.weight", "cls_score_ft.bias", "bbox_pred_ft.weight",
                                         "bbox_pred_ft.bias")},
    }
    return float(loss.detach()), t_label, t_roi, t_pred, pooled, gfeat


def cpu_workload(rank=0):
    head = build_head(torch.device("cpu"))
    w = {k: v.detach().clone() for k, v in head.box_predictor.state_dict().items()}
    meta = (head._base_classes_tensor, head._novel_classes_tensor, head._coco_indexer_tensor)
    g = _seeded(77 + rank)
    R = N_IMG * BATCH
    x = torch.relu(torch.randn(R, FEAT_DIM, generator=g))
    xw = torch.relu(torch.randn(R, FEAT_DIM, generator=g))
    gp = torch.randn(R, C, 14, 14, generator=g)
    return w, meta, x, xw, gp


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: torchvision's compiled CPU
    ROIAlign kernels + the restated Detectron2 / UniT glue) on all host cores.  Every TIMED step is the FULL workload
    -- label/sample, ROIAlign forward + backward over all 2 x 512 RoIs, transfer + loss + backward -- measured by wall
    clock; nothing is sampled or scaled.  Warm-up steps run the same code on 32 RoIs per image (they are untimed and
    only page the libraries in), so that K = 20 timed steps (~9 s each on 16 cores) finish in about three minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w, meta, x, xw, gp = cpu_workload()
    sets = [make_inputs(2000 + s) for s in range(2)]
    gen = _seeded(1000)
    for i in range(max(args.warmup, 1)):
        cpu_reference_step(sets[i % 2], w, meta, gen, x, xw, gp, 32)
    per_step, parts = [], [0.0, 0.0, 0.0]
    t_wall = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        _, tl, tr, tp, _, _ = cpu_reference_step(sets[i % 2], w, meta, gen, x, xw, gp, None)
        per_step.append(time.perf_counter() - t0)
        parts = [parts[0] + tl, parts[1] + tr, parts[2] + tp]
    t_wall = time.perf_counter() - t_wall
    ms = 1000.0 * t_wall / args.steps
    value = N_IMG / (ms / 1000.0)
    sample = (f"every timed step is the full workload (2 images x 512 RoIs, C=1024): wall {t_wall:.1f}s for {args.steps} "
              f"steps; mean per step: label+sample {1e3 * parts[0] / args.steps:.1f} ms, ROIAlign fwd+bwd "
              f"{1e3 * parts[1] / args.steps:.0f} ms, transfer+loss+backward {1e3 * parts[2] / args.steps:.1f} ms; "
              f"min/max step {min(per_step):.2f}/{max(per_step):.2f} s; warm-up steps use 32 RoIs/image (untimed)")
    line = {
        "impl": "reference", "metric": "RoI-stage images/sec (VOC R101-C4 FT train step)", "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config("f32"),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(dtype):
    return {
        "workload": "BASELINE.json configs[1]: VOC-RCNN-101-C4-split1 10-shot fine-tune RoI-head train step, "
                    "2 synthetic 800x1333 images per GPU (16 images on 8 GPUs), res4 [2,1024,50,84], 2000 proposals "
                    "+ GT per image -> 512 sampled RoIs/image, 15 base + 5 novel classes",
        "images_per_gpu": N_IMG, "rois_per_image": BATCH, "proposals_per_image": P_RPN, "io_dtype": dtype,
        "box_head": "res5 excluded (stock PyTorch, out of scope): fixed synthetic [1024,2048] box features and "
                    "fixed synthetic dL/dpooled feed the in-scope kernels",
        "l2": f"inputs rotate over {N_SETS} sets; 137 MB of features + 822 MB ROIAlign output per step exceed the "
              "126 MB L2",
        "parallelism": "images sharded across GPUs, no data-path collective; one NCCL all-reduce of the flat "
                       "0.83 MB cls_score_ft/bbox_pred_ft gradient bucket per step",
    }



# ------------------------------------------------------------------------------------------------- comparison arms
def _tv_style_inference_gpu(boxes, scores, n_img, per_img, K, score_thresh, nms_thresh, topk):
    """What a stock Detectron2 does on the GPU for fast_rcnn_inference: a per-image Python loop of
    mask -> nonzero -> torchvision.ops.batched_nms -> top-k (comparison arm only, library kernels)."""
    import torchvision

    out = []
    for i in range(n_img):
        b = boxes[i * per_img:(i + 1) * per_img].view(per_img, K, 4)
        sc = scores[i * per_img:(i + 1) * per_img, :K]
        mask = sc > score_thresh
        idx = mask.nonzero()
        bb, ss = b[mask], sc[mask]
        keep = torchvision.ops.batched_nms(bb, ss, idx[:, 1], nms_thresh)[:topk]
        out.append((bb[keep], ss[keep], idx[keep]))
    return out


def aux_comparison(device, wl, flush, ops):
    """aux.torchvision_cuda + aux.nms_batch_sweep: the kernels the reference actually runs on a GPU (torchvision's CUDA
    roi_align / batched_nms, the bars SURVEY.md section 2a names) timed next to ours on the same box, and the
    SURVEY 8(d) honest-caveat sweeps for NMS.  Library kernels appear ONLY here, as the thing compared against."""
    import torchvision

    g = _seeded(909)
    res = {}
    # ---- ROIAlign forward / backward at the bench shapes
    feats = wl.dev_sets[0][0].float()
    rois = torch.cat([torch.cat([torch.full((BATCH, 1), float(i)), _boxes(BATCH, IMG_HW[0], IMG_HW[1], g)], 1)
                      for i in range(N_IMG)]).to(device)
    gout = wl.grad_pooled.float()
    fwd_tv = lambda: torchvision.ops.roi_align(feats, rois, 14, 1 / 16, 0, True)
    bwd_tv = lambda: torch.ops.torchvision._roi_align_backward(gout, rois, 1 / 16, 14, 14, N_IMG, C, H, W, 0, True)
    fwd = lambda: ops.roi_align_forward(feats, rois, (14, 14), 1 / 16, 0, True, True)
    bwd = lambda: ops.roi_align_backward(gout, rois, feats.shape, 1 / 16, 0, True, True)
    for f in (fwd_tv, bwd_tv, fwd, bwd):
        f()
    res["roi_align_2x512_c1024_f32"] = {
        "torchvision_fwd_ms": time_kernel(fwd_tv, 5, flush), "ours_fwd_ms": time_kernel(fwd, 10, flush),
        "torchvision_bwd_ms": time_kernel(bwd_tv, 5, flush), "ours_bwd_ms": time_kernel(bwd, 10, flush),
        "what": "torch.ops.torchvision.roi_align / _roi_align_backward (CUDA) vs unit_roi_align_fwd / _bwd, "
                "[2,1024,50,84] fp32, 2 x 512 RoIs, CUDA events, L2 flushed"}
    # ---- batched NMS, one call, Nc candidates over 80 classes (SURVEY 8d NMS-boundary micro-bench)
    sweep = {}
    for nc in (1000, 4000, 20000, 80000):
        bx = _boxes(nc, IMG_HW[0], IMG_HW[1], g).to(device)
        sc = (0.05 + 0.95 * torch.rand(nc, generator=g) + torch.arange(nc) * 2.0 ** -24).to(device)  # tie-free
        ids = torch.randint(0, 80, (nc,), generator=g).to(device)
        tv = lambda: torchvision.ops.batched_nms(bx, sc, ids, 0.5)
        ours = lambda: ops.batched_nms(bx, sc, ids, 0.5)
        k_tv, k_ours = tv(), ours()
        t_tv, t_ours = time_kernel(tv, 5, flush), time_kernel(ours, 5, flush)
        alg = 36 * nc
        sweep[str(nc)] = {"torchvision_ms": t_tv, "ours_ms": t_ours, "keep_equal": bool(torch.equal(k_tv, k_ours)),
                          "kept": int(k_ours.numel()), "algorithmic_bytes": alg,
                          "ours_GBps": alg / (t_ours * 1e-3) / 1e9}
    res["batched_nms_one_call"] = {
        "by_candidates": sweep,
        "what": "torchvision.ops.batched_nms (CUDA) vs unit_batched_nms, boxes as the proposals, scores U(0.05,1) "
                "tie-free, 80 classes, IoU 0.5; both end with the same host read of the keep count; algorithmic bytes "
                "= 36 B per candidate (SURVEY 8d) -- KB to a few MB, i.e. launch/latency-bound, not an HBM roofline"}
    # ---- decode + filter + class-wise NMS + top-100 for a BATCH of images in one call (SURVEY 8d honest caveat)
    batch = {}
    K, per = 80, 1000
    for n_img in (1, 2, 16, 64):
        R = n_img * per
        scores = torch.softmax(4.0 * torch.randn(R, K + 1, generator=g), -1).to(device)
        deltas = (0.2 * torch.randn(R, 4 * K, generator=g)).to(device)
        pb = torch.cat([_boxes(per, IMG_HW[0], IMG_HW[1], g) for _ in range(n_img)]).to(device)
        off = ops.offsets_from_counts([per] * n_img, device)
        hw = torch.tensor([[float(IMG_HW[0]), float(IMG_HW[1])]] * n_img, device=device)

        def ours():
            _, boxes = ops.softmax_decode(None, deltas, pb, want_probs=False)
            return ops.detect(boxes, scores, off, hw, 0.05, 0.5, 100)

        def stock():
            _, boxes = ops.softmax_decode(None, deltas, pb, want_probs=False)  # same decoded boxes for both arms
            return _tv_style_inference_gpu(boxes, scores, n_img, per, K, 0.05, 0.5, 100)

        out = ours()
        stock()
        n_cand = int(out[5][4].sum().item())
        t_ours, t_stock = time_kernel(ours, 5, flush), time_kernel(stock, 3, flush)
        alg = R * (K + 1) * 4 + R * 4 * K * 4 + R * 16 + 36 * n_cand + 36 * n_cand
        batch[str(n_img)] = {"ours_ms": t_ours, "torchvision_loop_ms": t_stock, "candidates": n_cand,
                             "algorithmic_bytes": alg, "ours_GBps": alg / (t_ours * 1e-3) / 1e9,
                             "images_per_s": n_img / (t_ours * 1e-3)}
    res["decode_filter_nms_by_batch"] = {
        "by_images": batch,
        "what": "softmax_decode + detect (filter, grouped class-wise NMS, top-100) for n images x 1000 proposals x 80 "
                "classes in ONE call sequence vs the stock per-image loop (mask, nonzero, torchvision batched_nms, "
                "top-k) on the same decoded boxes; CUDA events, L2 flushed; bytes per SURVEY 8d (scores + deltas + "
                "proposals read, 36 B per candidate written and re-read)"}
    return res


# ------------------------------------------------------------------------------------------------- main arm
def run_ours(args):
    import torch.distributed as dist

    from unit_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the "
                         "CPU baseline")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_to_gpu_numa_node(local_rank) if not args.no_affinity else None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cuda.matmul.allow_tf32 = False
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    # Nsight Compute serialises and replays every kernel, which is illegal while a stream is capturing: under a
    # profiler the step runs eagerly (same kernels, one launch each)
    profiled = any(k in os.environ for k in ("NV_COMPUTE_PROFILER_PERFWORKS_DIR", "CUDA_INJECTION64_PATH",
                                             "NV_NSIGHT_INJECTION_TRANSPORT_TYPE"))
    wl = Workload(device, dtype, rank, use_graph=not (args.no_graph or profiled), prefetch=not args.no_prefetch,
                  upload_bf16=not args.e2e_upload_f32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(N_SETS):  # first touch of every input set (graph capture happens here), never timed
        wl.step(i)
    for i in range(args.warmup):
        wl.step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + wl.stage.graph_launches
    total_ms = timed(wl.step, args.steps)
    launches = _lib.launch_count() + wl.stage.graph_launches - launches0
    for i in range(max(args.warmup // 2, 1)):
        wl.step_e2e(i)
    wl.e2e_finish()
    e2e_ms = timed(wl.step_e2e, args.steps, finish=wl.e2e_finish)
    clocks = sampler.stop() if rank == 0 else None

    # dominant kernels, timed alone with CUDA events on the launching stream (L2 flushed between launches)
    roofline = None
    kernel_ms = {}
    if rank == 0:
        feats, props, tgts = wl.dev_sets[0]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        g = _seeded(5)
        rois = torch.cat([torch.cat([torch.full((BATCH, 1), float(i)), _boxes(BATCH, IMG_HW[0], IMG_HW[1], g)], 1)
                          for i in range(N_IMG)]).to(device)
        es = 2 if dtype == torch.bfloat16 else 4
        alg_bytes = N_IMG * C * H * W * es + N_IMG * BATCH * 20 + N_IMG * BATCH * C * 196 * es
        fwd = lambda: ops.roi_align_forward(feats, rois, (14, 14), 1 / 16, 0, True, True)
        bwd = lambda: ops.roi_align_backward(wl.grad_pooled, rois, feats.shape, 1 / 16, 0, True, True)
        for _ in range(3):
            fwd()
            bwd()
        kernel_ms["roi_align_fwd"] = time_kernel(fwd, 10, flush)
        kernel_ms["roi_align_bwd"] = time_kernel(bwd, 10, flush)
        peak, peak_src = measured_peak_gbs()
        dom = max(kernel_ms, key=kernel_ms.get)
        achieved = alg_bytes / (kernel_ms[dom] * 1e-3) / 1e9
        traffic = None
        try:  # measured DRAM bytes per launch of that kernel (ncu --set full, see profiles/)
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                traffic = int(json.load(f)[dom]["dram_bytes"]) if dtype == torch.float32 else None
        except Exception:
            traffic = None
        roofline = {
            "kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "per_kernel": {k: {"ms": v, "GBps": alg_bytes / (v * 1e-3) / 1e9, "frac": alg_bytes / (v * 1e-3) / 1e9 / peak}
                           for k, v in kernel_ms.items()},
        }

    # auxiliary: the inference side of the stage (decode / filter / NMS are launch-latency bound; reported
    # per launch, not as a roofline claim).  configs[0]: VOC, 2 images x 512 proposals; configs[3]-sized decode + NMS.
    aux = None
    if rank == 0 and not args.no_aux:
        from unit_b200 import layers
        aux = {}
        with torch.no_grad():
            wl.head.eval()
            g = _seeded(31)
            f, pr, _, _ = make_inputs(3000)
            props = [wl.Instances(IMG_HW, proposal_boxes=wl.Boxes(p[:512].to(device)),
                                  objectness_logits=torch.zeros(512, device=device)) for p in pr]
            feats = f.to(device)
            xi, xwi = wl.x[:1024], wl.xw[:1024]
            infer_stage = type(wl.stage)(wl.head, lambda pooled: (xi, xwi))
            def wall_ms(fn):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(20):
                    fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / 20 * 1e3

            ms_eager = wall_ms(lambda: infer_stage.infer(feats, props))
            ms = ms_eager if not wl.use_graph else wall_ms(lambda: infer_stage.infer_graphed(feats, props))
            aux["voc_inference"] = {"images_per_s": N_IMG / (ms * 1e-3), "ms_per_call": ms,
                                    "ms_per_call_eager": ms_eager, "cuda_graph": bool(wl.use_graph),
                                    "what": "RoIStage.infer, 2 images x 512 proposals, K=20: ROIAlign fwd -> transfer "
                                            "-> softmax+decode -> filter -> class-wise NMS -> top-100 (wall clock, "
                                            "includes the one D2H of the detection counts)"}
            # COCO-sized decode + filter + NMS: 2 images x 1000 proposals x 80 classes
            R, K = 2000, 80
            scores = torch.softmax(4.0 * torch.randn(R, K + 1, generator=g), -1).to(device)
            deltas = (0.2 * torch.randn(R, 4 * K, generator=g)).to(device)
            pb = torch.cat([_boxes(1000, IMG_HW[0], IMG_HW[1], g) for _ in range(2)]).to(device)
            off = ops.offsets_from_counts([1000, 1000], device)
            hw = torch.tensor([[float(IMG_HW[0]), float(IMG_HW[1])]] * 2, device=device)

            def dec_nms():
                _, boxes = ops.softmax_decode(None, deltas, pb, want_probs=False)
                return ops.detect(boxes, scores, off, hw, 0.05, 0.5, 100)

            for _ in range(3):
                out = dec_nms()
            n_cand = int(out[5][4].sum().item())
            aux["coco_decode_filter_nms"] = {
                "ms_per_call": time_kernel(dec_nms, 10, flush), "candidates": n_cand,
                "what": "softmax_decode + detect (parallel filter, grouped NMS, top-100) for 2 images x 1000 proposals "
                        "x 80 classes, CUDA events, L2 flushed"}
            # BASELINE.json configs[2] size: Matcher + fg/bg sampling, 64 images x (1000 proposals + GT) in one call
            from unit_b200 import layers as _layers
            lp, lt = [], []
            for _ in range(64):
                gtb = _boxes(4, IMG_HW[0], IMG_HW[1], g, 32.0)
                pb64 = torch.cat([_boxes(1000, IMG_HW[0], IMG_HW[1], g), gtb])
                lp.append(wl.Instances(IMG_HW, proposal_boxes=wl.Boxes(pb64.to(device)),
                                       objectness_logits=torch.zeros(len(pb64), device=device)))
                lt.append(wl.Instances(IMG_HW, gt_boxes=wl.Boxes(gtb.to(device)),
                                       gt_classes=torch.randint(0, 20, (4,), generator=g).to(device)))
            gen64 = _seeded(64)
            ms64 = wall_ms(lambda: _layers.label_and_sample(lp, lt, num_classes=20, batch_size_per_image=512,
                                                            positive_fraction=0.25, thresholds=[0.5], labels=[0, 1],
                                                            generator=gen64))
            aux["label_sample_batch64"] = {
                "ms_per_call": ms64, "images_per_s": 64 / (ms64 * 1e-3),
                "what": "layers.label_and_sample, 64 images x 1004 proposals: fused IoU+match, label, host randperm "
                        "draw, gather (wall clock incl. the one D2H of the fg/bg counts and building 64 Instances)"}
            # BASELINE.json configs[4] size: transferred mask logits -> class select + sigmoid -> paste, 100 detections
            D, KM = 100, 80
            mspec = ops.TransferSpec(KM, list(range(60)), list(range(60, 80)), device)
            mlog = torch.randn(D, KM, 28, 28, generator=g).to(device)
            msim = torch.softmax(torch.randn(D, 20, 60, generator=g), -1).to(device)
            mcls = torch.randint(0, KM, (D,), generator=g).to(device)
            mbox = _boxes(D, IMG_HW[0], IMG_HW[1], g, 24.0).to(device)

            def mask_path():
                _, probs = ops.mask_transfer(mlog, msim, mspec, None, mcls)
                return ops.mask_paste(probs[:, 0], mbox, IMG_HW, 0.5)

            for _ in range(3):
                mask_path()
            aux["coco_mask_transfer_paste"] = {
                "ms_per_call": time_kernel(mask_path, 10, flush),
                "what": "mask_transfer (per-RoI similarity, 60 base -> 20 novel, class select + sigmoid) + mask_paste "
                        "of 100 detections into 800x1333, CUDA events, L2 flushed"}
            # weak-image branch of base training: MIL loss + 3 OICR refinements (targets + weighted CE), fwd + grads
            Rw, Kw = 2 * 2000, 20
            wc, wd_ = torch.randn(Rw, Kw, generator=g).to(device), torch.randn(Rw, Kw, generator=g).to(device)
            wo = [torch.randn(Rw, Kw + 1, generator=g).to(device) for _ in range(3)]
            wb = torch.cat([_boxes(2000, IMG_HW[0], IMG_HW[1], g) for _ in range(2)]).to(device)
            woff = ops.offsets_from_counts([2000, 2000], device)
            wgt = torch.zeros(2, Kw)
            wgt[0, [2, 7, 11]] = 1
            wgt[1, [5]] = 1
            wgt = wgt.to(device)

            def weak_path():
                loss, probs, _ = ops.mil_loss(wc, wd_, woff, wgt, 1.0, max_rows=2000)
                for i in range(3):
                    if i:
                        probs, _ = ops.softmax_decode(wo[i - 1], None, None, want_boxes=False)
                    lab, wgh, _, _ = ops.oicr_targets(probs, wb, woff, wgt, [0.5], [0, 1], 0.1)
                    loss = loss + ops.weighted_ce_loss(wo[i], lab, wgh)
                return loss

            for _ in range(3):
                weak_path()
            aux["weak_losses"] = {
                "ms_per_call": time_kernel(weak_path, 10, flush),
                "what": "MIL image loss + 3 x (OICR pseudo-labelling + weighted CE) with gradients, 2 images x 2000 "
                        "proposals, K=20: 15 launches, eager, CUDA events"}
            wl.head.train()
            aux["torchvision_cuda"] = aux_comparison(device, wl, flush, ops)

    # CPU baseline on the box's host cores (rank 0, N = 1 only): the oracle port on a bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core again
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        w, meta, x, xw, gp = cpu_workload()
        hs = make_inputs(2000)
        gen = _seeded(1000)
        cpu_reference_step(hs, w, meta, gen, x, xw, gp, 8)  # warm-up (8 RoIs/image: pages the libraries in)
        t0 = time.perf_counter()
        _, tl, tr, tp, _, _ = cpu_reference_step(hs, w, meta, gen, x, xw, gp, None)
        wall = time.perf_counter() - t0
        cpu_baseline = {
            "value": N_IMG / wall, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"ONE full step measured by wall clock, nothing scaled ({wall:.1f}s: label+sample {tl * 1e3:.1f} ms, "
                      f"ROIAlign fwd+bwd over all 2x512 RoIs {tr:.2f}s, transfer+loss+backward {tp * 1e3:.1f} ms); "
                      f"`bench.py --impl reference` times K such steps",
        }

    if rank == 0:
        ms_step = total_ms / args.steps
        value = world * N_IMG / (ms_step / 1000.0)
        e2e_value = world * N_IMG / (e2e_ms / args.steps / 1000.0)
        line = {
            "metric": "RoI-stage images/sec (VOC R101-C4 FT train step)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": workload_config(args.dtype),
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": 4 + 4 * N_IMG * 2,
                    "upload_dtype": "bf16 features (BASELINE configs[1] dtype), widened to the step dtype on the device "
                                    "inside the timed region" if wl.upload_bf16 else args.dtype,
                    "pipeline": "inputs of step j+2 are copied (pinned host -> device, side stream) while step j "
                                "computes; the loss of step j is read after step j+1 is enqueued; K copies and K "
                                "loss reads inside the K timed steps"},
            "gpu_launches": int(launches),
            "gpu_launches_note": "kernels of libunit_b200.so executed in the timed region, directly or as nodes of "
                                 "the replayed CUDA graphs (cuBLAS GEMMs and ATen kernels not counted)",
            "cuda_graphs": bool(wl.use_graph), "label_prefetch": bool(wl.prefetch), "cpu_affinity": affinity,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "aux": aux,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- inference arm
def run_infer(args):
    """--mode infer: the inference side of the stage, weak scaling over the GPUs of one box, ending in the gather of
    the detections the reference's evaluator performs (data/evaluators.py:159 -> comm.gather): every rank runs
    RoIStage.infer_graphed on its own 2 images x 512 proposals (BASELINE.json configs[0] shapes, on the GPU) and hands
    the padded <= 100 detections per image to the host every step (the evaluator's process()); after the K steps ONE
    distributed.gather_detection_store all-gathers every rank's results over NCCL, as the reference gathers once in
    evaluate().  Both are inside the timed region.  Same JSON contract as the training arm."""
    import torch.distributed as dist

    from unit_b200 import _lib, distributed as udist
    from unit_b200.stage import RoIStage
    from unit_b200.structures import Boxes, Instances

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --mode infer needs a CUDA device")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cuda.matmul.allow_tf32 = False
    P_INF = 512
    head = build_head(device).eval()
    g = _seeded(77 + rank)
    x = torch.relu(torch.randn(N_IMG * P_INF, FEAT_DIM, generator=g)).to(device)
    xw = torch.relu(torch.randn(N_IMG * P_INF, FEAT_DIM, generator=g)).to(device)
    stage = RoIStage(head, lambda pooled: (x, xw))
    topk = head.box_predictor.test_topk_per_image
    host = []
    # as in the training arm the res4 features leave the host as bf16 (the dtype BASELINE configs[1] names) and are
    # widened on the device inside the timed region; --e2e-upload-f32 copies fp32 (twice the PCIe bytes)
    up_bf16 = not args.e2e_upload_f32
    for s_ in range(N_SETS):
        f, pr, _, _ = make_inputs(3000 + 10 * rank + s_)
        if up_bf16:
            f = f.bfloat16()
        host.append((f.pin_memory(), [p[:P_INF].contiguous().pin_memory() for p in pr]))
    dev_sets = []
    for f, pr in host:
        dev_sets.append((f.float().to(device), [Instances(IMG_HW, proposal_boxes=Boxes(p.to(device)),
                                                          objectness_logits=torch.zeros(P_INF, device=device))
                                                for p in pr]))
    stage_bf16 = [torch.empty_like(f, device=device) for f, _ in host] if up_bf16 else None
    copy_stream = torch.cuda.Stream(device=device)
    copied = [torch.cuda.Event() for _ in range(N_SETS)]
    freed = [torch.cuda.Event() for _ in range(N_SETS)]
    for e in freed:
        e.record(torch.cuda.current_stream(device))
    # Per step (as the reference's evaluator does in process()): this rank's padded detections go to the host (pinned,
    # asynchronous; read one step later).  Across ranks (as the reference does ONCE in evaluate(), data/evaluators.py:159
    # comm.gather): the rank's whole result store is all-gathered over NCCL at the end of the timed loop.
    width = 6 * topk + 1
    max_steps = max(args.steps, args.warmup + N_SETS, 8)
    store = torch.zeros((max_steps * N_IMG, width), dtype=torch.float32, device=device)
    host_slots = [torch.zeros((N_IMG, width), dtype=torch.float32).pin_memory() for _ in range(N_SETS)]
    slot_ready = [torch.cuda.Event() for _ in range(N_SETS)]
    state = {"n": 0, "pending": None, "dets": 0}

    def consume(k):  # the evaluator's host-side read of one step's detections
        if k is None:
            return
        slot_ready[k].synchronize()
        state["dets"] = int(host_slots[k][:, 6 * topk].sum())

    def emit(dets):
        db, ds, dc, _, cnt = dets
        j = state["n"] % max_steps
        k = state["n"] % N_SETS
        state["n"] += 1
        packed = udist.pack_detections(db, ds, dc, cnt)
        store[j * N_IMG:(j + 1) * N_IMG].copy_(packed)
        host_slots[k].copy_(packed, non_blocking=True)
        slot_ready[k].record(torch.cuda.current_stream(device))
        prev, state["pending"] = state["pending"], k
        consume(prev)

    def finish(steps):
        consume(state["pending"])
        state["pending"] = None
        allr = udist.gather_detection_store(store[:steps * N_IMG])       # ONE collective for the whole loop
        n_all = allr[:, :, 6 * topk].sum().item()                        # host read of the gathered result
        state["gathered"] = int(n_all)
        state["n"] = 0

    def step(i):
        with torch.no_grad():
            emit(stage.infer_graphed(*dev_sets[i % N_SETS], padded=True))

    issued = [-1]

    def issue(j):
        k = j % N_SETS
        copy_stream.wait_event(freed[k])
        with torch.cuda.stream(copy_stream):
            f, pr = host[k]
            if up_bf16:
                stage_bf16[k].copy_(f, non_blocking=True)
                dev_sets[k][0].copy_(stage_bf16[k])
            else:
                dev_sets[k][0].copy_(f, non_blocking=True)
            for inst, src in zip(dev_sets[k][1], pr):
                inst.proposal_boxes.tensor.copy_(src, non_blocking=True)
            copied[k].record(copy_stream)
        issued[0] = j

    e2e_n = [0]

    def step_e2e(_i):
        j = e2e_n[0]
        e2e_n[0] += 1
        main = torch.cuda.current_stream(device)
        while issued[0] < j + 2:
            issue(issued[0] + 1)
        k = j % N_SETS
        main.wait_event(copied[k])
        with torch.no_grad():
            dets = stage.infer_graphed(*dev_sets[k], padded=True)
        freed[k].record(main)
        emit(dets)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        finish(steps)  # last host read + the end-of-loop gather of every rank's detections: inside the timed region
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(N_SETS + args.warmup):
        step(i)
    finish(min(N_SETS + args.warmup, max_steps))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    total_ms = timed(step, args.steps)
    for i in range(3):
        step_e2e(i)
    finish(3)
    e2e_ms = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # kernels inside the replayed graph: count them once with an eager call
    n0 = _lib.launch_count()
    with torch.no_grad():
        stage.infer(*dev_sets[0])
    per_call = _lib.launch_count() - n0
    if rank == 0:
        ms_step = total_ms / args.steps
        h2d = int(host[0][0].numel() * host[0][0].element_size() + sum(p.numel() * 4 for p in host[0][1]))
        line = {
            "metric": "RoI-stage inference images/sec (VOC R101-C4, 512 proposals/img)",
            "value": world * N_IMG / (ms_step / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "infer",
            "config": {"workload": "BASELINE.json configs[0] shapes on the GPU: VOC-RCNN-101-C4-split1 RoI-stage inference, 2 "
                                   "synthetic 800x1333 images per GPU, 512 proposals/img, 15 base + 5 novel classes: ROIAlign "
                                   "fwd -> transfer -> softmax+decode -> filter -> class-wise NMS -> top-100, then the "
                                   "per-step copy of the detections to the host and ONE end-of-loop all-gather over NCCL",
                       "box_head": "res5 excluded (out of scope): fixed synthetic [1024,2048] box features",
                       "l2": f"inputs rotate over {N_SETS} sets (137 MB of features + 822 MB ROIAlign output per call)",
                       "parallelism": "images sharded across GPUs; the only collective is the final all-gather of "
                                      "detections (reference: data/evaluators.py:159)"},
            "e2e": {"value": world * N_IMG / (e2e_ms / args.steps / 1e3), "unit": "images/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(N_IMG * width * 4),
                    "upload_dtype": "bf16 features, widened on the device inside the timed region" if up_bf16 else "f32"},
            "gpu_launches": int(per_call * args.steps), "detections_last_step": state["dets"],
            "detections_gathered": state.get("gathered"), "clocks": clocks,
            "gathered_bytes_per_rank": int(args.steps * N_IMG * width * 4),
            "gather": "per step: the rank's padded detections are copied to pinned host memory and read one step later (the "
                      "evaluator's process()); after the K timed steps ONE all-gather of every rank's K x 2 x (6 x 100 + 1) "
                      "fp32 result store over NCCL + the host read of the gathered counts, inside the timed region",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--dtype", choices=["f32", "bf16"], default="f32")
    ap.add_argument("--mode", choices=["train", "infer"], default="train",
                    help="train: the fine-tune step (BASELINE metric, default); infer: inference + detection gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to its GPU's NUMA-local CPUs")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="do not start the labelling (graph A) of step i+1 while step i runs")
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary inference-side measurements")
    ap.add_argument("--e2e-upload-f32", action="store_true",
                    help="e2e arm: copy the features from the host as fp32 (34.4 MB/step) instead of bf16 (17.2 MB)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "infer":
        run_infer(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
