"""Multi-GPU plumbing of the RoI stage: one process per GPU, images sharded across ranks.

The path has NO data-path collective: every op is per image.  Two exchanges exist at the edges:
  * fine-tuning: ONE NCCL all-reduce per step over a single flat bucket that holds only the trainable RoI-head /
    transfer parameter gradients (VOC-FT: cls_score_ft + bbox_pred_ft = 0.83 MB fp32), replacing DDP's generic
    25 MB buckets (reference: DistributedDataParallel inside [D2] DefaultTrainer, engine/defaults.py:266-288);
  * inference: a final gather of the <= 100 detections per image (reference: comm.gather, data/evaluators.py:159).
Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin image sharding (image i -> rank i % world)."""
    return list(range(rank, n_items, world_size))


class FlatGradBucket:
    """Pre-flattened gradient bucket: ``param.grad`` of every registered parameter is a view into one buffer, so
    the backward kernels write straight into it and one all-reduce (average) covers the whole RoI head."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one trainable parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        self._offsets = []
        off = 0
        for p in self.params:
            self._offsets.append(off)
            off += p.numel()
        self.bind()

    def bind(self) -> None:
        """(Re-)install ``param.grad`` as views of the flat buffer.  ``optimizer.zero_grad()`` /
        ``Module.zero_grad()`` default to ``set_to_none=True``, which drops the views: backward would then allocate
        fresh ``.grad`` tensors and the all-reduce would run on an orphaned buffer.  Call ``bucket.zero_()`` instead
        of ``zero_grad()`` (or ``zero_grad(set_to_none=False)``); ``zero_`` re-binds, ``all_reduce_mean`` verifies."""
        for p, off in zip(self.params, self._offsets):
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def check_bound(self) -> None:
        es = self.flat.element_size()
        for p, off in zip(self.params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * es:
                raise RuntimeError(
                    "FlatGradBucket: a parameter's .grad is no longer a view of the flat bucket (zero_grad("
                    "set_to_none=True) or an external assignment replaced it); gradients would not be all-reduced. "
                    "Use bucket.zero_() / zero_grad(set_to_none=False), or call bucket.bind() before backward.")

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero_(self) -> None:
        """Clear the bucket and make sure every ``param.grad`` (still) aliases it."""
        self.bind()
        self.flat.zero_()

    def all_reduce_mean(self, async_op: bool = False):
        self.check_bound()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        self._avg = dist.get_backend() == "nccl"  # NCCL averages inside the collective; gloo needs sum + divide
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return work
        if not self._avg:
            self.flat.div_(dist.get_world_size())
        return None

    def finish(self, work) -> None:
        if work is not None:
            work.wait()
            if not self._avg:
                self.flat.div_(dist.get_world_size())


def gather_detections(boxes: torch.Tensor, scores: torch.Tensor, classes: torch.Tensor, counts: torch.Tensor,
                      topk: int):
    """Gather padded per-image detections from every rank onto all ranks.

    boxes [n,topk,4], scores [n,topk], classes [n,topk] (int64), counts [n] (int32); every rank must hold the same n.
    Returns lists (one entry per rank) of the same tensors."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [boxes], [scores], [classes], [counts]
    ws = dist.get_world_size()
    packed = torch.cat([boxes.reshape(boxes.shape[0], -1), scores, classes.to(scores.dtype),
                        counts.to(scores.dtype)[:, None]], dim=1).contiguous()
    out = [torch.empty_like(packed) for _ in range(ws)]
    dist.all_gather(out, packed)
    b, s, c, n = [], [], [], []
    for t in out:
        b.append(t[:, :4 * topk].reshape(-1, topk, 4))
        s.append(t[:, 4 * topk:5 * topk])
        c.append(t[:, 5 * topk:6 * topk].to(torch.int64))
        n.append(t[:, 6 * topk].to(torch.int32))
    return b, s, c, n


def pack_detections(boxes: torch.Tensor, scores: torch.Tensor, classes: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """Padded detections of n images as ONE [n, 6 * topk + 1] fp32 tensor (boxes | scores | classes | count): what a rank
    appends to its result store after every inference step and what ``gather_detection_store`` exchanges."""
    return torch.cat([boxes.reshape(boxes.shape[0], -1), scores, classes.to(scores.dtype),
                      counts.to(scores.dtype)[:, None]], dim=1)


def unpack_detections(packed: torch.Tensor, topk: int):
    """Inverse of ``pack_detections``: (boxes [n,topk,4], scores [n,topk], classes int64 [n,topk], counts int32 [n])."""
    return (packed[:, :4 * topk].reshape(-1, topk, 4), packed[:, 4 * topk:5 * topk],
            packed[:, 5 * topk:6 * topk].to(torch.int64), packed[:, 6 * topk].to(torch.int32))


def gather_detection_store(store: torch.Tensor) -> torch.Tensor:
    """The reference gathers the per-rank prediction lists ONCE, after the inference loop (data/evaluators.py:159,
    ``comm.gather`` inside ``evaluate()``).  Here every rank keeps its packed detections in one device tensor
    ``store [n_local_images, 6 * topk + 1]`` and this is that single exchange: one all-gather over NCCL (gloo on CPU)
    -> [world, n_local_images, 6 * topk + 1] on every rank.  All ranks must hold the same n_local_images."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return store.unsqueeze(0)
    store = store.contiguous()
    world = dist.get_world_size()
    out = torch.empty((world * store.shape[0],) + tuple(store.shape[1:]), dtype=store.dtype, device=store.device)
    dist.all_gather_into_tensor(out, store)  # concatenated along dim 0 (the form gloo and NCCL both accept)
    return out.view((world,) + tuple(store.shape))
