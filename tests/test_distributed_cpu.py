"""CPU, world_size 2, gloo: the N>1 plumbing (image sharding, flat gradient bucket all-reduce, detection gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unit_b200.distributed import (FlatGradBucket, gather_detection_store, gather_detections, pack_detections,
                                       shard_indices, unpack_detections)

    torch.manual_seed(0)
    lin1, lin2 = torch.nn.Linear(8, 5), torch.nn.Linear(8, 12)
    bucket = FlatGradBucket(list(lin1.parameters()) + list(lin2.parameters()))
    assert bucket.flat.numel() == 8 * 5 + 5 + 8 * 12 + 12
    x = torch.full((4, 8), float(rank + 1))
    (lin1(x).sum() + lin2(x).sum()).backward()
    assert lin1.weight.grad.data_ptr() == bucket.flat.data_ptr()   # grads were written straight into the bucket
    local = lin1.weight.grad.clone()
    bucket.all_reduce_mean()
    expect = local * (1 + 2) / 2 / (rank + 1)                       # grad is linear in x
    ok = torch.allclose(lin1.weight.grad, expect)
    # detections gather
    topk = 3
    boxes = torch.full((2, topk, 4), float(rank))
    scores = torch.full((2, topk), 0.5 + rank)
    classes = torch.full((2, topk), rank, dtype=torch.int64)
    counts = torch.tensor([1 + rank, 2], dtype=torch.int32)
    b, s, c, n = gather_detections(boxes, scores, classes, counts, topk)
    ok &= len(b) == world and b[1].eq(1).all().item() and c[1].eq(1).all().item() and n[1].tolist() == [2, 2]
    ok &= shard_indices(5, rank, world) == ([0, 2, 4] if rank == 0 else [1, 3])
    # end-of-loop gather of a rank's whole result store (the reference's single comm.gather)
    store = torch.cat([pack_detections(boxes, scores, classes, counts) for _ in range(3)])  # 3 steps x 2 images
    allr = gather_detection_store(store)
    ok &= tuple(allr.shape) == (world, 6, 6 * topk + 1)
    ub, us, uc, un = unpack_detections(allr[1], topk)
    ok &= ub.eq(1).all().item() and uc.eq(1).all().item() and un.tolist() == [2, 2] * 3 and gather_detection_store(store)[0, 0, -1].item() == 1 and us.eq(1.5).all().item()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world_size_2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_flat_bucket_survives_zero_grad():
    """zero_grad(set_to_none=True) drops the views; all_reduce_mean must refuse, zero_() must re-bind."""
    from unit_b200.distributed import FlatGradBucket

    lin = torch.nn.Linear(4, 3)
    bucket = FlatGradBucket(lin.parameters())
    lin.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError, match="no longer a view"):
        bucket.all_reduce_mean()
    bucket.zero_()
    lin(torch.ones(2, 4)).sum().backward()
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    assert torch.equal(bucket.flat[:12].view(3, 4), torch.full((3, 4), 2.0))
    bucket.all_reduce_mean()  # single process: nothing to reduce, but the views are verified
