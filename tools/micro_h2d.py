"""Pinned host -> device copy bandwidth per rank, all ranks copying at once (the e2e arm's limiter at N = 4 / 8).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/micro_h2d.py

Every rank copies a 34.4 MB pinned buffer (one step's fp32 features) to its GPU 50 times between two barriers; rank 0
prints per-rank and aggregate GB/s."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = {}
for name, nbytes in (("fp32_features_34MB", 2 * 1024 * 50 * 84 * 4), ("bf16_features_17MB", 2 * 1024 * 50 * 84 * 2)):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(5):
        d.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = 50
    t0 = time.perf_counter()
    for _ in range(iters):
        d.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([nbytes * iters / dt / 1e9], device=dev)
    allg = [torch.zeros_like(gbs) for _ in range(world)]
    if world > 1:
        dist.all_gather(allg, gbs)
    else:
        allg = [gbs]
    out[name] = {"per_rank_GBps": [round(float(x), 1) for x in allg], "aggregate_GBps": round(float(sum(allg)), 1),
                 "ms_per_copy": round(dt / iters * 1e3, 3)}
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}))
if world > 1:
    dist.destroy_process_group()
