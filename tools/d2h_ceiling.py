"""What can the host side take?  Every rank copies an RGBA8 image from its GPU into pinned host memory in a loop
(the read-back of bench.py's e2e leg, nothing else) and the aggregate rate is printed: the ceiling the views-sharded
e2e numbers of an 8-GPU box are measured against.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_ceiling.py [width height]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
src = torch.zeros((h, w), dtype=torch.int32, device="cuda")
dst = [torch.empty((h, w), dtype=torch.int32).pin_memory() for _ in range(2)]
stream = torch.cuda.Stream()
results = {}
for label, copies in (("one copy in flight", 1), ("two copies in flight", 2)):
    with torch.cuda.stream(stream):
        for i in range(20):
            dst[i & 1].copy_(src, non_blocking=True)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n = 400
        t0 = time.perf_counter()
        for i in range(n):
            dst[i & 1].copy_(src, non_blocking=True)
            if copies == 1:
                stream.synchronize()
        stream.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    results[label] = {"images_per_s": round(world * n / dt, 1), "GB_per_s": round(world * n * w * h * 4 / dt / 1e9, 2)}
if rank == 0:
    print(json.dumps({"ranks": world, "image": [w, h], "bytes": w * h * 4, **results}))
if world > 1:
    dist.destroy_process_group()
