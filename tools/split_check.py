"""Bin-row split on real GPUs (torchrun, one rank per GPU): every rank rasterises its rows of one
frame and stores them into rank 0's image through a CUDA-IPC peer mapping; rank 0 then compares the
composite with its own single-GPU frame, bit for bit, and the summed fragment counters.
    python -m torch.distributed.run --nproc-per-node N tools/split_check.py [config ...]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, multigpu, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
token = torch.zeros(1, device="cuda")
ok_all = True
for ci in [int(a) for a in sys.argv[1:]] or [0, 2]:
    sc = scenes.get_config(ci, 1.0)
    w, h = sc["width"], sc["height"]
    cfg, inst, cols, rects = api.prepare_frame(sc)
    part = api.LucidRenderer(w, h, 0, 0, device=local)
    part.set_scene(sc)
    # cost-balanced row-major bin ranges (rank 0's calibration frame), as bench.py --mode split uses
    part.render(cfg, inst, cols, rects)
    cost = torch.from_numpy(part.read_bin_costs().astype(np.float64)).cuda()
    dist.broadcast(cost, src=0)
    part.set_bin_range(*multigpu.split_bins(part.bin_count, world, cost.cpu().numpy())[rank])
    handle = [part.ipc_export_image() if rank == 0 else None]
    dist.broadcast_object_list(handle, src=0)
    peer = part.ipc_open_image(handle[0]) if rank != 0 else None
    part.render(cfg, inst, cols, rects)
    if peer is not None:
        part.composite_to(peer, w * 4)
    st = part.getStats()
    frags = torch.tensor([st["fragments"]], device="cuda", dtype=torch.int64)
    dist.all_reduce(frags)
    dist.all_reduce(token)  # every strip has landed
    torch.cuda.synchronize()
    if rank == 0:
        composite = part.read_image()
        full = api.LucidRenderer(w, h, 0, 0, device=local)
        full.set_scene(sc)
        img = np.zeros((h, w), np.uint32)
        full.render(cfg, inst, cols, rects, out=img)
        fs = full.getStats()
        same = bool(np.array_equal(composite, img))
        ok = same and int(frags.item()) == fs["fragments"]
        ok_all &= ok
        print(f"config {ci} world {world}: composite == single-GPU frame: {same}; fragments split {int(frags.item())} "
              f"full {fs['fragments']}: {'OK' if ok else 'MISMATCH'}", flush=True)
        full.close()
    dist.barrier()
    if peer is not None:
        part.ipc_close_image(peer)
    part.close()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
