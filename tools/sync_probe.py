"""How long does the per-frame cross-rank completion signal take?  torchrun --nproc-per-node N tools/sync_probe.py"""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
tok = torch.zeros(1, device="cuda")
work = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(20):
    dist.all_reduce(tok)
torch.cuda.synchronize()
for name, body in (("all_reduce only", lambda: dist.all_reduce(tok)),
                   ("fill(64MB) + all_reduce", lambda: (work.fill_(1), dist.all_reduce(tok))),
                   ("fill(64MB) only", lambda: work.fill_(1)),
                   ("barrier", lambda: dist.barrier())):
    s = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
    for k in range(50):
        s[k].record(stream)
        body()
        e[k].record(stream)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in zip(s, e))
    if rank == 0:
        print(f"world {world} {name:26s} median {ms[25]:.4f} ms  p90 {ms[45]:.4f} ms", flush=True)
dist.destroy_process_group()
