"""One GPU plays every rank of an N-way bin-range split in turn: stage times per range (what each device of the split
would spend), next to the full frame.  python tools/split_probe.py [config] [world] [--cull]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, multigpu, scenes  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 3
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cull = api.RENDER_CULL_INSTANCES if "--cull" in sys.argv else 0  # full frame below: always without
sc = scenes.get_config(ci)
cfg, inst, cols, rects = api.prepare_frame(sc)
r = api.LucidRenderer(sc["width"], sc["height"], 0, 0)
r.set_scene(sc)
names = ["setup", "count", "scan", "dispatch", "lists", "sort", "shade", "frame"]


def measure(flags=0):
    for _ in range(2):
        r.render(cfg, inst, cols, rects, flags=flags)
    t = []
    for _ in range(5):
        r.render(cfg, inst, cols, rects, flags=flags)
        t.append(r.stage_times())
    return np.median(np.array(t), axis=0)


full = measure()
print("full   ", dict(zip(names, np.round(full, 3).tolist())))
cost = r.read_bin_costs().astype(np.float64)
ranges = multigpu.split_bins(r.bin_count, world, cost)
tot = np.zeros(8)
for q, (lo, hi) in enumerate(ranges):
    r.set_bin_range(lo, hi)
    ms = measure(cull)
    st = r.getStats()
    tot += ms
    print(f"rank {q} bins [{lo},{hi}) cost share {cost[lo:hi].sum() / cost.sum():.3f}", dict(zip(names, np.round(ms, 3).tolist())),
          "fragments", st["fragments"], "hbt", st["half_block_tris"])
print("sum    ", dict(zip(names, np.round(tot, 3).tolist())))

# the feedback of bench.py --mode split, played on one GPU: every rank's bins are rescaled to the time the rank needed
iters = 0
for a in sys.argv:
    if a.startswith("--feedback="):
        iters = int(a.split("=")[1])
for mode in ("after-setup", "frame"):
    c = cost.copy()
    rg = multigpu.split_bins(r.bin_count, world, c)
    for it in range(iters):
        times = []
        for lo, hi in rg:
            r.set_bin_range(lo, hi)
            ms = measure(cull)
            times.append(float(ms[1:7].sum()) if mode == "after-setup" else float(ms[:7].sum()))
        frames = []
        for lo, hi in rg:
            r.set_bin_range(lo, hi)
            frames.append(float(measure(cull)[7]))
        print(f"[{mode}] iteration {it}: frame max {max(frames):.3f} mean {np.mean(frames):.3f} min {min(frames):.3f}  bounds {[lo for lo, _ in rg]}")
        for q, (lo, hi) in enumerate(rg):
            c[lo:hi] *= times[q] / max(float(c[lo:hi].sum()), 1e-9)
        rg = multigpu.split_bins(r.bin_count, world, c)
r.close()
