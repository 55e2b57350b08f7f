"""One GPU plays every rank of an N-way bin-range split in turn: stage times per range (what each device of the split
would spend), next to the full frame.  python tools/split_probe.py [config] [world] [--cull]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, multigpu, scenes  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 3
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cull = api.RENDER_CULL_INSTANCES if "--cull" in sys.argv else 0
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
r.close()
