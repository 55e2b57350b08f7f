"""Frames in flight: two renderer handles on two streams of ONE GPU render alternate frames.  Compares frames/s of
one handle (frames back to back on one stream) with two handles, for the full frame and for one rank's share of an
8-way bin-range split (where every kernel leaves most of the GPU idle at its tail).
python tools/inflight_probe.py [config] [frames]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, multigpu, scenes  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 3
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sc = scenes.get_config(ci)
cfg, inst, cols, rects = api.prepare_frame(sc)
NH = int(os.environ.get('PROBE_HANDLES', '3'))
streams = [torch.cuda.Stream() for _ in range(NH)]
rs = [api.LucidRenderer(sc["width"], sc["height"], 0, 0, stream=s.cuda_stream) for s in streams]
for r in rs:
    r.set_scene(sc)
flags = api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_NO_STAGE_TIMES | api.RENDER_CULL_INSTANCES


def run(handles, n):
    for k in range(6):
        rs[handles[k % len(handles)]].render(cfg, inst, cols, rects, flags=flags)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        rs[handles[k % len(handles)]].render(cfg, inst, cols, rects, flags=flags)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0)


rs[0].render(cfg, inst, cols, rects)
cost = rs[0].read_bin_costs().astype(np.float64)
print(f"config {ci} full frame:", ", ".join(f"{h} handles {run(list(range(h)), frames):.1f} frames/s" for h in range(1, NH + 1)))
ranges = multigpu.split_bins(rs[0].bin_count, 8, cost)
for q in (0, 3, 7):
    for r in rs:
        r.set_bin_range(*ranges[q])
    res = [run(list(range(h)), frames * 4) for h in range(1, NH + 1)]
    print(f"rank {q} of 8, bins {ranges[q]}:", ", ".join(f"{h + 1} handles {v:.1f} frames/s ({1e3 / v:.3f} ms)" for h, v in enumerate(res)))
for r in rs:
    r.close()
