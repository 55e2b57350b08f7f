"""Renders a few frames of one config for ncu: python tools/profile_frame.py <config> [frames] [scale]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
sc = scenes.get_config(ci, scale)
cfg, inst, cols, rects = api.prepare_frame(sc)
r = api.LucidRenderer(sc["width"], sc["height"], 0, 0)
r.set_scene(sc)
for _ in range(frames):
    r.render(cfg, inst, cols, rects)
print(r.getStats())
print(np.round(r.stage_times(), 3))
r.close()
