"""Per-view stage times of config 1 around the orbit, with and without an L2 flush before the frame."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

sc = scenes.get_config(1)
w, h = sc["width"], sc["height"]
inst, cols, rects = api.build_instances(sc["draw_calls"], sc["materials"])
stream = torch.cuda.current_stream()
r = api.LucidRenderer(w, h, 0, 0, stream=stream.cuda_stream)
r.set_scene(sc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for do_flush in (False, True):
    print("flush", do_flush)
    for view in range(0, 64, 8):
        cam = dict(sc["camera"])
        cam["rot_h"] = cam["rot_h"] + 2.0 * np.pi * view / 64
        cfg = api.make_config(api.make_camera(cam, w, h), len(inst), sc["background"])
        ts = []
        for _ in range(4):
            if do_flush:
                flush.fill_(1)
            r.render(cfg, inst, cols, rects, flags=api.RENDER_SKIP_INFO)
            ts.append(r.stage_times())
        ms = np.median(np.array(ts), axis=0)
        print(f"  view {view:2d}", np.round(ms, 3).tolist())
r.close()
