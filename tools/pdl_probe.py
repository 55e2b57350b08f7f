"""Frame time with and without per-stage events between the kernels (the events keep a launch from
overlapping its predecessor's tail):  python tools/pdl_probe.py [config ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ci in [int(a) for a in sys.argv[1:]] or [1]:
    sc = scenes.get_config(ci)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r = api.LucidRenderer(sc["width"], sc["height"], 0, 0, stream=stream.cuda_stream)
    r.set_scene(sc)
    for name, fl in (("stage events", 0), ("frame events only", api.RENDER_NO_STAGE_TIMES)):
        for cold in (False, True):
            ms = []
            for k in range(13):
                if cold:
                    flush.fill_(k)
                r.render(cfg, inst, cols, rects, flags=api.RENDER_ASYNC | api.RENDER_SKIP_INFO | fl)
                r.wait()
                ms.append(r.stage_times()[7])
            print(f"config {ci} {name:18s} {'cold L2' if cold else 'warm L2'}: frame {np.median(ms[3:]):.4f} ms")
    r.close()
