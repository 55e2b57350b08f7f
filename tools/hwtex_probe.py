"""Hardware texture filtering experiment (LUCID_HW_TEXTURE=1) against the CPU checker's software filter:
max abs image difference and stage times on the textured scenes.   LUCID_HW_TEXTURE=1 python tools/hwtex_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402
from tests import parity_util as pu  # noqa: E402

for name, sc, mvq in (("arch (small)", pu.small_scenes()["arch"], 1 << 20), ("config3", scenes.get_config(3), 4793490)):
    o = pu.run_oracle(sc, mvq=mvq, threads=os.cpu_count())
    r, img = pu.run_cuda(sc, mvq=mvq)
    d = np.abs(img.view(np.uint8).astype(np.int32) - o.read_image().view(np.uint8).astype(np.int32))[..., :3] if False else \
        np.abs(img.view(np.uint8).reshape(sc["height"], sc["width"], 4).astype(np.int32) -
               o.read_image().view(np.uint8).reshape(sc["height"], sc["width"], 4).astype(np.int32))
    hist = np.bincount(d.reshape(-1), minlength=8)[:8]
    same_counts = np.array_equal(r.read_frag_counts(), o.read_frag_counts())
    cfg, inst, cols, rects = api.prepare_frame(sc)
    t = []
    for _ in range(5):
        r.render(cfg, inst, cols, rects)
        t.append(r.stage_times())
    ms = np.median(np.array(t), axis=0)
    print(f"{name}: hw={os.environ.get('LUCID_HW_TEXTURE', '0')} max abs diff {d.max()} / 255, histogram of |diff| 0..7 {hist.tolist()}, "
          f"fragment counts equal {same_counts}, shade {ms[6]:.3f} ms frame {ms[7]:.3f} ms")
    r.close()
