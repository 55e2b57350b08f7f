"""Per-phase clock sums of k_raster_blocks (experiment build with -DRB_PHASE_CLOCKS):
   LUCID_B200_SO=variants/lucid_phase.so python tools/phase_probe.py [config ...]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

for ci in [int(a) for a in sys.argv[1:]] or [1]:
    sc = scenes.get_config(ci)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(sc["width"], sc["height"], 0, 0)
    r.set_scene(sc)
    lib = api.load_library()
    for _ in range(3):
        r.render(cfg, inst, cols, rects)
    ph = (C.c_ulonglong * 16)()
    n_warps = 148 * 5 * 4
    we = (C.c_ulonglong * n_warps)()
    lib.lucid_debug_phase_clocks(ph, we, 1)
    r.render(cfg, inst, cols, rects)
    ms = r.stage_times()
    lib.lucid_debug_phase_clocks(ph, we, 0)
    ph = np.array(list(ph), dtype=np.float64)
    we = np.array(list(we), dtype=np.float64) / 1e3
    tot = ph[:4].sum()
    print(f"== config {ci}: shade stage {ms[5]:.3f} ms, kernel span {(ph[9] - ph[8]) / 1e3:.1f} us, items {int(ph[5])}, "
          f"entries {int(ph[6])}")
    for k, name in enumerate(["work fetch", "key pass", "sort+ties", "shading"]):
        print(f"   {name:10s} {100 * ph[k] / tot:5.1f}% of warp cycles; {ph[k] / max(ph[5], 1):9.0f} cycles/item")
    print(f"   warp-cycles total {tot / n_warps / 1.965e3:.1f} us per warp (mean busy incl. stalls)")
    print("   warp finish times us: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" %
          (we.min(), np.percentile(we, 10), np.median(we), np.percentile(we, 90), we.max()))
    r.close()
