"""Comparators on the BASELINE scenes: the exact frame next to what hardware alpha blending in submission
order, weighted blended OIT and 4-layer MLAB make of the same samples -- how many pixels differ, by how much, and
what the comparator kernels took.
   python tools/compare_probe.py [config ...] > profiles/<tag>_comparators.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

NAMES = {api.COMPARE_HW_BLEND: "hw_blend_submission_order", api.COMPARE_WBOIT: "wboit", api.COMPARE_MLAB4: "mlab4"}
out = {}
for ci in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    sc = scenes.get_config(ci)
    w, h = sc["width"], sc["height"]
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(w, h, 0, 4793490)
    r.set_scene(sc)
    exact = np.zeros((h, w), np.uint32)
    for _ in range(3):
        r.render(cfg, inst, cols, rects, out=exact)
    frame_ms = float(r.stage_times()[7])
    st = api.decode_stats(r.read_info(), r.bin_count, w, h)
    ex = exact.view(np.uint8).reshape(h, w, 4)[..., :3].astype(np.int32)
    res = {"workload": sc["name"], "resolution": [w, h], "fragments": st["fragments"], "exact_frame_ms": round(frame_ms, 3), "modes": {}}
    for mode, name in NAMES.items():
        times = []
        for _ in range(3):
            img, ms = r.compare_render(mode, cfg)
            times.append(ms)
        d = np.abs(img.view(np.uint8).reshape(h, w, 4)[..., :3].astype(np.int32) - ex)
        mse = float((d.astype(np.float64) ** 2).mean())
        res["modes"][name] = {
            "kernel_ms": round(float(np.median(times)), 3),
            "pixels_off_by_more_than_1": round(float((d.max(axis=2) > 1).mean()), 5),
            "pixels_off_by_more_than_8": round(float((d.max(axis=2) > 8).mean()), 5),
            "mean_abs_error_255": round(float(d.mean()), 4),
            "max_abs_error_255": int(d.max()),
            "psnr_db": None if mse == 0 else round(10.0 * np.log10(255.0 ** 2 / mse), 2),
        }
    out["config%d" % ci] = res
    r.close()
print(json.dumps(out, indent=1))
