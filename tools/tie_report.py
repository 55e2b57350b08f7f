"""Depth-key ties on the BASELINE scenes ("report pixels whose block had depth-key ties separately"),
counted by the CPU checker (test infrastructure; no GPU needed): lists with a run of equal quantised depth, entries in
such runs, pixels covered by two or more entries of one run -- the only pixels whose colour can depend on the tie
order (the reference: arrival order of racing atomics; here: triangle index) -- and how many pixels actually change
when every run is rendered in reverse order.
   python tools/tie_report.py [config ...] > profiles/<tag>_depth_key_ties.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

out = {}
for ci in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]:
    sc = scenes.get_config(ci)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    imgs, st, mask, covered = [], None, None, 0
    for reverse in (False, True):
        o = Oracle(sc["width"], sc["height"], 0, 4793490, threads=os.cpu_count() or 1)
        o.set_tie_report(True)
        o.set_reverse_ties(reverse)
        o.set_scene(sc)
        o.render(cfg, inst, cols, rects)
        imgs.append(o.read_image())
        mask, st = o.read_tie_pixels()
        covered = int((o.read_frag_counts() > 0).sum())
        o.close()
    differ = imgs[0] != imgs[1]
    d = np.abs(imgs[0].view(np.uint8).astype(np.int32) - imgs[1].view(np.uint8).astype(np.int32))
    out["config%d" % ci] = {
        "workload": sc["name"], "resolution": [sc["width"], sc["height"]], "covered_pixels": covered,
        "lists_with_ties": st["lists"], "entries_in_tie_runs": st["entries"], "tie_pixels": st["pixels"],
        "tie_pixels_frac_of_covered": round(st["pixels"] / max(covered, 1), 6),
        "pixels_changed_by_reversed_ties": int(differ.sum()),
        "changed_outside_tie_pixels": int((differ & (mask == 0)).sum()),
        "max_channel_change_255": int(d.max()),
    }
print(json.dumps(out, indent=1))
