"""Depth-complexity histogram of a scene, counted by the CPU checker (test infrastructure):
    python tests/depth_histogram.py 2                       # BASELINE configs[2] as lucid_b200.scenes builds it
    python tests/depth_histogram.py hairball ribbon_width=0.05 distance=14
BASELINE.json asks configs[2] for "depth complexity > 64 per pixel" on the raster_high path with no bin over
the reference's limits (raster_high.glsl:80-83,140-141); this prints what a candidate actually has."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402
from oracle.binding import Oracle  # noqa: E402


def histogram(sc, mvq=4793490):
    o = Oracle(sc["width"], sc["height"], 0, mvq, threads=os.cpu_count() or 1)
    o.set_scene(sc)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    t0 = time.time()
    o.render(cfg, inst, cols, rects)
    dt = time.time() - t0
    fc = o.read_frag_counts()
    st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    levels = o.read_bin_levels()
    o.close()
    cov = fc[fc > 0]
    out = dict(covered_frac=round(float(cov.size) / fc.size, 4), mean=round(float(cov.mean()), 2) if cov.size else 0,
               median=float(np.median(cov)) if cov.size else 0, p99=float(np.percentile(cov, 99)) if cov.size else 0,
               max=int(fc.max()), over64_frac=round(float((fc > 64).sum()) / fc.size, 4),
               red_bins=int((levels == 5).sum()), oracle_s=round(dt, 1))
    out.update({k: st[k] for k in ("input_quads", "visible_small", "visible_large", "bin_quads", "high_bins", "low_bins",
                                   "fragments", "half_block_tris", "dropped_quads", "list_overflow",
                                   "max_quads_per_bin")})
    return out


if __name__ == "__main__":
    arg = sys.argv[1] if len(sys.argv) > 1 else "2"
    kw = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        kw[k] = float(v) if "." in v or "e" in v else int(v)
    sc = scenes.get_config(int(arg)) if arg.isdigit() else getattr(scenes, arg)(**kw)
    print(histogram(sc))
