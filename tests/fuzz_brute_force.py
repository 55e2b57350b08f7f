"""The checker's coverage (scanline machinery through setup, binning, LOW and HIGH raster paths) against the independent
brute-force rasteriser of tests/test_oracle.py (edge functions from view_proj) on many random soups and cameras:
    python tests/fuzz_brute_force.py [scenes] [seed]
They may differ only where a pixel centre lies numerically on an edge.  Not collected by pytest."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402
from tests import parity_util as pu  # noqa: E402
from tests.test_oracle import _brute_force_counts  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst_frac, worst_abs, frags, pixels_off = 0.0, 0, 0, 0
for k in range(n):
    w, h = [(320, 192), (256, 256), (480, 270), (200, 120)][k % 4]
    sc = scenes.quad_soup(num_quads=int(rng.integers(200, 3000)), width=w, height=h, distance=float(rng.uniform(22.0, 50.0)),
                          seed=int(rng.integers(1, 1 << 30)), min_edge=float(rng.uniform(0.05, 0.5)), max_edge=float(rng.uniform(0.6, 4.0)))
    sc["camera"]["rot_h"], sc["camera"]["rot_v"] = float(rng.uniform(0, 6.28)), float(rng.uniform(-1.2, 1.2))
    cfg, inst, cols, rects = api.prepare_frame(sc)
    o = pu.run_oracle(sc, threads=8)
    got = o.read_frag_counts().astype(np.int64)
    want = _brute_force_counts(sc, cfg, inst)
    diff = got != want
    worst_frac, worst_abs = max(worst_frac, float(diff.mean())), max(worst_abs, int(np.abs(got - want).max()))
    frags += int(want.sum())
    pixels_off += int(diff.sum())
    o.close()
print(f"fuzz: {n} scenes, {frags} fragments, {pixels_off} pixels differ ({pixels_off / max(frags, 1):.2e} per fragment), "
      f"worst scene {worst_frac:.4%} of pixels, largest per-pixel difference {worst_abs}")
sys.exit(1 if worst_frac > 0.004 or worst_abs > 2 else 0)
