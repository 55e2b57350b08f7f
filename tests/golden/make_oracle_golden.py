"""Writes tests/golden/oracle_golden.json: digests of the CPU oracle's outputs on the small parity
scenes.  They pin the oracle against silent drift (the reference itself ships no golden data for
this path) and let the GPU tests check the CUDA path against committed values.
Regenerate only when the oracle is deliberately changed:  python tests/golden/make_oracle_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from lucid_b200 import api  # noqa: E402
from tests import parity_util as pu  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def record(o):
    st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    bq, bt = o.read_bin_lists()
    _, counts = api.split_info(o.info, o.bin_count)
    return dict(stats={k: st[k] for k in ("input_quads", "visible_small", "visible_large", "rejected_frustum",
                                          "rejected_between_samples", "bin_quads", "bin_tris", "low_bins",
                                          "high_bins", "promoted_bins", "fragments", "half_block_tris")},
                image=digest(o.read_image()), frag_counts=digest(o.read_frag_counts()),
                bin_counts=digest(counts[:6]), bin_quads=digest(bq), bin_tris=digest(bt),
                quad_aabbs=digest(np.concatenate([o.read_quad_aabbs(0), o.read_quad_aabbs(1)])),
                tri_records=digest(pu.canonical_tri_records(np.concatenate([o.read_tri_records(0), o.read_tri_records(1)]))))


FULL_MVQ = 4793490  # the reference's MAX_VISIBLE_QUADS formula at the default memory budget


def main():
    """Small parity scenes, then BASELINE.json configs[0..3] at full size ("config0" .. "config3": about a
    minute of CPU for the two 4K scenes)."""
    from lucid_b200 import scenes
    out = {name: record(pu.run_oracle(sc)) for name, sc in pu.small_scenes().items()}
    for ci in range(4):
        out[f"config{ci}"] = record(pu.run_oracle(scenes.get_config(ci), mvq=FULL_MVQ, threads=os.cpu_count() or 8))
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", list(out))


if __name__ == "__main__":
    main()
