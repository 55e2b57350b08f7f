"""Writes tests/golden/comparators_golden.json: digests of the checker's three comparator images (hardware alpha
blending in submission order, weighted blended OIT, 4-layer MLAB; oracle/lucid_oracle.cpp comparePixel) on the parity
scenes and the order-recovery scene of tests/test_comparators.py.  They pin the checker's comparator modes against
silent drift; regenerate only when those modes are deliberately changed:
    python tests/golden/make_comparators_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import parity_util as pu  # noqa: E402
from tests.golden.make_oracle_golden import digest  # noqa: E402
from tests.test_comparators import _run_oracle, mixed_order_scene  # noqa: E402

MODES = ("hw_blend", "wboit", "mlab4")


def record(o):
    return {name: digest(o.read_compare_image(mode)) for mode, name in enumerate(MODES)}


def scenes_to_pin():
    out = pu.small_scenes()
    out["mixed_order"] = mixed_order_scene()
    return out


if __name__ == "__main__":
    out = {name: record(_run_oracle(sc)) for name, sc in scenes_to_pin().items()}
    with open(os.path.join(HERE, "comparators_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", list(out))
