"""Pins of the CPU checker against the reference's own shader source.

tests/golden/ref_shader_funcs.json.gz holds the outputs of functions compiled from the GLSL text of the
reference tree (oracle/build_ref_shaders.py: processInputQuad, storeTri, loadScanlineParams*, scanlineStep,
rasterBinStep, rasterHalfBlockCentroid / Bits, rasterBlockDepth, the sample reduction, shadeSample) on seeded
inputs.  The checker's own
functions (oracle_fn_* in oracle/lucid_oracle.cpp -- the ones its pipeline calls) must give the same words.
These are the functions that decide coverage: which quads survive, their bin AABBs, the plane, barycentric
and scanline equations, the bin-row and pixel-row spans, fragment counts, centroids and block depth keys.
The CUDA kernels are compared with the checker by the GPU suite, so the pin carries over to them.
"""
import ctypes as C
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.binding import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
vp = C.c_void_p


@pytest.fixture(scope="module")
def golden():
    with gzip.open(os.path.join(HERE, "golden", "ref_shader_funcs.json.gz"), "rt") as f:
        return json.load(f)


def ptr(a):
    return a.ctypes.data_as(vp)


def _lib(o):
    lib = o.lib
    lib.oracle_fn_process_quad.argtypes = [vp, vp, vp, vp]
    lib.oracle_fn_store_tri.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp]
    lib.oracle_fn_raster_rows.argtypes = [vp, C.c_float, C.c_float, C.c_int, vp]
    lib.oracle_fn_bin_rows.argtypes = [vp, vp]
    lib.oracle_fn_half_block.argtypes = [C.c_uint32, C.c_uint32, C.c_int, vp, C.c_float, C.c_float, C.c_float, vp]
    return lib


def test_checker_functions_match_the_reference_shader_functions(golden):
    counts = {"quads": 0, "visible": 0, "tris": 0, "bin_rows": 0, "raster_rows": 0, "half_blocks": 0}
    for case in golden["cases"]:
        o = Oracle(case["width"], case["height"], 0, 1 << 16, threads=1)
        lib = _lib(o)
        cfg = np.array(case["config_words"], np.uint32)
        try:
            for q in case["quads"]:
                pos = np.array(q["pos"], np.uint32).view(np.float32).reshape(4, 3)
                idx = np.array(q["idx"], np.uint32)
                o.set_scene({"positions": pos, "quads": idx.reshape(1, 4)})
                res = np.zeros(5, np.uint32)
                lib.oracle_fn_process_quad(o.h, ptr(cfg), ptr(idx), ptr(res))
                assert res.tolist() == q["process_quad"], ("processInputQuad", q["pos"], res.tolist(), q["process_quad"])
                counts["quads"] += 1
                for t in q.get("tris", []):
                    counts["visible"] += t["second"] == 0
                    tri = np.array(t["tri"], np.uint32).view(np.float32)
                    rec = np.zeros(21, np.uint32)
                    lib.oracle_fn_store_tri(o.h, ptr(cfg), ptr(tri), t["flags_id"], t["y_aabb"], ptr(rec))
                    assert rec.tolist() == t["record"], ("storeTri", t["tri"], rec.tolist(), t["record"])
                    counts["tris"] += 1
                    scan8 = rec[8:16].copy()
                    rows = np.zeros(2 + 256, np.int32)
                    lib.oracle_fn_bin_rows(ptr(scan8), ptr(rows))
                    assert rows[:len(t["bin_rows"])].tolist() == t["bin_rows"], ("bin rows", t["record"][8:16])
                    counts["bin_rows"] += (len(t["bin_rows"]) - 2) // 2
                    for rr in t["raster_rows"]:
                        spans = np.zeros(6, np.uint32)
                        lib.oracle_fn_raster_rows(ptr(scan8), rr["start"][0], rr["start"][1], 2, ptr(spans))
                        assert spans.tolist() == rr["spans"], ("rasterBinStep", rr["start"], t["record"][8:16])
                        counts["raster_rows"] += 8
                        for startx, want in zip(rr["half_block_startx"], rr["half_blocks"]):
                            depth_eq = rec[16:19].view(np.float32).copy()
                            cpx, cpy = np.float32(rr["start"][0] + startx + 3.25), np.float32(rr["start"][1] + 1.75)
                            o5 = np.zeros(5, np.uint32)
                            lib.oracle_fn_half_block(int(spans[0]), int(spans[1]), startx, ptr(depth_eq), float(cpx), float(cpy),
                                                     float(0x7FFFE if startx % 16 else 0x3FFFFE), ptr(o5))
                            assert o5.tolist() == want, ("half block", startx, spans.tolist(), o5.tolist(), want)
                            counts["half_blocks"] += 1
        finally:
            o.close()
    # the vectors exercise every path: rejections of all kinds but "other", large and small quads
    assert counts["quads"] >= 200 and counts["tris"] >= 150 and counts["half_blocks"] >= 1000, counts
    statuses = {q["process_quad"][0] for c in golden["cases"] for q in c["quads"]}
    assert {1, 2, 3, 0xFFFFFFFF} <= statuses
    sizes = {q["process_quad"][1] for c in golden["cases"] for q in c["quads"] if q["process_quad"][0] == 0xFFFFFFFF}
    assert sizes == {0, 1}


def test_checker_reduction_and_codec_match_the_reference(golden):
    """shading.glsl initReduceSamples / reduceSample / finishReduceSamples (3-entry insertion window, blending,
    background) and funcs.glsl encodeRGBA8 against the checker's Reducer and codec, bit for bit."""
    o = Oracle(64, 64, 0, 1 << 10, threads=1)
    lib = o.lib
    lib.oracle_fn_reduce_pixel.argtypes = [vp, vp, C.c_int, vp]
    lib.oracle_fn_encode_rgba8.argtypes = [vp]
    lib.oracle_fn_encode_rgba8.restype = C.c_uint32
    cfg = np.array(golden["cases"][0]["config_words"], np.uint32)
    try:
        longest = 0
        for r in golden["reduce"]:
            samples = np.array(r["samples"], np.uint32)
            o4 = np.zeros(4, np.uint32)
            lib.oracle_fn_reduce_pixel(ptr(cfg), ptr(samples), samples.size // 2, ptr(o4))
            assert o4.tolist() == r["rgba"], (samples.size // 2, o4.tolist(), r["rgba"])
            longest = max(longest, samples.size // 2)
        assert len(golden["reduce"]) >= 50 and longest > 64  # more than two rounds of 32
        for e in golden["encode_rgba8"]:
            c = np.array(e["rgba"], np.uint32)
            assert int(lib.oracle_fn_encode_rgba8(ptr(c))) == e["packed"]
        # quad_setup.glsl storeQuad: vertex colours / normals / uv0 + deltas per quad
        lib.oracle_fn_store_quad.argtypes = [C.c_uint32, vp, vp, vp, vp]
        assert len(golden["store_quad"]) >= 16
        for e in golden["store_quad"]:
            o16 = np.zeros(16, np.uint32)
            lib.oracle_fn_store_quad(e["flags"], ptr(np.array(e["colors"], np.uint32)), ptr(np.array(e["normals"], np.uint32)),
                                     ptr(np.array(e["uvs"], np.uint32)), ptr(o16))
            assert o16.tolist() == e["out"], (hex(e["flags"]), o16.tolist(), e["out"])
    finally:
        o.close()


def test_checker_shade_sample_matches_the_reference(golden):
    """shading.glsl shadeSample (+ getTriangle*, funcs.glsl finalShading / sRGB conversions / codecs): sample
    depth, perspective-correct barycentrics, uv and its analytic derivatives as handed to the sampler, instance
    and vertex colours, interpolated or flat normals, lighting, truncating RGBA8 encode.  The texture fetch is a
    probe on both sides (the Vulkan sampler is not part of the reference's source) and pow is the polynomial
    contract on both sides: what is pinned is the reference's arithmetic around them."""
    seen_flags, fetched, dropped = set(), 0, 0
    oracles = {}
    try:
        for e in golden["shade"]:
            case = golden["cases"][e["case"]]
            if e["case"] not in oracles:
                oracles[e["case"]] = Oracle(case["width"], case["height"], 0, 1 << 10, threads=1)
            o = oracles[e["case"]]
            o.lib.oracle_fn_shade_sample.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, vp]
            cfg = np.array(case["config_words"], np.uint32)
            rec, attrs = np.array(e["record"], np.uint32), np.array(e["attrs"], np.uint32)
            uv_rect, preset = np.array(e["uv_rect"], np.uint32), np.array(e["tex_preset"], np.uint32)
            o10 = np.zeros(10, np.uint32)
            o.lib.oracle_fn_shade_sample(o.h, ptr(cfg), ptr(rec), ptr(attrs), e["inst_color"], ptr(uv_rect), ptr(preset),
                                         e["pixel"][0], e["pixel"][1], e["second"], ptr(o10))
            assert o10.tolist() == e["out"], (hex(int(rec[19]) & 0xFFFF), o10.tolist(), e["out"])
            seen_flags.add(int(rec[19]) & 0xFFFF)
            fetched += e["out"][9] != 0
            dropped += e["out"][0] == 0
    finally:
        for o in oracles.values():
            o.close()
    assert len(golden["shade"]) >= 150 and len(seen_flags) >= 30 and fetched >= 40 and dropped >= 3


def test_checker_bin_counts_and_lists_match_the_reference_on_whole_scenes(golden):
    """Scene level: the reference's per-invocation functions of quad setup, bin counting and bin dispatch
    (processInputQuad, addVisibleTri / storeTri, countSmallQuadBins, countLargeTriBins, dispatchQuad,
    dispatchLargeTriSimple) run over every quad of a scene give the checker's visible-quad counts, per-bin quad and
    triangle counts and per-bin lists (as sorted sets) exactly; walking every listed triangle the way generateRowTris
    does (loadScanlineParamsRow, rasterBinStep, rasterHalfBlockBits) gives its per-pixel fragment counts."""
    from lucid_b200 import api
    from tests import parity_util as pu
    small = pu.small_scenes()
    assert len(golden["bin_scenes"]) >= 3
    large_seen = images = 0
    for e in golden["bin_scenes"]:
        sc = small[e["scene"]]
        # the reference's operation order for colour: that is the form the image digest was taken in
        o = pu.run_oracle(sc, mvq=e["max_visible_quads"], threads=4, reference_colour=True)
        try:
            assert [int(o.info[1]), int(o.info[2])] == e["visible"]
            _, counts = api.split_info(o.info, o.bin_count)
            assert counts[0].tolist() == e["quad_counts"]
            assert counts[3].tolist() == e["tri_counts"]
            bq, bt = o.read_bin_lists()
            assert [bq.size, bt.size] == e["list_entries"]
            for values, cnt, want in ((bq, counts[0], e["bin_quads_sha256"]), (bt, counts[3], e["bin_tris_sha256"])):
                canon = np.ascontiguousarray(pu.canonical_lists(values, cnt), np.uint32)
                assert hashlib.sha256(canon.tobytes()).hexdigest() == want  # every entry of every bin's list
            large_seen += e["visible"][1]
            # raster coverage: fragments per pixel from the reference's scanline / span / half-block functions
            fc = np.ascontiguousarray(o.read_frag_counts(), np.uint32)
            assert np.flatnonzero(o.read_bin_levels() == 4).tolist() == e["high_bins"]  # incl. promoted LOW bins
            assert int(o.info[60]) == e["fragments"] and int(fc.sum()) == e["fragments_in_image"]
            assert hashlib.sha256(fc.tobytes()).hexdigest() == e["frag_counts_sha256"]
            # the image (scenes without textures): the flow of raster_low / raster_high restated around the
            # reference's key, shading and reduction functions gives the checker's RGBA8 pixels
            if e.get("image_sha256"):
                img = np.ascontiguousarray(o.read_image(), np.uint32)
                assert hashlib.sha256(img.tobytes()).hexdigest() == e["image_sha256"]
                images += 1
        finally:
            o.close()
    assert large_seen > 0  # the large-triangle path (per bin row scan) is exercised
    assert images >= 2


def test_reference_library_matches_golden_when_available(golden):
    """Where the reference tree is mounted and oracle/_ref/libref_shaders.so is built, the library itself
    reproduces the committed vectors (guards against a stale JSON)."""
    path = os.path.join(HERE, "..", "oracle", "_ref", "libref_shaders.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_shaders.so not built (needs /root/reference)")
    lib = C.CDLL(path)
    lib.ref_process_quad.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.ref_store_tri.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
    for case in golden["cases"]:
        cfg = np.array(case["config_words"], np.uint32)
        for q in case["quads"]:
            pos = np.array(q["pos"], np.uint32)
            idx = np.array(q["idx"], np.uint32)
            res = np.zeros(5, np.uint32)
            lib.ref_process_quad(ptr(cfg), case["width"], case["height"], ptr(pos), ptr(idx), ptr(res))
            assert res.tolist() == q["process_quad"]
            for t in q.get("tris", []):
                rec = np.zeros(21, np.uint32)
                lib.ref_store_tri(ptr(cfg), ptr(np.array(t["tri"], np.uint32)), t["flags_id"], t["y_aabb"], ptr(rec))
                assert rec.tolist() == t["record"]
