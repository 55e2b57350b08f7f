"""Parity tests proper (B200): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs -- bit-exact for every integer product the north star names (bin counts,
per-bin lists, fragment counts, statistics) and <= 1/255 per channel for the blended image --
plus committed golden digests and size-independent properties at BASELINE.json's full sizes."""
import json
import os

import numpy as np
import pytest

from lucid_b200 import api, multigpu, scenes
from tests import parity_util as pu
from tests.golden.make_oracle_golden import digest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = ["soup", "soup_close", "planes", "meshlets", "hairball", "arch"]


@pytest.fixture(scope="module")
def small():
    return pu.small_scenes()


def _clean(bad):
    return {k: v for k, v in bad.items() if not k.startswith("_")}


@pytest.mark.parametrize("name", SCENES)
def test_matches_oracle(name, small):
    o = pu.run_oracle(small[name])
    r, img = pu.run_cuda(small[name])
    try:
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


@pytest.mark.parametrize("name", SCENES)
def test_matches_committed_golden(name, small):
    """Same digests as tests/golden/oracle_golden.json, without running the oracle."""
    with open(os.path.join(HERE, "golden", "oracle_golden.json")) as f:
        g = json.load(f)[name]
    r, img = pu.run_cuda(small[name])
    try:
        info = r.read_info()
        st = api.decode_stats(info, r.bin_count, r.width, r.height)
        assert {k: st[k] for k in g["stats"]} == g["stats"]
        _, counts = api.split_info(info, r.bin_count)
        assert digest(counts[:6]) == g["bin_counts"]
        bq, bt = r.read_bin_lists(st["bin_quads"], st["bin_tris"])
        bq, bt = pu.canonical_lists(bq, counts[0]), pu.canonical_lists(bt, counts[3])
        assert digest(bq) == g["bin_quads"] and digest(bt) == g["bin_tris"]
        assert digest(r.read_frag_counts()) == g["frag_counts"]
        ns, nl = st["visible_small"], st["visible_large"]
        assert digest(np.concatenate([r.read_quad_aabbs(0, ns), r.read_quad_aabbs(1, nl)])) == g["quad_aabbs"]
        assert digest(np.concatenate([r.read_tri_records(0, ns), r.read_tri_records(1, nl)])) == g["tri_records"]
        assert digest(img) == g["image"]  # the fp contract makes even the colours bit-identical
    finally:
        r.close()


@pytest.mark.parametrize("opts", [api.OPT_ADDITIVE_BLENDING, api.OPT_VISUALIZE_ERRORS, api.OPT_ALPHA_THRESHOLD])
@pytest.mark.parametrize("name", ["soup_close", "hairball", "arch"])
def test_render_options(name, opts, small):
    """LucidRenderOpt variants (lucid_renderer.cpp:153-183): additive blending, error
    visualisation (stats[2]) and the alpha-threshold early out."""
    o = pu.run_oracle(small[name], opts=opts)
    r, img = pu.run_cuda(small[name], opts=opts)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


def test_backface_culling(small):
    sc = small["soup"]
    cfg, inst, cols, rects = api.prepare_frame(sc)
    cfg.enable_backface_culling = 1
    from oracle.binding import Oracle
    o = Oracle(sc["width"], sc["height"], 0, 1 << 20, threads=8)
    o.set_scene(sc)
    o.render(cfg, inst, cols, rects)
    r = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    r.set_scene(sc)
    img = np.zeros((sc["height"], sc["width"]), np.uint32)
    r.render(cfg, inst, cols, rects, out=img, flags=api.RENDER_FRAG_COUNTS)
    try:
        st = r.getStats()
        assert st["rejected_backface"] > 0
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


def test_deterministic_and_idempotent(small):
    """Two renders of the same frame (and a render after a different frame) give identical bytes:
    no result depends on atomic arrival order."""
    sc = small["hairball"]
    r, img0 = pu.run_cuda(sc)
    try:
        info0 = r.read_info().copy()
        _, cnt0 = api.split_info(info0, r.bin_count)
        bq0, bt0 = r.read_bin_lists(int(api.decode_stats(info0, r.bin_count, r.width, r.height)["bin_quads"]), 0)
        bq0 = pu.canonical_lists(bq0, cnt0[0])
        other = dict(sc["camera"], rot_h=1.7)
        pu.run_cuda(sc, camera=other, renderer=r)
        _, img1 = pu.run_cuda(sc, renderer=r)
        info1 = r.read_info()
        bq1, _ = r.read_bin_lists(bq0.size, 0)
        bq1 = pu.canonical_lists(bq1, cnt0[0])
        assert np.array_equal(img0, img1)
        assert np.array_equal(info0[:64], info1[:64])
        # the six per-bin counter arrays are cleared every frame; the level lists are only
        # overwritten up to their counts (lucid_renderer.cpp:437)
        _, c0 = api.split_info(info0, r.bin_count)
        _, c1 = api.split_info(info1, r.bin_count)
        assert np.array_equal(c0[:6], c1[:6])
        n_low, n_high = int(info0[7]), int(info0[9])
        assert np.array_equal(c0[7][:n_low], c1[7][:n_low]) and np.array_equal(c0[9][:n_high], c1[9][:n_high])
        assert np.array_equal(bq0, bq1)
    finally:
        r.close()


def test_edge_cases():
    sc = scenes.quad_soup(num_quads=64, width=96, height=64)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(96, 64, 0, 1 << 16)
    try:
        with pytest.raises(api.LucidError):  # render before geometry
            r.render(cfg, inst, cols, rects)
        r.set_scene(sc)
        img = np.zeros((64, 96), np.uint32)
        r.render(cfg, inst[:0], cols[:0], rects[:0], out=img)  # empty instance list
        st = r.getStats()
        assert st["input_quads"] == 0 and st["empty_bins"] == r.bin_count
        assert (img == 0xFF1E1E00).all()
        bad_inst = inst.copy()
        bad_inst[0, 2] = 2000  # more than 1024 quads per instance
        with pytest.raises(api.LucidError):
            r.render(cfg, bad_inst, cols, rects)
        bad_inst = inst.copy()
        bad_inst[0, 0] = 10 ** 6  # index offset outside the index buffer
        with pytest.raises(api.LucidError):
            r.render(cfg, bad_inst, cols, rects)
    finally:
        r.close()
    # degenerate quads and MAX_VISIBLE_QUADS overflow follow the oracle
    sc2 = dict(sc)
    q = sc["quads"].copy()
    q[:8, 1] = q[:8, 0]
    q[:8, 3] = q[:8, 2]
    sc2["quads"] = q
    for scene, mvq in ((sc2, 1 << 16), (sc, 16)):
        o = pu.run_oracle(scene, mvq=mvq)
        r, img = pu.run_cuda(scene, mvq=mvq)
        try:
            assert _clean(pu.compare(r, img, o)) == {}
        finally:
            r.close()


def test_promotion_low_to_high():
    sc = scenes.planes(num_planes=200, width=128, height=96, plane_size=0.12, plane_dist=0.02)
    o = pu.run_oracle(sc)
    r, img = pu.run_cuda(sc)
    try:
        assert r.getStats()["promoted_bins"] >= 1
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


def test_high_bin_limits_paint_red():
    """More than 4096 triangles in one half-block (raster_high.glsl:140-141): the bin is red."""
    sc = scenes.planes(num_planes=2200, width=64, height=64, plane_size=0.05, plane_dist=0.002)
    o = pu.run_oracle(sc, mvq=1 << 16)
    r, img = pu.run_cuda(sc, mvq=1 << 16)
    try:
        assert (o.read_bin_levels() == 5).any()
        assert (img == 0x000000FF).any()
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


def test_bin_row_split_matches_full_frame(small):
    """a renderer that owns bin rows [a,b) reproduces exactly those rows; storing every
    strip into one image gives the single-GPU frame."""
    sc = small["soup_close"]
    full_r, full_img = pu.run_cuda(sc)
    full_r.close()
    nby = (sc["height"] + 31) // 32
    cfg, inst, cols, rects = api.prepare_frame(sc)
    # one device image shared by all "ranks": the composite is just stores into disjoint rows
    target = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    ptr, pitch = target.image_pointer()
    frags = 0
    try:
        for rows in multigpu.split_bin_rows(nby, 3):
            part = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20, bin_rows=rows)
            part.set_scene(sc)
            part.render(cfg, inst, cols, rects, out_device_ptr=ptr, out_pitch=pitch)
            frags += part.getStats()["fragments"]
            o = pu.run_oracle(sc, bin_rows=rows)
            pi = part.read_info()
            for lo, hi in ((0, 10), (32, 36), (60, 63)):  # counts, rejections, statistics
                assert np.array_equal(pi[lo:hi], o.info[lo:hi])
            part.close()
        assert np.array_equal(target.read_image(), full_img)
    finally:
        target.close()


@pytest.mark.parametrize("config", [0, 1])
def test_full_size_config_against_oracle(config):
    """BASELINE.json configs[0] and [1] at full size against the oracle (a few seconds of CPU)."""
    sc = scenes.get_config(config)
    o = pu.run_oracle(sc, mvq=4793490, threads=os.cpu_count())
    r, img = pu.run_cuda(sc, mvq=4793490)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


def test_full_size_hairball_properties():
    """configs[2] (5M triangles, 4K): size-independent properties -- verifyInfo offsets, list
    sortedness, the fragment statistic equals the sum of the per-pixel fragment image, no bin over
    the reference's limits, deterministic image."""
    sc = scenes.get_config(2)
    r, img = pu.run_cuda(sc, mvq=4793490)
    try:
        info = r.read_info()
        st = api.decode_stats(info, r.bin_count, r.width, r.height)
        assert r.verifyInfo(info) == []
        assert st["high_bins"] > 500 and st["dropped_quads"] == 0 and st["list_overflow"] == 0
        assert (img != 0x000000FF).all()  # no red (overflow) bins
        assert 3840 % 32 == 0 and 2160 % 32 != 0
        fc = r.read_frag_counts()
        assert int(fc.sum()) <= st["fragments"] <= int(fc.sum()) * 1.02
        _, counts = api.split_info(info, r.bin_count)
        bq, _ = r.read_bin_lists(st["bin_quads"], 0)
        offs, cnts = counts[1], counts[0]
        for b in np.argsort(cnts)[-20:]:  # list entries are unique visible small-quad slots
            seg = np.sort(bq[offs[b]:offs[b] + cnts[b]] & 0x0FFFFFFF)
            assert (np.diff(seg.astype(np.int64)) > 0).all() and seg[-1] < st["visible_small"]
        _, img2 = pu.run_cuda(sc, renderer=r)
        assert np.array_equal(img, img2)
    finally:
        r.close()


def test_full_size_architecture_properties():
    """configs[3] (10M triangles, textured, large wall/floor triangles, 4K): verifyInfo offsets, unique
    list entries, the fragment statistic against the per-pixel fragment image, no overflow, a deterministic
    image, and a bin-row strip rendered on its own reproducing exactly its rows of the full frame."""
    sc = scenes.get_config(3)
    r, img = pu.run_cuda(sc, mvq=4793490)
    try:
        info = r.read_info()
        st = api.decode_stats(info, r.bin_count, r.width, r.height)
        assert r.verifyInfo(info) == []
        assert st["visible_large"] > 1000 and st["bin_tris"] > 100000
        assert st["dropped_quads"] == 0 and st["list_overflow"] == 0 and st["invalid_pixels"] == 0
        assert (img != 0x000000FF).all()
        fc = r.read_frag_counts()
        assert int(fc.sum()) <= st["fragments"] <= int(fc.sum()) * 1.02
        _, counts = api.split_info(info, r.bin_count)
        bq, bt = r.read_bin_lists(st["bin_quads"], st["bin_tris"])
        mvq = 4793490
        for b in np.argsort(counts[3])[-20:]:  # large-triangle lists: unique triangles of large-quad slots
            seg = np.sort(bt[counts[4][b]:counts[4][b] + counts[3][b]])
            assert (np.diff(seg.astype(np.int64)) > 0).all()
            assert (seg >> 1).min() >= mvq - st["visible_large"] and (seg >> 1).max() < mvq
        assert int(counts[3].sum()) == st["bin_tris"] and int(counts[0].sum()) == st["bin_quads"]
        _, img2 = pu.run_cuda(sc, renderer=r)
        assert np.array_equal(img, img2)
        # rows [20, 41) on their own: same pixels, same per-bin counts
        rows = (20, 41)
        cfg, inst, cols, rects = api.prepare_frame(sc)
        r.set_bin_rows(*rows)
        strip = np.zeros_like(img)
        r.render(cfg, inst, cols, rects, out=strip)
        _, pc = api.split_info(r.read_info(), r.bin_count)
        y0, y1 = rows[0] * 32, rows[1] * 32
        assert np.array_equal(strip[y0:y1], img[y0:y1])
        bcx = (sc["width"] + 31) // 32
        for which in (0, 3):
            assert np.array_equal(pc[which][rows[0] * bcx:rows[1] * bcx], counts[which][rows[0] * bcx:rows[1] * bcx])
            assert pc[which][:rows[0] * bcx].sum() == 0 and pc[which][rows[1] * bcx:].sum() == 0
    finally:
        r.close()


def test_frames_without_stage_events_and_row_costs(small):
    """LUCID_RENDER_NO_STAGE_TIMES changes timing bookkeeping only; lucid_read_row_costs reports raster
    cost exactly for the bin rows that hold work."""
    sc = small["arch"]
    r, img = pu.run_cuda(sc)
    try:
        cfg, inst, cols, rects = api.prepare_frame(sc)
        img2 = np.zeros_like(img)
        r.render(cfg, inst, cols, rects, out=img2, flags=api.RENDER_NO_STAGE_TIMES)
        assert np.array_equal(img, img2)
        ms = r.stage_times()
        assert ms[7] > 0 and (ms[:7] == 0).all()
        r.render(cfg, inst, cols, rects, out=img2)
        ms = r.stage_times()
        assert ms[7] > 0 and ms[0] > 0 and ms[5] > 0
        cost = r.read_row_costs()
        _, counts = api.split_info(r.read_info(), r.bin_count)
        bcx, bcy = (sc["width"] + 31) // 32, (sc["height"] + 31) // 32
        work = (counts[0] + counts[3]).reshape(bcy, bcx).sum(axis=1)
        assert cost.shape == (bcy,) and ((cost > 0) == (work > 0)).all()
        # balanced boundaries from these costs cover every row exactly once
        parts = multigpu.split_bin_rows(bcy, 3, cost.astype(np.float64))
        assert parts[0][0] == 0 and parts[-1][1] == bcy and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    finally:
        r.close()


@pytest.mark.parametrize("view", [21, 42])
def test_orbit_views_against_oracle(view):
    """configs[4]: other views of the 64-view orbit of the 1M-triangle scene, full size, against the oracle."""
    sc = scenes.get_config(1)
    cam = dict(sc["camera"], rot_h=sc["camera"]["rot_h"] + 2.0 * np.pi * view / 64)
    o = pu.run_oracle(sc, mvq=4793490, threads=os.cpu_count(), camera=cam)
    r, img = pu.run_cuda(sc, mvq=4793490, camera=cam)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()


def test_bin_range_split_matches_full_frame(small):
    """Ownership finer than rows (lucid_set_bin_range): row-major bin ranges that cut through bin rows
    compose to the single-GPU frame, and the counts of owned bins equal the full frame's."""
    sc = small["arch"]
    full_r, full_img = pu.run_cuda(sc)
    _, full_counts = api.split_info(full_r.read_info(), full_r.bin_count)
    full_frags = full_r.getStats()["fragments"]
    cost = full_r.read_bin_costs().astype(np.float64)
    bc = full_r.bin_count
    full_r.close()
    assert (cost > 0).sum() > 0
    cfg, inst, cols, rects = api.prepare_frame(sc)
    target = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    ptr, pitch = target.image_pointer()
    frags = 0
    try:
        ranges = multigpu.split_bins(bc, 5, cost)
        assert ranges[0][0] == 0 and ranges[-1][1] == bc and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        bcx = (sc["width"] + 31) // 32
        assert any(a % bcx != 0 for a, _ in ranges[1:])  # at least one boundary inside a bin row
        for n, (lo, hi) in enumerate(ranges):
            part = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
            part.set_scene(sc)
            part.set_bin_range(lo, hi)
            if n % 2 == 0:  # the raster kernels store into the shared image ...
                part.render(cfg, inst, cols, rects, out_device_ptr=ptr, out_pitch=pitch)
            else:  # ... or the owned bins are copied there after the frame (lucid_composite_to)
                part.render(cfg, inst, cols, rects)
                part.composite_to(ptr, pitch)
            frags += part.getStats()["fragments"]
            pi = part.read_info()
            _, pc = api.split_info(pi, bc)
            for which in (0, 3):
                assert np.array_equal(pc[which][lo:hi], full_counts[which][lo:hi])
                assert pc[which][:lo].sum() == 0 and pc[which][hi:].sum() == 0
            o = pu.run_oracle(sc, bin_range=(lo, hi))  # the oracle with the same ownership
            for a, b in ((0, 10), (32, 36), (60, 63)):  # counts, rejections, statistics
                assert np.array_equal(pi[a:b], o.info[a:b])
            _, oc = api.split_info(o.info, bc)
            for which in (0, 1, 3, 4):
                assert np.array_equal(pc[which], oc[which])
            part.close()
        assert np.array_equal(target.read_image(), full_img)
        assert frags == full_frags
    finally:
        target.close()
