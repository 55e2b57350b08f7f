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
SCENES = ["soup", "soup_close", "planes", "meshlets", "hairball", "arch", "boxes"]
FULL_MVQ = 4793490


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
        assert _golden_mismatches(r, img, g) == []
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


@pytest.mark.parametrize("name", ["soup", "soup_close", "meshlets", "arch", "planes"])
@pytest.mark.parametrize("extra", [0, api.OPT_VISUALIZE_ERRORS])
def test_opaque_prepass_option(name, extra, small):
    """LUCID_OPT_OPAQUE_PREPASS (the TODO of shared/shading.glsl:31-32 as an option): samples behind the
    nearest INST_IS_OPAQUE sample of their pixel are dropped before sorting and shading.  Against the oracle's own
    pre-pass mode: per-pixel counts of the surviving samples, stats and image are identical; against the plain
    frame: setup and binning products are untouched, fewer fragments, and the image can only differ by more than
    1/255 where the reference's 3-entry window was too small (pixels VISUALIZE_ERRORS flags); 1/255 is what a
    flagged-opaque sample leaks when its interpolated vertex alpha truncates to 254."""
    sc = small[name]
    opts = api.OPT_OPAQUE_PREPASS | extra
    o = pu.run_oracle(sc, opts=opts)
    r, img = pu.run_cuda(sc, opts=opts)
    plain_r, plain_img = pu.run_cuda(sc, opts=extra)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
        st, st0 = api.decode_stats(r.read_info(), r.bin_count, r.width, r.height), plain_r.getStats()
        assert st["fragments"] <= st0["fragments"] and st["half_block_tris"] == st0["half_block_tris"]
        if name in ("soup", "meshlets", "arch"):
            assert st["fragments"] < st0["fragments"]
        assert (r.read_frag_counts() <= plain_r.read_frag_counts()).all()
        if extra == 0:
            ev_r, _ = pu.run_cuda(sc, opts=api.OPT_VISUALIZE_ERRORS)
            invalid = ev_r.getStats()["invalid_pixels"]
            ev_r.close()
            d = np.abs(img.view(np.uint8).astype(np.int32) - plain_img.view(np.uint8).astype(np.int32))
            differing = int((d.reshape(img.shape + (4,)).max(axis=-1) > 1).sum())
            assert differing <= invalid, (differing, invalid)
    finally:
        r.close()
        plain_r.close()


def test_opaque_prepass_closed_form():
    """Stacked planes with layer k opaque: k + 1 surviving samples wherever all layers overlap."""
    from tests.test_oracle import planes_with_opaque_layer
    k, n = 5, 12
    sc = planes_with_opaque_layer(k, n)
    r, img = pu.run_cuda(sc, opts=api.OPT_OPAQUE_PREPASS)
    r0, img0 = pu.run_cuda(sc)
    try:
        f1, f0 = r.read_frag_counts(), r0.read_frag_counts()
        assert (f1[f0 == n] == k + 1).all() and int(r.read_info()[60]) == int(f1.sum())
        assert np.abs(img.view(np.uint8).astype(np.int32) - img0.view(np.uint8).astype(np.int32)).max() <= 1
    finally:
        r.close()
        r0.close()


def test_opaque_prepass_is_ignored_where_it_would_change_the_image(small):
    """Additive blending has no transmittance and the alpha-threshold build stops on segment boundaries: the option
    is ignored there (include/lucid_abi.h), frames equal the ones without it."""
    sc = small["soup_close"]
    for other in (api.OPT_ADDITIVE_BLENDING, api.OPT_ALPHA_THRESHOLD):
        r, img = pu.run_cuda(sc, opts=other | api.OPT_OPAQUE_PREPASS)
        r0, img0 = pu.run_cuda(sc, opts=other)
        try:
            assert np.array_equal(img, img0) and np.array_equal(r.read_info()[60:63], r0.read_info()[60:63])
            assert np.array_equal(r.read_frag_counts(), r0.read_frag_counts())
        finally:
            r.close()
            r0.close()


def test_opaque_prepass_full_size_architecture():
    """configs[3] at full size with the pre-pass against the oracle's pre-pass mode: 357.6 M fragments become the
    ~10 % that lie in front of the nearest opaque sample."""
    sc = scenes.get_config(3)
    o = pu.run_oracle(sc, opts=api.OPT_OPAQUE_PREPASS, mvq=FULL_MVQ, threads=os.cpu_count())
    r, img = pu.run_cuda(sc, opts=api.OPT_OPAQUE_PREPASS, mvq=FULL_MVQ)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
        st = api.decode_stats(r.read_info(), r.bin_count, r.width, r.height)
        assert st["fragments"] < 357_598_648 // 4
    finally:
        r.close()


@pytest.mark.parametrize("name", ["soup", "soup_close", "meshlets", "arch", "hairball", "planes"])
def test_debug_raster_variant_finds_nothing_on_a_correct_frame(name, small):
    """LUCID_OPT_DEBUG_RASTER (the reference's raster_*_debug pipelines): the checks of the shaders' DEBUG_ENABLED code
    -- no block-list entry without coverage (raster_low.glsl:125-126), keys strictly increasing after the sort
    (raster_low.glsl:154-163) -- run over every list of the frame and record nothing; the frame is the plain one."""
    sc = small[name]
    r, img = pu.run_cuda(sc, opts=api.OPT_DEBUG_RASTER)
    r0, img0 = pu.run_cuda(sc)
    try:
        recs, n = r.read_debug_records()
        assert n == 0, recs[:4]
        assert np.array_equal(img, img0) and np.array_equal(r.read_info()[60:63], r0.read_info()[60:63])
    finally:
        r.close()
        r0.close()


def test_debug_raster_variant_records_a_violation(small, monkeypatch):
    """The recording path itself: with the test hook the sorted keys of the frame's first work item are swapped once
    before the check, which then writes DEBUG_RECORD(i, tri_count, prev_value, value) for position 1."""
    monkeypatch.setenv("LUCID_DEBUG_RASTER_INJECT", "1")
    sc = small["soup_close"]
    r, img = pu.run_cuda(sc, opts=api.OPT_DEBUG_RASTER)
    monkeypatch.delenv("LUCID_DEBUG_RASTER_INJECT")
    r0, img0 = pu.run_cuda(sc)
    try:
        recs, n = r.read_debug_records()
        assert n >= 1 and (recs[:, 0] == 2).all()  # LUCID_DEBUG_UNSORTED
        first = recs[recs[:, 3] == 1][0]
        assert first[4] > 3 and first[6] <= first[5]  # a list the sort applies to; value <= prev_value
        assert np.array_equal(img, img0)  # the swap is undone after the check
        with pytest.raises(RuntimeError):
            r0.read_debug_records()
    finally:
        r.close()
        r0.close()


@pytest.mark.parametrize("name", ["soup", "soup_close", "meshlets", "arch", "hairball", "planes"])
def test_compact_block_lists_render_the_same_frame(name, small):
    """LUCID_CREATE_COMPACT_LISTS: block lists in a pool sized by max_block_entries (count pass, one allocation per
    bin, fill pass) instead of a fixed 1 MiB per bin.  Everything the frame produces equals the oracle's, as with the
    fixed slots (LOW and HIGH bins, promoted bins, the pre-pass)."""
    sc = small[name]
    o = pu.run_oracle(sc)
    r, img = pu.run_cuda(sc, create_flags=api.CREATE_COMPACT_LISTS)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()
    if name in ("arch", "soup"):
        op = pu.run_oracle(sc, opts=api.OPT_OPAQUE_PREPASS)
        r, img = pu.run_cuda(sc, opts=api.OPT_OPAQUE_PREPASS, create_flags=api.CREATE_COMPACT_LISTS)
        try:
            assert _clean(pu.compare(r, img, op)) == {}
        finally:
            r.close()


def test_compact_block_lists_full_size_and_pool_limit():
    """configs[3] at full size with a pool of 2^25 entries (512 MB of lists where the fixed slots take 8 GiB): the
    oracle's frame; with a pool that is too small the bins that did not fit are red and the frame says so."""
    sc = scenes.get_config(3)
    o = pu.run_oracle(sc, mvq=FULL_MVQ, threads=os.cpu_count())
    r, img = pu.run_cuda(sc, mvq=FULL_MVQ, create_flags=api.CREATE_COMPACT_LISTS, max_block_entries=1 << 25)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
    finally:
        r.close()
    small_sc = scenes.quad_soup(num_quads=20_000, width=640, height=360, distance=12.0, seed=3)
    cfg, inst, cols, rects = api.prepare_frame(small_sc)
    r = api.LucidRenderer(640, 360, 0, 1 << 20, create_flags=api.CREATE_COMPACT_LISTS, max_block_entries=4096)
    try:
        r.set_scene(small_sc)
        out = np.zeros((360, 640), np.uint32)
        with pytest.raises(api.LucidError):
            r.render(cfg, inst, cols, rects, out=out)
        assert (r.read_image() == 0x000000FF).any()  # red bins
    finally:
        r.close()


def coplanar_stack(n=700, width=320, height=200):
    """n coplanar, nearly coincident half-transparent quads with a colour each: every interior block's list is one long
    run of equal depth keys, and the blend order -- by triangle index -- decides the image."""
    rng = np.random.default_rng(12)
    base = np.array([[-1.0, -0.6, 0.0], [1.0, -0.6, 0.0], [1.0, 0.6, 0.0], [-1.0, 0.6, 0.0]], np.float32)
    pos = np.concatenate([base + np.array([0.0007 * k, 0.0004 * k, 0.0], np.float32) for k in range(n)])
    quads = np.arange(4 * n, dtype=np.uint32).reshape(n, 4)
    cols = np.repeat(rng.integers(0, 1 << 24, n, dtype=np.uint32) | np.uint32(0x60000000), 4)  # alpha 96 / 255
    return scenes._scene(pos, quads, [(0, n, 0, scenes.INST_HAS_VERTEX_COLORS)], [((1.0, 1.0, 1.0), 1.0, (0.0, 0.0, 1.0, 1.0))],
                         dict(kind="orbit", center=(0.0, 0.0, 0.0), distance=3.0, rot_h=0.0, rot_v=0.0), width, height,
                         colors=cols, name="coplanar_stack")


@pytest.mark.parametrize("n", [40, 150, 300, 700])
def test_long_runs_of_depth_ties_follow_the_triangle_index(n):
    """Depth-key ties are ordered by triangle index (the canonical form of SURVEY 8c).  Coplanar stacks make runs as
    long as the list: short ones are fixed by one lane inside k_block_sort, runs of 96 entries or more (24 where the
    triangle indices are not in shared memory) are queued and
    ranked by a warp each in the sorted-entry stream (k_tie_runs) -- lists in shared memory (<= 512 entries), in the
    shared-memory key array with the list in global memory (<= 1024), and in the L2 key array (700 entries of a HIGH
    bin stay below that; the promoted bins of n = 700 cover the rest of the paths)."""
    sc = coplanar_stack(n)
    o = pu.run_oracle(sc)
    r, img = pu.run_cuda(sc)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
        assert r.getStats()["fragments"] > 10_000 * n // 40
    finally:
        r.close()


def test_timers_option(small):
    """LUCID_OPT_TIMERS (the reference's `_timers` shader variants, shared/timers.glsl, lucid_renderer.cpp:754-762):
    the phases' clock ticks land in LucidInfo.setup_timers / bin_dispatcher_timers / raster_timers in the reference's
    slot meaning; without the option they stay zero; results are the same either way."""
    sc = small["arch"]
    plain_r, plain_img = pu.run_cuda(sc)
    r, img = pu.run_cuda(sc, opts=api.OPT_TIMERS)
    try:
        assert np.array_equal(img, plain_img)
        st, st0 = r.getStats(), plain_r.getStats()
        for group in ("setup_timers", "bin_dispatcher_timers", "raster_timers"):
            assert all(v > 0 for v in st[group].values()), (group, st[group])
            assert all(v == 0 for v in st0[group].values())
        info, info0 = r.read_info(), plain_r.read_info()
        assert np.array_equal(info[0:36], info0[0:36]) and np.array_equal(info[60:63], info0[60:63])
        # shading is the bulk of the raster work on this scene
        rt = st["raster_timers"]
        assert rt["shade and reduce"] > rt["finish reduce"]
    finally:
        r.close()
        plain_r.close()


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


def _golden_mismatches(r, img, g):
    """The CUDA frame against one committed record of tests/golden/oracle_golden.json."""
    bad = []
    info = r.read_info()
    st = api.decode_stats(info, r.bin_count, r.width, r.height)
    if {k: st[k] for k in g["stats"]} != g["stats"]:
        bad.append("stats")
    _, counts = api.split_info(info, r.bin_count)
    bq, bt = r.read_bin_lists(st["bin_quads"], st["bin_tris"])
    ns, nl = st["visible_small"], st["visible_large"]
    for name, value in (("bin_counts", counts[:6]), ("bin_quads", pu.canonical_lists(bq, counts[0])),
                        ("bin_tris", pu.canonical_lists(bt, counts[3])), ("frag_counts", r.read_frag_counts()),
                        ("quad_aabbs", np.concatenate([r.read_quad_aabbs(0, ns), r.read_quad_aabbs(1, nl)])),
                        ("tri_records", pu.canonical_tri_records(np.concatenate([r.read_tri_records(0, ns),
                                                                                  r.read_tri_records(1, nl)]))),
                        ("image", img)):  # the fp contract makes even the colours bit-identical
        if digest(value) != g[name]:
            bad.append(name)
    return bad


@pytest.mark.parametrize("config", [0, 1, 2, 3])
def test_full_size_config_against_oracle(config):
    """BASELINE.json configs[0..3] at FULL size against the oracle run on the same inputs (the two 4K scenes
    take the oracle 10-20 s on the box's host cores), and against the digests committed for them."""
    sc = scenes.get_config(config)
    o = pu.run_oracle(sc, mvq=FULL_MVQ, threads=os.cpu_count())
    r, img = pu.run_cuda(sc, mvq=FULL_MVQ)
    try:
        assert _clean(pu.compare(r, img, o)) == {}
        with open(os.path.join(HERE, "golden", "oracle_golden.json")) as f:
            g = json.load(f)[f"config{config}"]
        assert _golden_mismatches(r, img, g) == []
    finally:
        r.close()


def test_full_size_hairball_is_to_spec():
    """configs[2] as BASELINE.json words it: "depth complexity > 64 per pixel, raster_high path" -- the median
    over the covered pixels is above 64, HIGH bins outnumber LOW bins, no bin is over the reference's limits
    (raster_high.glsl:80-83,140-141), lists longer than the shared-memory sort exist, and the frame is
    deterministic."""
    sc = scenes.get_config(2)
    r, img = pu.run_cuda(sc, mvq=FULL_MVQ)
    try:
        info = r.read_info()
        st = api.decode_stats(info, r.bin_count, r.width, r.height)
        assert r.verifyInfo(info) == []
        assert st["high_bins"] > st["low_bins"] and st["dropped_quads"] == 0 and st["list_overflow"] == 0
        assert (img != 0x000000FF).all()  # no red (overflow) bins
        fc = r.read_frag_counts()
        cov = fc[fc > 0]
        assert cov.size > 0.9 * fc.size and np.median(cov) > 64 and fc.max() > 200
        assert int(fc.sum()) <= st["fragments"] <= int(fc.sum()) * 1.02
        _, img2 = pu.run_cuda(sc, renderer=r)
        assert np.array_equal(img, img2)
    finally:
        r.close()


def test_full_size_architecture_split_in_eight_composes_to_the_oracle_frame():
    """configs[3] (10M triangles, textured, large wall/floor triangles, 4K) the way bench.py --mode split runs
    it on 8 GPUs: eight cost-balanced row-major bin ranges (they cut through bin rows), every range rendered
    on its own into one shared image.  The composite equals the ORACLE's full frame, the owned bins' counts
    equal the oracle's, and the fragment statistics add up to the oracle's."""
    sc = scenes.get_config(3)
    o = pu.run_oracle(sc, mvq=FULL_MVQ, threads=os.cpu_count())
    _, oc = api.split_info(o.info, o.bin_count)
    full_r, full_img = pu.run_cuda(sc, mvq=FULL_MVQ)
    try:
        assert _clean(pu.compare(full_r, full_img, o)) == {}
        cost = full_r.read_bin_costs().astype(np.float64)
        bc = full_r.bin_count
        ranges = multigpu.split_bins(bc, 8, cost)
        bcx = (sc["width"] + 31) // 32
        assert any(a % bcx != 0 for a, _ in ranges[1:])
        cfg, inst, cols, rects = api.prepare_frame(sc)
        # the full-frame renderer becomes the eight "ranks" in turn; a second handle only lends its image
        target = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 16)
        ptr, pitch = target.image_pointer()
        try:
            stats = np.zeros(3, np.int64)
            for lo, hi in ranges:
                full_r.set_bin_range(lo, hi)
                # with the instance selection of the split (k_instance_select over 4788 instances: frustum planes and
                # owned rows) and, on the ranges of few heavy bins, the 512-thread list kernel -- as bench.py runs it
                full_r.render(cfg, inst, cols, rects, out_device_ptr=ptr, out_pitch=pitch, flags=api.RENDER_CULL_INSTANCES)
                pi = full_r.read_info()
                _, pc = api.split_info(pi, bc)
                for which in (0, 3):
                    assert np.array_equal(pc[which][lo:hi], oc[which][lo:hi])
                    assert pc[which][:lo].sum() == 0 and pc[which][hi:].sum() == 0
                stats += pi[60:63].astype(np.int64)
            assert np.array_equal(stats, o.info[60:63].astype(np.int64))
            composite = target.read_image()
            assert np.array_equal(composite, full_img)
            d = np.abs(composite.view(np.uint8).astype(np.int32) - o.read_image().view(np.uint8).astype(np.int32))
            assert d.max() <= 1
        finally:
            target.close()
    finally:
        full_r.close()


def test_bin_list_overflow_is_contained():
    """More per-bin list entries than the lists hold (2 * max_visible_quads): nothing is rasterised from the
    unfilled lists, the frame is painted red and the call reports LUCID_E_LIMIT -- no out-of-bounds read."""
    sc = scenes.planes(num_planes=40, width=1280, height=720, plane_size=40.0, plane_dist=0.01)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(sc["width"], sc["height"], 0, 1000)  # capacity 2000 entries; one plane covers 920 bins
    try:
        r.set_scene(sc)
        img = np.zeros((sc["height"], sc["width"]), np.uint32)
        with pytest.raises(api.LucidError) as e:
            r.render(cfg, inst, cols, rects, out=img)
        assert "(-3)" in str(e.value) and "max_visible_quads" in str(e.value)
        r.wait()  # the error is reported once; the handle stays usable
        assert (r.read_image() == 0x000000FF).all()
        info = r.read_info()
        assert api.decode_stats(info, r.bin_count, r.width, r.height)["list_overflow"] == 1
        # a frame that fits renders normally afterwards
        small = scenes.planes(num_planes=1, width=1280, height=720, plane_size=0.5)
        r.set_scene(small)
        cfg2, inst2, cols2, rects2 = api.prepare_frame(small)
        r.render(cfg2, inst2, cols2, rects2, out=img)
        assert (img != 0x000000FF).all() and r.getStats()["fragments"] > 0
    finally:
        r.close()


def test_bad_vertex_references_are_rejected():
    """vertex_offset outside the vertex buffer is an argument error; a stale index inside a quad rejects that
    quad (counted as REJECTION_OTHER) instead of reading out of bounds."""
    sc = scenes.quad_soup(num_quads=64, width=96, height=64)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(96, 64, 0, 1 << 16)
    try:
        r.set_scene(sc)
        bad_inst = inst.copy()
        bad_inst[0, 1] = sc["positions"].shape[0]
        with pytest.raises(api.LucidError):
            r.render(cfg, bad_inst, cols, rects)
        r.render(cfg, inst, cols, rects)
        base = r.getStats()
        q = sc["quads"].copy()
        q[:5, 2] = 10 ** 7
        r.set_geometry(sc["positions"], q)
        r.render(cfg, inst, cols, rects)
        st = r.getStats()
        assert st["rejected_other"] == base["rejected_other"] + 5
    finally:
        r.close()


def test_image_pointer_names_the_device_resident_image(small):
    """lucid_image_pointer / lucid_ipc_export_image always name the image LUCID_MEM_NONE frames render into,
    also after frames that were read back to the host; lucid_read_image after a frame rendered into a caller's
    device image is a state error."""
    sc = small["soup_close"]
    r, img = pu.run_cuda(sc)  # a frame copied to the host (alternating images)
    other = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 16)
    try:
        p0, _ = r.image_pointer()
        cfg, inst, cols, rects = api.prepare_frame(sc)
        r.render(cfg, inst, cols, rects, out=np.zeros_like(img))
        assert r.image_pointer()[0] == p0
        r.render(cfg, inst, cols, rects)  # MEM_NONE: renders into the named image
        assert np.array_equal(r.read_image(), img)
        ptr, pitch = other.image_pointer()
        r.render(cfg, inst, cols, rects, out_device_ptr=ptr, out_pitch=pitch)
        with pytest.raises(api.LucidError):
            r.read_image()
        assert np.array_equal(other.read_image(), img)
    finally:
        r.close()
        other.close()


def test_instance_culling_of_the_split_changes_nothing_but_the_rejection_counters(small):
    """LUCID_RENDER_CULL_INSTANCES: instances whose bounding box projects outside the owned bin rows are skipped
    before their quads are loaded.  Visible quads, per-bin counts, lists and pixels of the owned bins are the ones
    of the plain split; only num_rejected_quads shrinks to the processed instances."""
    for name in ("arch", "meshlets"):
        sc = small[name]
        cfg, inst, cols, rects = api.prepare_frame(sc)
        nby = (sc["height"] + 31) // 32
        r = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
        try:
            r.set_scene(sc)
            skipped_any = False
            for rows in multigpu.split_bin_rows(nby, 4):
                r.set_bin_rows(*rows)
                a = np.zeros((sc["height"], sc["width"]), np.uint32)
                b = np.zeros_like(a)
                r.render(cfg, inst, cols, rects, out=a)
                ia = r.read_info().copy()
                bqa, bta = r.read_bin_lists(int(api.split_info(ia, r.bin_count)[1][0].sum()),
                                            int(api.split_info(ia, r.bin_count)[1][3].sum()))
                for _ in range(2):  # second frame: boxes come from the cache
                    r.render(cfg, inst, cols, rects, out=b, flags=api.RENDER_CULL_INSTANCES)
                ib = r.read_info()
                _, ca = api.split_info(ia, r.bin_count)
                _, cb = api.split_info(ib, r.bin_count)
                y0, y1 = rows[0] * 32, min(rows[1] * 32, sc["height"])
                assert np.array_equal(a[y0:y1], b[y0:y1])
                assert np.array_equal(ia[0:10], ib[0:10]) and np.array_equal(ia[60:63], ib[60:63])
                assert np.array_equal(ca[:6], cb[:6])
                bqb, btb = r.read_bin_lists(bqa.size, bta.size)
                assert np.array_equal(pu.canonical_lists(bqa, ca[0]), pu.canonical_lists(bqb, ca[0]))
                assert np.array_equal(pu.canonical_lists(bta, ca[3]), pu.canonical_lists(btb, ca[3]))
                assert (ib[32:36] <= ia[32:36]).all()
                skipped_any = skipped_any or (ib[32:36] < ia[32:36]).any()
            assert skipped_any  # some instance was dropped by its box
        finally:
            r.close()


def test_instance_culling_of_a_whole_frame_drops_instances_outside_the_frustum(small):
    """Without a split LUCID_RENDER_CULL_INSTANCES still drops the instances whose box lies outside one frustum plane
    (k_instance_select): their quads would all be rejected by processInputQuad's clip-mask test.  Everything but the
    rejection counters is the plain frame's; the kept instances go through the look-back chain in input order, so even
    the visible-quad slots -- and with them the triangle records -- are the same."""
    sc = small["arch"]  # the camera stands inside the room: clusters lie behind it and to its sides
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    try:
        r.set_scene(sc)
        a, b = np.zeros((sc["height"], sc["width"]), np.uint32), np.zeros((sc["height"], sc["width"]), np.uint32)
        r.render(cfg, inst, cols, rects, out=a)
        ia = r.read_info().copy()
        nvis = int(api.decode_stats(ia, r.bin_count, r.width, r.height)["visible_small"])
        recs_a = r.read_tri_records(0, nvis)
        for _ in range(2):
            r.render(cfg, inst, cols, rects, out=b, flags=api.RENDER_CULL_INSTANCES)
        ib = r.read_info()
        recs_b = r.read_tri_records(0, nvis)
        assert np.array_equal(a, b)
        assert np.array_equal(ia[0:10], ib[0:10]) and np.array_equal(ia[60:63], ib[60:63])
        assert np.array_equal(api.split_info(ia, r.bin_count)[1][:6], api.split_info(ib, r.bin_count)[1][:6])
        assert np.array_equal(pu.canonical_tri_records(recs_a), pu.canonical_tri_records(recs_b))
        assert (ib[32:36] <= ia[32:36]).all() and (ib[32:36] < ia[32:36]).any()
    finally:
        r.close()


def test_owned_bins_read_back_composes_the_frame_in_host_memory(small):
    """LUCID_RENDER_OWNED_BINS_ONLY: every range of a split copies only its own bins into a host image that has the
    layout of the whole frame (here: one array, the way the processes of a split share one in /dev/shm); after all
    ranges the array is the full frame.  Ranges are cut inside bin rows, so all three rectangle shapes occur."""
    sc = small["arch"]
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    try:
        r.set_scene(sc)
        full = np.zeros((sc["height"], sc["width"]), np.uint32)
        r.render(cfg, inst, cols, rects, out=full)
        n = r.bin_count
        nbx = (sc["width"] + 31) // 32
        cuts = [0, nbx // 2, nbx // 2 + 3, 2 * nbx + 5, 5 * nbx, n - 2, n]
        host = np.full_like(full, 0xDEADBEEF)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            r.set_bin_range(lo, hi)
            before = host.copy()
            r.render(cfg, inst, cols, rects, out=host, flags=api.RENDER_OWNED_BINS_ONLY)
            changed = host != before
            by, bx = np.nonzero(changed)
            owned = ((by // 32) * nbx + bx // 32)
            assert ((owned >= lo) & (owned < hi)).all()  # nothing outside the owned bins was written
        assert np.array_equal(host, full)
    finally:
        r.close()


def test_frames_in_flight_on_two_handles_render_the_same_frames(small):
    """Two handles on their own streams render alternate frames of an orbit (what bench.py does on every GPU), without
    programmatic dependent launch (LUCID_RENDER_NO_DEPENDENT_LAUNCH); every frame equals the one a single handle
    renders with the default launch mode, whichever handle rendered it."""
    sc = small["arch"]
    base = sc["camera"]  # kind "lookat": a camera that walks sideways
    cams = [dict(base, pos=(base["pos"][0] + 0.7 * k, base["pos"][1], base["pos"][2] - 0.4 * k)) for k in range(6)]
    frames = [api.prepare_frame(sc, c) for c in cams]
    ref = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    a = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    b = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20)
    try:
        for r in (ref, a, b):
            r.set_scene(sc)
        want = []
        for cfg, inst, cols, rects in frames:
            img = np.zeros((sc["height"], sc["width"]), np.uint32)
            ref.render(cfg, inst, cols, rects, out=img)
            want.append(img)
        got = [np.zeros((sc["height"], sc["width"]), np.uint32) for _ in frames]
        flags = api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES | api.RENDER_NO_DEPENDENT_LAUNCH | api.RENDER_CULL_INSTANCES
        for k, (cfg, inst, cols, rects) in enumerate(frames):
            (a, b)[k & 1].render(cfg, inst, cols, rects, out=got[k], flags=flags)
        a.wait()
        b.wait()
        for k in range(len(frames)):
            assert np.array_equal(got[k], want[k]), k
    finally:
        for r in (ref, a, b):
            r.close()


def test_texture_unit_filter_equals_its_restatement():
    """The filter is the B200 texture unit's (tex2DLod on RGBA8 mipmapped arrays).  The CPU checker restates its
    arithmetic in integers (8-bit weights split level -> x -> y, 16-bit unorm texels; fitted with tools/hwtex/):
    random textures, random coordinates incl. wrapped and negative ones, random lods incl. out-of-range ones --
    every float of every sample is bit-identical."""
    from oracle.binding import Oracle
    rng = np.random.default_rng(7)
    for (w, h, levels) in ((64, 32, 6), (4096, 4096, 3), (2, 2, 2)):
        chain = np.concatenate([rng.integers(0, 256, max(1, w >> l) * max(1, h >> l) * 4, dtype=np.uint8) for l in range(levels)])
        r = api.LucidRenderer(64, 64, 0, 1 << 10)
        o = Oracle(64, 64, 0, 1 << 10)
        try:
            r.set_texture(1, chain, w, h, levels)
            o.lib.oracle_set_texture(o.h, 1, chain.ctypes.data, w, h, levels)
            n = 200_000
            uvl = np.stack([rng.random(n, np.float32) * 3 - 1, rng.random(n, np.float32) * 3 - 1,
                            rng.random(n, np.float32) * (levels + 0.5) - 0.25], axis=1).astype(np.float32)
            uvl[:1000, 2] = 0.0
            uvl[1000:2000, :2] = (rng.integers(0, 4 * w, (1000, 2)) / np.float32(2 * w)).astype(np.float32)  # texel centres and edges
            got = r.debug_sample_texture(1, uvl)
            want = o.texture_samples(1, uvl)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), int((got != want).any(axis=1).sum())
        finally:
            r.close()
            o.close()


def test_frame_hand_over_flags(small):
    """lucid_signal / lucid_wait_flags / lucid_set_frame_gate: the device-side hand-over of the bin-row split, here
    between two renderers on one GPU (each on its own stream).  The gathering renderer sees the complete composite
    after its flag wait, a gated frame does not store before the image was released, and a wait nobody answers
    gives up with LUCID_E_STATE instead of hanging the device."""
    sc = small["soup_close"]
    full_r, full_img = pu.run_cuda(sc)
    full_r.close()
    cfg, inst, cols, rects = api.prepare_frame(sc)
    nby = (sc["height"] + 31) // 32
    rows = multigpu.split_bin_rows(nby, 2)
    gather = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20, bin_rows=rows[0])
    peer = api.LucidRenderer(sc["width"], sc["height"], 0, 1 << 20, bin_rows=rows[1])
    try:
        gather.set_scene(sc), peer.set_scene(sc)
        ptr, pitch = gather.image_pointer()
        flags = gather.sync_pointer()
        for frame in (1, 2, 3):
            # the peer may store into the shared image once frame - 1 was released; it signals its flag when done
            peer.set_frame_gate(flags, api.LucidRenderer.SYNC_RELEASED, frame - 1)
            peer.render(cfg, inst, cols, rects, out_device_ptr=ptr, out_pitch=pitch, flags=api.RENDER_ASYNC)
            peer.signal(flags, 1, frame)
            gather.render(cfg, inst, cols, rects, flags=api.RENDER_ASYNC)
            gather.wait_flags(flags, 1, 1, frame)
            img = gather.read_image()  # after the wait on the gathering stream: both strips are there
            assert np.array_equal(img, full_img)
            gather.signal(flags, api.LucidRenderer.SYNC_RELEASED, frame)
        peer.wait()
        # nobody ever signals flag 5: the wait gives up after five seconds and the next lucid_wait reports it
        gather.wait_flags(flags, 5, 1, 1)
        with pytest.raises(api.LucidError) as e:
            gather.wait()
        assert "(-4)" in str(e.value)
        gather.wait()  # reported once
    finally:
        gather.close()
        peer.close()


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
