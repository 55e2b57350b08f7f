"""Comparators of SURVEY 8 f4 (lucid_compare_render): the last frame's samples reduced the way hardware alpha blending
in submission order (the reference's SimpleRenderer, src/simple_renderer.cpp:69-132,134-196), weighted blended OIT and
4-layer MLAB would reduce them.

CPU part: the checker's three reductions (oracle/lucid_oracle.cpp comparePixel) against closed forms and an
independent float32 restatement written here.  GPU part: lucid_compare_render against the checker, bit for bit."""
import numpy as np
import pytest

from lucid_b200 import api, scenes
from oracle import binding as ob
from tests import parity_util as pu

f32 = np.float32
HW, WBOIT, MLAB4 = api.COMPARE_HW_BLEND, api.COMPARE_WBOIT, api.COMPARE_MLAB4
BG8 = 0xFF1E1E00  # the scenes' background: (0, 30, 30)


def rgba(r, g, b, a):
    return int(r) | (int(g) << 8) | (int(b) << 16) | (int(a) << 24)


def chans(c):
    return np.array([c & 255, (c >> 8) & 255, (c >> 16) & 255, (c >> 24) & 255], np.int64)


def sample(order, depth, color, opaque=False):
    return [order, int(np.array([depth], f32).view(np.uint32)[0]), color, int(opaque)]


def over_exact(colors_near_to_far, bg8):
    """float64 front-to-back blend of RGBA8 samples over the background."""
    out, trans = np.zeros(3), 1.0
    for c in colors_near_to_far:
        ch = chans(c) / 255.0
        out += ch[:3] * ch[3] * trans
        trans *= 1.0 - ch[3]
    return out + trans * chans(bg8)[:3] / 255.0


def hw_blend_restated(samples_in_order, bg8):
    """The fixed-function blend on an 8-bit target, one float32 operation at a time (no fused multiply-add here: the
    tolerance below covers the difference): dst = round8(src * a + dst * (1 - a))."""
    dst = chans(bg8)[:3].astype(np.int64)
    for c in samples_in_order:
        ch = chans(c).astype(f32) * f32(1.0 / 255.0)
        d = dst.astype(f32) * f32(1.0 / 255.0)
        out = ch[:3] * ch[3] + d * (f32(1.0) - ch[3])
        dst = np.floor(np.clip(out, 0, 1) * f32(255.0) + f32(0.5)).astype(np.int64)
    return dst


def test_hw_blend_depends_on_submission_order_and_rounds_every_blend():
    near, far = rgba(250, 10, 10, 128), rgba(10, 10, 250, 128)
    # far first, near second: the blend order hardware needs -- equals the exact blend up to the per-blend rounding
    good = ob.compare_pixel(HW, [sample(0, 0.2, far), sample(1, 0.5, near)], BG8)
    assert np.abs(chans(good)[:3] - over_exact([near, far], BG8) * 255.0).max() <= 1.0
    assert np.array_equal(chans(good)[:3], hw_blend_restated([far, near], BG8))
    # near first, far second: the far sample is blended OVER the near one
    bad = ob.compare_pixel(HW, [sample(0, 0.5, near), sample(1, 0.2, far)], BG8)
    assert np.array_equal(chans(bad)[:3], hw_blend_restated([near, far], BG8))
    assert np.abs(chans(bad)[:3] - over_exact([near, far], BG8) * 255.0).max() > 50
    # the order is the submission order, not the order of the sample array
    assert ob.compare_pixel(HW, [sample(1, 0.5, near), sample(0, 0.2, far)], BG8) == good
    assert chans(good)[3] == 255


def test_opaque_phase_depth_test_and_ties():
    wall, wall2, glass = rgba(200, 200, 200, 255), rgba(50, 60, 70, 255), rgba(0, 255, 0, 64)
    for mode in (HW, WBOIT, MLAB4):
        # a transparent sample behind the nearest opaque one fails the depth test; so does one at the same depth (`less`)
        assert ob.compare_pixel(mode, [sample(0, 0.1, glass), sample(1, 0.5, wall, True)], BG8) == wall
        assert ob.compare_pixel(mode, [sample(0, 0.5, glass), sample(1, 0.5, wall, True)], BG8) == wall
        # the nearest opaque sample wins whatever the submission order; among equal depths the first submitted
        assert ob.compare_pixel(mode, [sample(0, 0.5, wall, True), sample(1, 0.3, wall2, True)], BG8) == wall
        assert ob.compare_pixel(mode, [sample(0, 0.3, wall2, True), sample(1, 0.5, wall, True)], BG8) == wall
        assert ob.compare_pixel(mode, [sample(0, 0.5, wall2, True), sample(1, 0.5, wall, True)], BG8) == wall2
        # in front of it the transparent sample blends over the wall
        got = chans(ob.compare_pixel(mode, [sample(0, 0.9, glass), sample(1, 0.5, wall, True)], BG8))[:3]
        assert np.abs(got - over_exact([glass, wall], BG8) * 255.0).max() <= 1.0
        # no sample at all, and samples shadeSample dropped (colour 0): the background
        assert ob.compare_pixel(mode, np.zeros((0, 4), np.uint32), BG8) == BG8
        assert ob.compare_pixel(mode, [sample(0, 0.5, 0)], BG8) == BG8
    # the opaque phase ignores alpha (no blending, simple_renderer.cpp:80)
    assert ob.compare_pixel(HW, [sample(0, 0.5, rgba(9, 8, 7, 100), True)], BG8) == rgba(9, 8, 7, 255)
    # additive blending: src_alpha / one
    got = chans(ob.compare_pixel(HW, [sample(0, 0.5, rgba(100, 0, 200, 128)), sample(1, 0.7, rgba(100, 0, 200, 128))], BG8, True))
    assert np.abs(got[:3] - np.array([100, 30, 231])).max() <= 1  # 2 * 100 * 128/255 ; 30 ; min(255, 30 + 2 * 200 * 128/255)


def test_mlab4_is_exact_up_to_four_layers_and_merges_beyond():
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 4):
        for _ in range(20):
            depths = rng.permutation(np.linspace(0.1, 0.9, n)).astype(f32)
            cols = [rgba(*rng.integers(0, 256, 3), rng.integers(1, 256)) for _ in range(n)]
            got = chans(ob.compare_pixel(MLAB4, [sample(i, depths[i], cols[i]) for i in range(n)], BG8))[:3]
            near_to_far = [cols[i] for i in np.argsort(-depths)]
            assert np.abs(got - over_exact(near_to_far, BG8) * 255.0).max() <= 0.51
    # submitted far to near, nothing ever falls out of order: exact for any count
    n = 40
    depths = np.linspace(0.1, 0.9, n).astype(f32)
    cols = [rgba(*rng.integers(0, 256, 3), rng.integers(1, 128)) for _ in range(n)]
    got = chans(ob.compare_pixel(MLAB4, [sample(i, depths[i], cols[i]) for i in range(n)], BG8))[:3]
    assert np.abs(got - over_exact(cols[::-1], BG8) * 255.0).max() <= 0.51
    # submitted near to far, the fifth and later fragments are merged into the fourth layer in arrival order, which is
    # depth order here: still exact
    got = chans(ob.compare_pixel(MLAB4, [sample(i, depths[n - 1 - i], cols[n - 1 - i]) for i in range(n)], BG8))[:3]
    assert np.abs(got - over_exact(cols[::-1], BG8) * 255.0).max() <= 0.51
    # a random order of many layers is an approximation
    worst = 0.0
    for _ in range(20):
        perm = rng.permutation(n)
        got = chans(ob.compare_pixel(MLAB4, [sample(i, depths[perm[i]], cols[perm[i]]) for i in range(n)], BG8))[:3]
        worst = max(worst, np.abs(got - over_exact(cols[::-1], BG8) * 255.0).max())
    assert worst > 1.0


def test_wboit_closed_forms():
    glass = rgba(40, 200, 90, 77)
    # one sample: average colour = the sample's, coverage = its alpha -> the exact blend
    got = chans(ob.compare_pixel(WBOIT, [sample(0, 0.25, glass)], BG8))[:3]
    assert np.abs(got - over_exact([glass], BG8) * 255.0).max() <= 0.51
    # order independent
    rng = np.random.default_rng(6)
    n = 12
    depths = rng.uniform(0.05, 0.9, n).astype(f32)
    cols = [rgba(*rng.integers(0, 256, 3), rng.integers(1, 200)) for _ in range(n)]
    ref = chans(ob.compare_pixel(WBOIT, [sample(i, depths[i], cols[i]) for i in range(n)], BG8))[:3]
    for _ in range(5):
        perm = rng.permutation(n)
        got = chans(ob.compare_pixel(WBOIT, [sample(i, depths[perm[i]], cols[perm[i]]) for i in range(n)], BG8))[:3]
        assert np.abs(got - ref).max() <= 1  # float summation order
    # equal colours at any depths: the weighted average is that colour, the coverage 1 - prod(1 - a)
    same = [rgba(120, 60, 30, a) for a in (50, 100, 150)]
    got = chans(ob.compare_pixel(WBOIT, [sample(i, 0.1 + 0.2 * i, c) for i, c in enumerate(same)], BG8))[:3]
    assert np.abs(got - over_exact(same, BG8) * 255.0).max() <= 0.51
    # the restated weight (eq. 7 of McGuire & Bavoil on the ray position z = 1 / depth), in float64
    a = np.array([chans(c)[3] / 255.0 for c in cols])
    z = 1.0 / depths.astype(np.float64)
    w = a * np.clip(10.0 / (1e-5 + (z / 5.0) ** 2 + (z / 200.0) ** 6), 1e-2, 3e3)
    rgb = np.array([chans(c)[:3] / 255.0 for c in cols])
    avg = (rgb * (a * w)[:, None]).sum(axis=0) / max((a * w).sum(), 1e-5)
    trans = np.prod(1.0 - a)
    want = avg * (1.0 - trans) + chans(BG8)[:3] / 255.0 * trans
    assert np.abs(ref - want * 255.0).max() <= 0.52


def _channels(img):
    return img.view(np.uint8).reshape(img.shape[0], img.shape[1], 4).astype(np.int32)


def _run_oracle(scene, opts=0):
    cfg, inst, cols, rects = api.prepare_frame(scene)
    o = ob.Oracle(scene["width"], scene["height"], opts, 1 << 20, threads=8)
    o.set_comparators(True)
    o.set_scene(scene)
    o.render(cfg, inst, cols, rects)
    return o


def mixed_order_scene(width=480, height=270):
    """Instances that mix small and large quads, duplicated quads and opaque / transparent draw calls: the cases in
    which the visible-quad slots say least about the submission order (small quads are compacted upwards, large ones
    downwards, each in input order)."""
    rng = scenes.Rng(21)
    n = 1500
    centers = np.stack([rng.uniform(-3.0, 3.0, n), rng.uniform(-1.7, 1.7, n), rng.uniform(-3.0, 3.0, n)], axis=1)
    big = (np.arange(n) % 5) == 2
    half = np.where(big, rng.uniform(0.8, 2.5, n), rng.uniform(0.03, 0.12, n)).astype(np.float32)
    u, v = scenes._orthonormal_frames(rng, n)
    hu, hv = u * half[:, None], v * half[:, None]
    corners = np.stack([centers - hu - hv, centers + hu - hv, centers + hu + hv, centers - hu + hv], axis=1)
    positions = corners.reshape(-1, 3).astype(np.float32)
    quads = np.arange(n * 4, dtype=np.uint32).reshape(n, 4)
    quads[7::50] = quads[6::50]  # duplicated quads inside an instance
    draw_calls, materials = [], []
    rgb = np.stack([rng.f32(8) for _ in range(3)], axis=1)
    for i, off in enumerate(range(0, n, 250)):
        opaque = i in (1, 4)
        materials.append((tuple(float(c) for c in rgb[i]), 1.0 if opaque else (60 + 30 * i + 0.5) / 255.0, (0.0, 0.0, 1.0, 1.0)))
        draw_calls.append((i, min(250, n - off), off, scenes.INST_IS_OPAQUE if opaque else 0))
    camera = dict(kind="orbit", center=(0.0, 0.0, 0.0), distance=9.0, rot_h=0.3, rot_v=0.4)
    return scenes._scene(positions, quads, draw_calls, materials, camera, width, height, name="mixed_order")


def test_checker_images_on_scenes():
    small = pu.small_scenes()
    # '#planes' as the reference submits it -- nearest plane first -- is the worst case for hardware blending; the same
    # planes submitted back to front blend correctly (up to one rounding per layer)
    sc = small["planes"]
    o = _run_oracle(sc)
    exact = _channels(o.read_image())[..., :3]
    assert np.abs(_channels(o.read_compare_image(HW))[..., :3] - exact).max() > 40
    back_to_front = dict(sc)
    back_to_front["quads"] = np.ascontiguousarray(sc["quads"][::-1])
    o2 = _run_oracle(back_to_front)
    assert np.array_equal(o2.read_image(), o.read_image())
    assert np.abs(_channels(o2.read_compare_image(HW))[..., :3] - exact).max() <= 8
    assert np.abs(_channels(o2.read_compare_image(HW))[..., :3] - exact).mean() < 1.0
    # 32 layers of alpha 0.25: MLAB keeps the nearest four exactly, WBOIT does not depend on the order
    assert np.abs(_channels(o.read_compare_image(WBOIT)) - _channels(o2.read_compare_image(WBOIT))).max() <= 1
    # a frame of opaque instances only: every comparator shows the exact image
    sc = small["soup"]
    opaque = dict(sc)
    opaque["draw_calls"] = [(m, n, off, opts | scenes.INST_IS_OPAQUE) for (m, n, off, opts) in sc["draw_calls"]]
    opaque["materials"] = [(rgb, 1.0, rect) for (rgb, _, rect) in sc["materials"]]
    o3 = _run_oracle(opaque)
    for mode in (HW, WBOIT, MLAB4):
        assert np.array_equal(o3.read_compare_image(mode), o3.read_image())
    # few layers per pixel: MLAB4 equals the exact image; hardware blending does not
    o4 = _run_oracle(sc)
    assert np.abs(_channels(o4.read_compare_image(MLAB4)) - _channels(o4.read_image())).max() <= 1
    assert np.abs(_channels(o4.read_compare_image(HW)) - _channels(o4.read_image())).max() > 30
    # the comparators leave the frame itself alone
    assert np.array_equal(o4.read_image(), pu.run_oracle(sc).read_image())
    # a scene made to confuse the recovery of the submission order renders at all and has both kinds of quads
    o5 = _run_oracle(mixed_order_scene())
    assert o5.num_visible()[0] > 500 and o5.num_visible()[1] > 100


# ---- GPU ------------------------------------------------------------------------------------------------------------

GPU_SCENES = ["soup", "soup_close", "planes", "meshlets", "hairball", "arch", "boxes", "mixed_order"]


@pytest.fixture(scope="module")
def gpu_scenes():
    out = pu.small_scenes()
    out["mixed_order"] = mixed_order_scene()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_SCENES)
def test_cuda_comparators_equal_the_checker(name, gpu_scenes):
    sc = gpu_scenes[name]
    o = _run_oracle(sc)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r, img = pu.run_cuda(sc)
    try:
        assert np.array_equal(img, o.read_image())
        for mode in (HW, WBOIT, MLAB4):
            got, ms = r.compare_render(mode, cfg)
            want = o.read_compare_image(mode)
            diff = int((got != want).sum())
            assert diff == 0, f"{name} mode {mode}: {diff} pixels differ, max {np.abs(_channels(got) - _channels(want)).max()}/255"
            assert ms > 0.0
        # the frame is still there: a second pass gives the same image, and so does the exact path
        again, _ = r.compare_render(HW, cfg)
        assert np.array_equal(again, o.read_compare_image(HW))
        assert np.array_equal(r.read_image(), o.read_image())
    finally:
        r.close()


@pytest.mark.gpu
def test_cuda_hw_blend_additive_and_states(gpu_scenes):
    sc = gpu_scenes["soup_close"]
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(sc["width"], sc["height"], api.OPT_ADDITIVE_BLENDING, 1 << 20)
    try:
        r.set_scene(sc)
        with pytest.raises(api.LucidError):  # no frame yet
            r.compare_render(HW, cfg)
        r.render(cfg, inst, cols, rects)
        o = _run_oracle(sc, api.OPT_ADDITIVE_BLENDING)
        got, _ = r.compare_render(HW, cfg)
        assert np.array_equal(got, o.read_compare_image(HW))
        with pytest.raises(api.LucidError):  # the approximate-OIT comparators are defined for the normal blend
            r.compare_render(WBOIT, cfg)
        with pytest.raises(api.LucidError):
            r.compare_render(7, cfg)
    finally:
        r.close()
    r = api.LucidRenderer(sc["width"], sc["height"], api.OPT_OPAQUE_PREPASS, 1 << 20)
    try:
        r.set_scene(sc)
        r.render(cfg, inst, cols, rects)
        with pytest.raises(api.LucidError):  # the pre-pass drops samples before they reach the entry stream
            r.compare_render(HW, cfg)
    finally:
        r.close()


def _long_list_scene():
    return scenes.hairball(num_strands=12_000, segments=48, width=240, height=136, ribbon_width=0.1)


def test_long_list_scene_has_lists_over_1024_entries():
    """What the GPU test below relies on: the half-block lists of over 384 entries average over 1024."""
    import ctypes as C

    sc = _long_list_scene()
    cfg, inst, cols, rects = api.prepare_frame(sc)
    o = ob.Oracle(sc["width"], sc["height"], 0, 1 << 20, threads=8)
    o.lib.oracle_set_item_stats.argtypes = [C.c_void_p, C.c_int]
    o.lib.oracle_read_item_stats.argtypes = [C.c_void_p, C.c_void_p]
    o.lib.oracle_set_item_stats(o.h, 1)
    o.set_scene(sc)
    o.render(cfg, inst, cols, rects)
    st = (C.c_ulonglong * 24)()
    o.lib.oracle_read_item_stats(o.h, st)
    assert st[9] > 100 and st[14] / st[9] > 1024
    assert not (o.read_bin_levels() == 5).any()  # no bin over the reference's list limits (it would have no lists)


@pytest.mark.gpu
def test_cuda_comparators_on_long_lists():
    """Lists over 1024 entries are sorted in the L2-resident scratch: a dense hairball at low resolution."""
    sc = _long_list_scene()
    o = _run_oracle(sc)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r, img = pu.run_cuda(sc)
    try:
        assert np.array_equal(img, o.read_image())
        for mode in (HW, MLAB4):
            got, _ = r.compare_render(mode, cfg)
            assert np.array_equal(got, o.read_compare_image(mode))
    finally:
        r.close()


def test_checker_comparator_images_match_committed_digests():
    """tests/golden/comparators_golden.json (written by make_comparators_golden.py): the checker's comparator modes
    do not drift silently."""
    import json
    import os

    from tests.golden import make_comparators_golden as mk

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "comparators_golden.json")) as f:
        golden = json.load(f)
    scenes_ = mk.scenes_to_pin()
    assert set(golden) == set(scenes_)
    for name, sc in scenes_.items():
        assert mk.record(_run_oracle(sc)) == golden[name], name
