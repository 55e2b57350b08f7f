"""CPU tests of the oracle: the reference's runtime invariants (verifyInfo, sortedness by
construction, window == exact sort), committed golden digests, closed-form scenes and an independent
brute-force coverage rasteriser."""
import json
import os

import numpy as np
import pytest

from lucid_b200 import api, scenes
from oracle import binding
from tests import parity_util as pu
from tests.golden.make_oracle_golden import record

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def small():
    return pu.small_scenes()


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "oracle_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["soup", "soup_close", "planes", "meshlets", "hairball", "arch", "boxes", "hairball_mini"])
def test_oracle_matches_golden_and_invariants(name, small, golden):
    o = pu.run_oracle(small[name], threads=4)
    assert record(o) == golden[name]
    # LucidRenderer::verifyInfo invariants (lucid_renderer.cpp:649-680)
    assert api.verify_info(o.info, o.bin_count) == []
    st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    assert st["input_quads"] == st["visible_small"] + st["visible_large"] + st["rejected_other"] + \
        st["rejected_backface"] + st["rejected_frustum"] + st["rejected_between_samples"]
    assert st["empty_bins"] + st["low_bins"] + st["high_bins"] - st["promoted_bins"] == o.bin_count
    # fragment statistic == sum of the per-pixel counts of pixels inside the view, up to the
    # fragments of bins that stick out of the viewport (height 360/270/540 are not multiples of 32)
    inside = int(o.read_frag_counts().sum())
    assert inside <= st["fragments"]
    if o.height % 32 == 0 and o.width % 32 == 0:
        assert inside == st["fragments"]


def test_oracle_is_thread_count_independent(small):
    a = pu.run_oracle(small["soup_close"], threads=1)
    b = pu.run_oracle(small["soup_close"], threads=8)
    assert np.array_equal(a.read_image(), b.read_image())
    assert np.array_equal(a.info, b.info)


def test_window_equals_exact_sort_where_reference_claims_exactness(small):
    """'Exact' OIT: the 3-entry window must reproduce a true per-pixel depth sort wherever the
    VISUALIZE_ERRORS counter (stats[2]) reports no overflow."""
    for name in ("soup", "planes", "arch", "hairball"):
        vis = pu.run_oracle(small[name], opts=api.OPT_VISUALIZE_ERRORS, threads=4)
        invalid = api.decode_stats(vis.info, vis.bin_count, vis.width, vis.height)["invalid_pixels"]
        o = pu.run_oracle(small[name], threads=4)
        mismatch = int((o.read_image() != o.read_exact_image()).sum())
        # a pixel can only deviate from the exact sort if the window overflowed there
        assert mismatch <= invalid, (name, mismatch, invalid)
        if name in ("soup", "planes"):
            assert invalid == 0 and mismatch == 0
        # pixels flagged by VISUALIZE_ERRORS are painted red, everything else is unchanged
        changed = vis.read_image() != o.read_image()
        assert int(changed.sum()) <= invalid
        assert (vis.read_image()[changed] == 0xFF0000FF).all()


def test_planes_closed_form(small):
    """#planes (scene_setup.cpp:198-222): N parallel quads seen head on.  Every pixel inside the
    smallest plane holds exactly N fragments; colour follows the analytic N-layer over blend of the
    truncated RGBA8 samples."""
    sc = small["planes"]
    o = pu.run_oracle(sc, threads=4)
    fc = o.read_frag_counts()
    cy, cx = sc["height"] // 2, sc["width"] // 2
    assert fc[cy, cx] == 32
    assert fc.max() == 32
    # fragments per pixel are non-increasing away from the centre along a row (nested squares)
    row = fc[cy, cx:]
    assert (np.diff(row.astype(np.int64)) <= 0).all()
    # analytic blend at the centre pixel from the oracle's own per-sample shading probe
    lib = binding.load()
    import ctypes as C
    samples = []
    for tri in range(64):  # tris 2k, 2k+1 of plane k; the centre lies in exactly one of each pair
        d = C.c_float()
        col = lib.oracle_shade_probe(o.h, cx, cy, tri + (o.max_visible_quads - 32) * 2, C.byref(d))
        samples.append((d.value, col, tri))
    # keep the triangle of each quad that actually covers the pixel: use the frag image == 32 and
    # pick per plane the sample whose barycentrics are valid, i.e. the one the renderer used
    img = o.read_image()[cy, cx]
    exact = o.read_exact_image()[cy, cx]
    assert img == exact


def planes_with_opaque_layer(k=5, n=12, width=320, height=200):
    """#planes with plane k made opaque (its own INST_IS_OPAQUE draw call): k + 1 samples are in front of or on the
    opaque layer at every pixel it covers."""
    sc = scenes.planes(num_planes=n, width=width, height=height)
    sc["materials"] = [((1.0, 1.0, 1.0), 0.25, (0.0, 0.0, 1.0, 1.0)), ((1.0, 1.0, 1.0), 1.0, (0.0, 0.0, 1.0, 1.0))]
    vc = scenes.INST_HAS_VERTEX_COLORS
    sc["draw_calls"] = [(0, k, 0, vc), (1, 1, k, vc | scenes.INST_IS_OPAQUE), (0, n - k - 1, k + 1, vc)]
    return sc


def _max_channel_diff(a, b):
    return np.abs(a.view(np.uint8).astype(np.int32) - b.view(np.uint8).astype(np.int32)).reshape(a.shape + (4,)).max(axis=-1)


def test_opaque_prepass_mode(small):
    """LUCID_OPT_OPAQUE_PREPASS in the checker (shared/shading.glsl:31-32 TODO as an option): closed form on stacked
    planes; on the parity scenes the exact per-pixel-sort image stays within 1/255 (a flagged-opaque sample whose
    interpolated vertex alpha truncates to 254 leaks 1/255 of what lies behind it without the option), the window
    image differs by more only where the window was too small, everything before the raster stage is untouched, and
    the option is ignored under additive blending."""
    k, n = 5, 12
    sc = planes_with_opaque_layer(k, n)
    plain, pre = pu.run_oracle(sc, threads=4), pu.run_oracle(sc, opts=api.OPT_OPAQUE_PREPASS, threads=4)
    f0, f1 = plain.read_frag_counts(), pre.read_frag_counts()
    cy, cx = sc["height"] // 2, sc["width"] // 2
    assert f0[cy, cx] == n and f1[cy, cx] == k + 1
    inside = f0 == n  # pixels inside the smallest plane: all layers present
    assert (f1[inside] == k + 1).all()
    assert _max_channel_diff(plain.read_image(), pre.read_image()).max() <= 1
    assert int(pre.info[60]) == int(f1.sum()) < int(plain.info[60])
    for name in ("soup", "meshlets", "arch", "hairball"):
        plain = pu.run_oracle(small[name], threads=8)
        pre = pu.run_oracle(small[name], opts=api.OPT_OPAQUE_PREPASS, threads=8)
        assert np.array_equal(plain.info[:60], pre.info[:60]) and plain.info[61] == pre.info[61]
        assert _max_channel_diff(plain.read_exact_image(), pre.read_exact_image()).max() <= 1
        f0, f1 = plain.read_frag_counts(), pre.read_frag_counts()
        assert (f1 <= f0).all() and int(pre.info[60]) == int(f1.sum())
        if name == "hairball":  # no opaque instance: nothing changes
            assert np.array_equal(f0, f1)
        else:
            assert f1.sum() < f0.sum()
        vis = pu.run_oracle(small[name], opts=api.OPT_VISUALIZE_ERRORS, threads=8)
        invalid = api.decode_stats(vis.info, vis.bin_count, vis.width, vis.height)["invalid_pixels"]
        assert int((_max_channel_diff(plain.read_image(), pre.read_image()) > 1).sum()) <= invalid
    add = pu.run_oracle(small["soup"], opts=api.OPT_ADDITIVE_BLENDING, threads=8)
    add_pre = pu.run_oracle(small["soup"], opts=api.OPT_ADDITIVE_BLENDING | api.OPT_OPAQUE_PREPASS, threads=8)
    assert np.array_equal(add.read_image(), add_pre.read_image()) and np.array_equal(add.info, add_pre.info)


def test_pow_contract_accuracy():
    """orc_pow is the shared polynomial pow; it must stay within 3e-6 relative of libm on the
    sRGB ranges so colours stay well inside the 1/255 tolerance."""
    lib = binding.load()
    xs = np.concatenate([np.linspace(0.0031308, 1.0, 4000), np.linspace(1.0, 3.0, 500)]).astype(np.float32)
    for y in (1.0 / 2.4, 2.4):
        got = np.array([lib.oracle_pow(float(x), float(np.float32(y))) for x in xs], np.float64)
        ref = np.power(xs.astype(np.float64), float(np.float32(y)))
        assert np.max(np.abs(got - ref) / ref) < 3e-6


def _brute_force_counts(scene, cfg, inst):
    """Independent coverage: project with view_proj in float64, edge functions at pixel centres."""
    w, h = scene["width"], scene["height"]
    m = np.array([[getattr(cfg.view_proj_matrix[c], k) for k in "xyzw"] for c in range(4)], np.float64).T
    pos = scene["positions"].astype(np.float64)
    clip = np.concatenate([pos, np.ones((len(pos), 1))], axis=1) @ m.T
    assert (clip[:, 3] > 1e-3).all(), "scene must be in front of the camera for the brute-force check"
    sx = (clip[:, 0] / clip[:, 3] + 1.0) * (w * 0.5)
    sy = (clip[:, 1] / clip[:, 3] + 1.0) * (h * 0.5)
    counts = np.zeros((h, w), np.int64)
    quads = scene["quads"]
    for q in quads:
        for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3])):
            x = sx[list(tri)]
            y = sy[list(tri)]
            x0, x1 = int(np.floor(x.min())), int(np.ceil(x.max()))
            y0, y1 = int(np.floor(y.min())), int(np.ceil(y.max()))
            x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, w - 1), min(y1, h - 1)
            if x1 < x0 or y1 < y0:
                continue
            px, py = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
            e0 = (x[1] - x[0]) * (py - y[0]) - (y[1] - y[0]) * (px - x[0])
            e1 = (x[2] - x[1]) * (py - y[1]) - (y[2] - y[1]) * (px - x[1])
            e2 = (x[0] - x[2]) * (py - y[2]) - (y[0] - y[2]) * (px - x[2])
            inside = ((e0 >= 0) & (e1 >= 0) & (e2 >= 0)) | ((e0 <= 0) & (e1 <= 0) & (e2 <= 0))
            counts[y0:y1 + 1, x0:x1 + 1] += inside
    return counts


def test_coverage_against_independent_brute_force():
    """The scanline machinery (fixed-point-like fp32 edge walking through setup, binning and the
    two raster paths) must agree with plain edge functions except on pixels whose centre lies
    (numerically) on an edge."""
    sc = scenes.quad_soup(num_quads=1500, width=320, height=192, distance=26.0, seed=21, min_edge=0.3,
                          max_edge=2.0)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    o = pu.run_oracle(sc, threads=4)
    got = o.read_frag_counts().astype(np.int64)
    want = _brute_force_counts(sc, cfg, inst)
    diff = got != want
    assert want.sum() > 20000
    assert diff.mean() < 0.004, f"{diff.sum()} of {diff.size} pixels differ"
    assert abs(int(got.sum()) - int(want.sum())) < 0.002 * want.sum()
    assert np.abs(got - want).max() <= 2


def test_high_path_and_promotion_are_exercised(small):
    o = pu.run_oracle(small["hairball"], threads=4)
    levels = o.read_bin_levels()
    assert (levels == 4).sum() >= 30 and (levels == 2).sum() >= 10
    # a scene that overflows 256 triangles in one 8x8 block of a LOW bin -> promotion
    sc = scenes.planes(num_planes=200, width=128, height=96, plane_size=0.12, plane_dist=0.02)
    o = pu.run_oracle(sc, threads=2)
    st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    assert st["promoted_bins"] >= 1
    assert api.verify_info(o.info, o.bin_count) == []


def test_edge_cases():
    # empty instance list
    sc = scenes.quad_soup(num_quads=64, width=96, height=64)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    o = binding.Oracle(96, 64)
    o.set_scene(sc)
    o.render(cfg, inst[:0], cols[:0], rects[:0])
    st = api.decode_stats(o.info, o.bin_count, 96, 64)
    assert st["input_quads"] == 0 and st["empty_bins"] == o.bin_count
    assert (o.read_image() == 0xFF1E1E00).all()
    # degenerate quads (repeated indices / coincident vertices) are rejected as "other"
    sc2 = dict(sc)
    q = sc["quads"].copy()
    q[:8, 1] = q[:8, 0]
    q[:8, 3] = q[:8, 2]
    sc2["quads"] = q
    o2 = pu.run_oracle(sc2, threads=1)
    assert api.decode_stats(o2.info, o2.bin_count, 96, 64)["rejected_other"] == 8
    # MAX_VISIBLE_QUADS overflow: later quads are dropped and counted, never written
    o3 = pu.run_oracle(sc, threads=1, mvq=16)
    st3 = api.decode_stats(o3.info, o3.bin_count, 96, 64)
    assert st3["visible_small"] + st3["visible_large"] == 16 and st3["dropped_quads"] > 0


def test_bin_row_split_composes_to_the_full_frame(small):
    """per-bin results of a rank that owns rows [a,b) equal the single-GPU results."""
    from lucid_b200 import multigpu
    sc = small["soup_close"]
    full = pu.run_oracle(sc, threads=4)
    nby = (sc["height"] + 31) // 32
    img = np.zeros_like(full.read_image())
    frags = 0
    for rows in multigpu.split_bin_rows(nby, 3):
        part = pu.run_oracle(sc, threads=4, bin_rows=rows)
        y0, y1 = multigpu.strip_pixel_rows(rows, sc["height"])
        img[y0:y1] = part.read_image()[y0:y1]
        frags += int(part.info[60])
        _, cf = api.split_info(full.info, full.bin_count)
        _, cp = api.split_info(part.info, part.bin_count)
        bx = (sc["width"] + 31) // 32
        sl = slice(rows[0] * bx, rows[1] * bx)
        assert np.array_equal(cf[0][sl], cp[0][sl]) and np.array_equal(cf[3][sl], cp[3][sl])
    assert np.array_equal(img, full.read_image())
    assert frags == int(full.info[60])


def test_bin_range_split_composes_to_the_full_frame(small):
    """Ownership finer than rows: row-major bin ranges that cut through bin rows reproduce the full
    frame's counts in their bins, nothing elsewhere, and their pixels compose to the full image."""
    sc = small["arch"]
    full = pu.run_oracle(sc)
    bc, bcx = full.bin_count, (sc["width"] + 31) // 32
    _, fc = api.split_info(full.info, bc)
    cuts = [0, bcx + 3, 3 * bcx + bcx // 2, bc]
    image = np.zeros_like(full.read_image())
    frags = 0
    for lo, hi in zip(cuts, cuts[1:]):
        part = pu.run_oracle(sc, bin_range=(lo, hi))
        _, pc = api.split_info(part.info, bc)
        for which in (0, 3):
            assert np.array_equal(pc[which][lo:hi], fc[which][lo:hi])
            assert pc[which][:lo].sum() == 0 and pc[which][hi:].sum() == 0
        frags += int(part.info[60])
        img = part.read_image()
        for b in range(lo, hi):
            by, bx = divmod(b, bcx)
            image[by * 32:(by + 1) * 32, bx * 32:(bx + 1) * 32] = img[by * 32:(by + 1) * 32, bx * 32:(bx + 1) * 32]
    assert np.array_equal(image, full.read_image())
    assert frags == int(full.info[60])


def test_depth_key_ties_are_reported_and_bound_the_order_dependence(small):
    """"report pixels whose block had depth-key ties separately".  The reference leaves entries of equal
    quantised depth in the arrival order of racing atomics; checker and kernels order them by triangle index.  The
    checker marks the pixels covered by two or more entries of one run of equal keys: rendering every run in REVERSE
    order may change marked pixels only.  On the parity scenes it changes none at all (the per-pixel window orders
    samples by their own depth); on a coplanar stack -- equal sample depths too -- it changes every covered pixel."""
    from tests.test_gpu_parity import coplanar_stack

    def both_orders(sc):
        cfg, inst, cols, rects = api.prepare_frame(sc)
        out = []
        for reverse in (False, True):
            o = binding.Oracle(sc["width"], sc["height"], 0, 1 << 20, threads=8)
            o.set_tie_report(True)
            o.set_reverse_ties(reverse)
            o.set_scene(sc)
            o.render(cfg, inst, cols, rects)
            mask, st = o.read_tie_pixels()
            out.append((o.read_image(), mask, st, o.read_frag_counts()))
        return out

    for name in ("meshlets", "hairball", "arch"):
        (img, mask, st, _), (img_r, mask_r, st_r, _) = both_orders(small[name])
        assert st == st_r and np.array_equal(mask, mask_r)
        assert st["lists"] > 0 and st["entries"] >= 2 * st["lists"] and st["pixels"] == int(mask.sum())
        differ = img != img_r
        assert not (differ & (mask == 0)).any()
        assert int(differ.sum()) == 0
    (img, mask, st, frags), (img_r, _, _, _) = both_orders(coplanar_stack(40))
    covered = frags > 0
    assert np.array_equal(mask != 0, covered)
    assert np.array_equal(img != img_r, covered)


def test_boxes_scene_is_the_reference_s(small):
    """'#boxes' (BoxesSetup::updateScene, src/scene_setup.cpp:159-196): the quads are exactly what the reference's
    quad generator (restated in oracle/quadgen_oracle.cpp and pinned against the reference's own source) makes of
    addBox's triangles -- order and vertex rotation included; every ray through closed boxes enters and leaves, so
    (away from silhouette pixels) every pixel holds an even number of fragments; no bin overflows."""
    from oracle import quadgen_binding as qb

    sc = small["boxes"]
    assert sc["quads"].shape == (6000, 4) and sc["positions"].shape == (8000, 3)
    want = qb.run(qb.load_oracle(), sc["positions"], scenes.box_triangles(1000), 4.0, mode=0)
    assert np.array_equal(np.asarray(want["quads"], np.uint32), sc["quads"]) and want["num_degenerate"] == 0
    if qb.reference_available():  # the reference's own src/quad_generator.cpp, where /root/reference is mounted
        ref = qb.run(qb.load_reference(), sc["positions"], scenes.box_triangles(1000), 4.0)
        assert np.array_equal(np.asarray(ref["quads"], np.uint32), sc["quads"])
    # colour ramp: FColor(float3(x, y, z) / 9, 1) truncated to bytes; the last box is white
    assert int(sc["colors"][0]) == 0xFF000000 and int(sc["colors"][-1]) == 0xFFFFFFFF
    o = pu.run_oracle(sc, threads=4)
    st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    assert st["input_quads"] == 6000 and st["promoted_bins"] == 0 and not (o.read_bin_levels() == 5).any()
    covered = o.read_frag_counts()
    covered = covered[covered > 0]
    assert covered.size > 200_000 and (covered % 2 == 0).mean() > 0.999
