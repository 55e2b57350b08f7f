"""Shared helpers of the parity tests: run the same frame through the CUDA path (C ABI) and the
CPU oracle and compare every intermediate product the north star names."""
from __future__ import annotations

import numpy as np

from lucid_b200 import api, scenes
from oracle.binding import Oracle

INFO = api.LUCID_INFO_U32_SIZE


def small_scenes():
    """name -> scene; sized so the oracle needs at most a few seconds each."""
    out = {}
    out["soup"] = scenes.quad_soup(num_quads=10_000, width=1280, height=720)
    out["soup_close"] = scenes.quad_soup(num_quads=6_000, width=640, height=360, distance=14.0, seed=11,
                                         name="soup_close")
    out["planes"] = scenes.planes(num_planes=32, width=640, height=360)
    out["meshlets"] = scenes.meshlet_patches(num_patches=24, width=960, height=540, distance=40.0)
    out["hairball"] = scenes.hairball(num_strands=2_500, segments=48, width=480, height=270, ribbon_width=0.08)
    out["arch"] = scenes.architecture(num_small=30_000, num_large=60, width=960, height=540, atlas_opaque=256,
                                      atlas_trans=128, levels=5)
    # small enough for the reference-source harness (32768 visible quads, tests/golden/make_ref_shader_golden.py): HIGH bins
    out["hairball_mini"] = scenes.hairball(num_strands=600, segments=48, width=160, height=96, ribbon_width=0.5)
    out["boxes"] = scenes.boxes()  # the reference's '#boxes' (src/scene_setup.cpp:159-196), 1280x720
    return out


def run_oracle(scene, opts=0, camera=None, bin_rows=None, threads=8, mvq=1 << 20, bin_range=None,
               reference_colour=False):
    """reference_colour: colour arithmetic in the reference's operation order (the form pinned against the
    reference's GLSL) instead of the product's colour contract (the form the kernels reproduce bit for bit)."""
    cfg, inst, cols, rects = api.prepare_frame(scene, camera)
    o = Oracle(scene["width"], scene["height"], opts, mvq, threads=threads)
    o.set_reference_colour(reference_colour)
    if bin_rows:
        o.set_bin_rows(*bin_rows)
    if bin_range:
        o.set_bin_range(*bin_range)
    o.set_scene(scene)
    o.render(cfg, inst, cols, rects)
    return o


def run_cuda(scene, opts=0, camera=None, bin_rows=None, mvq=1 << 20, renderer=None, create_flags=0, max_block_entries=0):
    cfg, inst, cols, rects = api.prepare_frame(scene, camera)
    r = renderer or api.LucidRenderer(scene["width"], scene["height"], opts, mvq, bin_rows=bin_rows,
                                      create_flags=create_flags, max_block_entries=max_block_entries)
    if renderer is None:
        r.set_scene(scene)
    img = np.zeros((scene["height"], scene["width"]), np.uint32)
    r.render(cfg, inst, cols, rects, out=img, flags=api.RENDER_FRAG_COUNTS)
    return r, img


FLOAT_COLS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 14, 16, 17, 18]  # float words of a 21-word triangle record


def canonical_tri_records(rec):
    """A NaN has many encodings (x86 produces 0xffc00000, the GPU 0x7fffffff) and both sides treat every one of
    them alike (fmin/fmax drop it, float->int gives 0): the float words of the records compare with one encoding.
    Full-size configs[2] has one such word -- a degenerate edge gives 0 * inf in storeTri (quad_setup.glsl:327-333)."""
    rec = np.array(rec, np.uint32, copy=True)
    f = rec[:, FLOAT_COLS]
    f[np.isnan(f.view(np.float32))] = 0x7FC00000
    rec[:, FLOAT_COLS] = f
    return rec


def canonical_lists(values, counts):
    """Sorts the entries of every bin's segment (segments are laid out in bin order)."""
    values = np.asarray(values)
    bins = np.repeat(np.arange(len(counts)), np.asarray(counts, np.int64))
    order = np.lexsort((values & 0x0FFFFFFF, bins))
    return values[order]


def compare(r, img, o, check_image=True):
    """Returns a dict of mismatch descriptions (empty = parity)."""
    bad = {}
    info_c, info_o = r.read_info(), o.info
    bc = o.bin_count
    hc, cc = api.split_info(info_c, bc)
    ho, co = api.split_info(info_o, bc)

    def words(name, a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape or not np.array_equal(a, b):
            n = int((a != b).sum()) if a.shape == b.shape else -1
            bad[name] = f"{n} differing entries (shape {a.shape} vs {b.shape})"

    words("num_input_quads", hc[0:1], ho[0:1])
    words("num_visible_quads", hc[1:3], ho[1:3])
    words("num_rejected_quads", hc[32:36], ho[32:36])
    n_small, n_large = int(ho[1]), int(ho[2])
    if "num_visible_quads" not in bad:
        for which, n in ((0, n_small), (1, n_large)):
            words(f"quad_aabbs[{which}]", r.read_quad_aabbs(which, n), o.read_quad_aabbs(which))
            tc, to = canonical_tri_records(r.read_tri_records(which, n)), canonical_tri_records(o.read_tri_records(which))
            for lo, hi, nm in ((0, 8, "bary"), (8, 16, "scan"), (16, 20, "depth"), (20, 21, "normal")):
                words(f"tri_{nm}[{which}]", tc[:, lo:hi], to[:, lo:hi])
            words(f"quad_attrs[{which}]", r.read_quad_attrs(which, n), o.read_quad_attrs(which))
    for idx, nm in ((0, "bin_quad_counts"), (1, "bin_quad_offsets"), (2, "bin_quad_offsets_temp"),
                    (3, "bin_tri_counts"), (4, "bin_tri_offsets"), (5, "bin_tri_offsets_temp")):
        words(nm, cc[idx], co[idx])
    words("bin_level_counts", hc[5:10], ho[5:10])
    n_low, n_high = int(ho[7]), int(ho[9])
    words("low_bins", cc[7][:n_low], co[7][:n_low])
    words("high_bins", cc[9][:n_high], co[9][:n_high])
    if "bin_quad_counts" not in bad and "bin_tri_counts" not in bad:
        # canonical form: every bin's list as a sorted set -- the order inside a
        # list comes from atomic arrival in the reference and in the CUDA path alike
        bq_o, bt_o = o.read_bin_lists()
        bq_c, bt_c = r.read_bin_lists(bq_o.size, bt_o.size)
        words("bin_quads", canonical_lists(bq_c, co[0]), canonical_lists(bq_o, co[0]))
        words("bin_tris", canonical_lists(bt_c, co[3]), canonical_lists(bt_o, co[3]))
    words("stats", hc[60:63], ho[60:63])
    words("frag_counts", r.read_frag_counts(), o.read_frag_counts())
    if check_image:
        ic = img.view(np.uint8).reshape(o.height, o.width, 4).astype(np.int32)
        io = o.read_image().view(np.uint8).reshape(o.height, o.width, 4).astype(np.int32)
        d = np.abs(ic - io)
        if d.max() > 1:  # <= 1/255 per channel (north star tolerance)
            bad["image"] = f"max abs diff {int(d.max())}/255 at {int((d > 1).sum())} channel values"
        elif d.max() > 0:
            bad_exact = int((d > 0).sum())
            bad.setdefault("_note", f"image differs by 1/255 at {bad_exact} channel values")
    errs = api.verify_info(info_c, bc)
    if errs:
        bad["verifyInfo"] = "; ".join(errs[:4])
    return bad


if __name__ == "__main__":
    import sys
    import time

    names = sys.argv[1:] or list(small_scenes().keys())
    all_scenes = small_scenes()
    for name in names:
        sc = all_scenes[name]
        t0 = time.time()
        o = run_oracle(sc)
        t1 = time.time()
        r, img = run_cuda(sc)
        t2 = time.time()
        bad = compare(r, img, o)
        st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
        print(f"== {name}: oracle {t1 - t0:.2f}s cuda {t2 - t1:.2f}s stage_ms {np.round(r.stage_times(), 3).tolist()}")
        print("   ", {k: st[k] for k in ("input_quads", "visible_small", "visible_large", "bin_quads", "bin_tris",
                                         "low_bins", "high_bins", "promoted_bins", "fragments", "half_block_tris")})
        print("   ", "PARITY OK" if not [k for k in bad if not k.startswith("_")] else "MISMATCH", bad)
        r.close()
