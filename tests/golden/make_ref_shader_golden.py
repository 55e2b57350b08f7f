#!/usr/bin/env python
"""Golden vectors from the reference's own shader functions.

oracle/build_ref_shaders.py compiles processInputQuad, storeTri (quad_setup.glsl), loadScanlineParamsRow /
loadScanlineParamsBin (shared/scanline.glsl), scanlineStep (bin_counter.glsl), rasterBinStep,
rasterHalfBlockCentroid / Bits and rasterBlockDepth (shared/raster.glsl) from the GLSL text under
/root/reference into oracle/_ref/libref_shaders.so.  This script feeds them seeded inputs and writes inputs
and outputs to tests/golden/ref_shader_funcs.json.gz; tests/test_ref_shader_pins.py replays the inputs through
the CPU checker's functions (oracle_fn_*) and demands identical words.  Only runs where the reference tree
is mounted (the build container); the JSON travels.

    python oracle/build_ref_shaders.py && python tests/golden/make_ref_shader_golden.py
"""
import ctypes as C
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from lucid_b200 import api  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_shaders.so"))
vp = C.c_void_p
lib.ref_process_quad.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
lib.ref_store_tri.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
lib.ref_raster_rows.argtypes = [vp, C.c_float, C.c_float, C.c_int, vp]
lib.ref_bin_rows.argtypes = [vp, vp]
lib.ref_half_block.argtypes = [C.c_uint32, C.c_uint32, C.c_int, vp, C.c_float, C.c_float, C.c_float, vp]
lib.ref_reduce_pixel.argtypes = [vp, vp, C.c_int, vp]
lib.ref_shade_sample.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, vp]
lib.ref_store_quad.argtypes = [C.c_uint32, vp, vp, vp, vp]
lib.ref_bin_scene.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
lib.ref_shade_image.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_uint32, vp]
lib.ref_frag_counts.argtypes = [C.c_int, C.c_int, vp, vp, vp]
lib.ref_frag_counts.restype = C.c_uint64
lib.ref_encode_rgba8.argtypes = [vp]
lib.ref_encode_rgba8.restype = C.c_uint32


def ptr(a):
    return a.ctypes.data_as(vp)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).tolist()


rng = np.random.default_rng(20261017)
cameras = [
    ({"kind": "orbit", "center": [0.0, 0.0, 0.0], "distance": 30.0, "rot_h": 0.5, "rot_v": 0.8}, 1280, 720, 0),
    ({"kind": "orbit", "center": [1.0, -2.0, 0.5], "distance": 12.0, "rot_h": 2.1, "rot_v": -0.4}, 1920, 1080, 1),
    ({"kind": "orbit", "center": [0.0, 0.0, 0.0], "distance": 3.0, "rot_h": -1.0, "rot_v": 0.2}, 3840, 2160, 0),
    ({"kind": "orbit", "center": [0.0, 0.0, 0.0], "distance": 0.6, "rot_h": 4.0, "rot_v": 1.1}, 640, 360, 0),
]
out = {"about": "outputs of the reference's shader functions (nadult/lucid data/shaders, compiled by "
                "oracle/build_ref_shaders.py) on seeded inputs; floats are stored as their binary32 bit patterns",
       "cases": []}
n_visible = 0
for cam_spec, width, height, backface in cameras:
    cam = api.make_camera(cam_spec, width, height)
    cfg = api.make_config(cam, 1, (0.0, 30.0 / 255.0, 30.0 / 255.0, 1.0))
    cfg.enable_backface_culling = backface
    cfg_words = np.frombuffer(bytes(cfg), np.uint32).copy()
    origin = np.frombuffer(bytes(cfg), np.float32)[32:35].copy()
    case = {"camera": cam_spec, "width": width, "height": height, "config_words": cfg_words.tolist(), "quads": []}
    for k in range(70):
        # a planar-ish quad of log-uniform size somewhere around the scene; every 7th close to the eye
        # (near-plane crossers), every 11th degenerate
        centre = rng.uniform(-8, 8, 3) if k % 7 else origin + rng.uniform(-0.4, 0.4, 3)
        size = float(np.exp(rng.uniform(np.log(0.02), np.log(6.0))))
        u, v = rng.normal(size=3), rng.normal(size=3)
        u /= np.linalg.norm(u)
        v -= u * np.dot(u, v)
        v /= np.linalg.norm(v)
        pos = np.array([centre, centre + u * size, centre + (u + v) * size + rng.normal(size=3) * 0.01 * size,
                        centre + v * size], np.float32)
        idx = np.array([0, 1, 2, 3], np.uint32)
        if k % 11 == 5:
            idx[3] = idx[2]  # a triangle
        if k % 11 == 9:
            pos[1] = pos[0]  # coincident positions
        res = np.zeros(5, np.uint32)
        lib.ref_process_quad(ptr(cfg_words), width, height, ptr(pos), ptr(idx), ptr(res))
        q = {"pos": bits(pos), "idx": idx.tolist(), "process_quad": res.tolist()}
        if res[0] == 0xFFFFFFFF:
            n_visible += 1
            q["tris"] = []
            for second in range(2):
                if (int(res[2]) >> (30 + second)) & 1:
                    continue
                tri = np.array([pos[idx[0]] - origin, pos[idx[1 + second]] - origin, pos[idx[2 + second]] - origin], np.float32)
                flags_id = (0x004 if k % 3 == 0 else 0x200) | (k << 16)
                rec = np.zeros(21, np.uint32)
                lib.ref_store_tri(ptr(cfg_words), ptr(tri), flags_id, int(res[3 + second]), ptr(rec))
                t = {"second": second, "tri": bits(tri), "flags_id": flags_id, "y_aabb": int(res[3 + second]),
                     "record": rec.tolist()}
                scan8 = rec[8:16].copy()
                # bin rows (binning) and pixel rows (raster) of the bins the triangle's y range touches
                rows = np.zeros(2 + 256, np.int32)
                lib.ref_bin_rows(ptr(scan8), ptr(rows))
                n_rows = max(0, min(int(rows[1]) - int(rows[0]) + 1, 128))
                t["bin_rows"] = rows[:2 + 2 * n_rows].tolist()
                ymin = int(res[3 + second]) & 0xFFFF
                bx0 = ((int(res[2]) >> 0) & 0x7F) * 32
                t["raster_rows"] = []
                for n_probe, start_y in enumerate(float((ymin // 32) * 32 + 4 * g) for g in (0, 5)):
                    for start_x in ((float(bx0),) if n_probe else (float(bx0), float(bx0 + 32))):
                        spans = np.zeros(6, np.uint32)
                        lib.ref_raster_rows(ptr(scan8), start_x, start_y, 2, ptr(spans))
                        hb = []
                        startxs = (0, 8) if (k + second) % 2 else (16, 24)
                        for startx in startxs:
                            depth_eq = rec[16:19].view(np.float32).copy()
                            cpx, cpy = np.float32(start_x + startx + 3.25), np.float32(start_y + 1.75)
                            o5 = np.zeros(5, np.uint32)
                            lib.ref_half_block(int(spans[0]), int(spans[1]), startx, ptr(depth_eq), float(cpx), float(cpy),
                                               float(0x7FFFE if startx % 16 else 0x3FFFFE), ptr(o5))
                            hb.append(o5.tolist())
                        t["raster_rows"].append({"start": [start_x, start_y], "spans": spans.tolist(),
                                                 "half_block_startx": list(startxs), "half_blocks": hb})
                q["tris"].append(t)
        case["quads"].append(q)
    out["cases"].append(case)
# shading.glsl: per-pixel reduction (3-entry insertion window + front-to-back blending) on sample streams
# that are sorted near to far (larger inverse depth first) up to local inversions, as the block sort leaves them
out["reduce"] = []
cfg_words = np.array(out["cases"][0]["config_words"], np.uint32)
for k in range(60):
    n = int(rng.integers(0, 90))
    depth = np.sort(rng.uniform(0.01, 4.0, n).astype(np.float32))[::-1].copy()
    for _ in range(n // 4):  # local inversions, mostly within the window's reach, sometimes beyond it
        i = int(rng.integers(0, max(n - 1, 1)))
        j = min(n - 1, i + int(rng.choice([1, 1, 2, 2, 3, 5])))
        depth[i], depth[j] = depth[j], depth[i]
    if n > 4 and k % 5 == 0:
        depth[2] = depth[1]  # equal depths
    colour = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    alpha_kind = rng.integers(0, 6, n)
    colour[alpha_kind == 0] |= 0xFF000000  # opaque
    colour[alpha_kind == 1] &= 0x00FFFFFF  # alpha 0 but colour bits set
    if n > 2 and k % 7 == 0:
        colour[1] = 0  # "no sample" (shadeSample returned 0)
    samples = np.empty(2 * n, np.uint32)
    samples[0::2], samples[1::2] = colour, depth.view(np.uint32)
    o4 = np.zeros(4, np.uint32)
    lib.ref_reduce_pixel(ptr(cfg_words), ptr(samples), n, ptr(o4))
    out["reduce"].append({"samples": samples.tolist(), "rgba": o4.tolist()})
# shading.glsl: shadeSample on the visible triangles above with every combination of instance flags, random
# vertex attributes and a preset colour standing in for the texture fetch (whose arguments are recorded)
out["shade"] = []
INST = {"vcolors": 0x001, "vnormals": 0x004, "tex_opaque": 0x010, "uv_rect": 0x020, "albedo": 0x040, "color": 0x200}
n_shade = 0
for case in out["cases"]:
    cw = np.array(case["config_words"], np.uint32)
    for q in case["quads"]:
        for t in q.get("tris", []):
            if n_shade >= 160:
                break
            k = n_shade
            n_shade += 1
            flags = 0
            for bit, name in enumerate(INST):
                if (k * 2654435761 >> (8 + bit)) & 1:
                    flags |= INST[name]
            rec = np.array(t["record"], np.uint32)
            instance_id = k % 200
            rec[19] = flags | (instance_id << 16)
            if flags & INST["vnormals"]:
                rec[20] = 0
            attrs = rng.integers(0, 1 << 32, 16, dtype=np.uint64).astype(np.uint32)
            attrs[8:16] = rng.uniform(-2.0, 3.0, 8).astype(np.float32).view(np.uint32)  # uv0, uv1-uv0, uv2-uv0, uv3-uv0
            inst_color = int(rng.integers(0, 1 << 32)) | (0xFF000000 if k % 3 == 0 else 0)
            if k % 17 == 0:
                inst_color &= 0x00FFFFFF  # alpha 0: the sample is dropped
            uv_rect = rng.uniform(0.0, 1.0, 4).astype(np.float32)
            preset = rng.uniform(0.0, 1.0, 4).astype(np.float32)
            # a pixel inside the triangle's y range and bin column
            ymin, ymax = t["y_aabb"] & 0xFFFF, t["y_aabb"] >> 16
            py = int(rng.integers(ymin, max(ymax, ymin) + 1))
            px = (q["process_quad"][2] & 0x7F) * 32 + int(rng.integers(0, 64))
            o10 = np.zeros(10, np.uint32)
            lib.ref_shade_sample(ptr(cw), ptr(rec), ptr(attrs), inst_color, ptr(uv_rect), ptr(preset), px, py,
                                 t["second"], ptr(o10))
            out["shade"].append({"case": out["cases"].index(case), "record": rec.tolist(), "attrs": attrs.tolist(),
                                 "inst_color": inst_color, "uv_rect": bits(uv_rect), "tex_preset": bits(preset),
                                 "pixel": [px, py], "second": t["second"], "out": o10.tolist()})
# quad_setup.glsl: storeQuad (vertex attribute repack) for every combination of the three attribute flags
out["store_quad"] = []
for k in range(24):
    flags = (0x001 if k & 1 else 0) | (0x004 if k & 2 else 0) | (0x040 if k & 4 else 0) | (0x200 if k & 8 else 0)
    cols = rng.integers(0, 1 << 32, 4, dtype=np.uint64).astype(np.uint32)
    nrms = rng.integers(0, 1 << 30, 4, dtype=np.uint64).astype(np.uint32)
    uvs = rng.uniform(-4.0, 4.0, 8).astype(np.float32)
    o16 = np.zeros(16, np.uint32)
    lib.ref_store_quad(flags, ptr(cols), ptr(nrms), ptr(uvs), ptr(o16))
    out["store_quad"].append({"flags": flags, "colors": cols.tolist(), "normals": nrms.tolist(), "uvs": bits(uvs),
                              "out": o16.tolist()})
# Scene level: setup -> bin counts -> bin lists of whole (small) scenes through the reference's per-invocation
# functions run one after the other (oracle/build_ref_shaders.py ref_bin_scene).  The scenes are the seeded
# procedural ones of tests/parity_util.py, so only their names and the reference's results are stored.
from tests import parity_util as pu  # noqa: E402
out["bin_scenes"] = []
small = pu.small_scenes()
for name in ("soup_close", "arch", "soup", "meshlets", "boxes", "planes", "hairball_mini"):
    sc = small[name]
    cfg, inst, inst_cols, inst_rects = api.prepare_frame(sc)
    cw = np.frombuffer(bytes(cfg), np.uint32).copy()
    pos = np.ascontiguousarray(sc["positions"], np.float32)
    quads = np.ascontiguousarray(sc["quads"], np.uint32)
    assert int(inst[:, 2].sum()) == quads.shape[0] and (inst[:, 1] == 0).all()  # instances cover the index buffer in order
    bc = ((sc["width"] + 31) // 32) * ((sc["height"] + 31) // 32)
    counts = np.zeros(2 * bc, np.int32)
    lists = np.zeros(32768 * 72, np.uint32)
    n = np.zeros(4, np.uint32)
    # per input quad: instance flags | instance id << 16 (instances are consecutive slices of the index buffer)
    flags_id = np.repeat(inst[:, 3].astype(np.uint32) | (np.arange(inst.shape[0], dtype=np.uint32) << 16), inst[:, 2])
    opt = lambda key, dt: None if sc.get(key) is None else np.ascontiguousarray(sc[key], dt)  # noqa: E731
    vcol, vnrm, vuv = opt("colors", np.uint32), opt("normals", np.uint32), opt("uvs", np.float32)
    rc = lib.ref_bin_scene(ptr(cw), sc["width"], sc["height"], ptr(pos), pos.shape[0], ptr(quads), quads.shape[0],
                           ptr(flags_id), None if vcol is None else ptr(vcol), None if vnrm is None else ptr(vnrm),
                           None if vuv is None else ptr(vuv), ptr(counts), ptr(lists), ptr(n))
    assert rc == 0, (name, rc)
    nq, nt = int(n[2]), int(n[3])
    bq = pu.canonical_lists(lists[:nq], counts[:bc])
    bt = pu.canonical_lists(lists[nq:nq + nt], counts[bc:])
    # raster coverage of the same frame: every listed triangle walked like generateRowTris, pixels of
    # rasterHalfBlockBits counted
    frag = np.zeros(sc["width"] * sc["height"], np.uint32)
    is_high = np.zeros(bc, np.uint8)
    frag_total = int(lib.ref_frag_counts(sc["width"], sc["height"], ptr(counts), ptr(frag), ptr(is_high)))
    print("   fragments", frag_total, "inside the image", int(frag.sum()), "max per pixel", int(frag.max()),
          "HIGH bins", int(is_high.sum()))
    # the image, for scenes without textures (the sampler is not part of the reference's source)
    image_sha = None
    if not any(int(f) & 0x040 for f in inst[:, 3]):
        bg = sc["background"]
        bg8 = sum(int(min(max(float(bg[i]), 0.0), 1.0) * 255.0 + 0.5) << (8 * i) for i in range(3)) | 0xFF000000
        image = np.zeros(sc["width"] * sc["height"], np.uint32)
        ic = np.ascontiguousarray(inst_cols, np.uint32)
        ir = np.ascontiguousarray(inst_rects, np.float32)
        assert lib.ref_shade_image(ptr(cw), sc["width"], sc["height"], ptr(counts), ptr(is_high), ptr(ic), ptr(ir),
                                   inst.shape[0], bg8, ptr(image)) == 0
        image_sha = hashlib.sha256(image.tobytes()).hexdigest()
        print("   image digest", image_sha[:16], "non-background pixels", int((image != bg8).sum()))
    out["bin_scenes"].append({"scene": name, "max_visible_quads": 32768, "visible": [int(n[0]), int(n[1])],
                              "image_sha256": image_sha,
                              "fragments": frag_total, "fragments_in_image": int(frag.sum()),
                              "high_bins": np.flatnonzero(is_high).tolist(),
                              "frag_counts_sha256": hashlib.sha256(frag.tobytes()).hexdigest(),
                              "quad_counts": counts[:bc].tolist(), "tri_counts": counts[bc:].tolist(),
                              "list_entries": [nq, nt],
                              "bin_quads_sha256": hashlib.sha256(np.ascontiguousarray(bq, np.uint32).tobytes()).hexdigest(),
                              "bin_tris_sha256": hashlib.sha256(np.ascontiguousarray(bt, np.uint32).tobytes()).hexdigest()})
    print("   bin scene", name, "visible small/large", int(n[0]), int(n[1]), "list entries", nq, nt)
out["encode_rgba8"] = []
for k in range(64):
    c = rng.uniform(0.0, 1.0, 4).astype(np.float32)
    if k % 8 == 0:
        c[k % 4] = np.float32(1.0)
    if k % 8 == 4:
        c[k % 4] = np.float32((k + 1) / 255.0)  # exactly on a step
    out["encode_rgba8"].append({"rgba": bits(c), "packed": int(lib.ref_encode_rgba8(ptr(c)))})

path = os.path.join(HERE, "ref_shader_funcs.json.gz")
with gzip.GzipFile(path, "wb", mtime=0) as f:
    f.write(json.dumps(out, separators=(",", ":")).encode())
statuses = [q["process_quad"][0] for c in out["cases"] for q in c["quads"]]
print(path, os.path.getsize(path), "bytes;", len(statuses), "quads,", n_visible, "visible; rejection types seen:",
      sorted(set(s for s in statuses if s != 0xFFFFFFFF)))
