"""Fuzz of the checker's functions (quad setup, triangle records, bin rows, pixel rows, half-block keys, shadeSample,
the per-pixel reduction) against the reference's shader functions (needs oracle/_ref/libref_shaders.so,
i.e. the build container):  python tests/fuzz_ref_pins.py [quads] [seed]
Not collected by pytest; the committed vectors (tests/golden/ref_shader_funcs.json.gz) are the regression pin."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

vp = C.c_void_p
ref = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_shaders.so"))
ref.ref_process_quad.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
ref.ref_store_tri.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
ref.ref_raster_rows.argtypes = [vp, C.c_float, C.c_float, C.c_int, vp]
ref.ref_bin_rows.argtypes = [vp, vp]
ref.ref_half_block.argtypes = [C.c_uint32, C.c_uint32, C.c_int, vp, C.c_float, C.c_float, C.c_float, vp]
ref.ref_reduce_pixel.argtypes = [vp, vp, C.c_int, vp]
ref.ref_shade_sample.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, vp]
INST_BITS = [0x001, 0x004, 0x010, 0x020, 0x040, 0x200]  # vertex colours / normals, tex opaque, uv rect, albedo, instance colour
ptr = lambda a: a.ctypes.data_as(vp)  # noqa: E731
n_quads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = {"process_quad": 0, "store_tri": 0, "bin_rows": 0, "raster_rows": 0, "half_block": 0, "shade_sample": 0, "reduce_pixel": 0}
seen = {"visible": 0, "tris": 0, "half_blocks": 0, "shaded": 0, "reduced": 0}
for cam_i in range(4):
    w, h = [(1280, 720), (1920, 1080), (3840, 2160), (640, 360)][cam_i]
    spec = {"kind": "orbit", "center": rng.uniform(-2, 2, 3).tolist(), "distance": float(rng.uniform(0.3, 40)),
            "rot_h": float(rng.uniform(0, 6.28)), "rot_v": float(rng.uniform(-1.2, 1.2))}
    cfg = api.make_config(api.make_camera(spec, w, h), 1, (0.0, 0.1, 0.1, 1.0))
    cfg.enable_backface_culling = cam_i & 1
    cw = np.frombuffer(bytes(cfg), np.uint32).copy()
    origin = np.frombuffer(bytes(cfg), np.float32)[32:35].copy()
    o = Oracle(w, h, 0, 1 << 16, threads=1)
    lib = o.lib
    lib.oracle_fn_process_quad.argtypes = [vp, vp, vp, vp]
    lib.oracle_fn_store_tri.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp]
    lib.oracle_fn_raster_rows.argtypes = [vp, C.c_float, C.c_float, C.c_int, vp]
    lib.oracle_fn_bin_rows.argtypes = [vp, vp]
    lib.oracle_fn_half_block.argtypes = [C.c_uint32, C.c_uint32, C.c_int, vp, C.c_float, C.c_float, C.c_float, vp]
    lib.oracle_fn_reduce_pixel.argtypes = [vp, vp, C.c_int, vp]
    lib.oracle_fn_shade_sample.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    for k in range(n_quads // 4):
        near = k % 5 == 0
        centre = origin + rng.uniform(-1, 1, 3) * (0.3 if near else 1.0) if near else rng.uniform(-12, 12, 3)
        size = float(np.exp(rng.uniform(np.log(0.005), np.log(20.0))))
        u, v = rng.normal(size=3), rng.normal(size=3)
        u /= np.linalg.norm(u)
        v -= u * np.dot(u, v)
        v /= np.linalg.norm(v)
        aspect = float(np.exp(rng.uniform(-4, 0)))  # slivers too
        pos = np.array([centre, centre + u * size, centre + u * size + v * size * aspect, centre + v * size * aspect], np.float32)
        idx = np.array([0, 1, 2, 3], np.uint32)
        o.set_scene({"positions": pos, "quads": idx.reshape(1, 4)})
        a, b = np.zeros(5, np.uint32), np.zeros(5, np.uint32)
        ref.ref_process_quad(ptr(cw), w, h, ptr(pos), ptr(idx), ptr(a))
        lib.oracle_fn_process_quad(o.h, ptr(cw), ptr(idx), ptr(b))
        if a.tolist() != b.tolist():
            bad["process_quad"] += 1
            if bad["process_quad"] <= 3:
                print("processInputQuad differs", pos.tolist(), a.tolist(), b.tolist())
            continue
        if a[0] != 0xFFFFFFFF:
            continue
        seen["visible"] += 1
        for second in range(2):
            if (int(a[2]) >> (30 + second)) & 1:
                continue
            tri = np.array([pos[0] - origin, pos[1 + second] - origin, pos[2 + second] - origin], np.float32)
            ra, rb = np.zeros(21, np.uint32), np.zeros(21, np.uint32)
            ref.ref_store_tri(ptr(cw), ptr(tri), 0x200, int(a[3 + second]), ptr(ra))
            lib.oracle_fn_store_tri(o.h, ptr(cw), ptr(tri), 0x200, int(a[3 + second]), ptr(rb))
            seen["tris"] += 1
            if ra.tolist() != rb.tolist():
                bad["store_tri"] += 1
                if bad["store_tri"] <= 3:
                    print("storeTri differs", tri.tolist(), ra.tolist(), rb.tolist())
                continue
            scan8 = ra[8:16].copy()
            ba, bb = np.zeros(258, np.int32), np.zeros(258, np.int32)
            ref.ref_bin_rows(ptr(scan8), ptr(ba))
            lib.oracle_fn_bin_rows(ptr(scan8), ptr(bb))
            bad["bin_rows"] += ba.tolist() != bb.tolist()
            ymin = int(a[3 + second]) & 0xFFFF
            for gx in range(int(a[2]) & 0x7F, ((int(a[2]) >> 14) & 0x7F) + 1):
                sa, sb = np.zeros(24, np.uint32), np.zeros(24, np.uint32)
                ref.ref_raster_rows(ptr(scan8), float(gx * 32), float((ymin // 32) * 32), 8, ptr(sa))
                lib.oracle_fn_raster_rows(ptr(scan8), float(gx * 32), float((ymin // 32) * 32), 8, ptr(sb))
                bad["raster_rows"] += sa.tolist() != sb.tolist()
                # rasterHalfBlockCentroid / Bits + rasterBlockDepth on the first group's spans, LOW and HIGH key ranges
                depth_eq = ra[16:19].copy()
                for startx in (0, 8, 16, 24):
                    ha, hb = np.zeros(5, np.uint32), np.zeros(5, np.uint32)
                    args = (int(sa[0]), int(sa[1]), startx, ptr(depth_eq), float(gx * 32 + startx + 3.25),
                            float((ymin // 32) * 32 + 1.75), float(0x7FFFE if startx % 16 else 0x3FFFFE))
                    ref.ref_half_block(*args, ptr(ha))
                    lib.oracle_fn_half_block(*args, ptr(hb))
                    seen["half_blocks"] += 1
                    bad["half_block"] += ha.tolist() != hb.tolist()
                if gx > (int(a[2]) & 0x7F) + 3:
                    break
            # shadeSample with random instance flags, vertex attributes and a preset colour for the texture fetch
            rec = ra.copy()
            flags = sum(b for b in INST_BITS if rng.integers(0, 2))
            rec[19] = flags | (int(rng.integers(0, 200)) << 16)
            if flags & 0x004:
                rec[20] = 0
            attrs = rng.integers(0, 1 << 32, 16, dtype=np.uint64).astype(np.uint32)
            attrs[8:16] = rng.uniform(-2.0, 3.0, 8).astype(np.float32).view(np.uint32)
            inst_color = int(rng.integers(0, 1 << 32)) | (0xFF000000 if k % 3 == 0 else 0)
            if k % 17 == 0:
                inst_color &= 0x00FFFFFF
            uv_rect, preset = rng.uniform(0.0, 1.0, 4).astype(np.float32), rng.uniform(0.0, 1.0, 4).astype(np.float32)
            ymax = int(a[3 + second]) >> 16
            py = int(rng.integers(ymin, max(ymax, ymin) + 1))
            px = (int(a[2]) & 0x7F) * 32 + int(rng.integers(0, 64))
            oa, ob = np.zeros(10, np.uint32), np.zeros(10, np.uint32)
            ref.ref_shade_sample(ptr(cw), ptr(rec), ptr(attrs), inst_color, ptr(uv_rect), ptr(preset), px, py, second, ptr(oa))
            lib.oracle_fn_shade_sample(o.h, ptr(cw), ptr(rec), ptr(attrs), inst_color, ptr(uv_rect), ptr(preset), px, py,
                                       second, ptr(ob))
            seen["shaded"] += 1
            if oa.tolist() != ob.tolist():
                bad["shade_sample"] += 1
                if bad["shade_sample"] <= 3:
                    print("shadeSample differs", hex(flags), oa.tolist(), ob.tolist())
        if k % 8 == 0:
            # reduceSample: a stream sorted near to far up to local inversions, as the block sort leaves it
            n = int(rng.integers(0, 120))
            depth = np.sort(rng.uniform(0.01, 4.0, n).astype(np.float32))[::-1].copy()
            for _ in range(n // 4):
                i = int(rng.integers(0, max(n - 1, 1)))
                j = min(n - 1, i + int(rng.choice([1, 1, 2, 2, 3, 5])))
                depth[i], depth[j] = depth[j], depth[i]
            colour = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
            kind = rng.integers(0, 6, n)
            colour[kind == 0] |= 0xFF000000
            colour[kind == 1] &= 0x00FFFFFF
            samples = np.empty(2 * n, np.uint32)
            samples[0::2], samples[1::2] = colour, depth.view(np.uint32)
            pa, pb = np.zeros(4, np.uint32), np.zeros(4, np.uint32)
            ref.ref_reduce_pixel(ptr(cw), ptr(samples), n, ptr(pa))
            lib.oracle_fn_reduce_pixel(ptr(cw), ptr(samples), n, ptr(pb))
            seen["reduced"] += 1
            bad["reduce_pixel"] += pa.tolist() != pb.tolist()
    o.close()
print("fuzz:", n_quads, "quads,", seen, "mismatches", bad)
sys.exit(1 if any(bad.values()) else 0)
