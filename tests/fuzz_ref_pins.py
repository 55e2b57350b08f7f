"""Fuzz of the checker's functions against the reference's shader functions (needs oracle/_ref/libref_shaders.so,
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
ptr = lambda a: a.ctypes.data_as(vp)  # noqa: E731
n_quads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = {"process_quad": 0, "store_tri": 0, "bin_rows": 0, "raster_rows": 0}
seen = {"visible": 0, "tris": 0}
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
                if gx > (int(a[2]) & 0x7F) + 3:
                    break
    o.close()
print("fuzz:", n_quads, "quads,", seen, "mismatches", bad)
sys.exit(1 if any(bad.values()) else 0)
