"""ctypes binding of the CPU oracle (oracle/liblucid_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (lucid_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "liblucid_oracle.so")
_lib = None
INFO_WORDS = 1152


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lucid_oracle.cpp")
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "liblucid_oracle.so"], check=True,
                       capture_output=True)
    return SO_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        build()
    lib = C.CDLL(SO_PATH)
    vp = C.c_void_p
    lib.oracle_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int]
    lib.oracle_create.restype = vp
    lib.oracle_destroy.argtypes = [vp]
    lib.oracle_set_threads.argtypes = [vp, C.c_int]
    lib.oracle_set_bin_rows.argtypes = [vp, C.c_int, C.c_int]
    lib.oracle_set_bin_range.argtypes = [vp, C.c_int, C.c_int]
    lib.oracle_set_geometry.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int]
    lib.oracle_set_texture.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    lib.oracle_render.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int]
    lib.oracle_stage_ms.argtypes = [vp, vp]
    lib.oracle_bin_count.argtypes = [vp]
    for name in ("oracle_read_info", "oracle_read_image", "oracle_read_exact_image", "oracle_read_image_float",
                 "oracle_read_frag_counts", "oracle_read_bin_levels"):
        getattr(lib, name).argtypes = [vp, vp]
    for name in ("oracle_read_quad_aabbs", "oracle_read_quad_input_ids", "oracle_read_tri_records",
                 "oracle_read_quad_attrs"):
        getattr(lib, name).argtypes = [vp, C.c_int, vp]
    lib.oracle_read_bin_lists.argtypes = [vp, vp, vp]
    lib.oracle_pow.argtypes = [C.c_float, C.c_float]
    lib.oracle_pow.restype = C.c_float
    lib.oracle_log2.argtypes = [C.c_float]
    lib.oracle_log2.restype = C.c_float
    lib.oracle_shade_probe.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_float)]
    lib.oracle_shade_probe.restype = C.c_uint32
    lib.oracle_set_reference_colour.argtypes = [vp, C.c_int]
    lib.oracle_set_comparators.argtypes = [vp, C.c_int]
    lib.oracle_set_tie_report.argtypes = [vp, C.c_int]
    lib.oracle_set_reverse_ties.argtypes = [vp, C.c_int]
    lib.oracle_read_tie_pixels.argtypes = [vp, vp, vp]
    lib.oracle_read_compare_image.argtypes = [vp, C.c_int, vp]
    lib.oracle_fn_compare_pixel.argtypes = [C.c_int, vp, C.c_int, C.c_uint32, C.c_int]
    lib.oracle_fn_compare_pixel.restype = C.c_uint32
    lib.oracle_texture_samples.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    for name in ("oracle_final_shade_fast", "oracle_final_shade_reference"):
        getattr(lib, name).argtypes = [C.c_float, C.c_float]
        getattr(lib, name).restype = C.c_float
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def compare_pixel(mode: int, samples, bg8: int, additive: bool = False) -> int:
    """One pixel through a comparator: samples = rows of (submission order, depth as float bits, RGBA8, opaque)."""
    s = np.ascontiguousarray(samples, np.uint32).reshape(-1, 4)
    return int(load().oracle_fn_compare_pixel(mode, _ptr(s), s.shape[0], bg8, int(additive)))


class Oracle:
    """CPU restatement of the pipeline with the same inputs as lucid_render()."""

    def __init__(self, width, height, opts=0, max_visible_quads=4793490, threads=1):
        self.lib = load()
        self.h = C.c_void_p(self.lib.oracle_create(width, height, opts, max_visible_quads))
        self.width, self.height = width, height
        self.bin_count = self.lib.oracle_bin_count(self.h)
        self.max_visible_quads = max_visible_quads
        self.lib.oracle_set_threads(self.h, threads)
        self._keep = []

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bin_rows(self, begin, end):
        self.lib.oracle_set_bin_rows(self.h, begin, end)

    def texture_samples(self, slot, uvl) -> np.ndarray:
        """(u, v, lod) triples through the restated texture-unit filter (lod clamped to the chain)."""
        uvl = np.ascontiguousarray(uvl, np.float32).reshape(-1, 3)
        out = np.zeros((uvl.shape[0], 4), np.float32)
        self.lib.oracle_texture_samples(self.h, slot, _ptr(uvl), uvl.shape[0], _ptr(out))
        return out

    def set_reference_colour(self, on: bool):
        """True: colour arithmetic in the reference's operation order (what the GLSL pins compare); False
        (default): the product's colour contract, which the CUDA kernels reproduce bit for bit."""
        self.lib.oracle_set_reference_colour(self.h, int(on))

    def set_bin_range(self, begin, end):
        self.lib.oracle_set_bin_range(self.h, begin, end)

    def set_tie_report(self, on: bool):
        """Mark the pixels whose colour can depend on the order of depth-key ties while rendering (read_tie_pixels)."""
        self.lib.oracle_set_tie_report(self.h, int(on))

    def set_reverse_ties(self, on: bool):
        """Test hook: every run of equal depth keys in reverse order (the canonical order is the triangle index)."""
        self.lib.oracle_set_reverse_ties(self.h, int(on))

    def read_tie_pixels(self):
        """(uint8[h, w] mask of the pixels covered by two or more entries of one run of equal depth keys -- the only
        pixels whose colour can depend on the tie order --, dict(lists, entries, pixels))."""
        mask = np.zeros((self.height, self.width), np.uint8)
        st = np.zeros(3, np.uint64)
        self.lib.oracle_read_tie_pixels(self.h, _ptr(mask), _ptr(st))
        return mask, dict(lists=int(st[0]), entries=int(st[1]), pixels=int(st[2]))

    def set_comparators(self, on: bool):
        """Also reduce the frame's samples the way the comparators of SURVEY 8 f4 would (read_compare_image)."""
        self.lib.oracle_set_comparators(self.h, int(on))

    def read_compare_image(self, mode: int):
        """mode 0: hardware alpha blending in submission order, 1: weighted blended OIT, 2: 4-layer MLAB."""
        out = np.zeros((self.height, self.width), np.uint32)
        self.lib.oracle_read_compare_image(self.h, mode, _ptr(out))
        return out

    def set_scene(self, scene):
        pos = np.ascontiguousarray(scene["positions"], np.float32)
        quads = np.ascontiguousarray(scene["quads"], np.uint32)
        cols = None if scene.get("colors") is None else np.ascontiguousarray(scene["colors"], np.uint32)
        uvs = None if scene.get("uvs") is None else np.ascontiguousarray(scene["uvs"], np.float32)
        nrm = None if scene.get("normals") is None else np.ascontiguousarray(scene["normals"], np.uint32)
        self._keep = [pos, quads, cols, uvs, nrm]
        self.lib.oracle_set_geometry(self.h, _ptr(pos), pos.shape[0], _ptr(cols), _ptr(uvs), _ptr(nrm), _ptr(quads),
                                     quads.shape[0])
        for name, slot in (("opaque", 0), ("transparent", 1)):
            if name in scene.get("textures", {}):
                w, h, levels, data = scene["textures"][name]
                data = np.ascontiguousarray(data, np.uint8)
                self.lib.oracle_set_texture(self.h, slot, _ptr(data), w, h, levels)

    def render(self, config, instances, colors, uv_rects=None, stages=7):
        instances = np.ascontiguousarray(instances, np.int32)
        colors = np.ascontiguousarray(colors, np.uint32)
        uv_rects = None if uv_rects is None else np.ascontiguousarray(uv_rects, np.float32)
        self.lib.oracle_render(self.h, C.byref(config), _ptr(instances), _ptr(colors), _ptr(uv_rects),
                               instances.shape[0], stages)
        self.info = self.read_info()
        return self

    def stage_ms(self):
        out = np.zeros(4, np.float64)
        self.lib.oracle_stage_ms(self.h, _ptr(out))
        return out

    def read_info(self):
        out = np.zeros(INFO_WORDS + 10 * self.bin_count, np.uint32)
        self.lib.oracle_read_info(self.h, _ptr(out))
        return out

    def num_visible(self):
        return int(self.info[1]), int(self.info[2])

    def read_quad_aabbs(self, which):
        out = np.zeros(self.num_visible()[which], np.uint32)
        self.lib.oracle_read_quad_aabbs(self.h, which, _ptr(out))
        return out

    def read_quad_input_ids(self, which):
        out = np.zeros(self.num_visible()[which], np.uint32)
        self.lib.oracle_read_quad_input_ids(self.h, which, _ptr(out))
        return out

    def read_tri_records(self, which):
        out = np.zeros((self.num_visible()[which] * 2, 21), np.uint32)
        self.lib.oracle_read_tri_records(self.h, which, _ptr(out))
        return out

    def read_quad_attrs(self, which):
        out = np.zeros((self.num_visible()[which], 16), np.uint32)
        self.lib.oracle_read_quad_attrs(self.h, which, _ptr(out))
        return out

    def read_bin_lists(self):
        counts = self.info[INFO_WORDS:].reshape(10, self.bin_count).view(np.int32)
        bq = np.zeros(int(counts[0].sum()), np.uint32)
        bt = np.zeros(int(counts[3].sum()), np.uint32)
        self.lib.oracle_read_bin_lists(self.h, _ptr(bq), _ptr(bt))
        return bq, bt

    def _img(self, fn, dtype=np.uint32, ch=None):
        shape = (self.height, self.width) if ch is None else (self.height, self.width, ch)
        out = np.zeros(shape, dtype)
        fn(self.h, _ptr(out))
        return out

    def read_image(self):
        return self._img(self.lib.oracle_read_image)

    def read_exact_image(self):
        return self._img(self.lib.oracle_read_exact_image)

    def read_image_float(self):
        return self._img(self.lib.oracle_read_image_float, np.float32, 3)

    def read_frag_counts(self):
        return self._img(self.lib.oracle_read_frag_counts)

    def read_bin_levels(self):
        out = np.zeros(self.bin_count, np.uint8)
        self.lib.oracle_read_bin_levels(self.h, _ptr(out))
        return out
