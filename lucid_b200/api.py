"""ctypes binding of the C ABI (include/lucid_b200.h, include/lucid_host.h) and a Python mirror of
the reference's LucidRenderer interface (src/lucid_renderer.h:21-96): construct with options and
a view size, render(context), getStats(), verifyInfo().

There is no fallback: if the CUDA library is missing or no GPU is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

LUCID_INFO_U32_SIZE = 1152
COUNTS_PER_BIN = 10
MEM_HOST, MEM_DEVICE, MEM_NONE = 0, 1, 2
RENDER_ASYNC, RENDER_SKIP_INFO, RENDER_FRAG_COUNTS, RENDER_NO_STAGE_TIMES, RENDER_CULL_INSTANCES = 1, 2, 4, 8, 16
RENDER_OWNED_BINS_ONLY = 32  # MEM_HOST read-back of the owned bins only (bin-row split)
RENDER_NO_DEPENDENT_LAUNCH = 64  # several handles busy on one device: no programmatic dependent launch

OPT_DEBUG_RASTER = 1 << 3  # the reference's raster_*_debug pipelines: lucid_read_debug_records
OPT_TIMERS = 1 << 4
OPT_ADDITIVE_BLENDING = 1 << 5
OPT_VISUALIZE_ERRORS = 1 << 6
OPT_ALPHA_THRESHOLD = 1 << 7
OPT_OPAQUE_PREPASS = 1 << 8  # extension, include/lucid_abi.h
COMPARE_HW_BLEND, COMPARE_WBOIT, COMPARE_MLAB4 = 0, 1, 2  # lucid_compare_render modes (include/lucid_abi.h)


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class InstanceData(C.Structure):
    _fields_ = [("index_offset", C.c_int32), ("vertex_offset", C.c_int32), ("num_quads", C.c_int32),
                ("flags", C.c_uint32)]


class Lighting(C.Structure):
    _fields_ = [("ambient_color", Vec4), ("sun_color", Vec4), ("sun_dir", Vec4), ("sun_power", C.c_float),
                ("ambient_power", C.c_float), ("_pad", C.c_float * 2)]


class Frustum(C.Structure):
    _fields_ = [("ws_origins", Vec4 * 4), ("ws_dirs", Vec4 * 4), ("ws_origin0", Vec4), ("ws_dir0", Vec4),
                ("ws_dirx", Vec4), ("ws_diry", Vec4)]


class LucidConfig(C.Structure):
    _fields_ = [("frustum", Frustum), ("view_proj_matrix", Vec4 * 4), ("lighting", Lighting),
                ("background_color", Vec4), ("enable_backface_culling", C.c_uint32),
                ("num_instances", C.c_int32), ("instance_packet_size", C.c_int32), ("_pad", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("target", C.c_float * 3), ("target_up", C.c_float * 3),
                ("fov_rad", C.c_float), ("z_near", C.c_float), ("z_far", C.c_float),
                ("viewport_width", C.c_int), ("viewport_height", C.c_int)]


class DrawCall(C.Structure):
    _fields_ = [("material_id", C.c_int32), ("num_quads", C.c_int32), ("quad_offset", C.c_int32),
                ("opts", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("diffuse", C.c_float * 3), ("opacity", C.c_float), ("uv_rect", C.c_float * 4)]


class CreateInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("opts", C.c_uint32),
                ("max_visible_quads", C.c_int32), ("max_dispatches", C.c_int32), ("device", C.c_int32),
                ("stream", C.c_void_p), ("bin_row_begin", C.c_int32), ("bin_row_end", C.c_int32),
                ("max_block_entries", C.c_uint32), ("flags", C.c_uint32)]


CREATE_COMPACT_LISTS = 1


assert C.sizeof(LucidConfig) == 352 and C.sizeof(InstanceData) == 16

_lib = None

# every symbol include/lucid_b200.h and include/lucid_host.h declare
C_ABI_SYMBOLS = [
    "lucid_create", "lucid_destroy", "lucid_last_error", "lucid_set_geometry", "lucid_set_texture",
    "lucid_set_bin_rows", "lucid_render", "lucid_wait", "lucid_read_info", "lucid_bin_count",
    "lucid_stage_times", "lucid_stage_times_at", "lucid_read_row_costs", "lucid_set_bin_range", "lucid_read_bin_costs", "lucid_composite_to", "lucid_read_quad_aabbs", "lucid_read_tri_records", "lucid_read_quad_attrs",
    "lucid_read_bin_lists", "lucid_read_frag_counts", "lucid_read_image", "lucid_image_pointer",
    "lucid_ipc_export_image", "lucid_ipc_open_image", "lucid_ipc_close_image",
    "lucid_debug_sample_texture", "lucid_sync_pointer", "lucid_ipc_export_sync", "lucid_signal", "lucid_wait_flags", "lucid_set_frame_gate",
    "lucid_host_orbit_camera", "lucid_host_default_lighting", "lucid_host_make_config",
    "lucid_host_camera_matrices", "lucid_host_build_instances", "lucid_host_packet_size", "lucid_host_cluster_order",
    "lucid_quadgen", "lucid_quadgen_last_error", "lucid_read_debug_records", "lucid_compare_render",
]


def load_library(build_if_needed: bool = True):
    """Loads lucid_b200/_lucid_b200.so (building it with nvcc first when sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LUCID_B200_SO") or _build.SO_PATH  # override: experiment builds only
    if build_if_needed and path == _build.SO_PATH and _build.needs_build():
        try:
            _build.build()
        except Exception as e:
            # never run a stale binary silently after a source edit that does not compile
            if not os.path.exists(path) or os.environ.get("LUCID_B200_ALLOW_STALE") != "1":
                raise RuntimeError(f"lucid_b200: sources are newer than {path} and the build failed: {e}") from e
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m lucid_b200.build` (needs nvcc); there is no CPU fallback")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.lucid_create.argtypes = [C.POINTER(CreateInfo), C.POINTER(vp)]
    lib.lucid_destroy.argtypes = [vp]
    lib.lucid_destroy.restype = None
    lib.lucid_last_error.argtypes = [vp]
    lib.lucid_last_error.restype = C.c_char_p
    lib.lucid_set_geometry.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32]
    lib.lucid_set_texture.argtypes = [vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32]
    lib.lucid_set_bin_rows.argtypes = [vp, C.c_int32, C.c_int32]
    lib.lucid_render.argtypes = [vp, C.POINTER(LucidConfig), vp, vp, vp, C.c_int32, vp, C.c_size_t, C.c_int32,
                                 C.c_uint32]
    lib.lucid_wait.argtypes = [vp]
    lib.lucid_read_info.argtypes = [vp, vp, C.c_size_t]
    lib.lucid_bin_count.argtypes = [vp]
    lib.lucid_read_row_costs.argtypes = [vp, vp, C.c_int32]
    lib.lucid_read_bin_costs.argtypes = [vp, vp, C.c_int32]
    lib.lucid_set_bin_range.argtypes = [vp, C.c_int32, C.c_int32]
    lib.lucid_composite_to.argtypes = [vp, vp, C.c_size_t]
    lib.lucid_stage_times.argtypes = [vp, C.POINTER(C.c_float * 8)]
    lib.lucid_stage_times_at.argtypes = [vp, C.c_int32, C.POINTER(C.c_float * 8)]
    lib.lucid_read_quad_aabbs.argtypes = [vp, C.c_int32, vp, C.c_int32]
    lib.lucid_read_tri_records.argtypes = [vp, C.c_int32, vp, C.c_int32]
    lib.lucid_read_quad_attrs.argtypes = [vp, C.c_int32, vp, C.c_int32]
    lib.lucid_read_bin_lists.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
    lib.lucid_read_frag_counts.argtypes = [vp, vp]
    lib.lucid_read_image.argtypes = [vp, vp, C.c_size_t]
    lib.lucid_image_pointer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.lucid_ipc_export_image.argtypes = [vp, vp]
    lib.lucid_ipc_open_image.argtypes = [vp, vp, C.POINTER(vp)]
    lib.lucid_ipc_close_image.argtypes = [vp, vp]
    lib.lucid_debug_sample_texture.argtypes = [vp, C.c_int32, vp, C.c_int32, vp]
    lib.lucid_sync_pointer.argtypes = [vp, C.POINTER(vp)]
    lib.lucid_ipc_export_sync.argtypes = [vp, vp]
    lib.lucid_signal.argtypes = [vp, vp, C.c_uint32]
    lib.lucid_wait_flags.argtypes = [vp, vp, C.c_int32, C.c_uint32]
    lib.lucid_set_frame_gate.argtypes = [vp, vp, C.c_uint32]
    lib.lucid_compare_render.argtypes = [vp, C.c_int32, C.POINTER(LucidConfig), vp, C.c_size_t, C.POINTER(C.c_float)]
    _host_prototypes(lib)
    _lib = lib
    return lib


def _host_prototypes(lib):
    vp = C.c_void_p
    lib.lucid_host_orbit_camera.argtypes = [C.POINTER(C.c_float * 3), C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(Camera)]
    lib.lucid_host_orbit_camera.restype = None
    lib.lucid_host_default_lighting.argtypes = [C.POINTER(Lighting)]
    lib.lucid_host_default_lighting.restype = None
    lib.lucid_host_make_config.argtypes = [C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(C.c_float * 4), C.c_int,
                                           C.c_int, C.c_int, C.POINTER(LucidConfig)]
    lib.lucid_host_make_config.restype = None
    lib.lucid_host_camera_matrices.argtypes = [C.POINTER(Camera), C.POINTER(C.c_float * 16), C.POINTER(C.c_float * 16)]
    lib.lucid_host_camera_matrices.restype = None
    lib.lucid_host_build_instances.argtypes = [C.POINTER(DrawCall), C.c_int, C.POINTER(Material), C.c_int, vp, vp,
                                               vp, C.c_int]
    lib.lucid_host_packet_size.argtypes = [C.c_int, C.c_int]
    lib.lucid_host_cluster_order.argtypes = [vp, C.c_int32, vp, C.c_int32, vp]


_host_lib = None


def load_host_library():
    """The host-side input preparation (include/lucid_host.h) from lucid_b200/_lucid_host.so: plain C++,
    no CUDA code in the file, so a process that only prepares inputs (the CPU reference arm of bench.py, the
    CPU tests) never maps the kernels' library.  The same functions are also linked into _lucid_b200.so for
    C and C++ callers."""
    global _host_lib
    if _host_lib is not None:
        return _host_lib
    path = _build.HOST_SO_PATH
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "host", "lucid_host.cpp")):
        _build.build_host()
    lib = C.CDLL(path)
    _host_prototypes(lib)
    _host_lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
# host-side input preparation (pure CPU)

DEFAULT_FOV = 60.0 * np.pi / 180.0  # src/lucid_app.cpp:95
DEFAULT_DEPTH = (1.0 / 16.0, 1024.0)


def make_camera(spec: dict, width: int, height: int) -> Camera:
    lib = load_host_library()
    cam = Camera()
    fov = float(spec.get("fov", DEFAULT_FOV))
    zn, zf = spec.get("depth", DEFAULT_DEPTH)
    if spec["kind"] == "orbit":
        center = (C.c_float * 3)(*spec["center"])
        lib.lucid_host_orbit_camera(C.byref(center), spec["distance"], spec["rot_h"], spec["rot_v"], fov, zn, zf,
                                    width, height, C.byref(cam))
    else:
        cam.pos[:] = spec["pos"]
        cam.target[:] = spec["target"]
        cam.target_up[:] = spec["up"]
        cam.fov_rad, cam.z_near, cam.z_far = fov, zn, zf
        cam.viewport_width, cam.viewport_height = width, height
    return cam


def make_config(camera: Camera, num_instances: int, background=(0.0, 30.0 / 255.0, 30.0 / 255.0, 1.0),
                backface_culling: bool = False, max_dispatches: int = 256, lighting: Lighting | None = None):
    lib = load_host_library()
    if lighting is None:
        lighting = Lighting()
        lib.lucid_host_default_lighting(C.byref(lighting))
    cfg = LucidConfig()
    bg = (C.c_float * 4)(*background)
    lib.lucid_host_make_config(C.byref(camera), C.byref(lighting), C.byref(bg), int(backface_culling),
                               num_instances, max_dispatches, C.byref(cfg))
    return cfg


def build_instances(draw_calls, materials):
    """uploadInstances: returns (instances[n] structured, colors u32[n], uv_rects f32[n,4])."""
    lib = load_host_library()
    dcs = (DrawCall * len(draw_calls))()
    for i, (mat, nq, off, opts) in enumerate(draw_calls):
        dcs[i] = DrawCall(mat, nq, off, opts)
    mats = (Material * len(materials))()
    for i, (diffuse, opacity, rect) in enumerate(materials):
        mats[i].diffuse[:] = diffuse
        mats[i].opacity = opacity
        mats[i].uv_rect[:] = rect
    cap = sum((nq + 1023) // 1024 for _, nq, _, _ in draw_calls) + 1
    inst = np.zeros((cap, 4), np.int32)
    colors = np.zeros(cap, np.uint32)
    rects = np.zeros((cap, 4), np.float32)
    n = lib.lucid_host_build_instances(dcs, len(draw_calls), mats, len(materials), _ptr(inst), _ptr(colors),
                                       _ptr(rects), cap)
    if n < 0:
        raise ValueError("lucid_host_build_instances failed")
    return inst[:n].copy(), colors[:n].copy(), rects[:n].copy()


def prepare_frame(scene: dict, camera_spec: dict | None = None, max_dispatches: int = 256):
    """scene dict (lucid_b200.scenes) -> (LucidConfig, instances, colors, uv_rects)"""
    inst, colors, rects = build_instances(scene["draw_calls"], scene["materials"])
    cam = make_camera(camera_spec or scene["camera"], scene["width"], scene["height"])
    cfg = make_config(cam, len(inst), scene["background"], max_dispatches=max_dispatches)
    return cfg, inst, colors, rects


# ---------------------------------------------------------------------------------------------


class LucidError(RuntimeError):
    pass


class LucidRenderer:
    """Mirror of the reference class (src/lucid_renderer.h): exConstruct -> __init__,
    render(ctx) -> render(...), getStats(), verifyInfo()."""

    max_instances = 64 * 1024
    max_instance_quads = 1024

    def __init__(self, width: int, height: int, opts: int = 0, max_visible_quads: int = 0, device: int = 0,
                 stream: int | None = None, bin_rows: tuple[int, int] | None = None, max_dispatches: int = 0,
                 max_block_entries: int = 0, create_flags: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        ci = CreateInfo(width, height, opts, max_visible_quads, max_dispatches, device, stream,
                        bin_rows[0] if bin_rows else 0, bin_rows[1] if bin_rows else 0, max_block_entries, create_flags)
        rc = self._lib.lucid_create(C.byref(ci), C.byref(self._h))
        if rc != 0:
            raise LucidError(f"lucid_create failed ({rc}): {self._lib.lucid_last_error(None).decode()}")
        self.width, self.height, self.opts = width, height, opts
        self.bin_count_x, self.bin_count_y = (width + 31) // 32, (height + 31) // 32
        self.bin_count = self.bin_count_x * self.bin_count_y
        self.max_visible_quads = max_visible_quads or 4793490
        self._keep = []

    def close(self):
        if self._h:
            self._lib.lucid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise LucidError(f"{what} failed ({rc}): {self._lib.lucid_last_error(self._h).decode()}")

    def set_scene(self, scene: dict):
        self.set_geometry(scene["positions"], scene["quads"], scene.get("colors"), scene.get("uvs"),
                          scene.get("normals"))
        for name, slot in (("opaque", 0), ("transparent", 1)):
            if name in scene.get("textures", {}):
                w, h, levels, data = scene["textures"][name]
                self.set_texture(slot, data, w, h, levels)

    def set_geometry(self, positions, quads, colors=None, uvs=None, normals=None):
        positions = np.ascontiguousarray(positions, np.float32)
        quads = np.ascontiguousarray(quads, np.uint32)
        colors = None if colors is None else np.ascontiguousarray(colors, np.uint32)
        uvs = None if uvs is None else np.ascontiguousarray(uvs, np.float32)
        normals = None if normals is None else np.ascontiguousarray(normals, np.uint32)
        self._check(self._lib.lucid_set_geometry(self._h, _ptr(positions), positions.shape[0], _ptr(colors),
                                                 _ptr(uvs), _ptr(normals), _ptr(quads), quads.shape[0], MEM_HOST),
                    "lucid_set_geometry")

    def set_geometry_device(self, positions_ptr, num_verts, quads_ptr, num_quads, colors_ptr=None, uvs_ptr=None,
                            normals_ptr=None):
        self._check(self._lib.lucid_set_geometry(self._h, positions_ptr, num_verts, colors_ptr, uvs_ptr, normals_ptr,
                                                 quads_ptr, num_quads, MEM_DEVICE), "lucid_set_geometry")

    def set_texture(self, slot, data, width, height, levels):
        data = np.ascontiguousarray(data, np.uint8)
        self._check(self._lib.lucid_set_texture(self._h, slot, _ptr(data), width, height, levels), "lucid_set_texture")

    def debug_sample_texture(self, slot, uvl) -> np.ndarray:
        """(u, v, lod) triples -> RGBA floats, fetched by the texture unit as the shading kernel does."""
        uvl = np.ascontiguousarray(uvl, np.float32).reshape(-1, 3)
        out = np.zeros((uvl.shape[0], 4), np.float32)
        self._check(self._lib.lucid_debug_sample_texture(self._h, slot, _ptr(uvl), uvl.shape[0], _ptr(out)),
                    "lucid_debug_sample_texture")
        return out

    def set_bin_rows(self, begin, end):
        self._check(self._lib.lucid_set_bin_rows(self._h, begin, end), "lucid_set_bin_rows")

    def render(self, config: LucidConfig, instances, colors, uv_rects=None, out: np.ndarray | None = None,
               out_device_ptr: int | None = None, out_pitch: int | None = None, flags: int = 0):
        """Renders one frame.  out: host uint32[h,w] array to receive RGBA8; out_device_ptr: device
        (or peer) pointer the raster kernels store into directly."""
        instances = np.ascontiguousarray(instances, np.int32)
        colors = np.ascontiguousarray(colors, np.uint32)
        uv_rects = None if uv_rects is None else np.ascontiguousarray(uv_rects, np.float32)
        if out_device_ptr is not None:
            mem, ptr, pitch = MEM_DEVICE, C.c_void_p(out_device_ptr), out_pitch or self.width * 4
        elif isinstance(out, int):  # address of a (pinned) host image, e.g. torch tensor.data_ptr()
            mem, ptr, pitch = MEM_HOST, C.c_void_p(out), out_pitch or self.width * 4
        elif out is not None:
            assert out.dtype == np.uint32 and out.shape == (self.height, self.width) and out.flags.c_contiguous
            mem, ptr, pitch = MEM_HOST, _ptr(out), self.width * 4
        else:
            mem, ptr, pitch = MEM_NONE, None, 0
        self._keep = [instances, colors, uv_rects, out]
        self._check(self._lib.lucid_render(self._h, C.byref(config), _ptr(instances), _ptr(colors), _ptr(uv_rects),
                                           instances.shape[0], ptr, pitch, mem, flags), "lucid_render")

    def wait(self):
        self._check(self._lib.lucid_wait(self._h), "lucid_wait")

    def compare_render(self, mode: int, config: LucidConfig):
        """The last frame's samples under a comparator's per-pixel rule (COMPARE_*): SimpleRenderer's hardware alpha
        blending in submission order (src/simple_renderer.cpp:69-132), weighted blended OIT, 4-layer MLAB.
        Returns (uint32[h, w] RGBA8 image, device milliseconds of the comparator's kernels)."""
        img = np.zeros((self.height, self.width), np.uint32)
        ms = C.c_float(0.0)
        self._check(self._lib.lucid_compare_render(self._h, mode, C.byref(config), _ptr(img), self.width * 4, C.byref(ms)),
                    "lucid_compare_render")
        return img, float(ms.value)

    # ---- results -------------------------------------------------------------------------
    def read_info(self) -> np.ndarray:
        out = np.zeros(LUCID_INFO_U32_SIZE + COUNTS_PER_BIN * self.bin_count, np.uint32)
        self._check(self._lib.lucid_read_info(self._h, _ptr(out), out.size), "lucid_read_info")
        return out

    def stage_times(self, frames_back: int = 0) -> np.ndarray:
        ms = (C.c_float * 8)()
        self._check(self._lib.lucid_stage_times_at(self._h, frames_back, C.byref(ms)), "lucid_stage_times")
        return np.array(ms[:], np.float32)

    def composite_to(self, device_ptr: int, pitch: int | None = None):
        """Copies the owned bins of the last frame (rendered into the renderer's own image) to a device /
        peer image as whole bin rows, asynchronously on the render stream."""
        self._check(self._lib.lucid_composite_to(self._h, C.c_void_p(device_ptr), pitch or self.width * 4),
                    "lucid_composite_to")

    def set_bin_range(self, begin, end):
        """Own the bins [begin, end) in row-major order (finer than whole rows)."""
        self._check(self._lib.lucid_set_bin_range(self._h, begin, end), "lucid_set_bin_range")

    def read_bin_costs(self) -> np.ndarray:
        """Warp cycles the raster kernels spent per bin in the last frame."""
        out = np.zeros(self.bin_count, np.uint64)
        self._check(self._lib.lucid_read_bin_costs(self._h, _ptr(out), out.size), "lucid_read_bin_costs")
        return out

    def read_row_costs(self) -> np.ndarray:
        """Warp cycles the raster kernels spent per bin row in the last frame (bin-row split balancing)."""
        out = np.zeros((self.height + 31) // 32, np.uint64)
        self._check(self._lib.lucid_read_row_costs(self._h, _ptr(out), out.size), "lucid_read_row_costs")
        return out

    def read_image(self) -> np.ndarray:
        out = np.zeros((self.height, self.width), np.uint32)
        self._check(self._lib.lucid_read_image(self._h, _ptr(out), self.width * 4), "lucid_read_image")
        return out

    def read_image_into(self, host_ptr: int, pitch: int | None = None):
        self._check(self._lib.lucid_read_image(self._h, C.c_void_p(host_ptr), pitch or self.width * 4),
                    "lucid_read_image")

    def read_debug_records(self, max_records: int = 4096):
        """OPT_DEBUG_RASTER: (records [n, 7] uint32 = check id, thread, work item, four values; number produced)."""
        out = np.zeros((max_records, 7), np.uint32)
        n = C.c_int32(0)
        self._check(self._lib.lucid_read_debug_records(self._h, _ptr(out), max_records, C.byref(n)), "lucid_read_debug_records")
        return out[:min(n.value, max_records)], n.value

    def read_frag_counts(self) -> np.ndarray:
        out = np.zeros((self.height, self.width), np.uint32)
        self._check(self._lib.lucid_read_frag_counts(self._h, _ptr(out)), "lucid_read_frag_counts")
        return out

    def read_quad_aabbs(self, which, count):
        out = np.zeros(count, np.uint32)
        self._check(self._lib.lucid_read_quad_aabbs(self._h, which, _ptr(out), count), "lucid_read_quad_aabbs")
        return out

    def read_tri_records(self, which, num_quads):
        out = np.zeros((num_quads * 2, 21), np.uint32)
        self._check(self._lib.lucid_read_tri_records(self._h, which, _ptr(out), num_quads), "lucid_read_tri_records")
        return out

    def read_quad_attrs(self, which, num_quads):
        out = np.zeros((num_quads, 16), np.uint32)
        self._check(self._lib.lucid_read_quad_attrs(self._h, which, _ptr(out), num_quads), "lucid_read_quad_attrs")
        return out

    def read_bin_lists(self, num_bin_quads, num_bin_tris):
        bq = np.zeros(num_bin_quads, np.uint32)
        bt = np.zeros(num_bin_tris, np.uint32)
        self._check(self._lib.lucid_read_bin_lists(self._h, _ptr(bq), bq.size, _ptr(bt), bt.size),
                    "lucid_read_bin_lists")
        return bq, bt

    def image_pointer(self):
        ptr, pitch = C.c_void_p(), C.c_size_t()
        self._check(self._lib.lucid_image_pointer(self._h, C.byref(ptr), C.byref(pitch)), "lucid_image_pointer")
        return ptr.value, pitch.value

    def ipc_export_image(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._lib.lucid_ipc_export_image(self._h, buf), "lucid_ipc_export_image")
        return buf.raw

    def ipc_open_image(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = C.create_string_buffer(handle, 64)
        self._check(self._lib.lucid_ipc_open_image(self._h, buf, C.byref(ptr)), "lucid_ipc_open_image")
        return ptr.value

    def ipc_close_image(self, ptr: int):
        self._check(self._lib.lucid_ipc_close_image(self._h, C.c_void_p(ptr)), "lucid_ipc_close_image")

    # ---- frame hand-over of the bin-row split (include/lucid_b200.h) ---------------------------------
    SYNC_RELEASED = 32

    def sync_pointer(self) -> int:
        ptr = C.c_void_p()
        self._check(self._lib.lucid_sync_pointer(self._h, C.byref(ptr)), "lucid_sync_pointer")
        return ptr.value

    def ipc_export_sync(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._lib.lucid_ipc_export_sync(self._h, buf), "lucid_ipc_export_sync")
        return buf.raw

    def signal(self, flags_ptr: int, index: int, value: int):
        self._check(self._lib.lucid_signal(self._h, C.c_void_p(flags_ptr + 4 * index), value), "lucid_signal")

    def wait_flags(self, flags_ptr: int, first: int, count: int, value: int):
        self._check(self._lib.lucid_wait_flags(self._h, C.c_void_p(flags_ptr + 4 * first), count, value), "lucid_wait_flags")

    def set_frame_gate(self, flags_ptr: int, index: int, value: int):
        self._check(self._lib.lucid_set_frame_gate(self._h, C.c_void_p(flags_ptr + 4 * index), value), "lucid_set_frame_gate")

    # ---- the reference's host-side decoders --------------------------------------------------
    def getStats(self, info: np.ndarray | None = None) -> dict:
        """LucidRenderer::getStats (src/lucid_renderer.cpp:709-836) as a dict."""
        return decode_stats(self.read_info() if info is None else info, self.bin_count, self.width, self.height)

    def verifyInfo(self, info: np.ndarray | None = None) -> list[str]:
        """LucidRenderer::verifyInfo (src/lucid_renderer.cpp:629-707): returns error strings."""
        return verify_info(self.read_info() if info is None else info, self.bin_count)


# LucidInfo word offsets (include/lucid_abi.h)
_OFF = dict(num_input_quads=0, num_visible_quads=1, num_counted_quads=3, bin_level_counts=5, a_small_bins=10,
            a_high_bins=11, num_binning_dispatches=14, bin_level_dispatches=17, num_rejected_quads=32,
            setup_timers=36, raster_timers=44, bin_dispatcher_timers=52, stats=60, temp=1088)


def split_info(info: np.ndarray, bin_count: int):
    head = info[:LUCID_INFO_U32_SIZE]
    counts = info[LUCID_INFO_U32_SIZE:LUCID_INFO_U32_SIZE + COUNTS_PER_BIN * bin_count].reshape(COUNTS_PER_BIN,
                                                                                                  bin_count)
    return head, counts.view(np.int32)


def decode_stats(info: np.ndarray, bin_count: int, width: int, height: int) -> dict:
    head, counts = split_info(info, bin_count)
    h = head.view(np.int32)
    levels = h[_OFF["bin_level_counts"]:_OFF["bin_level_counts"] + 5]
    rejected = head[_OFF["num_rejected_quads"]:_OFF["num_rejected_quads"] + 4]
    stats = head[_OFF["stats"]:_OFF["stats"] + 4]
    return dict(
        input_quads=int(h[0]), visible_small=int(h[1]), visible_large=int(h[2]),
        rejected_other=int(rejected[0]), rejected_backface=int(rejected[1]), rejected_frustum=int(rejected[2]),
        rejected_between_samples=int(rejected[3]),
        bin_quads=int(counts[0].sum()), bin_tris=int(counts[3].sum()),
        max_quads_per_bin=int(counts[0].max(initial=0)), max_tris_per_bin=int(counts[3].max(initial=0)),
        empty_bins=int(levels[0]), low_bins=int(levels[2]), high_bins=int(levels[4]),
        promoted_bins=int(levels.sum()) - bin_count,
        fragments=int(stats[0]), half_block_tris=int(stats[1]), invalid_pixels=int(stats[2]),
        avg_fragments_per_pixel=float(stats[0]) / (width * height),
        dropped_quads=int(h[_OFF["temp"]]), list_overflow=int(h[_OFF["temp"] + 1]),
        # LUCID_OPT_TIMERS: clock ticks >> 4 per phase, in the reference's slot meaning (lucid_renderer.cpp:754-762)
        setup_timers=dict(zip(("init & finish", "process input quads", "store tri data", "store quad data"),
                              (int(v) for v in head[_OFF["setup_timers"]:_OFF["setup_timers"] + 4]))),
        bin_dispatcher_timers=dict(zip(("count small quads", "count large tris", "dispatch small quads", "dispatch large tris"),
                                       (int(v) for v in head[_OFF["bin_dispatcher_timers"]:_OFF["bin_dispatcher_timers"] + 4]))),
        raster_timers=dict(zip(("generate rows", "generate blocks", "unpack samples", "shade and reduce", "finish reduce"),
                               (int(v) for v in head[_OFF["raster_timers"]:_OFF["raster_timers"] + 5]))),
    )


def verify_info(info: np.ndarray, bin_count: int) -> list[str]:
    _, c = split_info(info, bin_count)
    errors = []
    for name, (cnt, off, tmp) in (("quad", (0, 1, 2)), ("tri", (3, 4, 5))):
        counts, offsets, temp = c[cnt].astype(np.int64), c[off].astype(np.int64), c[tmp].astype(np.int64)
        bad = np.nonzero(offsets[1:] != offsets[:-1] + counts[:-1])[0]
        errors += [f"Invalid bin {name} offset [{i + 1}]" for i in bad[:32]]
        bad = np.nonzero(temp != offsets + counts)[0]
        errors += [f"Invalid temp bin {name} offset [{i}]" for i in bad[:32]]
    return errors
