"""Host-side input preparation (C++ in lucid_b200/host) against golden vectors produced by the
reference's own libfwk camera code (tests/golden/ref_camera.json, made by make_ref_camera.py),
and the instance slicing rules of LucidRenderer::uploadInstances."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from lucid_b200 import api, scenes

HERE = os.path.dirname(os.path.abspath(__file__))


def _v(v):
    return np.array([v.x, v.y, v.z], np.float64)


def _all_cases():
    with open(os.path.join(HERE, "golden", "ref_camera.json")) as f:
        return json.load(f)


def _cases():
    return [c for c in _all_cases() if c["kind"] != "color"]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c['kind']}-{c['args'][-2]}x{c['args'][-1]}")
def test_config_matches_reference_camera(case):
    a = case["args"]
    if case["kind"] == "orbit":
        spec = dict(kind="orbit", center=a[0:3], distance=a[3], rot_h=a[4], rot_v=a[5], fov=math.radians(a[6]),
                    depth=(a[7], a[8]))
    else:
        spec = dict(kind="lookat", pos=a[0:3], target=a[3:6], up=a[6:9], fov=math.radians(a[9]), depth=(a[10], a[11]))
    w, h = int(a[-2]), int(a[-1])
    cam = api.make_camera(spec, w, h)
    cfg = api.make_config(cam, 7)
    # tolerance: a few float ulps of the magnitudes involved (libfwk uses a general 4x4 inverse,
    # the host code an analytic rigid inverse)
    for i in range(4):
        np.testing.assert_allclose(_v(cfg.frustum.ws_dirs[i]), case[f"dir{i}"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(_v(cfg.frustum.ws_origins[i]), case[f"origin{i}"], rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(_v(cfg.frustum.ws_dir0), case["dir0"], atol=2e-6)
    np.testing.assert_allclose(_v(cfg.frustum.ws_dirx), case["dirx"], atol=5e-9, rtol=2e-4)
    np.testing.assert_allclose(_v(cfg.frustum.ws_diry), case["diry"], atol=5e-9, rtol=2e-4)
    vp = np.array([[cfg.view_proj_matrix[c].x, cfg.view_proj_matrix[c].y, cfg.view_proj_matrix[c].z,
                    cfg.view_proj_matrix[c].w] for c in range(4)], np.float64).reshape(-1)
    np.testing.assert_allclose(vp, case["view_proj"], rtol=2e-5, atol=2e-5)
    assert cfg.num_instances == 7 and cfg.instance_packet_size == 1


def test_reference_camera_binary_matches_golden_when_available():
    """In the build container (reference mounted) the golden file must equal a fresh run."""
    binary = os.path.join(HERE, "..", "oracle", "_ref", "ref_camera")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref/ref_camera not built (no reference tree here)")
    import subprocess
    case = _cases()[0]
    a = case["args"]
    txt = subprocess.run([binary, "orbit"] + [repr(float(x)) for x in a[:-2]] + [str(int(a[-2])), str(int(a[-1]))],
                         capture_output=True, text=True, check=True).stdout
    got = {ln.split()[0]: [float(v) for v in ln.split()[1:]] for ln in txt.strip().splitlines()}
    assert got["dir0"] == case["dir0"] and got["view_proj"] == case["view_proj"]


def test_frustum_rays_hit_projected_pixels():
    """dir0 + px*dirx + py*diry must be the ray through pixel (px,py) of view_proj (self-consistency)."""
    cam = api.make_camera(dict(kind="orbit", center=(0, 0, 0), distance=30.0, rot_h=0.5, rot_v=0.8), 1280, 720)
    cfg = api.make_config(cam, 1)
    org = _v(cfg.frustum.ws_origin0)
    m = np.array([[getattr(cfg.view_proj_matrix[c], k) for k in "xyzw"] for c in range(4)], np.float64).T
    for px, py in ((0.0, 0.0), (640.0, 360.0), (1280.0, 720.0), (100.5, 600.25)):
        d = _v(cfg.frustum.ws_dir0) + px * _v(cfg.frustum.ws_dirx) + py * _v(cfg.frustum.ws_diry)
        p = org + 17.0 * d
        clip = m @ np.array([p[0], p[1], p[2], 1.0])
        sx = (clip[0] / clip[3] + 1.0) * 640.0
        sy = (clip[1] / clip[3] + 1.0) * 360.0
        assert abs(sx - px) < 2e-2 and abs(sy - py) < 2e-2, (px, py, sx, sy)


def test_default_lighting_and_packet_size():
    lib = api.load_host_library()
    light = api.Lighting()
    lib.lucid_host_default_lighting(C.byref(light))
    assert abs(light.sun_power - 2.5) < 1e-7 and abs(light.ambient_power - 0.4) < 1e-7
    assert abs(light.sun_dir.x - 0.842121) < 1e-6 and abs(light.ambient_color.z - 0.6) < 1e-7
    # clamp(num_instances / (max_dispatches / 2), 1, 2), lucid_renderer.cpp:424-426
    assert lib.lucid_host_packet_size(10, 256) == 1
    assert lib.lucid_host_packet_size(255, 256) == 1
    assert lib.lucid_host_packet_size(256, 256) == 2
    assert lib.lucid_host_packet_size(5000, 256) == 2


def test_build_instances_slices_and_colours():
    # one draw call of 2500 quads -> 1024 + 1024 + 452; colour truncates x255; white opaque has no INST_HAS_COLOR
    dcs = [(0, 2500, 10, scenes.INST_HAS_VERTEX_COLORS), (1, 0, 0, 0), (1, 7, 4000, 0)]
    mats = [((1.0, 0.5, 0.25), 0.5, (0.0, 0.0, 1.0, 1.0)), ((1.0, 1.0, 1.0), 1.0, (0.25, 0.5, 0.125, 0.125))]
    inst, cols, rects = api.build_instances(dcs, mats)
    assert inst.shape == (4, 4)
    assert inst[:, 2].tolist() == [1024, 1024, 452, 7]
    assert inst[:, 0].tolist() == [40, 40 + 4096, 40 + 8192, 16000]
    assert (inst[:, 1] == 0).all()
    c0 = 255 | (127 << 8) | (63 << 16) | (127 << 24)
    assert cols.tolist() == [c0, c0, c0, 0xFFFFFFFF]
    f = inst[:, 3].astype(np.uint32)
    assert (f[:3] == (scenes.INST_HAS_VERTEX_COLORS | scenes.INST_HAS_COLOR)).all() and f[3] == 0
    assert rects[3].tolist() == [0.25, 0.5, 0.125, 0.125]


def test_material_colours_match_libfwk():
    """The instance colour of uploadInstances is u32(IColor(FColor(diffuse, opacity))) (src/lucid_renderer.cpp:364-365):
    libfwk's conversion, compiled from the reference tree (oracle/_ref/ref_camera color ...), against
    lucid_host_build_instances."""
    cases = [c for c in _all_cases() if c["kind"] == "color"]
    assert len(cases) >= 8
    for c in cases:
        r, g, b, opacity = c["args"]
        _, cols, _ = api.build_instances([(0, 1, 0, 0)], [((r, g, b), opacity, (0.0, 0.0, 1.0, 1.0))])
        assert int(cols[0]) == c["color"], (c["args"], hex(int(cols[0])), hex(c["color"]))


def test_struct_layouts():
    assert C.sizeof(api.LucidConfig) == 352 and C.sizeof(api.InstanceData) == 16
    assert api.LucidConfig.view_proj_matrix.offset == 192 and api.LucidConfig.lighting.offset == 256
    assert api.LucidConfig.background_color.offset == 320 and api.LucidConfig.num_instances.offset == 340


def test_c_abi_exports_every_declared_symbol():
    lib = api.load_library()
    declared = set()
    import re
    for hdr in ("lucid_b200.h", "lucid_host.h", "lucid_quadgen.h"):
        with open(os.path.join(HERE, "..", "include", hdr)) as f:
            txt = f.read()
        declared |= set(re.findall(r"\b(lucid_[a-z0-9_]+)\s*\(", txt))
    declared -= {"lucid_renderer"}
    assert declared == set(api.C_ABI_SYMBOLS), declared ^ set(api.C_ABI_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    # the host-only library (what input preparation in Python loads) carries every lucid_host_* symbol
    host = api.load_host_library()
    for name in declared:
        if name.startswith("lucid_host_"):
            assert getattr(host, name) is not None


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.LucidError) as e:
        api.LucidRenderer(640, 360)
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_bad_create_arguments():
    lib = api.load_library()
    h = C.c_void_p()
    ci = api.CreateInfo(0, 0, 0, 0, 0, 0, None, 0, 0, 0)
    assert lib.lucid_create(C.byref(ci), C.byref(h)) == -1
    ci = api.CreateInfo(8192, 100, 0, 0, 0, 0, None, 0, 0, 0)
    assert lib.lucid_create(C.byref(ci), C.byref(h)) == -3  # 7-bit bin coordinates
    assert b"4096" in lib.lucid_last_error(None)


def test_committed_dram_traffic_file_serves_the_bench_line():
    """bench.py takes roofline.traffic / sm_issue_active_pct from profiles/dram_traffic.json (written by
    tools/ncu_traffic.py from an ncu --set full capture) -- but only while the kernel sources are the ones the
    capture was taken with: the file records their digest, and a line produced after a kernel edit carries
    `traffic: null` with the reason instead of a stale figure."""
    import json
    import bench
    path = os.path.join(HERE, "..", "profiles", "dram_traffic.json")
    with open(path) as f:
        d = json.load(f)
    assert len(d["kernel_source_hash"]) == 16 and len(bench.kernel_source_hash()) == 16
    for cfg in ("config3",):
        for stage in ("setup", "bin_count", "bin_dispatch", "raster"):
            assert d[cfg][stage] > 0
            assert 0 < d[cfg]["issue_active_pct"][stage] <= 100
        assert "source" in d[cfg] and sum(d[cfg]["per_kernel"].values()) > 0
    if d["kernel_source_hash"] != bench.kernel_source_hash():
        import warnings
        warnings.warn("profiles/dram_traffic.json is older than the kernel sources: bench.py reports traffic = null "
                      "until tools/gpu_round.sh has taken a new capture")


def test_shared_host_images_are_one_memory():
    """multigpu.SharedHostImages: what one process of a split writes into the shared frame, the others see."""
    from lucid_b200 import multigpu
    name = "lucid_b200_test_%d" % os.getpid()
    a = multigpu.SharedHostImages(name, 64, 32, 2, create=True)
    try:
        b = multigpu.SharedHostImages(name, 64, 32, 2, create=False)
        a.array[1, 3, 5] = 0xCAFE
        assert b.array[1, 3, 5] == 0xCAFE and b.pointer(1) - b.pointer(0) == 64 * 32 * 4
        b.close()
    finally:
        a.close()
    assert not os.path.exists("/dev/shm/" + name)


def test_public_headers_are_plain_c():
    """The drop-in boundary is a C ABI: every header under include/ except the C++ facade compiles as C99 on its own
    (plain pointers and sizes, no C++ or CUDA types in a signature)."""
    import subprocess

    inc = os.path.join(HERE, "..", "include")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    for hdr in ("lucid_abi.h", "lucid_b200.h", "lucid_host.h", "lucid_quadgen.h"):
        res = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c", "-"],
                             input=f'#include "{hdr}"\n', capture_output=True, text=True)
        assert res.returncode == 0, hdr + ": " + res.stderr


def test_bench_roofline_arithmetic():
    """bench.algorithmic_bytes: SURVEY 8(d)'s byte formulas with this implementation's record sizes, and the
    depth-complexity summary of the bench line's counters."""
    import bench

    stats = dict(input_quads=1000, visible_small=600, visible_large=40, bin_quads=900, bin_tris=300, low_bins=7, high_bins=3)
    plain = bench.algorithmic_bytes(stats, {}, 1920, 1080)
    n_vis, t_bin = 640, 2 * 900 + 300
    assert plain["setup"] == 64 * 1000 + (4 + 192) * n_vis
    assert plain["bin_count"] == 4 * n_vis + 64 * 40
    assert plain["bin_dispatch"] == plain["bin_count"] + 4 * 1200
    assert plain["raster"] == 4 * 1200 + 96 * t_bin + 4 * 10 * 1024
    full = bench.algorithmic_bytes(stats, dict(colors=1, normals=1, uvs=1), 1920, 1080)
    assert full["setup"] - plain["setup"] == 64 * n_vis and full["raster"] - plain["raster"] == 32 * t_bin
    frag = np.zeros((4, 8), np.uint32)
    frag[0, :4] = [1, 2, 3, 100]
    d = bench.depth_complexity(frag)
    assert d == {"covered_frac": 0.125, "median": 2.5, "p99": float(np.percentile([1, 2, 3, 100], 99)), "max": 100}
    assert bench.depth_complexity(np.zeros((2, 2), np.uint32))["covered_frac"] == 0.0
