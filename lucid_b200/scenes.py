"""Deterministic synthetic scenes for the BASELINE.json configs.

Every scene is a dict of numpy arrays in exactly the form the render call consumes:
positions float32[nv,3], quads uint32[nq,4] (two triangles (v0,v1,v2),(v0,v2,v3) per quad,
quad_generator.cpp:192 convention), optional per-vertex colors (RGBA8), tex coords and 10-10-10
normals (scene.cpp:338-343), and a list of draw calls + materials that
lucid_host_build_instances() slices into <= 1024-quad instances the way
LucidRenderer::uploadInstances does (lucid_renderer.cpp:352-429).

Random numbers come from a counter-based generator (PCG output hash of seed/stream/index), so a
scene is a pure function of its seed and every element can be produced independently (vectorised);
floats are (u >> 8) * 2^-24 as SURVEY 8d prescribes.
"""
from __future__ import annotations

import math

import numpy as np

INST_HAS_VERTEX_COLORS = 0x001
INST_HAS_VERTEX_TEX_COORDS = 0x002
INST_HAS_VERTEX_NORMALS = 0x004
INST_IS_OPAQUE = 0x008
INST_TEX_OPAQUE = 0x010
INST_HAS_UV_RECT = 0x020
INST_HAS_ALBEDO_TEXTURE = 0x040
INST_HAS_COLOR = 0x200

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _pcg_hash(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """32-bit output for (seed, stream, idx): one LCG step of a per-index state + PCG XSH-RR."""
    with np.errstate(over="ignore"):
        idx = idx.astype(np.uint64)
        inc = np.uint64(((stream << 1) | 1) & 0xFFFFFFFFFFFFFFFF)
        state = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)) & _M64
        state = (state * np.uint64(6364136223846793005) + inc) & _M64
        state = (state ^ (state >> np.uint64(29))) & _M64
        state = (state * np.uint64(6364136223846793005) + inc) & _M64
        xorshifted = (((state >> np.uint64(18)) ^ state) >> np.uint64(27)).astype(np.uint32)
        rot = (state >> np.uint64(59)).astype(np.uint32)
        out = (xorshifted >> rot) | (xorshifted << ((np.uint32(32) - rot) & np.uint32(31)))
    return out.astype(np.uint32)


class Rng:
    """rng.u32(n) / rng.f32(n): each call consumes a fresh stream id."""

    def __init__(self, seed: int):
        self.seed = seed
        self.stream = 0

    def u32(self, n: int) -> np.ndarray:
        self.stream += 1
        return _pcg_hash(self.seed, self.stream, np.arange(n, dtype=np.uint64))

    def f32(self, n: int) -> np.ndarray:
        return ((self.u32(n) >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)).astype(np.float32)

    def uniform(self, lo: float, hi: float, n: int) -> np.ndarray:
        return (np.float32(lo) + self.f32(n) * np.float32(hi - lo)).astype(np.float32)

    def normal_dirs(self, n: int) -> np.ndarray:
        """uniform unit vectors"""
        z = self.uniform(-1.0, 1.0, n)
        phi = self.uniform(0.0, 2.0 * math.pi, n)
        r = np.sqrt(np.maximum(0.0, 1.0 - z * z)).astype(np.float32)
        return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)


def encode_normal_uint(n: np.ndarray) -> np.ndarray:
    q = (np.float32(512.0) + n.astype(np.float32) * np.float32(511.0)).astype(np.uint32) & np.uint32(0x3FF)
    return (q[:, 0] | (q[:, 1] << np.uint32(10)) | (q[:, 2] << np.uint32(20))).astype(np.uint32)


def _orthonormal_frames(rng: Rng, n: int):
    a = rng.normal_dirs(n)
    b = rng.normal_dirs(n)
    u = a
    v = np.cross(u, b)
    ln = np.linalg.norm(v, axis=1, keepdims=True)
    bad = ln[:, 0] < 1e-3
    v = np.where(bad[:, None], np.cross(u, np.array([0.0, 0.0, 1.0], np.float32) + 0.5 * u[:, ::-1]), v)
    v = (v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)).astype(np.float32)
    return u.astype(np.float32), v


def _scene(positions, quads, draw_calls, materials, camera, width, height, *, colors=None, uvs=None,
           normals=None, textures=None, name=""):
    return dict(
        name=name,
        positions=np.ascontiguousarray(positions, np.float32),
        quads=np.ascontiguousarray(quads, np.uint32),
        colors=None if colors is None else np.ascontiguousarray(colors, np.uint32),
        uvs=None if uvs is None else np.ascontiguousarray(uvs, np.float32),
        normals=None if normals is None else np.ascontiguousarray(normals, np.uint32),
        draw_calls=draw_calls,  # list of (material_id, num_quads, quad_offset, opts)
        materials=materials,  # list of (diffuse rgb, opacity, uv_rect)
        camera=camera,  # dict(kind="orbit", center, distance, rot_h, rot_v) or kind="lookat"
        width=width,
        height=height,
        textures=textures or {},
        background=(0.0, 30.0 / 255.0, 30.0 / 255.0, 1.0),
    )


def _slice_draw_calls(num_quads: int, per_instance: int = 1024):
    out = []
    for off in range(0, num_quads, per_instance):
        out.append((off, min(per_instance, num_quads - off)))
    return out


def quad_soup(num_quads=50_000, seed=1, width=1280, height=720, extent=10.0, min_edge=0.05,
              max_edge=1.0, distance=30.0, alpha_even=127, name="config1_soup"):
    """Config 1: random planar quads in a cube, alternating opaque / alpha=0.5 instances."""
    rng = Rng(seed)
    n = num_quads
    centers = np.stack([rng.uniform(-extent, extent, n) for _ in range(3)], axis=1)
    ln_lo, ln_hi = math.log(min_edge), math.log(max_edge)
    ea = np.exp(rng.uniform(ln_lo, ln_hi, n)).astype(np.float32)
    eb = np.exp(rng.uniform(ln_lo, ln_hi, n)).astype(np.float32)
    u, v = _orthonormal_frames(rng, n)
    hu = u * (0.5 * ea)[:, None]
    hv = v * (0.5 * eb)[:, None]
    corners = np.stack([centers - hu - hv, centers + hu - hv, centers + hu + hv, centers - hu + hv], axis=1)
    positions = corners.reshape(-1, 3).astype(np.float32)
    quads = np.arange(n * 4, dtype=np.uint32).reshape(n, 4)

    slices = _slice_draw_calls(n)
    rgb = np.stack([rng.f32(len(slices)) for _ in range(3)], axis=1)
    draw_calls, materials = [], []
    for i, (off, cnt) in enumerate(slices):
        if i % 2 == 1:
            opacity, opts = 1.0, INST_IS_OPAQUE
        else:
            opacity, opts = (alpha_even + 0.5) / 255.0, 0
        materials.append((tuple(float(c) for c in rgb[i]), opacity, (0.0, 0.0, 1.0, 1.0)))
        draw_calls.append((i, cnt, off, opts))
    camera = dict(kind="orbit", center=(0.0, 0.0, 0.0), distance=distance, rot_h=0.5, rot_v=0.8)
    return _scene(positions, quads, draw_calls, materials, camera, width, height, name=name)


def meshlet_patches(num_patches=489, seed=2, width=1920, height=1080, grid=32, extent=20.0,
                    distance=50.0, rot_h=0.5, rot_v=0.6, name="config2_meshlets"):
    """Config 2: grid x grid-quad height-field patches with vertex colours and vertex normals."""
    rng = Rng(seed)
    g = grid
    nv_p = (g + 1) * (g + 1)
    centers = np.stack([rng.uniform(-extent, extent, num_patches) for _ in range(3)], axis=1)
    sizes = rng.uniform(1.0, 4.0, num_patches)
    u, v = _orthonormal_frames(rng, num_patches)
    w = np.cross(u, v).astype(np.float32)
    freq = rng.uniform(1.0, 4.0, num_patches)
    phase = rng.uniform(0.0, 6.2831853, num_patches)
    amp = rng.uniform(0.02, 0.15, num_patches) * sizes

    s = np.linspace(-0.5, 0.5, g + 1, dtype=np.float32)
    gx, gy = np.meshgrid(s, s, indexing="xy")
    gx = gx.reshape(-1)
    gy = gy.reshape(-1)
    # height h(x,y) = amp * sin(f*2pi*x + p) * cos(f*2pi*y)
    ax = (freq[:, None] * np.float32(2 * math.pi)) * gx[None, :] + phase[:, None]
    ay = (freq[:, None] * np.float32(2 * math.pi)) * gy[None, :]
    hgt = amp[:, None] * np.sin(ax) * np.cos(ay)
    px = gx[None, :] * sizes[:, None]
    py = gy[None, :] * sizes[:, None]
    positions = (centers[:, None, :] + px[:, :, None] * u[:, None, :] + py[:, :, None] * v[:, None, :]
                 + hgt[:, :, None] * w[:, None, :]).astype(np.float32)
    # analytic normals of the height field in the patch frame
    k = freq[:, None] * np.float32(2 * math.pi) / sizes[:, None]
    dhdx = amp[:, None] * np.cos(ax) * np.cos(ay) * k
    dhdy = -amp[:, None] * np.sin(ax) * np.sin(ay) * k
    nrm = (w[:, None, :] - dhdx[:, :, None] * u[:, None, :] - dhdy[:, :, None] * v[:, None, :])
    nrm = nrm / np.linalg.norm(nrm, axis=2, keepdims=True)
    normals = encode_normal_uint(nrm.reshape(-1, 3))
    positions = positions.reshape(-1, 3)

    cu = rng.u32(num_patches * nv_p)
    colors = (cu | np.uint32(0xFF000000)).astype(np.uint32)  # opaque vertex colours

    jj, ii = np.meshgrid(np.arange(g, dtype=np.uint32), np.arange(g, dtype=np.uint32), indexing="ij")
    v0 = (jj * (g + 1) + ii).reshape(-1)
    local = np.stack([v0, v0 + 1, v0 + 1 + (g + 1), v0 + (g + 1)], axis=1).astype(np.uint32)
    quads = (local[None, :, :] + (np.arange(num_patches, dtype=np.uint32) * nv_p)[:, None, None]).reshape(-1, 4)

    alphas = [64, 127, 191]
    pick = rng.u32(num_patches)
    draw_calls, materials = [], []
    qpp = g * g
    for i in range(num_patches):
        opts = INST_HAS_VERTEX_COLORS | INST_HAS_VERTEX_NORMALS
        if pick[i] & 1:
            opacity, opts = 1.0, opts | INST_IS_OPAQUE
        else:
            opacity = (alphas[int(pick[i] >> 1) % 3] + 0.5) / 255.0
        materials.append(((1.0, 1.0, 1.0), opacity, (0.0, 0.0, 1.0, 1.0)))
        for off, cnt in _slice_draw_calls(qpp):
            draw_calls.append((i, cnt, i * qpp + off, opts))
    camera = dict(kind="orbit", center=(0.0, 0.0, 0.0), distance=distance, rot_h=rot_h, rot_v=rot_v)
    return _scene(positions, quads, draw_calls, materials, camera, width, height, colors=colors,
                  normals=normals, name=name)


def hairball(num_strands=39_063, segments=64, seed=3, width=3840, height=2160, radius=10.0,
             ribbon_width=None, distance=22.0, alpha=127, start_frac=0.35, name="config3_hairball"):
    """Config 3: curly ribbons inside a sphere; dense overlap drives bins into raster_high."""
    rng = Rng(seed)
    ns, sg = num_strands, segments
    if ribbon_width is None:
        ribbon_width = 0.002 * radius
    start = rng.normal_dirs(ns) * (rng.f32(ns) ** np.float32(1.0 / 3.0))[:, None] * np.float32(start_frac * radius)
    d = rng.normal_dirs(ns)
    step = np.float32(1.6 * radius / sg)
    pts = np.empty((ns, sg + 1, 3), np.float32)
    pts[:, 0] = start
    cur = start.copy()
    for s in range(sg):
        curl = rng.normal_dirs(ns)
        d = d + np.float32(0.45) * curl
        # steer back inside the ball
        rr = np.linalg.norm(cur, axis=1, keepdims=True)
        d = d - cur * (np.maximum(rr - 0.8 * radius, 0.0) / radius * 2.0 / np.maximum(rr, 1e-6))
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        cur = (cur + d * step).astype(np.float32)
        pts[:, s + 1] = cur
    side = np.cross(np.diff(pts, axis=1, append=pts[:, -1:] * 2 - pts[:, -2:-1]), rng.normal_dirs(ns)[:, None, :])
    side = side / np.maximum(np.linalg.norm(side, axis=2, keepdims=True), 1e-9) * np.float32(0.5 * ribbon_width)
    left = (pts - side).astype(np.float32)
    right = (pts + side).astype(np.float32)
    positions = np.stack([left, right], axis=2).reshape(-1, 3)  # per strand: (sg+1) x 2 verts
    base = (np.arange(ns, dtype=np.uint32) * np.uint32((sg + 1) * 2))[:, None]
    k = np.arange(sg, dtype=np.uint32)[None, :] * np.uint32(2)
    quads = np.stack([base + k, base + k + 1, base + k + 3, base + k + 2], axis=2).reshape(-1, 4)

    nq = ns * sg
    slices = _slice_draw_calls(nq)
    rgb = np.stack([rng.uniform(0.3, 1.0, len(slices)) for _ in range(3)], axis=1)
    draw_calls, materials = [], []
    for i, (off, cnt) in enumerate(slices):
        materials.append((tuple(float(c) for c in rgb[i]), (alpha + 0.5) / 255.0, (0.0, 0.0, 1.0, 1.0)))
        draw_calls.append((i, cnt, off, 0))
    camera = dict(kind="orbit", center=(0.0, 0.0, 0.0), distance=distance, rot_h=0.5, rot_v=0.6)
    return _scene(positions, quads, draw_calls, materials, camera, width, height, name=name)


def _procedural_atlas(size: int, levels: int, seed: int, holes: bool) -> np.ndarray:
    """RGBA8 mip chain (box filtered), tightly packed level after level."""
    rng = Rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    cell = 64
    checker = (((x // cell) + (y // cell)) & 1).astype(np.float32)
    tile = (x // cell) * 131 + (y // cell) * 71
    base = np.stack([((tile * 37) % 255), ((tile * 59) % 255), ((tile * 83) % 255)], axis=2).astype(np.float32)
    noise = (rng.u32(size * size) & np.uint32(63)).astype(np.float32).reshape(size, size)
    rgb = np.clip(base * (0.6 + 0.4 * checker[:, :, None]) + noise[:, :, None] - 32.0, 0, 255)
    if holes:
        cx = (x % cell) - cell / 2
        cy = (y % cell) - cell / 2
        a = np.where(cx * cx + cy * cy < (cell * 0.3) ** 2, 0.0, 200.0).astype(np.float32)
    else:
        a = np.full((size, size), 255.0, np.float32)
    img = np.concatenate([rgb, a[:, :, None]], axis=2)
    out = []
    cur = img
    for _ in range(levels):
        out.append(np.clip(np.floor(cur + 0.5), 0, 255).astype(np.uint8).reshape(-1))
        if cur.shape[0] > 1:
            cur = 0.25 * (cur[0::2, 0::2] + cur[1::2, 0::2] + cur[0::2, 1::2] + cur[1::2, 1::2])
    return np.concatenate(out)


def architecture(num_small=4_900_000, num_large=2_000, seed=4, width=3840, height=2160,
                 atlas_opaque=4096, atlas_trans=2048, levels=6, name="config4_architecture"):
    """Config 4: large wall/floor quads (large-triangle path) + clustered small textured clutter."""
    rng = Rng(seed)
    room = 40.0
    # large axis-aligned-ish quads
    nl = num_large
    lc = np.stack([rng.uniform(-room, room, nl), rng.uniform(-4.0, 12.0, nl), rng.uniform(-room, room, nl)], axis=1)
    lu, lv = _orthonormal_frames(rng, nl)
    la = rng.uniform(4.0, 20.0, nl)
    lb = rng.uniform(2.0, 10.0, nl)
    # clustered clutter
    ns = num_small
    ncl = max(1, ns // 2048)
    cc = np.stack([rng.uniform(-room, room, ncl), rng.uniform(-3.0, 8.0, ncl), rng.uniform(-room, room, ncl)], axis=1)
    cr = rng.uniform(0.5, 3.0, ncl)
    cid = (np.arange(ns) // 2048).clip(0, ncl - 1)
    off = rng.normal_dirs(ns) * (rng.f32(ns) ** np.float32(1.0 / 3.0))[:, None]
    sc = (cc[cid] + off * cr[cid][:, None]).astype(np.float32)
    su, sv = _orthonormal_frames(rng, ns)
    sa = np.exp(rng.uniform(math.log(0.02), math.log(0.3), ns)).astype(np.float32)
    sb = np.exp(rng.uniform(math.log(0.02), math.log(0.3), ns)).astype(np.float32)

    centers = np.concatenate([lc, sc]).astype(np.float32)
    u = np.concatenate([lu, su])
    v = np.concatenate([lv, sv])
    ea = np.concatenate([la, sa]).astype(np.float32)
    eb = np.concatenate([lb, sb]).astype(np.float32)
    n = nl + ns
    hu = u * (0.5 * ea)[:, None]
    hv = v * (0.5 * eb)[:, None]
    positions = np.stack([centers - hu - hv, centers + hu - hv, centers + hu + hv, centers - hu + hv],
                         axis=1).reshape(-1, 3).astype(np.float32)
    quads = np.arange(n * 4, dtype=np.uint32).reshape(n, 4)
    rep = np.concatenate([np.maximum(1.0, ea / 2.0), np.ones(0, np.float32)])[:n]
    uv_corner = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    uvs = (uv_corner[None, :, :] * np.stack([np.maximum(1.0, np.round(ea)), np.maximum(1.0, np.round(eb))], axis=1)[:, None, :])
    uvs = uvs.reshape(-1, 2).astype(np.float32)

    slices = _slice_draw_calls(n)
    ni = len(slices)
    rgb = np.stack([rng.uniform(0.5, 1.0, ni) for _ in range(3)], axis=1)
    rect = np.stack([np.floor(rng.f32(ni) * 8) / 8, np.floor(rng.f32(ni) * 8) / 8], axis=1)
    pick = rng.u32(ni)
    draw_calls, materials = [], []
    for i, (o, cnt) in enumerate(slices):
        opts = INST_HAS_VERTEX_TEX_COORDS | INST_HAS_ALBEDO_TEXTURE | INST_HAS_UV_RECT
        if pick[i] & 1:
            opts |= INST_TEX_OPAQUE | INST_IS_OPAQUE
            opacity = 1.0
        else:
            opacity = (191 + 0.5) / 255.0
        materials.append((tuple(float(c) for c in rgb[i]), opacity,
                          (float(rect[i, 0]), float(rect[i, 1]), 0.125, 0.125)))
        draw_calls.append((i, cnt, o, opts))
    textures = {
        "opaque": (atlas_opaque, atlas_opaque, levels, _procedural_atlas(atlas_opaque, levels, seed * 17 + 1, False)),
        "transparent": (atlas_trans, atlas_trans, levels, _procedural_atlas(atlas_trans, levels, seed * 17 + 2, True)),
    }
    camera = dict(kind="lookat", pos=(-30.0, 3.0, -28.0), target=(10.0, 1.0, 12.0), up=(0.0, 1.0, 0.0))
    return _scene(positions, quads, draw_calls, materials, camera, width, height, uvs=uvs,
                  textures=textures, name=name)


def planes(num_planes=32, width=1280, height=720, opacity=0.25, plane_size=2.0, plane_dist=0.1,
           name="planes"):
    """The reference's '#planes' known-answer scene (scene_setup.cpp:198-222): stacked parallel
    quads, plane z scaled by 5% each, hue ramp colours, viewed head on."""
    pos, quads, cols = [], [], []
    for z in range(num_planes):
        size = plane_size * (1.0 + z * 0.05)
        t = z / max(1, num_planes - 1)
        # hsvToRgb(t, 1, 1)
        h6 = t * 6.0
        c = [abs(h6 - 3.0) - 1.0, 2.0 - abs(h6 - 2.0), 2.0 - abs(h6 - 4.0)]
        rgb = [min(max(v, 0.0), 1.0) for v in c]
        x0, y0 = -0.5 * size, -0.5 * size
        zz = z * plane_dist
        base = len(pos)
        pos += [(x0, y0, zz), (x0 + size, y0, zz), (x0 + size, y0 + size, zz), (x0, y0 + size, zz)]
        quads.append((base, base + 1, base + 2, base + 3))
        ic = int(rgb[0] * 255) | (int(rgb[1] * 255) << 8) | (int(rgb[2] * 255) << 16) | (255 << 24)
        cols += [ic] * 4
    materials = [((1.0, 1.0, 1.0), opacity, (0.0, 0.0, 1.0, 1.0))]
    draw_calls = [(0, num_planes, 0, INST_HAS_VERTEX_COLORS)]
    camera = dict(kind="lookat", pos=(0.0, 0.0, -5.0), target=(0.005, 0.005, 0.0), up=(0.0, 1.0, 0.0))
    return _scene(np.array(pos, np.float32), np.array(quads, np.uint32), draw_calls, materials, camera,
                  width, height, colors=np.array(cols, np.uint32), name=name)


def boxes(dims=(10, 10, 10), box_size=0.5, box_dist=0.1, opacity=0.5, width=1280, height=720, seed=123,
          name="boxes"):
    """The reference's '#boxes' scene (BoxesSetup::updateScene, src/scene_setup.cpp:159-196: dims jittered colour
    cubes, scene opacity 0.5, OrbitingCamera({}, 10, 0.5, 0.8)).  Corners in fwk's Box::corners order (bit i of the
    corner number picks max on axis i, libfwk/include/fwk/math/box.h:163-172), the twelve triangles of addBox
    (scene_setup.cpp:134-145) paired into the six face quads Scene::generateQuads(4.0) makes of them, in its order and vertex
    rotation (checked against the restated quad generator in tests/test_oracle.py).  The jitter of +-0.1 comes from this module's counter-based
    generator instead of fwk's Random (std::mt19937_64 + libstdc++ distributions)."""
    dims = tuple(int(d) for d in dims)
    rng = Rng(seed)
    n = dims[0] * dims[1] * dims[2]
    step = np.float32(box_size + box_dist)
    grid = np.stack(np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij"), axis=3)
    grid = grid.reshape(-1, 3).astype(np.float32)  # x outermost, z innermost, as the reference's loops
    offset = -np.array(dims, np.float32) * step * np.float32(0.5)
    jitter = np.stack([rng.uniform(-0.1, 0.1, n) for _ in range(3)], axis=1)
    pos = offset[None, :] + grid * step + jitter
    bits = np.array([[(c >> a) & 1 for a in range(3)] for c in range(8)], np.float32)
    positions = (pos[:, None, :] + bits[None, :, :] * np.float32(box_size)).reshape(-1, 3).astype(np.float32)
    faces = np.array([[3, 1, 0, 2], [7, 5, 1, 3], [7, 3, 2, 6], [0, 4, 6, 2], [0, 1, 5, 4], [4, 5, 7, 6]], np.uint32)
    quads = (faces[None, :, :] + (np.arange(n, dtype=np.uint32) * np.uint32(8))[:, None, None]).reshape(-1, 4)
    col_scale = 1.0 / np.maximum(np.array(dims, np.float32) - 1.0, 1.0)
    rgb = (grid * col_scale[None, :] * np.float32(255.0)).astype(np.uint32)  # IColor(FColor): truncation
    col = rgb[:, 0] | (rgb[:, 1] << np.uint32(8)) | (rgb[:, 2] << np.uint32(16)) | np.uint32(0xFF000000)
    colors = np.repeat(col, 8)
    # scene_opacity < 1 clears DrawCallOpt::is_opaque (src/lucid_app.cpp:646-650)
    draw_calls = [(0, int(quads.shape[0]), 0, INST_HAS_VERTEX_COLORS)]
    materials = [((1.0, 1.0, 1.0), opacity, (0.0, 0.0, 1.0, 1.0))]
    camera = dict(kind="orbit", center=(0.0, 0.0, 0.0), distance=10.0, rot_h=0.5, rot_v=0.8)
    return _scene(positions, quads, draw_calls, materials, camera, width, height, colors=colors, name=name)


def box_triangles(num_boxes: int) -> np.ndarray:
    """The triangle list addBox writes for `num_boxes` boxes (src/scene_setup.cpp:136-143), int32[12 n, 3]."""
    tris = np.array([[0, 2, 3], [0, 3, 1], [1, 3, 7], [1, 7, 5], [2, 6, 7], [2, 7, 3],
                     [0, 6, 2], [0, 4, 6], [0, 5, 4], [0, 1, 5], [4, 7, 6], [4, 5, 7]], np.int32)
    return (tris[None, :, :] + (np.arange(num_boxes, dtype=np.int32) * 8)[:, None, None]).reshape(-1, 3)


def get_config(index: int, scale: float = 1.0):
    """BASELINE.json configs[index]; scale < 1 shrinks primitive counts for tests."""
    if index == 0:
        return quad_soup(num_quads=max(64, int(50_000 * scale)))
    if index == 1:
        return meshlet_patches(num_patches=max(2, int(489 * scale)))
    if index == 2:
        # BASELINE configs[2]: "depth complexity > 64 per pixel, raster_high path".  The ball fills the 4K frame
        # (94 % of the pixels covered) and the ribbons are wide enough for a median of 72 fragments per covered pixel
        # (p99 182, max 228; tests/depth_histogram.py) with no bin over the reference's HIGH limits
        # (raster_high.glsl:80-83,140-141).
        return hairball(num_strands=max(16, int(39_063 * scale)), ribbon_width=0.08, distance=14.0, start_frac=0.7)
    if index == 3:
        return architecture(num_small=max(2048, int(4_900_000 * scale)), num_large=max(8, int(2_000 * scale)))
    if index == 4:
        return meshlet_patches(num_patches=max(2, int(489 * scale)), name="config5_orbit")
    raise ValueError(index)
