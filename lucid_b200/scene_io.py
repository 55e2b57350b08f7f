"""Reader / writer of the reference's `.scene` files and the `Scene` -> draw-call step (the caller side
of the hot path).

  Scene::load / save          src/scene.cpp:119-179 (signature "SCENE", three counts, the vertex arrays, meshes,
                              materials, textures)
  SceneMesh / SceneMaterial / SceneTexture load & save        src/scene.cpp:17-99
  fwk::BaseStream             libfwk/src/io/stream.cpp:120-176 (variable-length sizes: one byte below 248, else
                              248 + index of the highest non-zero byte followed by that many + 1 little-endian bytes;
                              vectors = size + raw elements; pack() = the raw bytes of its arguments back to back)
  Scene::draws                src/scene.cpp:483-512 (one SceneDrawCall per mesh, options from the scene's vertex
                              arrays and the mesh's material)
  Scene::updatePrimitiveOffsets: a mesh's quads follow those of the meshes before it in the quad index buffer

`load_scene(path)` returns the dict every other part of this package consumes (lucid_b200.scenes): positions, quads,
optional colours / uvs / 10-10-10 normals, draw calls, materials, the opaque and the transparent albedo atlas, and a
default orbit camera around the bounding box -- so a hairball or san-miguel converted by the reference's tools can be
rendered as is.  `save_scene` writes the same format from such a dict: the synthetic BASELINE scenes can be loaded
by the reference application for a cross-check on a machine that can run it.
"""
from __future__ import annotations

import io
import struct

import numpy as np

from . import scenes

RGBA8_UNORM = 20  # index in fwk's VColorFormat (libfwk/include/fwk/vulkan_base.h:173-186)
MAP_TYPES = ("albedo", "normal", "pbr")  # DEFINE_ENUM(SceneMapType, ...), src/scene.h:13


class SceneFormatError(ValueError):
    pass


# ---- fwk::BaseStream primitives ------------------------------------------------------------------------------


def _write_size(f, n: int):
    """BaseStream::saveSize, stream.cpp:120-141"""
    if n < 0:
        raise ValueError("negative size")
    if n < 248:
        f.write(bytes([n]))
        return
    raw = struct.pack("<Q", n)
    max_byte = max(j for j in range(8) if raw[j])
    f.write(bytes([248 + max_byte]) + raw[:max_byte + 1])


def _read_size(f) -> int:
    """BaseStream::loadSize, stream.cpp:143-160"""
    small = _read(f, 1)[0]
    if small < 248:
        return small
    raw = _read(f, small - 247)
    return int.from_bytes(raw, "little")


def _read(f, n: int) -> bytes:
    data = f.read(n)
    if len(data) != n:
        raise SceneFormatError(f"unexpected end of stream ({len(data)} of {n} bytes)")
    return data


def _write_vector(f, a: np.ndarray | None, dtype, width: int):
    if a is None:
        _write_size(f, 0)
        return
    a = np.ascontiguousarray(a, dtype).reshape(-1, width) if width > 1 else np.ascontiguousarray(a, dtype).reshape(-1)
    _write_size(f, a.shape[0])
    f.write(a.tobytes())


def _read_vector(f, dtype, width: int) -> np.ndarray:
    n = _read_size(f)
    item = np.dtype(dtype).itemsize * width
    a = np.frombuffer(_read(f, n * item), dtype)
    return a.reshape(n, width).copy() if width > 1 else a.copy()


def _write_string(f, s: str):
    raw = s.encode()
    _write_size(f, len(raw))
    f.write(raw)


def _read_string(f) -> str:
    return _read(f, _read_size(f)).decode(errors="replace")


# ---- Scene::load ------------------------------------------------------------------------------------------------


def load_scene(path_or_file, width: int = 1920, height: int = 1080, name: str | None = None, quads: str = "file",
               square_weight: float = 4.0, device: int = 0, cluster: bool = False) -> dict:
    """Reads a `.scene` file (Scene::load, src/scene.cpp:119-151) into a scene dict.

    quads="file" takes every mesh's quads as stored; quads="gpu" pairs the mesh's triangles again on the GPU
    (Scene::generateQuads, src/scene.cpp:237-247, through lucid_quadgen: needs the CUDA library) -- for files that carry
    triangles only, or to re-pair with another squareness weight (the reference's converter uses the input scene's
    quad_squareness, its procedural scenes 4.0: src/scene_convert.cpp:430, scene_setup.cpp:186).

    cluster=True lists every mesh's quads in Morton order, so that the 1024-quad instances uploadInstances cuts are
    spatially compact (lucid_b200.clustering; the goal of the reference's meshPartition, src/meshlet.cpp:68-222)."""
    if quads not in ("file", "gpu"):
        raise ValueError('quads must be "file" or "gpu"')
    f = open(path_or_file, "rb") if isinstance(path_or_file, str) else path_or_file
    try:
        if _read(f, 5) != b"SCENE":
            raise SceneFormatError('expected signature "SCENE"')
        num_meshes, num_materials, num_textures = struct.unpack("<iii", _read(f, 12))
        if num_meshes <= 0 or num_materials <= 0 or num_textures < 0:
            raise SceneFormatError("a scene needs at least one mesh and one material")
        positions = _read_vector(f, np.float32, 3)
        colors = _read_vector(f, np.uint32, 1)  # IColor: r | g << 8 | b << 16 | a << 24
        tex_coords = _read_vector(f, np.float32, 2)
        _read_vector(f, np.float32, 3)  # normals
        _read_vector(f, np.float32, 3)  # tangents
        qnormals = _read_vector(f, np.uint32, 1)
        _read_vector(f, np.uint32, 1)  # quantized tangents
        bbox = np.frombuffer(_read(f, 24), np.float32).reshape(2, 3).copy()

        meshes = []
        for _ in range(num_meshes):
            material_id, colors_opaque = struct.unpack("<i?", _read(f, 5))
            tris = _read_vector(f, np.int32, 3)  # the renderer itself only reads the quads
            mesh_quads = _read_vector(f, np.int32, 4)
            (num_degenerate,) = struct.unpack("<i", _read(f, 4))
            _read(f, 24)  # mesh bounding box
            if not 0 <= material_id < num_materials:
                raise SceneFormatError("mesh refers to a material that does not exist")
            meshes.append(dict(material_id=material_id, colors_opaque=colors_opaque, quads=mesh_quads, tris=tris,
                               num_degenerate_quads=num_degenerate))

        materials = []
        for _ in range(num_materials):
            mname = _read_string(f)
            diffuse = struct.unpack("<fff", _read(f, 12))
            (opacity,) = struct.unpack("<f", _read(f, 4))
            maps = {}
            for mt in MAP_TYPES:
                texture_id, is_opaque, is_clamped, x0, y0, x1, y1 = struct.unpack("<i??ffff", _read(f, 22))
                if texture_id != -1 and not 0 <= texture_id < num_textures:
                    raise SceneFormatError("material refers to a texture that does not exist")
                maps[mt] = dict(texture_id=texture_id, is_opaque=is_opaque, is_clamped=is_clamped, uv_rect=(x0, y0, x1, y1))
            materials.append(dict(name=mname, diffuse=diffuse, opacity=opacity, maps=maps))

        textures = []
        for _ in range(num_textures):
            tname = _read_string(f)
            map_type, is_opaque, is_clamped, is_atlas, num_levels, fmt = struct.unpack("<B???BB", _read(f, 6))
            mips = []
            for _ in range(num_levels):
                w, h, byte_size = struct.unpack("<iii", _read(f, 12))
                if w <= 0 or h <= 0 or byte_size <= 0:
                    raise SceneFormatError("bad mip level")
                mips.append((w, h, np.frombuffer(_read(f, byte_size), np.uint8).copy()))
            textures.append(dict(name=tname, map_type=map_type, is_opaque=is_opaque, is_clamped=is_clamped,
                                 is_atlas=is_atlas, format=fmt, mips=mips))
    finally:
        if isinstance(path_or_file, str):
            f.close()
    if quads == "gpu":
        from . import quadgen
        quadgen.generate_quads(meshes, positions, square_weight, device)
    scene = scene_from_parts(positions, colors, tex_coords, qnormals, bbox, meshes, materials, textures, width, height,
                             name or (path_or_file if isinstance(path_or_file, str) else "scene"))
    if cluster:
        from . import clustering
        scene = clustering.cluster_scene(scene)
    return scene


def scene_from_parts(positions, colors, tex_coords, qnormals, bbox, meshes, materials, textures, width, height, name):
    """Scene::draws (src/scene.cpp:483-512) + updatePrimitiveOffsets + textureAtlasPair (:514-526)."""
    nv = positions.shape[0]
    has_colors = colors.size == nv and nv > 0
    has_uvs = tex_coords.shape[0] == nv and nv > 0
    has_normals = qnormals.size == nv and nv > 0
    scene_opts = (scenes.INST_HAS_VERTEX_COLORS if has_colors else 0) | (scenes.INST_HAS_VERTEX_TEX_COORDS if has_uvs else 0) | \
                 (scenes.INST_HAS_VERTEX_NORMALS if has_normals else 0)
    draw_calls, quad_arrays, offset = [], [], 0
    for mesh in meshes:
        mat = materials[mesh["material_id"]]
        albedo = mat["maps"]["albedo"]
        # SceneMaterial::isOpaque (:45-48) and Map::usesUvRect (src/scene.h:44)
        is_opaque = mat["opacity"] == 1.0 and (albedo["texture_id"] == -1 or albedo["is_opaque"]) and mesh["colors_opaque"]
        uses_rect = (not albedo["is_clamped"]) and tuple(albedo["uv_rect"]) != (0.0, 0.0, 1.0, 1.0)
        opts = scene_opts | (scenes.INST_IS_OPAQUE if is_opaque else 0) | (scenes.INST_TEX_OPAQUE if albedo["is_opaque"] else 0) | \
            (scenes.INST_HAS_UV_RECT if uses_rect else 0)
        if has_uvs:
            opts |= scenes.INST_HAS_ALBEDO_TEXTURE if albedo["texture_id"] != -1 else 0
            opts |= 0x080 if mat["maps"]["normal"]["texture_id"] != -1 else 0  # has_normal_tex
            opts |= 0x100 if mat["maps"]["pbr"]["texture_id"] != -1 else 0  # has_pbr_tex
        nq = mesh["quads"].shape[0]
        draw_calls.append((mesh["material_id"], nq, offset, opts))
        quad_arrays.append(mesh["quads"].astype(np.uint32))
        offset += nq
    out_materials = []
    for mat in materials:
        x0, y0, x1, y1 = mat["maps"]["albedo"]["uv_rect"]
        out_materials.append((tuple(float(c) for c in mat["diffuse"]), float(mat["opacity"]), (x0, y0, x1 - x0, y1 - y0)))
    # the first opaque and the first transparent albedo texture are the two atlases the renderer binds
    tex = {}
    for t in textures:
        if t["map_type"] != 0 or t["format"] != RGBA8_UNORM or not t["mips"]:
            continue
        key = "opaque" if t["is_opaque"] else "transparent"
        if key not in tex:
            w, h, _ = t["mips"][0]
            tex[key] = (w, h, len(t["mips"]), np.concatenate([m[2] for m in t["mips"]]))
    center = tuple(float(c) for c in (bbox[0] + bbox[1]) * 0.5)
    extent = float(np.linalg.norm(bbox[1] - bbox[0]))
    camera = dict(kind="orbit", center=center, distance=max(extent * 0.75, 1e-3), rot_h=0.5, rot_v=0.6)
    sc = scenes._scene(positions, np.concatenate(quad_arrays) if quad_arrays else np.zeros((0, 4), np.uint32), draw_calls,
                       out_materials, camera, width, height, colors=colors if has_colors else None,
                       uvs=tex_coords if has_uvs else None, normals=qnormals if has_normals else None, textures=tex, name=name)
    sc["bounding_box"] = bbox
    return sc


# ---- Scene::save ---------------------------------------------------------------------------------------------------


def save_scene(path_or_file, scene: dict, with_tris: bool = False, with_quads: bool = True):
    """Writes a scene dict in the reference's format (Scene::save, src/scene.cpp:153-179): one mesh per draw call.
    with_tris also writes every mesh's triangles (the two of each quad, degenerate ones left out) as the reference's
    files carry them; with_quads=False leaves the quads out: a triangle-only file for load_scene(quads="gpu")."""
    f = open(path_or_file, "wb") if isinstance(path_or_file, str) else path_or_file
    try:
        pos = np.ascontiguousarray(scene["positions"], np.float32)
        tex_names = [k for k in ("opaque", "transparent") if k in scene.get("textures", {})]
        f.write(b"SCENE")
        f.write(struct.pack("<iii", len(scene["draw_calls"]), len(scene["materials"]), len(tex_names)))
        _write_vector(f, pos, np.float32, 3)
        _write_vector(f, scene.get("colors"), np.uint32, 1)
        _write_vector(f, scene.get("uvs"), np.float32, 2)
        _write_vector(f, None, np.float32, 3)  # float normals / tangents are not kept by the quad path
        _write_vector(f, None, np.float32, 3)
        _write_vector(f, scene.get("normals"), np.uint32, 1)
        _write_vector(f, None, np.uint32, 1)
        bbox = np.stack([pos.min(axis=0), pos.max(axis=0)]).astype(np.float32) if pos.size else np.zeros((2, 3), np.float32)
        f.write(bbox.tobytes())
        quads = np.ascontiguousarray(scene["quads"], np.int32)
        for mat_id, nq, off, opts in scene["draw_calls"]:
            q = quads[off:off + nq]
            f.write(struct.pack("<i?", mat_id, True))
            tris = None
            if with_tris and q.size:
                second = q[q[:, 2] != q[:, 3]]  # (a, b, c, c) stands for one triangle
                tris = np.concatenate([q[:, [0, 1, 2]], second[:, [0, 2, 3]]])
            _write_vector(f, tris, np.int32, 3)
            _write_vector(f, q if with_quads else None, np.int32, 4)
            f.write(struct.pack("<i", int((q[:, 2] == q[:, 3]).sum()) if (with_quads and q.size) else 0))
            v = pos[q.reshape(-1)] if q.size else np.zeros((1, 3), np.float32)
            f.write(np.stack([v.min(axis=0), v.max(axis=0)]).astype(np.float32).tobytes())
        # which texture a material uses follows from the draw calls that use the material
        mat_opts = {}
        for mat_id, _, _, opts in scene["draw_calls"]:
            mat_opts[mat_id] = mat_opts.get(mat_id, 0) | opts
        for i, (diffuse, opacity, rect) in enumerate(scene["materials"]):
            _write_string(f, f"material{i}")
            f.write(struct.pack("<ffff", *diffuse, opacity))
            opts = mat_opts.get(i, 0)
            textured = bool(opts & scenes.INST_HAS_ALBEDO_TEXTURE)
            tex_opaque = bool(opts & scenes.INST_TEX_OPAQUE)
            tex_id = tex_names.index("opaque" if tex_opaque else "transparent") if textured else -1
            x0, y0, sx, sy = rect
            f.write(struct.pack("<i??ffff", tex_id, tex_opaque, not (opts & scenes.INST_HAS_UV_RECT), x0, y0, x0 + sx, y0 + sy))
            for _ in MAP_TYPES[1:]:
                f.write(struct.pack("<i??ffff", -1, False, True, 0.0, 0.0, 1.0, 1.0))
        for key in tex_names:
            w, h, levels, data = scene["textures"][key]
            data = np.ascontiguousarray(data, np.uint8)
            _write_string(f, key)
            f.write(struct.pack("<B???BB", 0, key == "opaque", False, True, levels, RGBA8_UNORM))
            off = 0
            for l in range(levels):
                lw, lh = max(1, w >> l), max(1, h >> l)
                f.write(struct.pack("<iii", lw, lh, lw * lh * 4))
                f.write(data[off:off + lw * lh * 4].tobytes())
                off += lw * lh * 4
    finally:
        if isinstance(path_or_file, str):
            f.close()


def roundtrip(scene: dict) -> dict:
    buf = io.BytesIO()
    save_scene(buf, scene)
    buf.seek(0)
    out = load_scene(buf, scene["width"], scene["height"], scene.get("name", "scene"))
    out["camera"], out["background"] = scene["camera"], scene["background"]
    return out
