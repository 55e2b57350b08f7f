"""`.scene` reader / writer (lucid_b200/scene_io.py) against the reference's format (src/scene.cpp:17-179,
libfwk/src/io/stream.cpp:120-176) and the Scene -> draw call step (src/scene.cpp:483-512)."""
import io
import json
import os
import struct

import numpy as np
import pytest

from lucid_b200 import api, scene_io, scenes
from tests import parity_util as pu
from tests.golden.make_oracle_golden import digest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_size_encoding_is_fwk_s():
    """BaseStream::saveSize / loadSize: one byte below 248, else 248 + index of the highest non-zero byte and that
    many + 1 little-endian bytes."""
    cases = {0: b"\x00", 5: b"\x05", 247: b"\xf7", 248: b"\xf8\xf8", 255: b"\xf8\xff", 256: b"\xf9\x00\x01",
             65535: b"\xf9\xff\xff", 65536: b"\xfa\x00\x00\x01", 4902000: b"\xfa" + struct.pack("<I", 4902000)[:3],
             (1 << 32) + 7: b"\xfc\x07\x00\x00\x00\x01"}
    for n, want in cases.items():
        f = io.BytesIO()
        scene_io._write_size(f, n)
        assert f.getvalue() == want, (n, f.getvalue())
        f.seek(0)
        assert scene_io._read_size(f) == n


def test_layout_of_a_tiny_scene():
    """Byte for byte: signature, counts, vectors, one mesh, one material with three maps."""
    sc = scenes.planes(num_planes=1, width=64, height=64)
    f = io.BytesIO()
    scene_io.save_scene(f, sc)
    raw = f.getvalue()
    assert raw[:5] == b"SCENE" and struct.unpack("<iii", raw[5:17]) == (1, 1, 0)
    assert raw[17] == 4  # four positions
    assert np.array_equal(np.frombuffer(raw[18:18 + 48], np.float32).reshape(4, 3), sc["positions"])
    assert raw[66] == 4 and raw[67 + 16] == 0  # four vertex colours, no tex coords
    # material record: name, diffuse, opacity, then 3 x (int, bool, bool, 4 floats) = 66 bytes
    tail = raw[-(1 + 9 + 16 + 66):]
    assert tail[0] == 9 and tail[1:10] == b"material0"
    assert struct.unpack("<ffff", tail[10:26]) == (1.0, 1.0, 1.0, 0.25)
    assert struct.unpack("<i??ffff", tail[26:48]) == (-1, False, True, 0.0, 0.0, 1.0, 1.0)


@pytest.mark.parametrize("name", ["soup", "meshlets", "arch", "hairball", "boxes"])
def test_round_trip_gives_the_same_frame(name):
    """save -> load reproduces geometry, draw calls, materials and atlases, and the checker renders the reloaded
    scene to the committed digests of the original."""
    sc = pu.small_scenes()[name]
    back = scene_io.roundtrip(sc)
    for key in ("positions", "quads", "colors", "uvs", "normals"):
        a, b = sc.get(key), back.get(key)
        assert (a is None) == (b is None), key
        if a is not None:
            assert np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)), key
    assert [tuple(d) for d in sc["draw_calls"]] == [tuple(d) for d in back["draw_calls"]]
    for (d0, o0, r0), (d1, o1, r1) in zip(sc["materials"], back["materials"]):
        assert np.allclose(d0, d1, rtol=0, atol=1e-7) and abs(o0 - o1) < 1e-7 and np.allclose(r0, r1, rtol=0, atol=1e-7)
    for key in sc["textures"]:
        assert sc["textures"][key][:3] == back["textures"][key][:3]
        assert np.array_equal(sc["textures"][key][3], back["textures"][key][3])
    o = pu.run_oracle(back, threads=4)
    with open(os.path.join(HERE, "golden", "oracle_golden.json")) as f:
        g = json.load(f)[name]
    st = api.decode_stats(o.info, o.bin_count, o.width, o.height)
    assert {k: st[k] for k in g["stats"]} == g["stats"]
    assert digest(o.read_image()) == g["image"] and digest(o.read_frag_counts()) == g["frag_counts"]


def test_truncated_and_foreign_files_are_rejected():
    sc = scenes.planes(num_planes=2, width=64, height=64)
    f = io.BytesIO()
    scene_io.save_scene(f, sc)
    raw = f.getvalue()
    with pytest.raises(scene_io.SceneFormatError):
        scene_io.load_scene(io.BytesIO(b"MODEL" + raw[5:]))
    with pytest.raises(scene_io.SceneFormatError):
        scene_io.load_scene(io.BytesIO(raw[:len(raw) // 2]))


def test_triangles_are_written_and_read_back():
    """with_tris: the two triangles of every quad, in the reference's mesh record; the loader keeps them."""
    sc = pu.small_scenes()["meshlets"]
    f = io.BytesIO()
    scene_io.save_scene(f, sc, with_tris=True)
    f.seek(0)
    out = scene_io.load_scene(f, sc["width"], sc["height"])
    assert np.array_equal(out["quads"], sc["quads"])  # quads="file": as stored
    f.seek(0)
    # triangle-only file: no quads to take
    g = io.BytesIO()
    scene_io.save_scene(g, sc, with_tris=True, with_quads=False)
    g.seek(0)
    only = scene_io.load_scene(g, sc["width"], sc["height"])
    assert only["quads"].shape[0] == 0
    with pytest.raises(ValueError):
        scene_io.load_scene(io.BytesIO(g.getvalue()), quads="cpu")


@pytest.mark.gpu
def test_triangle_only_scene_is_paired_on_the_gpu_and_renders_the_same():
    """f2 + f3 together: a scene file that carries triangles only is paired by lucid_quadgen when it is loaded
    (Scene::generateQuads) and renders the fragments of the scene it was written from."""
    sc = pu.small_scenes()["meshlets"]
    g = io.BytesIO()
    scene_io.save_scene(g, sc, with_tris=True, with_quads=False)
    g.seek(0)
    paired = scene_io.load_scene(g, sc["width"], sc["height"], quads="gpu")
    paired["camera"], paired["background"] = sc["camera"], sc["background"]
    nq0, nq1 = sc["quads"].shape[0], paired["quads"].shape[0]
    assert nq0 <= nq1 <= 1.05 * nq0  # a height-field patch pairs back into (almost) as many quads as it was made of
    r0, img0 = pu.run_cuda(sc)
    r1, img1 = pu.run_cuda(paired)
    try:
        a, b = r0.read_frag_counts(), r1.read_frag_counts()
        assert abs(int(a.sum()) - int(b.sum())) <= 2e-3 * a.sum()
        assert np.count_nonzero(a != b) <= 5e-3 * np.count_nonzero(a)
    finally:
        r0.close()
        r1.close()
