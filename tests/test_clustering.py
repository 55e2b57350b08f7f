"""Spatially coherent instances (lucid_b200/clustering.py): quads of a draw call in Morton order, so the
1024-quad slices uploadInstances cuts (src/lucid_renderer.cpp:352-429) are compact -- the goal of the reference's
meshPartition experiment (src/meshlet.cpp:68-222)."""
import numpy as np
import pytest

from lucid_b200 import api, clustering, scenes
from oracle.binding import Oracle
from tests import parity_util as pu


def _oracle(sc, bin_rows=None):
    cfg, inst, cols, rects = api.prepare_frame(sc)
    o = Oracle(sc["width"], sc["height"], 0, 1 << 20, threads=8)
    o.set_tie_report(True)
    if bin_rows:
        o.set_bin_rows(*bin_rows)
    o.set_scene(sc)
    o.render(cfg, inst, cols, rects)
    return o, inst


def test_morton_keys_and_order():
    pts = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    k = clustering.morton_keys(pts, pts.min(axis=0), pts.max(axis=0))
    assert k[0] == 0 and k[1] == 0x3FFFFFFF  # all 30 bits set at the far corner
    assert k[2] == 0x09249249 and k[3] == 0x12492492 and k[4] == 0x24924924  # x, y, z bits interleaved in that order
    # neighbours on a grid stay together: the first quarter of the curve is one octant pair
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), axis=3).reshape(-1, 3).astype(np.float32)
    order = np.argsort(clustering.morton_keys(g, g.min(axis=0), g.max(axis=0)), kind="stable")
    first = g[order[:512]]
    assert (first.max(axis=0) - first.min(axis=0)).tolist() == [7.0, 7.0, 7.0]


def test_clustered_scene_is_the_same_scene_in_another_order():
    sc = pu.small_scenes()["soup"]
    cl = clustering.cluster_scene(sc)
    assert cl["draw_calls"] == sc["draw_calls"] and cl["positions"] is sc["positions"]
    for (_, n, off, _) in sc["draw_calls"]:
        a = np.sort(sc["quads"][off:off + n].view([("q", np.uint32, 4)]).ravel(), order="q")
        b = np.sort(cl["quads"][off:off + n].view([("q", np.uint32, 4)]).ravel(), order="q")
        assert np.array_equal(a, b)
    assert not np.array_equal(sc["quads"], cl["quads"])
    # idempotent: a clustered scene is already in order
    assert np.array_equal(clustering.cluster_scene(cl)["quads"], cl["quads"])


def test_instances_become_compact_and_reach_fewer_bin_rows():
    # one draw call of randomly ordered quads: every 1024-quad slice spans the whole cube
    sc = scenes.quad_soup(num_quads=20_000, width=640, height=360)
    sc["draw_calls"] = [(0, 20_000, 0, 0)]
    sc["materials"] = [sc["materials"][0]]
    cl = clustering.cluster_scene(sc)
    area, area_cl = clustering.box_surface_area(clustering.instance_boxes(sc)), clustering.box_surface_area(clustering.instance_boxes(cl))
    assert area.shape == area_cl.shape == (20,)
    assert area_cl.sum() < 0.4 * area.sum()
    # the frame is the same frame: fragment counts per pixel, statistics, visible quads; colour except where the
    # depth-key tie convention (triangle index) speaks -- the marked pixels of either frame
    o, _ = _oracle(sc)
    oc, _ = _oracle(cl)
    assert np.array_equal(o.read_frag_counts(), oc.read_frag_counts())
    assert np.array_equal(o.info[1:3], oc.info[1:3]) and np.array_equal(o.info[60:62], oc.info[60:62])
    marked = (o.read_tie_pixels()[0] != 0) | (oc.read_tie_pixels()[0] != 0)
    assert not ((o.read_image() != oc.read_image()) & ~marked).any()
    # bin-row split over 4 devices: instances with a visible quad in a device's rows
    rows = (sc["height"] + 31) // 32
    edges = np.linspace(0, rows, 5).round().astype(int)

    def reached(scene):
        total = 0
        for r in range(4):
            orc, inst = _oracle(scene, bin_rows=(int(edges[r]), int(edges[r + 1])))
            first = np.concatenate([[0], np.cumsum(inst[:, 2])])
            ids = np.concatenate([orc.read_quad_input_ids(0), orc.read_quad_input_ids(1)])
            total += np.unique(np.searchsorted(first, ids, side="right") - 1).size
        return total

    plain, clustered = reached(sc), reached(cl)
    assert plain == 4 * 20  # in random order every instance reaches every device
    assert clustered <= 0.65 * plain


@pytest.mark.gpu
def test_cuda_renders_the_clustered_scene_like_the_checker():
    sc = clustering.cluster_scene(pu.small_scenes()["soup_close"])
    o = pu.run_oracle(sc)
    r, img = pu.run_cuda(sc)
    try:
        assert {k: v for k, v in pu.compare(r, img, o).items() if not k.startswith("_")} == {}
    finally:
        r.close()


def test_load_scene_can_cluster_while_loading():
    import io

    from lucid_b200 import scene_io

    sc = scenes.quad_soup(num_quads=5_000, width=320, height=200)
    sc["draw_calls"], sc["materials"] = [(0, 5_000, 0, 0)], [sc["materials"][0]]  # one mesh: five instances
    f = io.BytesIO()
    scene_io.save_scene(f, sc)
    f.seek(0)
    plain = scene_io.load_scene(f, sc["width"], sc["height"])
    f.seek(0)
    clustered = scene_io.load_scene(f, sc["width"], sc["height"], cluster=True)
    assert np.array_equal(clustered["quads"], clustering.cluster_scene(plain)["quads"])
    assert clustering.box_surface_area(clustering.instance_boxes(clustered)).sum() < \
        clustering.box_surface_area(clustering.instance_boxes(plain)).sum()


def test_native_order_equals_the_numpy_statement():
    """lucid_host_cluster_order (C++, include/lucid_host.h) against clustering.cluster_order_numpy, element for element."""
    small = pu.small_scenes()
    for name in ("soup", "meshlets", "hairball", "arch", "boxes"):
        sc = small[name]
        assert np.array_equal(clustering.cluster_order(sc["positions"], sc["quads"]),
                              clustering.cluster_order_numpy(sc["positions"], sc["quads"])), name
    # degenerate inputs: no quads; all centroids equal (input order is kept); an index outside the vertex array
    pos = np.zeros((4, 3), np.float32)
    assert clustering.cluster_order(pos, np.zeros((0, 4), np.uint32)).size == 0
    same = np.tile(np.arange(4, dtype=np.uint32), (5, 1))
    assert clustering.cluster_order(pos, same).tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        clustering.cluster_order(pos, np.array([[0, 1, 2, 4]], np.uint32))
