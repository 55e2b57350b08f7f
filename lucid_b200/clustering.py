"""Spatially coherent instances ("real meshlet clustering for locality").

`uploadInstances` (src/lucid_renderer.cpp:352-429) cuts every draw call into instances of 1024 consecutive quads, so
what an instance covers on screen is decided by the order of the quads in the index buffer.  The reference's own
experiment in this direction (`meshPartition`, src/meshlet.cpp:68-222: greedy region growing over triangle adjacency
that keeps the partition's bounding box small, then pairwise merging by bounding-box surface area; `meshletTest` only
visualises the result, nothing feeds the renderer) is a sequential CPU pass.  Here the goal is stated directly --
every 1024-quad slice of a draw call should have a small bounding box -- and reached by ordering the quads of a draw
call along a Morton curve through their centroids: one key per quad, one sort, data-parallel, and a slice of a
space-filling curve is compact whatever the mesh connectivity is (soups included).

What it buys on this path: `k_instance_select` (LUCID_RENDER_CULL_INSTANCES) drops an instance whose box misses the
frustum or the rows a device owns in the bin split, and a CTA of `k_quad_cull` reads one instance -- both work per
instance, so compact instances mean fewer instances per device and neighbouring quads in neighbouring slots.

The order of quads inside a draw call never changes coverage, fragment counts or the exact blend except through the
depth-key tie convention (triangle index: DESIGN.md section 3) -- tests/test_clustering.py holds it to that.
"""
from __future__ import annotations

import numpy as np

MAX_INSTANCE_QUADS = 1024


def _spread3(v: np.ndarray) -> np.ndarray:
    """10-bit integers -> every bit followed by two zero bits (the classic Morton magic numbers)."""
    v = v.astype(np.uint32) & np.uint32(0x3FF)
    v = (v | (v << np.uint32(16))) & np.uint32(0x030000FF)
    v = (v | (v << np.uint32(8))) & np.uint32(0x0300F00F)
    v = (v | (v << np.uint32(4))) & np.uint32(0x030C30C3)
    v = (v | (v << np.uint32(2))) & np.uint32(0x09249249)
    return v


def morton_keys(points: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    """30-bit Morton codes of points inside the box [lo, hi] (10 bits per axis)."""
    extent = np.maximum(hi - lo, np.float32(1e-30))
    q = np.clip((points - lo) / extent * 1024.0, 0, 1023).astype(np.uint32)
    return _spread3(q[:, 0]) | (_spread3(q[:, 1]) << np.uint32(1)) | (_spread3(q[:, 2]) << np.uint32(2))


def cluster_order(positions: np.ndarray, quads: np.ndarray) -> np.ndarray:
    """Permutation that puts the quads in Morton order of their centroids (stable: equal keys keep their order):
    lucid_host_cluster_order of the C++ host library (include/lucid_host.h)."""
    from . import api

    positions = np.ascontiguousarray(positions, np.float32)
    quads = np.ascontiguousarray(quads, np.uint32)
    order = np.zeros(quads.shape[0], np.int32)
    rc = api.load_host_library().lucid_host_cluster_order(api._ptr(positions), positions.shape[0], api._ptr(quads),
                                                          quads.shape[0], api._ptr(order))
    if rc != 0:
        raise ValueError("cluster_order: a quad refers to a vertex that does not exist")
    return order.astype(np.int64)


def cluster_order_numpy(positions: np.ndarray, quads: np.ndarray) -> np.ndarray:
    """The same order stated in numpy (what tests hold the C++ function to)."""
    if quads.shape[0] == 0:
        return np.zeros(0, np.int64)
    centroids = positions[quads.reshape(-1)].reshape(-1, 4, 3).mean(axis=1, dtype=np.float64).astype(np.float32)
    keys = morton_keys(centroids, centroids.min(axis=0), centroids.max(axis=0))
    return np.argsort(keys, kind="stable")


def cluster_scene(scene: dict) -> dict:
    """A copy of the scene whose draw calls list their quads in Morton order (draw calls, materials and everything
    per vertex stay as they are; only rows of `quads` inside each draw call's range move)."""
    quads = np.array(scene["quads"], np.uint32, copy=True)
    positions = np.asarray(scene["positions"], np.float32)
    for (_, num_quads, quad_offset, _) in scene["draw_calls"]:
        part = quads[quad_offset:quad_offset + num_quads]
        quads[quad_offset:quad_offset + num_quads] = part[cluster_order(positions, part)]
    out = dict(scene)
    out["quads"] = np.ascontiguousarray(quads)
    out["name"] = scene.get("name", "") + "_clustered"
    return out


def instance_boxes(scene: dict) -> np.ndarray:
    """float32[n, 2, 3]: bounding boxes of the instances uploadInstances makes of the scene (1024-quad slices)."""
    positions = np.asarray(scene["positions"], np.float32)
    quads = np.asarray(scene["quads"], np.uint32)
    boxes = []
    for (_, num_quads, quad_offset, _) in scene["draw_calls"]:
        for off in range(quad_offset, quad_offset + num_quads, MAX_INSTANCE_QUADS):
            v = positions[quads[off:min(off + MAX_INSTANCE_QUADS, quad_offset + num_quads)].reshape(-1)]
            boxes.append((v.min(axis=0), v.max(axis=0)))
    return np.array(boxes, np.float32).reshape(-1, 2, 3)


def box_surface_area(boxes: np.ndarray) -> np.ndarray:
    d = boxes[:, 1] - boxes[:, 0]
    return 2.0 * (d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0])
