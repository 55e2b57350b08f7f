"""Seeded triangle meshes for the quad-pairing tests (test infrastructure)."""
from __future__ import annotations

import numpy as np


def grid(nx=40, ny=30, jitter=0.25, seed=1, random_diagonals=True, shuffle=False):
    """Height-field patch: (nx+1)(ny+1) vertices, 2 nx ny triangles, consistent winding."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx + 1, dtype=np.float32), np.arange(ny + 1, dtype=np.float32), indexing="xy")
    pos = np.stack([xs, ys, np.zeros_like(xs)], -1).reshape(-1, 3)
    pos += (rng.random(pos.shape, dtype=np.float32) - 0.5) * np.float32(jitter)
    idx = lambda x, y: y * (nx + 1) + x
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    cx, cy = cx.ravel(), cy.ravel()
    a, b, c, d = idx(cx, cy), idx(cx + 1, cy), idx(cx + 1, cy + 1), idx(cx, cy + 1)
    flip = rng.random(len(a)) < 0.5 if random_diagonals else np.zeros(len(a), bool)
    t0 = np.where(flip[:, None], np.stack([a, b, d], -1), np.stack([a, b, c], -1))
    t1 = np.where(flip[:, None], np.stack([b, c, d], -1), np.stack([a, c, d], -1))
    tris = np.stack([t0, t1], 1).reshape(-1, 3).astype(np.int32)
    if shuffle:
        tris = tris[rng.permutation(len(tris))]
        tris = np.stack([np.roll(t, int(r)) for t, r in zip(tris, rng.integers(0, 3, len(tris)))]).astype(np.int32)
    return pos, tris


def sphere(rings=24, sectors=48, seed=2):
    """Closed UV sphere with pole fans."""
    rng = np.random.default_rng(seed)
    pos = [[0.0, 0.0, 1.0]]
    for r in range(1, rings):
        th = np.pi * r / rings
        for s in range(sectors):
            ph = 2 * np.pi * s / sectors
            pos.append([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
    pos.append([0.0, 0.0, -1.0])
    pos = np.asarray(pos, np.float32) * np.float32(3.0)
    pos += (rng.random(pos.shape, dtype=np.float32) - 0.5) * np.float32(0.01)
    v = lambda r, s: 1 + (r - 1) * sectors + s % sectors
    tris = []
    for s in range(sectors):
        tris.append([0, v(1, s), v(1, s + 1)])
    for r in range(1, rings - 1):
        for s in range(sectors):
            tris.append([v(r, s), v(r + 1, s), v(r + 1, s + 1)])
            tris.append([v(r, s), v(r + 1, s + 1), v(r, s + 1)])
    last = len(pos) - 1
    for s in range(sectors):
        tris.append([last, v(rings - 1, s + 1), v(rings - 1, s)])
    return pos, np.asarray(tris, np.int32)


def delaunay(n=1500, seed=3):
    """Irregular planar triangulation (vertex degrees 3..12), wound counter-clockwise."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    p2 = rng.random((n, 2))
    tri = Delaunay(p2).simplices.astype(np.int32)
    a, b, c = p2[tri[:, 0]], p2[tri[:, 1]], p2[tri[:, 2]]
    area = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    tri[area < 0] = tri[area < 0][:, ::-1]
    pos = np.concatenate([p2, 0.1 * rng.random((n, 1))], 1).astype(np.float32) * np.float32(10)
    return pos, np.ascontiguousarray(tri)


def soup(n=500, seed=4):
    """Triangles that share no vertices: no node at all, every output quad is degenerate."""
    rng = np.random.default_rng(seed)
    pos = rng.random((3 * n, 3), dtype=np.float32)
    return pos, np.arange(3 * n, dtype=np.int32).reshape(-1, 3)


def non_manifold(seed=5):
    """What real assets contain: duplicated triangles, an edge shared by three faces, a pair joined along two edges
    (the same three vertices in reverse), inconsistent winding (no reversed edge to pair along), a triangle listing a
    vertex twice."""
    rng = np.random.default_rng(seed)
    pos, tris = grid(12, 9, seed=seed)
    tris = tris.tolist()
    tris += [tris[10], tris[11], tris[40]]  # duplicates
    nv = len(pos)
    extra = rng.random((6, 3), dtype=np.float32) * np.float32(4.0) + np.float32(2.0)
    pos = np.concatenate([pos, extra])
    e0, e1 = tris[20][0], tris[20][1]
    tris += [[e1, e0, nv], [e1, e0, nv + 1], [e0, e1, nv + 2]]  # fan around one edge
    tris += [[nv + 3, nv + 4, nv + 5], [nv + 5, nv + 4, nv + 3]]  # two faces of one triangle
    tris += [tris[60][::-1]]  # flipped copy
    tris += [[nv, nv, nv + 1]]  # repeated vertex (zero-area; squareness stays finite: lengths are not zero pairwise)
    order = rng.permutation(len(tris))
    return pos, np.asarray(tris, np.int32)[order]


def strip(n=64, seed=6):
    """One long triangle strip: the pairing graph is a path, the worst case for a round-synchronous selection."""
    rng = np.random.default_rng(seed)
    pos = np.stack([np.arange(n + 2) // 2, np.arange(n + 2) % 2, np.zeros(n + 2)], -1).astype(np.float32)
    pos += (rng.random(pos.shape, dtype=np.float32) - 0.5) * np.float32(0.2)
    tris = [[i, i + 1, i + 2] if i % 2 == 0 else [i + 1, i, i + 2] for i in range(n)]
    return pos, np.asarray(tris, np.int32)


CASES = {
    "grid": lambda: grid(),
    "grid_regular": lambda: grid(32, 32, jitter=0.0, random_diagonals=False),
    "grid_shuffled": lambda: grid(25, 20, shuffle=True, seed=7),
    "sphere": lambda: sphere(),
    "delaunay": lambda: delaunay(),
    "soup": lambda: soup(),
    "non_manifold": lambda: non_manifold(),
    "strip": lambda: strip(),
    "single": lambda: (np.eye(3, dtype=np.float32), np.array([[0, 1, 2]], np.int32)),
    "pair": lambda: (np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.int32)),
}
