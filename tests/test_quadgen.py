"""Triangle -> quad pairing: reference's src/quad_generator.cpp vs the CPU restatement vs the CUDA path.

CPU tests: the restatement (oracle/quadgen_oracle.cpp, mode 0) equals the reference's own code compiled from
/root/reference (oracle/_ref/libref_quadgen.so) output for output where that library exists, and the digests committed
in tests/golden/ref_quadgen.json (made from it by tests/golden/make_ref_quadgen.py) everywhere; the per-element rules the
CUDA kernels execute (lucid_b200/csrc/quadgen_rules.h) give the same graph when run over a mesh on the CPU; the
round-synchronous selection (mode 1) is a valid pairing with a quad count close to the reference's.
GPU tests: lucid_quadgen() through the C ABI equals the restatement -- graph word for word, quads equal to mode 1.
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import pytest

from oracle import quadgen_binding as qb
from tests import quadgen_meshes as qm

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_quadgen.json")
GRAPH_KEYS = ("neighbours", "tri_quads", "node_tris", "node_verts", "node_conflicts", "squareness")
# quads of the round-synchronous selection / quads of the reference, upper bound per mesh (measured: <= 1.010)
COUNT_BOUND = 1.015


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]


def digests(out: dict) -> dict:
    d = {k: digest(out[k]) for k in GRAPH_KEYS + ("quads",)}
    d["num_quads"], d["num_degenerate"], d["num_nodes"] = len(out["quads"]), out["num_degenerate"], len(out["node_tris"])
    return d


def same(a: np.ndarray, b: np.ndarray) -> bool:
    if a.dtype == np.float32:
        a, b = a.view(np.uint32), b.view(np.uint32)
    return a.shape == b.shape and np.array_equal(a, b)


def check_pairing(tris: np.ndarray, out: dict):
    """Every triangle in exactly one output quad; each quad a node of the graph or a degenerate (a, b, c, c); output in
    the order of the first triangle; no node left whose two triangles are both unpaired (maximal)."""
    nt = len(tris)
    sel = out["selected"] == 2
    nodes = out["node_tris"][sel]
    used = np.zeros(nt, np.int32)
    np.add.at(used, nodes.ravel(), 1)
    assert used.max(initial=0) <= 1, "a triangle is in two quads"
    assert len(out["quads"]) == nt - len(nodes)
    assert out["num_degenerate"] == nt - 2 * len(nodes)
    first = np.ones(nt, bool)
    first[nodes[:, 1]] = False
    mate_node = np.full(nt, -1)
    ids = np.flatnonzero(sel)
    mate_node[nodes[:, 0]] = ids
    mate_node[nodes[:, 1]] = ids
    expect = []
    for t in np.flatnonzero(first):
        q = mate_node[t]
        expect.append(out["node_verts"][q] if q >= 0 else [tris[t][0], tris[t][1], tris[t][2], tris[t][2]])
    assert np.array_equal(np.asarray(expect, np.int32).reshape(-1, 4), out["quads"])
    conf = out["node_conflicts"]
    for q in np.flatnonzero(~sel):
        if q in conf[q]:
            continue
        a, b = out["node_tris"][q]
        assert used[a] or used[b], f"node {q} could still be paired"


@pytest.mark.parametrize("name", sorted(qm.CASES))
def test_restatement_equals_the_reference(name):
    pos, tris = qm.CASES[name]()
    o = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=0)
    with open(GOLDEN) as f:
        golden = json.load(f)
    assert digests(o) == golden[name], "restatement differs from the committed outputs of the reference's code"
    if qb.reference_available():
        r = qb.run(qb.load_reference(), pos, tris, 4.0)
        for k in GRAPH_KEYS + ("quads",):
            assert same(o[k], r[k]), k
        assert o["num_degenerate"] == r["num_degenerate"]


@pytest.mark.parametrize("weight", [0.0, 1.0, 16.0])
def test_restatement_equals_the_reference_for_other_weights(weight):
    if not qb.reference_available():
        pytest.skip("oracle/_ref/libref_quadgen.so is built only where /root/reference is mounted")
    for name in ("delaunay", "sphere", "non_manifold"):
        pos, tris = qm.CASES[name]()
        o, r = qb.run(qb.load_oracle(), pos, tris, weight), qb.run(qb.load_reference(), pos, tris, weight)
        assert same(o["quads"], r["quads"]), (name, weight)


@pytest.mark.parametrize("name", sorted(qm.CASES))
def test_kernel_rules_give_the_reference_graph(name):
    """quadgen_rules.h (closed per-element rules) against the restated sequential loops."""
    pos, tris = qm.CASES[name]()
    o, k = qb.run(qb.load_oracle(), pos, tris), qb.run_rules(pos, tris)
    for key in GRAPH_KEYS:
        assert same(o[key], k[key]), key


@pytest.mark.parametrize("name", sorted(qm.CASES))
def test_round_selection_is_a_valid_pairing_close_to_the_reference(name):
    pos, tris = qm.CASES[name]()
    ref = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=0)
    out = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=1)
    for key in GRAPH_KEYS:
        assert same(ref[key], out[key])
    check_pairing(tris, out)
    assert len(out["quads"]) <= COUNT_BOUND * len(ref["quads"]), (len(out["quads"]), len(ref["quads"]))


def test_round_selection_on_a_large_irregular_mesh():
    pos, tris = qm.delaunay(60_000, seed=21)
    ref = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=0)
    out = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=1)
    assert len(out["quads"]) <= 1.005 * len(ref["quads"])
    assert out["rounds"] < 400


def test_c_abi_exports_quadgen():
    import ctypes as C

    from lucid_b200 import build
    lib = C.CDLL(build.SO_PATH)
    for sym in ("lucid_quadgen", "lucid_quadgen_last_error"):
        getattr(lib, sym)


# ---- GPU ---------------------------------------------------------------------------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(qm.CASES))
def test_cuda_equals_restatement(name):
    from lucid_b200 import quadgen
    pos, tris = qm.CASES[name]()
    out = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=1)
    quads, info = quadgen.gen_quads(pos, tris, 4.0, with_graph=True)
    for key in GRAPH_KEYS:
        assert same(out[key], info[key]), key
    assert np.array_equal(out["selected"] == 2, info["selected"] == 2)
    assert np.array_equal(quads, out["quads"])
    assert (info["num_degenerate"], info["rounds"], info["num_augmented"]) == (out["num_degenerate"], out["rounds"], out["num_augmented"])


@pytest.mark.gpu
@pytest.mark.parametrize("weight", [0.0, 16.0])
def test_cuda_equals_restatement_other_weights(weight):
    from lucid_b200 import quadgen
    pos, tris = qm.delaunay(20_000, seed=31)
    out = qb.run(qb.load_oracle(), pos, tris, weight, mode=1)
    quads, info = quadgen.gen_quads(pos, tris, weight)
    assert np.array_equal(quads, out["quads"]) and info["rounds"] == out["rounds"]


@pytest.mark.gpu
def test_cuda_large_mesh_against_restatement_and_reference_count():
    """2 M triangles: equal to the restatement's round selection, quad count within 0.5 % of the reference algorithm."""
    from lucid_b200 import quadgen
    pos, tris = qm.grid(1000, 1000, jitter=0.5, seed=41)
    ref = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=0)
    out = qb.run(qb.load_oracle(), pos, tris, 4.0, mode=1)
    quads, info = quadgen.gen_quads(pos, tris, 4.0)
    assert np.array_equal(quads, out["quads"])
    assert len(quads) <= 1.005 * len(ref["quads"])


@pytest.mark.gpu
def test_cuda_paired_mesh_renders_like_the_triangle_mesh():
    """The point of the pairing: the paired quads cover the triangles' pixels.  A patch rendered from its triangles as
    degenerate quads and from the GPU's quads has the same per-pixel fragment counts (up to samples on shared edges)."""
    from lucid_b200 import api, quadgen, scenes
    from tests import parity_util as pu
    sc = scenes.meshlet_patches(num_patches=40, grid=16, width=640, height=360, seed=9, extent=6.0)
    quads0 = sc["quads"]
    tris = np.concatenate([quads0[:, [0, 1, 2]], quads0[:, [0, 2, 3]]]).astype(np.int32)
    tris = tris[tris[:, 1] != tris[:, 2]]
    keep = tris[:, 0] != tris[:, 2]
    tris = np.ascontiguousarray(tris[keep])
    quads, info = quadgen.gen_quads(sc["positions"], tris, 4.0)
    assert len(quads) < 0.6 * len(tris)

    def frag_counts(q):
        s = dict(sc)
        s["quads"] = np.ascontiguousarray(q, np.int32)
        s["draw_calls"] = [(0, len(q), 0, 0)]  # (material, quads, first quad, instance flags)
        s["colors"] = s["uvs"] = s["normals"] = None
        r, _ = pu.run_cuda(s)
        try:
            return r.read_frag_counts()
        finally:
            r.close()

    degenerate = np.concatenate([tris, tris[:, 2:3]], 1)
    a, b = frag_counts(degenerate), frag_counts(quads)
    # the same triangles, but with their vertices rotated inside the quads: edge set-up rounds differently, so a sample
    # exactly on an edge may fall to the other side
    assert a.sum() > 5_000 and abs(int(a.sum()) - int(b.sum())) <= 2e-3 * a.sum()
    assert np.count_nonzero(a != b) <= 2e-3 * np.count_nonzero(a)


@pytest.mark.gpu
def test_cuda_rejects_bad_input():
    from lucid_b200 import quadgen
    pos, tris = qm.CASES["pair"]()
    bad = tris.copy()
    bad[1, 2] = 99
    with pytest.raises(ValueError):
        quadgen.gen_quads(pos, bad)
    quads, info = quadgen.gen_quads(pos, np.zeros((0, 3), np.int32))
    assert len(quads) == 0
