"""Triangle -> quad pairing on the GPU: host side of include/lucid_quadgen.h.

Mirrors the reference's call site Scene::generateQuads (src/scene.cpp:237-247): every mesh of a scene is paired on its
own, `quads` and `num_degenerate_quads` are stored with the mesh.  The work itself is lucid_b200/csrc/quadgen.cu; there
is no CPU fallback -- without the CUDA library this module raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


class QuadgenResult(C.Structure):
    _fields_ = [("num_quads", C.c_int32), ("num_degenerate", C.c_int32), ("num_nodes", C.c_int32), ("rounds", C.c_int32),
                ("device_ms", C.c_float), ("num_augmented", C.c_int32), ("reserved", C.c_int32 * 2)]


class QuadgenGraph(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("neighbours", "tri_quads", "node_tris", "node_verts", "node_conflicts",
                                          "squareness", "selected")]


def _lib():
    lib = api.load_library()
    if not getattr(lib, "_quadgen_bound", False):
        lib.lucid_quadgen.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_void_p,
                                      C.POINTER(QuadgenResult), C.POINTER(QuadgenGraph)]
        lib.lucid_quadgen.restype = C.c_int
        lib.lucid_quadgen_last_error.restype = C.c_char_p
        lib._quadgen_bound = True
    return lib


def gen_quads(positions, tris, square_weight: float = 4.0, device: int = 0, with_graph: bool = False):
    """triNeighbours + quadNodes + genQuads of one mesh.  -> (quads [Q, 4] int32, info dict); with_graph adds the
    pairing graph's arrays (the reference's intermediate products) to the dict."""
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    nt = len(tris)
    out = np.empty((max(nt, 1), 4), np.int32)
    res = QuadgenResult()
    graph, arrays = None, {}
    if with_graph:
        cap = 3 * nt // 2 + 2
        arrays = dict(neighbours=np.full((max(nt, 1), 3), -1, np.int32), tri_quads=np.full((max(nt, 1), 3), -1, np.int32),
                      node_tris=np.zeros((cap, 2), np.int32), node_verts=np.zeros((cap, 4), np.int32),
                      node_conflicts=np.zeros((cap, 4), np.int32), squareness=np.zeros(cap, np.float32),
                      selected=np.zeros(cap, np.uint8))
        graph = QuadgenGraph(*[a.ctypes.data for a in arrays.values()])
    lib = _lib()
    rc = lib.lucid_quadgen(positions.ctypes.data, len(positions), tris.ctypes.data, nt, float(square_weight), device,
                           out.ctypes.data, C.byref(res), C.byref(graph) if graph is not None else None)
    if rc != 0:
        msg = lib.lucid_quadgen_last_error().decode()
        raise (ValueError if rc == -1 else RuntimeError)(msg)
    info = dict(num_degenerate=res.num_degenerate, num_nodes=res.num_nodes, rounds=res.rounds, device_ms=res.device_ms,
                num_augmented=res.num_augmented)
    if with_graph:
        n = res.num_nodes
        for k, a in arrays.items():
            info[k] = a[:nt] if k in ("neighbours", "tri_quads") else a[:n]
    return out[:res.num_quads].copy(), info


def generate_quads(meshes, positions, square_weight: float = 4.0, device: int = 0):
    """Scene::generateQuads: meshes is a list of dicts with "tris"; sets "quads" and "num_degenerate_quads" on each."""
    for mesh in meshes:
        quads, info = gen_quads(positions, mesh["tris"], square_weight, device)
        mesh["quads"], mesh["num_degenerate_quads"] = quads, info["num_degenerate"]
    return meshes
