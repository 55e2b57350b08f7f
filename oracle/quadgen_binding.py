"""ctypes binding of the triangle -> quad pairing checkers.

TEST INFRASTRUCTURE: imported only by tests/ and bench-side CPU legs.  Three libraries share one call signature:
  oracle/libquadgen_oracle.so        the restatement (mode 0 = the reference's algorithm, mode 1 = the round-synchronous
                                     selection the CUDA path computes)
  oracle/_ref/libref_quadgen.so      the REFERENCE's own src/quad_generator.cpp compiled by build_ref_quadgen.py
                                     (exists only where /root/reference was mounted at build time)
  tests/cpp/libquadgen_rules.so      the per-element rules header of the CUDA kernels run on the CPU (graph only)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libquadgen_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_quadgen.so")
RULES_SO = os.path.join(os.path.dirname(_HERE), "tests", "cpp", "libquadgen_rules.so")
RULES_SRC = os.path.join(os.path.dirname(_HERE), "tests", "cpp", "quadgen_rules_harness.cpp")
RULES_HDR = os.path.join(os.path.dirname(_HERE), "lucid_b200", "csrc", "quadgen_rules.h")
_libs = {}


def _signature(fn):
    vp = C.c_void_p
    fn.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    fn.restype = C.c_int
    return fn


def _older(target, *sources):
    return not os.path.exists(target) or any(os.path.getmtime(target) < os.path.getmtime(s) for s in sources)


def load_oracle():
    if "oracle" not in _libs:
        if _older(ORACLE_SO, os.path.join(_HERE, "quadgen_oracle.cpp")):
            subprocess.run(["make", "-C", _HERE, "-s", "libquadgen_oracle.so"], check=True, capture_output=True)
        _libs["oracle"] = _signature(C.CDLL(ORACLE_SO).quadgen_oracle)
    return _libs["oracle"]


def reference_available() -> bool:
    return os.path.exists(REF_SO)


def load_reference():
    if "ref" not in _libs:
        _libs["ref"] = _signature(C.CDLL(REF_SO).ref_quadgen)
    return _libs["ref"]


def load_rules():
    if "rules" not in _libs:
        if _older(RULES_SO, RULES_SRC, RULES_HDR):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-march=x86-64-v3", "-shared",
                            "-o", RULES_SO, RULES_SRC], check=True, capture_output=True)
        fn = C.CDLL(RULES_SO).quadgen_rules_graph
        vp = C.c_void_p
        fn.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
        fn.restype = C.c_int
        _libs["rules"] = fn
    return _libs["rules"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(fn, positions, tris, square_weight=4.0, mode=0):
    """-> dict(quads, num_degenerate, rounds, neighbours, tri_quads, node_tris, node_verts, node_conflicts, squareness,
    selected)"""
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    nt = len(tris)
    cap = 2 * nt + 4
    out = np.zeros((max(nt, 1), 4), np.int32)
    counts = np.zeros(8, np.int32)
    nb, tq = np.zeros((max(nt, 1), 3), np.int32), np.zeros((max(nt, 1), 3), np.int32)
    ntris, nverts, nconf = np.zeros((cap, 2), np.int32), np.zeros((cap, 4), np.int32), np.zeros((cap, 4), np.int32)
    sq, sel = np.zeros(cap, np.float32), np.zeros(cap, np.uint8)
    rc = fn(_p(positions), len(positions), _p(tris), nt, square_weight, mode, _p(out), _p(counts), _p(nb), _p(tq),
            _p(ntris), _p(nverts), _p(nconf), _p(sq), _p(sel))
    assert rc == 0
    n = int(counts[2])
    return dict(quads=out[:counts[0]].copy(), num_degenerate=int(counts[1]), rounds=int(counts[3]), num_augmented=int(counts[4]), neighbours=nb[:nt],
                tri_quads=tq[:nt], node_tris=ntris[:n], node_verts=nverts[:n], node_conflicts=nconf[:n],
                squareness=sq[:n], selected=sel[:n])


def run_rules(positions, tris):
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    nt = len(tris)
    cap = 2 * nt + 4
    nb, tq = np.zeros((max(nt, 1), 3), np.int32), np.zeros((max(nt, 1), 3), np.int32)
    ntris, nverts, nconf = np.zeros((cap, 2), np.int32), np.zeros((cap, 4), np.int32), np.zeros((cap, 4), np.int32)
    sq = np.zeros(cap, np.float32)
    n = load_rules()(_p(positions), _p(tris), nt, _p(nb), _p(tq), _p(ntris), _p(nverts), _p(nconf), _p(sq))
    return dict(neighbours=nb[:nt], tri_quads=tq[:nt], node_tris=ntris[:n], node_verts=nverts[:n],
                node_conflicts=nconf[:n], squareness=sq[:n])
