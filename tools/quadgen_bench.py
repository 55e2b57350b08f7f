"""Triangle -> quad pairing on a B200 against the reference's CPU code.

  python tools/quadgen_bench.py [--tris 10000000] [--cpu-tris 1000000] [--mesh grid|delaunay]

GPU: lucid_quadgen through the C ABI on the whole mesh -- device_ms (kernels, CUDA events) and the wall time of the
call with its host copies.  CPU: the reference's own src/quad_generator.cpp (oracle/_ref/libref_quadgen.so, one thread:
the algorithm is sequential) when that library travelled with the snapshot, else the restatement, on a bounded sample
of the same mesh.  Prints one JSON line.  Test / measurement infrastructure (loads oracle/).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lucid_b200 import quadgen  # noqa: E402
from oracle import quadgen_binding as qb  # noqa: E402
from tests import quadgen_meshes as qm  # noqa: E402


def mesh(kind, ntris, seed):
    if kind == "grid":
        n = int(round((ntris / 2) ** 0.5))
        return qm.grid(n, n, jitter=0.5, seed=seed)
    return qm.delaunay(ntris // 2, seed=seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--cpu-tris", type=int, default=1_000_000)
    ap.add_argument("--mesh", default="grid")
    args = ap.parse_args()
    pos, tris = mesh(args.mesh, args.tris, 51)
    quadgen.gen_quads(*mesh(args.mesh, 20_000, 52))  # context, module load
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        quads, info = quadgen.gen_quads(pos, tris, 4.0)
        wall = time.perf_counter() - t0
        if best is None or info["device_ms"] < best[0]:
            best = (info["device_ms"], wall, info)
    dev_ms, wall, info = best
    cpos, ctris = mesh(args.mesh, args.cpu_tris, 51)
    kind = "reference" if qb.reference_available() else "port"
    fn = qb.load_reference() if kind == "reference" else qb.load_oracle()
    t0 = time.perf_counter()
    ref = qb.run(fn, cpos, ctris, 4.0)
    cpu_s = time.perf_counter() - t0
    small_quads, _ = quadgen.gen_quads(cpos, ctris, 4.0)
    print(json.dumps({
        "what": "triangle -> quad pairing", "mesh": args.mesh, "triangles": len(tris), "quads": len(quads),
        "degenerate": info["num_degenerate"], "nodes": info["num_nodes"], "rounds": info["rounds"],
        "augmented": info["num_augmented"], "device_ms": round(dev_ms, 3), "call_wall_ms": round(wall * 1e3, 1),
        "mtris_per_s_device": round(len(tris) / dev_ms / 1e3, 1), "mtris_per_s_call": round(len(tris) / wall / 1e6, 1),
        "cpu_baseline": {"kind": kind, "cores": 1, "sample": f"{len(ctris)} triangles of the same mesh generator",
                         "seconds": round(cpu_s, 3), "mtris_per_s": round(len(ctris) / cpu_s / 1e6, 3),
                         "quads": len(ref["quads"]), "gpu_quads_same_sample": len(small_quads)},
        "algorithmic_bytes": int(12 * len(tris) + 12 * len(pos) + 16 * len(quads)),
    }))


if __name__ == "__main__":
    main()
