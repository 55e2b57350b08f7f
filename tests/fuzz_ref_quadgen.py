"""Fuzz of the restated quad generator (oracle/quadgen_oracle.cpp, mode 0) against the reference's own
src/quad_generator.cpp (oracle/_ref/libref_quadgen.so, i.e. the build container), output for output, on random meshes:
    python tests/fuzz_ref_quadgen.py [rounds] [seed]
Not collected by pytest; tests/golden/ref_quadgen.json is the regression pin."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import quadgen_binding as qb  # noqa: E402
from tests import quadgen_meshes as qm  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
oracle, ref = qb.load_oracle(), qb.load_reference()
from tests.test_quadgen import GRAPH_KEYS, same  # noqa: E402

KEYS = GRAPH_KEYS + ("quads",)  # what the reference's code outputs: neighbours, nodes, squareness, conflicts, tri_quads, quads
bad = tris_total = 0
for r in range(rounds):
    kind = r % 6
    seed = int(rng.integers(1, 1 << 30))
    if kind == 0:
        pos, tris = qm.grid(int(rng.integers(2, 60)), int(rng.integers(2, 60)), jitter=float(rng.uniform(0, 0.45)), seed=seed,
                            random_diagonals=bool(rng.integers(0, 2)), shuffle=bool(rng.integers(0, 2)))
    elif kind == 1:
        pos, tris = qm.delaunay(int(rng.integers(10, 3000)), seed=seed)
    elif kind == 2:
        pos, tris = qm.soup(int(rng.integers(1, 800)), seed=seed)
    elif kind == 3:
        pos, tris = qm.sphere(int(rng.integers(3, 40)), int(rng.integers(3, 80)), seed=seed)
    elif kind == 4:
        pos, tris = qm.non_manifold(seed=seed)
    else:
        pos, tris = qm.strip(int(rng.integers(1, 300)), seed=seed)
    weight = float(rng.choice([0.0, 0.5, 1.0, 4.0, 16.0]))
    a, b = qb.run(oracle, pos, tris, weight), qb.run(ref, pos, tris, weight)
    tris_total += len(tris)
    diff = [k for k in KEYS if not same(np.asarray(a[k]), np.asarray(b[k]))]
    if diff or a["num_degenerate"] != b["num_degenerate"]:
        bad += 1
        print("differs:", kind, seed, weight, len(tris), diff)
print("fuzz:", rounds, "meshes,", tris_total, "triangles, mismatching meshes:", bad)
sys.exit(1 if bad else 0)
