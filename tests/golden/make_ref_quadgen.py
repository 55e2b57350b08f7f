"""Writes tests/golden/ref_quadgen.json: digests of what the REFERENCE's src/quad_generator.cpp (compiled by
oracle/build_ref_quadgen.py into oracle/_ref/libref_quadgen.so) returns for the seeded meshes of tests/quadgen_meshes.py.
Run where /root/reference is mounted:  python tests/golden/make_ref_quadgen.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import quadgen_binding as qb  # noqa: E402
from tests import quadgen_meshes as qm  # noqa: E402
from tests.test_quadgen import digests  # noqa: E402

if not qb.reference_available():
    sys.exit("oracle/_ref/libref_quadgen.so missing: run `make -C oracle ref` with the reference tree mounted")
out = {name: digests(qb.run(qb.load_reference(), *make(), 4.0)) for name, make in sorted(qm.CASES.items())}
with open(os.path.join(ROOT, "tests", "golden", "ref_quadgen.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", len(out), "cases")
