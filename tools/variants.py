"""Builds experiment variants of the library (compile-time defines) into gpurun_out/variants/ and
prints the command that times them on the GPU box:  python tools/variants.py name=DEF1,DEF2 ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import build  # noqa: E402

out_dir = os.path.join(build.ROOT, "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    path = os.path.join(out_dir, f"lucid_{name}.so")
    build.build(force=True, defines=[d for d in defs.split(",") if d], out=path)
    print("built", path)
