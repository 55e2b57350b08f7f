"""Static SASS instruction count per enclosing source function, from the same dump ncu_funcs.py reads:
python tools/ncu_codesize.py dump.csv   (each SASS row is attributed to the CUDA line above it)"""
import csv
import os
import sys

sys.argv = [sys.argv[0], sys.argv[1]]
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
import importlib.util
spec = importlib.util.spec_from_file_location("nf", os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_funcs.py"))
import io, contextlib
nf = importlib.util.module_from_spec(spec)
with contextlib.redirect_stdout(io.StringIO()):
    spec.loader.exec_module(nf)
cur = None
fn = None
size = {}
executed = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) < 10 or r[0] == "Line No":
        continue
    if r[2] == "-":
        fn = (os.path.basename(cur), nf.enclosing(cur, int(r[0])))
    elif r[2].startswith("0x") and fn:
        size[fn] = size.get(fn, 0) + 1
tot = sum(size.values())
print("SASS instructions listed:", tot, "(collapsed '...' rows are not counted)")
for k, v in sorted(size.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{v:6d}  {k[0]}:{k[1]}")
