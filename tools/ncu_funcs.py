"""Like ncu_lines.py but aggregated per enclosing source function (by scanning the .cu/.cuh text)."""
import csv
import re
import sys
import os

path = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(path)))
func_cache = {}


def funcs_of(fpath):
    if fpath in func_cache:
        return func_cache[fpath]
    out = []
    try:
        local = fpath
        if not os.path.exists(local):
            for sub in ("lucid_b200/csrc", "lucid_b200/host", "include"):
                cand = os.path.join(root, sub, os.path.basename(fpath))
                if os.path.exists(cand):
                    local = cand
        lines = open(local).read().splitlines()
    except OSError:
        lines = []
    for i, ln in enumerate(lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline|void|int)\b.*?(\w+)\s*\(", ln)
        if m and not ln.startswith("\t") and "=" not in ln.split("(")[0]:
            out.append((i, m.group(1)))
    func_cache[fpath] = out
    return out


def enclosing(fpath, line):
    name = "?"
    for start, n in funcs_of(fpath):
        if start <= line:
            name = n
        else:
            break
    return name


cur = None
hdr = None
agg = {}
tot_i = tot_s = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    try:
        inst = int(r[hdr["Instructions Executed"]])
        smp = int(r[hdr["# Samples"]])
        tinst = int(r[hdr["Thread Instructions Executed"]])
    except ValueError:
        continue
    key = (os.path.basename(cur), enclosing(cur, int(r[0])))
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += inst
    a[1] += smp
    a[2] += tinst
    tot_i += inst
    tot_s += smp
print(f"total warp instructions {tot_i:,} samples {tot_s:,}")
for (f, fn), (inst, smp, tinst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{100 * inst / tot_i:5.1f}% inst {100 * smp / max(tot_s, 1):5.1f}% smp  avg threads {tinst / max(inst, 1):4.1f}  {f}:{fn}")
