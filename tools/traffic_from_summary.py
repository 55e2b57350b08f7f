"""profiles/dram_traffic.json from the text summaries tools/ncu_summary.py wrote (same content as
tools/ncu_traffic.py, for when the .ncu-rep files were not kept):
   python tools/traffic_from_summary.py config1=profiles/r1i_ncu_full_config1.txt ... > profiles/dram_traffic.json"""
import json
import re
import sys

STAGE = {"k_quad_cull": "setup", "k_tri_setup": "setup", "k_bin_count": "bin_count", "k_bin_dispatch": "bin_dispatch",
         "k_raster_bins": "raster", "k_raster_blocks": "raster"}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
out = {}
for spec in sys.argv[1:]:
    name, _, path = spec.partition("=")
    stages, kernels, issue = {}, {}, {}
    cur, vals = None, {}

    def flush():
        if cur is None:
            return
        total = int(vals.get("dram__bytes_read.sum", 0) + vals.get("dram__bytes_write.sum", 0))
        kernels[cur] = kernels.get(cur, 0) + total
        if cur in STAGE:
            st = STAGE[cur]
            stages[st] = stages.get(st, 0) + total
            acc = issue.setdefault(st, [0.0, 0.0])
            acc[0] += vals.get("gpu__time_duration.sum", 0) * vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
            acc[1] += vals.get("gpu__time_duration.sum", 0)

    for line in open(path):
        if line.startswith("=="):
            flush()
            cur, vals = line.split()[1].split("(")[0].split("::")[-1], {}
            continue
        m = re.match(r"\s+(\S+)\s+([\d.,]+)\s*(\S*)", line)
        if not m:
            continue
        v = float(m.group(2).replace(",", ""))
        unit = m.group(3)
        v *= BYTES.get(unit, TIME.get(unit, 1.0)) if (m.group(1).startswith("dram__bytes") or m.group(1).startswith("gpu__time")) else 1.0
        vals[m.group(1)] = v
    flush()
    stages["issue_active_pct"] = {k: round(a / b, 1) for k, (a, b) in issue.items() if b > 0}
    stages["per_kernel"] = kernels
    stages["source"] = path.split("/")[-1] + " (ncu --set full, one frame, cold-cache serialised replays)"
    out[name] = stages
print(json.dumps(out, indent=1))
