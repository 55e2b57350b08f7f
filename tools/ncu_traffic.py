"""DRAM bytes (read + write) per stage of one captured frame, from `ncu --set full` reports:
   python tools/ncu_traffic.py config1=a.ncu-rep config3=b.ncu-rep > profiles/dram_traffic.json
Stages follow bench.py's roofline keys; a stage's traffic is summed over its kernels."""
import csv
import io
import json
import subprocess
import sys

STAGE = {"k_instance_select": "setup", "k_quad_cull": "setup", "k_tri_setup": "setup", "k_bin_count": "bin_count",
         "k_bin_dispatch": "bin_dispatch", "k_raster_bins": "raster", "k_block_sort": "raster", "k_tie_runs": "raster", "k_block_shade": "raster"}
out = {}
for spec in sys.argv[1:]:
    name, _, path = spec.partition("=")
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    stages, kernels, issue = {}, {}, {}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    for r in rows[2:]:
        # "void k_block_sort<0>(Params, unsigned int)" -> k_block_sort
        kname = r[idx["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]]
        kernels[kname] = kernels.get(kname, 0) + int(total)
        if kname in STAGE:
            stages[STAGE[kname]] = stages.get(STAGE[kname], 0) + int(total)
            # SM issue-slot utilisation of the stage: time-weighted over its kernels
            t = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * tscale[units[idx["gpu__time_duration.sum"]]]
            ia = float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", ""))
            acc = issue.setdefault(STAGE[kname], [0.0, 0.0])
            acc[0] += t * ia
            acc[1] += t
    stages["issue_active_pct"] = {k: round(v[0] / v[1], 1) for k, v in issue.items() if v[1] > 0}
    stages["per_kernel"] = kernels
    stages["source"] = path.split("/")[-1] + " (ncu --set full, one frame, cold-cache serialised replays)"
    out[name] = stages
# the digest of the kernel sources the capture was taken with: bench.py only quotes a capture that matches
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
out["kernel_source_hash"] = bench.kernel_source_hash()
print(json.dumps(out, indent=1))
