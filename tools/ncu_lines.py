"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by source line:
python tools/ncu_lines.py dump.csv [top_n]  -> instructions executed and stall samples per line."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
agg = {}
total_inst = 0
total_samples = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue  # only cuda-source summary rows (address "-")
    try:
        inst = int(r[hdr["Instructions Executed"]])
        samples = int(r[hdr["# Samples"]])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0])
    a[0] += inst
    a[1] += samples
    total_inst += inst
    total_samples += samples
print(f"total instructions {total_inst:,}  samples {total_samples:,}")
for (f, ln, src), (inst, samples) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * inst / max(total_inst, 1):5.1f}% inst {100 * samples / max(total_samples, 1):5.1f}% smp  {f}:{ln}  {src}")
