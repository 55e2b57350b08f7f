"""Full-size probe: stage times + counters of the CUDA path on the BASELINE configs (parity against the
checker is the GPU test suite's job).  usage: python tools/gpu_probe.py [config ...] [--scale=S] [--opts=BITS] [--compact]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
scale, opts, create_flags, mbe = 1.0, 0, 0, 0
for a in sys.argv[1:]:
    if a == "--compact":
        create_flags, mbe = api.CREATE_COMPACT_LISTS, 0
    if a.startswith("--scale="):
        scale = float(a.split("=")[1])
    if a.startswith("--opts="):
        opts = int(a.split("=")[1], 0)
configs = [int(a) for a in args] or [0, 1]
for ci in configs:
    t0 = time.time()
    sc = scenes.get_config(ci, scale)
    cfg, inst, cols, rects = api.prepare_frame(sc)
    r = api.LucidRenderer(sc["width"], sc["height"], opts, 0, create_flags=create_flags, max_block_entries=mbe)
    r.set_scene(sc)
    img = np.zeros((sc["height"], sc["width"]), np.uint32)
    for _ in range(3):
        r.render(cfg, inst, cols, rects, out=img, flags=api.RENDER_FRAG_COUNTS)
    times = []
    for _ in range(5):
        r.render(cfg, inst, cols, rects)
        times.append(r.stage_times())
    ms = np.median(np.array(times), axis=0)
    st = r.getStats()
    print(f"== config {ci} ({sc['name']}) opts {opts:#x} gen+setup {time.time() - t0:.1f}s")
    print("   stage_ms", dict(zip(["setup", "count", "scan", "dispatch", "lists", "sort", "shade", "frame"],
                                  np.round(ms, 3).tolist())))
    print("   stats", json.dumps(st))
    print("   verifyInfo", r.verifyInfo()[:3], "frag image sum", int(r.read_frag_counts().sum()))
    try:
        from PIL import Image
        os.makedirs("gpurun_out", exist_ok=True)
        Image.fromarray(img.view(np.uint8).reshape(sc["height"], sc["width"], 4)[:, :, :3]).save(
            f"gpurun_out/config{ci}.png")
    except Exception as e:
        print("   (no image saved)", e)
    r.close()
