"""Where does end-to-end frame time go?  python tools/e2e_probe.py [config]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucid_b200 import api, scenes  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sc = scenes.get_config(ci)
w, h = sc["width"], sc["height"]
cfg, inst, cols, rects = api.prepare_frame(sc)
r = api.LucidRenderer(w, h, 0, 0)
r.set_scene(sc)
pinned = [torch.empty((h, w), dtype=torch.int32).pin_memory() for _ in range(2)]
N = 200


def run(name, fn):
    for k in range(5):
        fn(k)
    r.wait()
    t0 = time.perf_counter()
    for k in range(N):
        fn(k)
    t1 = time.perf_counter()
    r.wait()
    t2 = time.perf_counter()
    print(f"{name:40s} submit {1e3 * (t1 - t0) / N:.3f} ms/frame  total {1e3 * (t2 - t0) / N:.3f} ms/frame")


run("device only, async, skip info", lambda k: r.render(cfg, inst, cols, rects, flags=api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_NO_STAGE_TIMES))
run("device only, async, with info", lambda k: r.render(cfg, inst, cols, rects, flags=api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES))
run("host pinned, async", lambda k: r.render(cfg, inst, cols, rects, out=pinned[k & 1].data_ptr(), flags=api.RENDER_ASYNC | api.RENDER_NO_STAGE_TIMES))
run("host pinned, async, skip info", lambda k: r.render(cfg, inst, cols, rects, out=pinned[k & 1].data_ptr(), flags=api.RENDER_ASYNC | api.RENDER_SKIP_INFO | api.RENDER_NO_STAGE_TIMES))
run("host pinned, sync", lambda k: r.render(cfg, inst, cols, rects, out=pinned[k & 1].data_ptr()))
cam = api.make_camera(sc["camera"], w, h)
t0 = time.perf_counter()
for k in range(N):
    api.make_config(api.make_camera(sc["camera"], w, h), len(inst), sc["background"])
print(f"python config build {1e3 * (time.perf_counter() - t0) / N:.3f} ms")
d = torch.empty((h, w), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(50):
    pinned[0].copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 50
print(f"D2H {w * h * 4 / 1e6:.1f} MB pinned: {1e3 * dt:.3f} ms = {w * h * 4 / dt / 1e9:.1f} GB/s")
r.close()
