"""Fuzz of the host-side input preparation (lucid_host_orbit_camera / lucid_host_make_config: FrustumInfo, view-projection)
against the reference's own libfwk camera code (oracle/_ref/ref_camera, i.e. the build container), with the
tolerances of tests/test_host.py::test_config_matches_reference_camera:
    python tests/fuzz_ref_camera.py [cameras] [seed]
Not collected by pytest; tests/golden/ref_camera.json is the regression pin."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tests.test_host import test_config_matches_reference_camera as check  # noqa: E402

BIN = os.path.join(HERE, "..", "oracle", "_ref", "ref_camera")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for k in range(n):
    w, h = [(1280, 720), (1920, 1080), (3840, 2160), (640, 360), (2560, 1330), (333, 777)][k % 6]
    fov, znear, zfar = float(rng.uniform(20, 100)), float(rng.choice([0.0625, 0.1, 0.5])), float(rng.choice([500, 1024, 4000]))
    if k % 2 == 0:
        kind = "orbit"
        args = [*rng.uniform(-5, 5, 3).tolist(), float(rng.uniform(0.5, 80)), float(rng.uniform(0, 6.28)), float(rng.uniform(-1.4, 1.4)),
                fov, znear, zfar, w, h]
    else:
        kind = "lookat"
        pos = rng.uniform(-40, 40, 3)
        target = pos + rng.normal(size=3) * float(rng.uniform(0.5, 50))
        args = [*pos.tolist(), *target.tolist(), 0.0, 1.0, 0.0, fov, znear, zfar, w, h]
    args = [float(np.float32(a)) for a in args[:-2]] + [w, h]
    txt = subprocess.run([BIN, kind] + [repr(a) for a in args[:-2]] + [str(w), str(h)], check=True, capture_output=True,
                         text=True).stdout
    case = {"kind": kind, "args": args}
    for line in txt.strip().splitlines():
        parts = line.split()
        case[parts[0]] = [float(v) for v in parts[1:]]
    try:
        check(case)
    except AssertionError as e:
        bad += 1
        if bad <= 3:
            print("differs:", kind, args, str(e)[:300])
# material colour of uploadInstances: u32(IColor(FColor(diffuse, opacity))) (src/lucid_renderer.cpp:364-365), exact
from lucid_b200 import api  # noqa: E402

bad_colors = 0
for k in range(n):
    c = rng.uniform(-0.2, 1.3, 4) if k % 4 == 0 else rng.integers(0, 256, 4) / 255.0 + rng.choice([0.0, 1e-7, -1e-7])
    c = [float(np.float32(v)) for v in c]
    want = int(subprocess.run([BIN, "color"] + [repr(v) for v in c], check=True, capture_output=True, text=True).stdout.split()[1])
    _, cols, _ = api.build_instances([(0, 1, 0, 0)], [((c[0], c[1], c[2]), c[3], (0.0, 0.0, 1.0, 1.0))])
    if int(cols[0]) != want:
        bad_colors += 1
        if bad_colors <= 3:
            print("colour differs:", c, hex(int(cols[0])), hex(want))
print("fuzz:", n, "cameras, outside the tolerances:", bad, ";", n, "material colours, differing:", bad_colors)
sys.exit(1 if bad or bad_colors else 0)
