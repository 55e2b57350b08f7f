"""Generates tests/golden/ref_camera.json by running oracle/_ref/ref_camera -- the REFERENCE's own
libfwk camera/frustum code compiled from /root/reference (oracle/build_ref.sh).  Run in the build
container (the reference tree does not exist on the GPU box); the JSON travels with the repo."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "..", "..", "oracle", "_ref", "ref_camera")

CASES = [
    ("orbit", [0, 0, 0, 30, 0.5, 0.8, 60, 0.0625, 1024, 1280, 720]),
    ("orbit", [0, 0, 0, 50, 0.5, 0.6, 60, 0.0625, 1024, 1920, 1080]),
    ("orbit", [0, 0, 0, 22, 0.5, 0.6, 60, 0.0625, 1024, 3840, 2160]),
    ("orbit", [1.5, -2, 3, 10, 2.5, -0.4, 45, 0.1, 500, 640, 360]),
    ("orbit", [0, 0, 0, 10, 0.5, 0.8, 60, 0.0625, 1024, 2560, 1330]),
    ("lookat", [-30, 3, -28, 10, 1, 12, 0, 1, 0, 60, 0.0625, 1024, 3840, 2160]),
    ("lookat", [0, 0, -5, 0.005, 0.005, 0, 0, 1, 0, 60, 0.0625, 1024, 1280, 720]),
]


# material colours (diffuse rgb, opacity) for the IColor(FColor(...)) conversion of uploadInstances
COLORS = [(1.0, 1.0, 1.0, 1.0), (1.0, 0.5, 0.25, 0.5), (0.0, 0.0, 0.0, 0.0), (0.999, 0.001, 0.5019608, 0.7490196),
          (0.2, 0.4, 0.6, 0.8), (1.5, -0.25, 0.0039215, 1.0), (0.3333333, 0.6666667, 0.9960785, 0.25),
          (0.0627451, 0.1254902, 0.2509804, 0.5019608)]


def main():
    out = []
    for kind, args in CASES:
        txt = subprocess.run([BIN, kind] + [repr(float(a)) if i < len(args) - 2 else str(int(a))
                                            for i, a in enumerate(args)], check=True, capture_output=True,
                             text=True).stdout
        rec = {"kind": kind, "args": args}
        for line in txt.strip().splitlines():
            parts = line.split()
            rec[parts[0]] = [float(v) for v in parts[1:]]
        out.append(rec)
    for c in COLORS:
        txt = subprocess.run([BIN, "color"] + [repr(float(v)) for v in c], check=True, capture_output=True, text=True).stdout
        out.append({"kind": "color", "args": list(c), "color": int(txt.split()[1])})
    with open(os.path.join(HERE, "ref_camera.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
