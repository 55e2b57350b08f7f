import struct, sys
import numpy as np
sys.path.insert(0, "tools/hwtex")
import model
# probe 2
data = open("gpurun_out/hwtex_probe2.bin", "rb").read()
n = struct.unpack("i", data[:4])[0]
uvl = np.frombuffer(data[4:4 + n * 12], np.float32).reshape(n, 3)
oa = np.frombuffer(data[4 + n * 12:4 + n * 28], np.float32).reshape(n, 4)
ob = np.frombuffer(data[4 + n * 28:4 + n * 44], np.float32).reshape(n, 4)
a0 = np.zeros((8, 8, 4), np.int64); a1 = np.zeros((4, 4, 4), np.int64); b0 = np.zeros((8, 8, 4), np.int64); b1 = np.zeros((4, 4, 4), np.int64)
a0[3, 3] = 255; b1[1, 1] = 255
for nm, lv, out in (("A", [a0, a1], oa), ("B", [b0, b1], ob)):
    m = model.trilinear16(lv, uvl[:, 0], uvl[:, 1], uvl[:, 2])
    g16 = np.round(out.astype(np.float64) * 65535).astype(np.int64)
    print("probe2 texture", nm, "mismatches", int((m != g16).any(axis=1).sum()), "of", n)
# probe 1 trilinear + bilinear
data = open("gpurun_out/hwtex_probe.bin", "rb").read()
pos = 0; tests = {}; tex = {}
while pos < len(data):
    if data[pos:pos + 8] in (b"texdata3", b"texdata4"):
        tag = data[pos:pos + 8].decode(); pos += 8
        k = 512 * 2 * 4 if tag == "texdata3" else sum((64 >> l) ** 2 * 4 for l in range(4))
        tex[tag] = np.frombuffer(data[pos:pos + k], np.uint8); pos += k; continue
    name = data[pos:pos + 16].split(b"\0")[0].decode(); k = struct.unpack("i", data[pos + 16:pos + 20])[0]; pos += 20
    u = np.frombuffer(data[pos:pos + k * 12], np.float32).reshape(k, 3); pos += k * 12
    o = np.frombuffer(data[pos:pos + k * 16], np.float32).reshape(k, 4); pos += k * 16
    tests[name] = (u, o)
levels = []; off = 0
for l in range(4):
    w = 64 >> l
    levels.append(tex["texdata4"][off:off + w * w * 4].reshape(w, w, 4).astype(np.int64)); off += w * w * 4
for name in ("bilinear", "trilinear"):
    u, o = tests[name]
    m = model.trilinear16(levels, u[:, 0], u[:, 1], u[:, 2])
    g16 = np.round(o.astype(np.float64) * 65535).astype(np.int64)
    bad = (m != g16).any(axis=1)
    print(name, "mismatches", int(bad.sum()), "of", len(bad), "max |diff|", np.abs(m - g16).max())
    f32 = (m.astype(np.float32) / np.float32(65535.0)).astype(np.float32)
    print("   float bits equal on matching samples:", bool((f32[~bad] == o[~bad]).all()))
    if bad.any():
        for i in np.nonzero(bad)[0][:5]:
            print("   ", u[i].tolist(), g16[i].tolist(), m[i].tolist())
