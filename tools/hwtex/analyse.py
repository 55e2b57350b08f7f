"""Reads gpurun_out/hwtex_probe.bin (tools/hwtex/probe.cu) and fits the texture unit's filter."""
import struct
import sys

import numpy as np

data = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/hwtex_probe.bin", "rb").read()
pos = 0
tests, tex = {}, {}
while pos < len(data):
    if data[pos:pos + 8] in (b"texdata3", b"texdata4"):
        tag = data[pos:pos + 8].decode()
        pos += 8
        n = 512 * 2 * 4 if tag == "texdata3" else sum((64 >> l) ** 2 * 4 for l in range(4))
        tex[tag] = np.frombuffer(data[pos:pos + n], np.uint8)
        pos += n
        continue
    name = data[pos:pos + 16].split(b"\0")[0].decode()
    n = struct.unpack("i", data[pos + 16:pos + 20])[0]
    pos += 20
    uvl = np.frombuffer(data[pos:pos + n * 12], np.float32).reshape(n, 3)
    pos += n * 12
    out = np.frombuffer(data[pos:pos + n * 16], np.float32).reshape(n, 4)
    pos += n * 16
    tests[name] = (uvl, out)

for name in ("stair_x8", "stair_x4096", "stair_lod"):
    uvl, out = tests[name]
    r = out[:, 0].astype(np.float64)
    vals = np.unique(r)
    print(f"== {name}: {vals.size} distinct outputs; first {vals[:4]}, last {vals[-3:]}")
    k = r * 256.0
    print("   outputs are multiples of 1/256:", np.allclose(k, np.round(k), atol=1e-9), " max level", k.max())
    # where do the steps happen, in units of the ideal weight t in [0, 1]?
    t = np.arange(r.size) / (r.size - 1)
    steps = np.nonzero(np.diff(np.round(k)))[0]
    frac_at_step = (t[steps + 1] * 256.0) % 1.0
    print("   step positions (t*256 mod 1): min %.4f max %.4f mean %.4f  (0.5 = round to nearest, 0 = floor)" %
          (frac_at_step.min(), frac_at_step.max(), frac_at_step.mean()))

# ---- values: texel pairs (a, b), weight i / 256
uvl, out = tests["values_x"]
t3 = tex["texdata3"].reshape(2, 512, 4)
k = np.repeat(np.arange(256), 257)
i = np.tile(np.arange(257), 256)
for ch in range(4):
    a = t3[0, 2 * k, ch].astype(np.int64)
    b = t3[0, 2 * k + 1, ch].astype(np.int64)
    got = out[:, ch].astype(np.float64)
    g16 = got * 65535.0
    print(f"== values ch{ch}: out*65535 integral: {np.allclose(g16, np.round(g16), atol=2e-3)}  max dev {np.abs(g16 - np.round(g16)).max():.4f}")
    g16 = np.round(g16).astype(np.int64)
    exact = (a * 257 * (256 - i) + b * 257 * i)  # / 256
    for nm, model in (("floor", exact // 256), ("rn", (exact + 128) // 256), ("rn8: ((a*(256-i)+b*i+128)>>8)*257", ((a * (256 - i) + b * i + 128) // 256) * 257)):
        print(f"   model {nm}: mismatches {int((model != g16).sum())} of {g16.size}")
    d = g16 * 256 - exact
    print("   out16*256 - exact: min", d.min(), "max", d.max())

# ---- bilinear on the random 64x64 chain
levels = []
off = 0
for l in range(4):
    w = 64 >> l
    levels.append(tex["texdata4"][off:off + w * w * 4].reshape(w, w, 4).astype(np.int64))
    off += w * w * 4


def bilinear16(level, u, v, mode, coord_bits=None):
    """returns out16 per channel (n, 4) under a model; u, v float32 arrays"""
    tl = levels[level]
    W = tl.shape[1]
    H = tl.shape[0]
    if coord_bits is None:
        x = u.astype(np.float64) * W - 0.5
        y = v.astype(np.float64) * H - 0.5
        xf, yf = np.floor(x), np.floor(y)
        a = np.floor((x - xf) * 256.0 + 0.5).astype(np.int64)
        b = np.floor((y - yf) * 256.0 + 0.5).astype(np.int64)
        x0, y0 = xf.astype(np.int64), yf.astype(np.int64)
    else:
        # fixed-point coordinate: u * W with coord_bits fractional bits, then -0.5, weight = RN to 8 bits
        s = 1 << coord_bits
        xi = np.floor(u.astype(np.float64) * W * s + 0.5).astype(np.int64) - s // 2
        yi = np.floor(v.astype(np.float64) * H * s + 0.5).astype(np.int64) - s // 2
        x0, y0 = xi >> coord_bits, yi >> coord_bits
        fx, fy = xi & (s - 1), yi & (s - 1)
        sh = coord_bits - 8
        a = (fx + (1 << (sh - 1))) >> sh if sh > 0 else fx
        b = (fy + (1 << (sh - 1))) >> sh if sh > 0 else fy
    carry_x, carry_y = a >> 8, b >> 8  # weight 256 = next texel with weight 0
    a, b = a & 255, b & 255
    x0, y0 = x0 + carry_x, y0 + carry_y
    x1, y1 = x0 + 1, y0 + 1
    x0, x1, y0, y1 = x0 % W, x1 % W, y0 % H, y1 % H
    t00, t10, t01, t11 = tl[y0, x0] * 257, tl[y0, x1] * 257, tl[y1, x0] * 257, tl[y1, x1] * 257
    a, b = a[:, None], b[:, None]
    if mode == "seq":
        top = (t00 * (256 - a) + t10 * a + 128) >> 8
        bot = (t01 * (256 - a) + t11 * a + 128) >> 8
        return (top * (256 - b) + bot * b + 128) >> 8
    if mode == "seq_y_first":
        l = (t00 * (256 - b) + t01 * b + 128) >> 8
        r = (t10 * (256 - b) + t11 * b + 128) >> 8
        return (l * (256 - a) + r * a + 128) >> 8
    s = t00 * (256 - a) * (256 - b) + t10 * a * (256 - b) + t01 * (256 - a) * b + t11 * a * b
    return (s + 32768) >> 16


uvl, out = tests["bilinear"]
g16 = np.round(out.astype(np.float64) * 65535.0).astype(np.int64)
print("== bilinear: out*65535 integral dev", np.abs(out.astype(np.float64) * 65535.0 - g16).max())
for mode in ("combined", "seq", "seq_y_first"):
    for cb in (None, 8, 12, 16, 20):
        m = bilinear16(0, uvl[:, 0], uvl[:, 1], mode, cb)
        bad = (m != g16).any(axis=1)
        print(f"   model {mode:12s} coord_bits {cb}: samples with a mismatch {int(bad.sum())} of {bad.size}, max |diff| {np.abs(m - g16).max()}")

m = bilinear16(0, uvl[:, 0], uvl[:, 1], "combined", None)
d = (m - g16)
print("diff histogram (combined, double coords):", np.unique(np.clip(np.abs(d), 0, 12), return_counts=True))
# look at channel 0 of a few samples in detail
W = 64
x = uvl[:, 0].astype(np.float64) * W - 0.5
y = uvl[:, 1].astype(np.float64) * W - 0.5
fa, fb = (x - np.floor(x)) * 256, (y - np.floor(y)) * 256
for i in range(6):
    print(f"  u={uvl[i,0]:.6f} v={uvl[i,1]:.6f} fa={fa[i]:.3f} fb={fb[i]:.3f} got {g16[i].tolist()} model {m[i].tolist()}")

# effective weights: fit (a, b) continuous per sample from the 4 channels
from scipy.optimize import least_squares
tl = levels[0]
x0 = np.floor(x).astype(np.int64) % 64
y0 = np.floor(y).astype(np.int64) % 64
x1, y1 = (x0 + 1) % 64, (y0 + 1) % 64
res = []
for i in range(400):
    t00, t10, t01, t11 = [tl[yy, xx].astype(np.float64) * 257 for yy, xx in ((y0[i], x0[i]), (y0[i], x1[i]), (y1[i], x0[i]), (y1[i], x1[i]))]
    def f(p):
        a, b = p
        return (t00 * (1 - a) * (1 - b) + t10 * a * (1 - b) + t01 * (1 - a) * b + t11 * a * b) - g16[i]
    sol = least_squares(f, [fa[i] / 256, fb[i] / 256])
    res.append((fa[i], sol.x[0] * 256, fb[i], sol.x[1] * 256, np.abs(sol.fun).max()))
res = np.array(res)
print("effective a vs ideal fa: mean |a_eff - fa| %.3f, mean |a_eff - round(fa)| %.3f" % (np.abs(res[:, 1] - res[:, 0]).mean(), np.abs(res[:, 1] - np.round(res[:, 0])).mean()))
print("effective b vs ideal fb: mean |b_eff - fb| %.3f, mean |b_eff - round(fb)| %.3f" % (np.abs(res[:, 3] - res[:, 2]).mean(), np.abs(res[:, 3] - np.round(res[:, 2])).mean()))
print("residual max", res[:, 4].max(), "median", np.median(res[:, 4]))
print(np.round(res[:8], 3))

print("---- neighbourhood search for the 2x2 footprint")
for i in range(8):
    best = None
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            xx0, yy0 = (x0[i] + dx) % 64, (y0[i] + dy) % 64
            xx1, yy1 = (xx0 + 1) % 64, (yy0 + 1) % 64
            t00, t10, t01, t11 = [tl[yy, xx].astype(np.float64) * 257 for yy, xx in ((yy0, xx0), (yy0, xx1), (yy1, xx0), (yy1, xx1))]
            def f(p):
                a, b = p
                return (t00 * (1 - a) * (1 - b) + t10 * a * (1 - b) + t01 * (1 - a) * b + t11 * a * b) - g16[i]
            sol = least_squares(f, [0.5, 0.5], bounds=([0, 0], [1, 1]))
            r = np.abs(sol.fun).max()
            if best is None or r < best[0]:
                best = (r, dx, dy, sol.x[0] * 256, sol.x[1] * 256)
    print(f"  sample {i}: u={uvl[i,0]:.5f} v={uvl[i,1]:.5f} fa={fa[i]:.2f} fb={fb[i]:.2f} best residual {best[0]:.2f} at offset ({best[1]},{best[2]}) a={best[3]:.2f} b={best[4]:.2f}")

print("---- four free weights per sample (exactly determined from the 4 channels)")
rows = []
for i in range(3000):
    T = np.stack([tl[yy, xx].astype(np.float64) * 257 for yy, xx in ((y0[i], x0[i]), (y0[i], x1[i]), (y1[i], x0[i]), (y1[i], x1[i]))], axis=1)  # (4 ch, 4 texels)
    if abs(np.linalg.det(T)) < 1e12:
        continue
    w = np.linalg.solve(T, g16[i].astype(np.float64))
    a8, b8 = np.floor(fa[i] + 0.5), np.floor(fb[i] + 0.5)
    ideal = np.array([(256 - a8) * (256 - b8), a8 * (256 - b8), (256 - a8) * b8, a8 * b8]) / 65536.0
    rows.append(np.concatenate([[fa[i], fb[i]], w * 256, ideal * 256, [np.linalg.cond(T)]]))
rows = np.array(rows)
good = rows[rows[:, -1] < 50]
print("well conditioned samples:", len(good))
np.set_printoptions(suppress=True, linewidth=200)
print(np.round(good[:12, :10], 3))
print("sum of effective weights*256: mean %.4f std %.4f" % (good[:, 2:6].sum(axis=1).mean(), good[:, 2:6].sum(axis=1).std()))
dev = good[:, 2:6] - good[:, 6:10]
print("effective - ideal (in 1/256 units): std", dev.std(axis=0), " max", np.abs(dev).max(axis=0))
fr = (good[:, 2:6] * 1) % 1
print("fractional part of effective weight*256 (histogram over 10 bins):", np.histogram(fr.reshape(-1), bins=10, range=(0, 1))[0])

print("---- rule R1: w11 = RN(a*b/256), w10 = a - w11, w01 = b - w11, w00 = 256 - a - b + w11")


def weights8(u, v, W, H):
    x = u.astype(np.float64) * W - 0.5
    y = v.astype(np.float64) * H - 0.5
    xf, yf = np.floor(x), np.floor(y)
    a = np.floor((x - xf) * 256.0 + 0.5).astype(np.int64)
    b = np.floor((y - yf) * 256.0 + 0.5).astype(np.int64)
    x0, y0 = xf.astype(np.int64) + (a >> 8), yf.astype(np.int64) + (b >> 8)
    return x0, y0, a & 255, b & 255


def bilinearR1(level, u, v):
    tl = levels[level]
    H, W = tl.shape[:2]
    x0, y0, a, b = weights8(u, v, W, H)
    x1, y1 = (x0 + 1) % W, (y0 + 1) % H
    x0, y0 = x0 % W, y0 % H
    w11 = (a * b + 128) >> 8
    w10, w01 = a - w11, b - w11
    w00 = 256 - a - b + w11
    s = (tl[y0, x0] * w00[:, None] + tl[y0, x1] * w10[:, None] + tl[y1, x0] * w01[:, None] + tl[y1, x1] * w11[:, None]) * 257
    return s  # 24-bit value: out16 = (s + 128) >> 8


s = bilinearR1(0, uvl[:, 0], uvl[:, 1])
m = (s + 128) >> 8
bad = (m != g16).any(axis=1)
print("bilinear R1: samples with a mismatch", int(bad.sum()), "of", bad.size, " max |diff|", np.abs(m - g16).max())
idx = np.nonzero(bad)[0][:6]
for i in idx:
    print(f"   u={uvl[i,0]!r} v={uvl[i,1]!r} fa={fa[i]:.4f} fb={fb[i]:.4f} got {g16[i].tolist()} model {m[i].tolist()}")
# float conversion
f32 = (m.astype(np.float32) / np.float32(65535.0)).astype(np.float32)
f64 = (m.astype(np.float64) / 65535.0).astype(np.float32)
mul = (m.astype(np.float32) * np.float32(1.0 / 65535.0)).astype(np.float32)
ok = ~bad
print("float conversion: out == f32(m)/65535f:", int((f32[ok] == out[ok]).all(axis=1).sum()), " == RN(m/65535):", int((f64[ok] == out[ok]).all(axis=1).sum()),
      " == m*(1/65535f):", int((mul[ok] == out[ok]).all(axis=1).sum()), "of", int(ok.sum()))

print("---- trilinear")
uvl, out = tests["trilinear"]
g16 = np.round(out.astype(np.float64) * 65535.0).astype(np.int64)
lod = np.clip(uvl[:, 2].astype(np.float64), 0.0, 3.0)
for gname, gam in (("floor", np.floor((lod - np.floor(lod)) * 256.0)), ("rn", np.floor((lod - np.floor(lod)) * 256.0 + 0.5))):
    gam = gam.astype(np.int64)
    l0 = np.floor(lod).astype(np.int64) + (gam >> 8)
    gam = gam & 255
    l1 = np.minimum(l0 + 1, 3)
    l0 = np.minimum(l0, 3)
    S0 = np.zeros((len(lod), 4), np.int64)
    S1 = np.zeros((len(lod), 4), np.int64)
    for l in range(4):
        sel = l0 == l
        S0[sel] = bilinearR1(l, uvl[sel, 0], uvl[sel, 1])
        sel = l1 == l
        S1[sel] = bilinearR1(l, uvl[sel, 0], uvl[sel, 1])
    g = gam[:, None]
    models = {
        "lerp of out16": ((((S0 + 128) >> 8) * (256 - g) + ((S1 + 128) >> 8) * g) + 128) >> 8,
        "lerp of 24-bit sums": (S0 * (256 - g) + S1 * g + 32768) >> 16,
    }
    for nm, m in models.items():
        bad = (m != g16).any(axis=1)
        print(f"   gamma {gname:5s} {nm:20s}: mismatching samples {int(bad.sum())} of {bad.size}, max |diff| {np.abs(m - g16).max()}")

gam = np.floor((lod - np.floor(lod)) * 256.0).astype(np.int64)
l0 = np.minimum(np.floor(lod).astype(np.int64), 3)
l1 = np.minimum(l0 + 1, 3)
L0 = np.zeros((len(lod), 4), np.int64)
L1 = np.zeros((len(lod), 4), np.int64)
for l in range(4):
    sel = l0 == l
    L0[sel] = (bilinearR1(l, uvl[sel, 0], uvl[sel, 1]) + 128) >> 8
    sel = l1 == l
    L1[sel] = (bilinearR1(l, uvl[sel, 0], uvl[sel, 1]) + 128) >> 8
inside = (uvl[:, 2] > 0.02) & (uvl[:, 2] < 2.98)
den = (L1 - L0).astype(np.float64)
geff = np.where(np.abs(den) > 3000, (g16 - L0) / np.where(den == 0, 1, den), np.nan) * 256.0
print("effective gamma*256 per channel vs floor(frac*256): first samples")
for i in np.nonzero(inside)[0][:10]:
    print(f"   lod={uvl[i,2]:.5f} frac*256={(lod[i]-np.floor(lod[i]))*256:.3f} geff={np.round(geff[i],3).tolist()}")
