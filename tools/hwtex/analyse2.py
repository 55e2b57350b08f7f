"""Single-texel trilinear weights from tools/hwtex/probe2.cu."""
import struct
import numpy as np

data = open("gpurun_out/hwtex_probe2.bin", "rb").read()
n = struct.unpack("i", data[:4])[0]
uvl = np.frombuffer(data[4:4 + n * 12], np.float32).reshape(n, 3)
oa = np.frombuffer(data[4 + n * 12:4 + n * 28], np.float32).reshape(n, 4)[:, 0].astype(np.float64)
ob = np.frombuffer(data[4 + n * 28:4 + n * 44], np.float32).reshape(n, 4)[:, 0].astype(np.float64)
a = np.round((uvl[:, 0].astype(np.float64) * 8 - 2.5) * 256).astype(np.int64)
b = np.round((uvl[:, 1].astype(np.float64) * 8 - 2.5) * 256).astype(np.int64)
g = np.round(uvl[:, 2].astype(np.float64) * 256).astype(np.int64)
va = oa * 65535.0  # = (w * 255 * 257 + 128) >> 8  -> w * 255.996
wa = va / (255 * 257 / 256.0)
wb = ob * 65535.0 / (255 * 257 / 256.0)
print("level-0 texel weight *256 integral:", np.abs(wa - np.round(wa)).max(), " level-1:", np.abs(wb - np.round(wb)).max())
wa, wb = np.round(wa).astype(np.int64), np.round(wb).astype(np.int64)
# level 0: texel (3,3) is the '11' corner: ideal weight a*b*(256-g) / 65536
w11 = (a * b + 128) >> 8
for nm, model in (("RN(w11*(256-g)/256)", (w11 * (256 - g) + 128) >> 8), ("floor", (w11 * (256 - g)) >> 8), ("RN(a*b*(256-g)/65536)", (a * b * (256 - g) + 32768) >> 16),
                  ("w11 - RN(w11*g/256)", w11 - ((w11 * g + 128) >> 8)), ("w11 - floor(w11*g/256)", w11 - ((w11 * g) >> 8)), ("w11 - ceil", w11 - ((w11 * g + 255) >> 8))):
    print(f"  level0 model {nm:28s} mismatches {int((model != wa).sum())} of {n}")
sel = (a == 128) & (b == 200)
print("a=128 b=200 (w11 = %d): g, weight:" % (((128 * 200 + 128) >> 8)), list(zip(g[sel][:12].tolist(), wa[sel][:12].tolist())))
for (aa, bb) in ((256, 256), (256, 128), (128, 128), (64, 192)):
    sel = (a == aa) & (b == bb)
    print(f"a={aa} b={bb}: (g, level-0 texel weight, level-1 texel weight):", list(zip(g[sel][:16].tolist(), wa[sel][:16].tolist(), wb[sel][:16].tolist())))
