import itertools, struct, sys
import numpy as np
data = open("gpurun_out/hwtex_probe2.bin", "rb").read()
n = struct.unpack("i", data[:4])[0]
uvl = np.frombuffer(data[4:4 + n * 12], np.float32).reshape(n, 3)
oa = np.frombuffer(data[4 + n * 12:4 + n * 28], np.float32).reshape(n, 4)[:, 0].astype(np.float64)
ob = np.frombuffer(data[4 + n * 28:4 + n * 44], np.float32).reshape(n, 4)[:, 0].astype(np.float64)
wa = np.round(oa * 65535.0 / (255 * 257 / 256.0)).astype(np.int64)
wb = np.round(ob * 65535.0 / (255 * 257 / 256.0)).astype(np.int64)
g = np.floor(uvl[:, 2].astype(np.float64) * 256).astype(np.int64)


def marg(u, W):
    x = u.astype(np.float64) * W - 0.5
    xf = np.floor(x)
    A = np.floor((x - xf) * 256.0 + 0.5).astype(np.int64)
    x0 = xf.astype(np.int64) + (A >> 8)
    return x0, A & 255


rn = lambda v: (v + 128) >> 8
for nm, W, tx, ty, target, zlevel in (("A(level0 texel 3,3)", 8, 3, 3, wa, 0), ("B(level1 texel 1,1)", 4, 1, 1, wb, 1)):
    x0, A = marg(uvl[:, 0], W)
    y0, B = marg(uvl[:, 1], W)
    ix = tx - x0  # 0 or 1: which corner the marked texel is (else weight 0)
    iy = ty - y0
    inside = (ix >= 0) & (ix <= 1) & (iy >= 0) & (iy <= 1)
    results = []
    for px, py, pz, nest in itertools.product((0, 1), (0, 1), (0, 1), ("xy_z", "xz_y", "yz_x")):
        Px = A if px else 256 - A      # primary marginal in x: weight of corner 1 (px=1) or corner 0
        Py = B if py else 256 - B
        Gz = g if zlevel == 1 else 256 - g  # weight of this level
        Pz = Gz if pz else 256 - Gz
        if nest == "xy_z":
            Pxy = rn(Px * Py); Pxz = rn(Px * Pz); Pyz = rn(Py * Pz); Pxyz = rn(Pxy * Pz)
        elif nest == "xz_y":
            Pxz = rn(Px * Pz); Pxy = rn(Px * Py); Pyz = rn(Py * Pz); Pxyz = rn(Pxz * Py)
        else:
            Pyz = rn(Py * Pz); Pxy = rn(Px * Py); Pxz = rn(Px * Pz); Pxyz = rn(Pyz * Px)
        # cell (i, j, k) where i = 1 means "x at primary corner", etc.  inclusion-exclusion
        def cell(i, j, k):
            # weight with x in primary set if i else complement ...
            t = 0
            for si in ((1,) if i else (0, 1)):
                for sj in ((1,) if j else (0, 1)):
                    for sk in ((1,) if k else (0, 1)):
                        sign = (-1) ** ((0 if i else si) + (0 if j else sj) + (0 if k else sk))
                        term = {(0, 0, 0): 256, (1, 0, 0): Px, (0, 1, 0): Py, (0, 0, 1): Pz, (1, 1, 0): Pxy, (1, 0, 1): Pxz, (0, 1, 1): Pyz, (1, 1, 1): Pxyz}[(si, sj, sk)]
                        t = t + sign * term
            return t
        # the marked texel: corner ix (1 = x1 texel).  primary corner in x is corner 1 if px else corner 0
        i_sel = np.where(ix == 1, px, 1 - px)
        j_sel = np.where(iy == 1, py, 1 - py)
        k_sel = pz  # this level is the primary z cell iff pz == 1
        w = np.zeros(n, np.int64)
        for i in (0, 1):
            for j in (0, 1):
                sel = inside & (i_sel == i) & (j_sel == j)
                if sel.any():
                    w[sel] = cell(i, j, k_sel)[sel]
        results.append((int((w != target).sum()), px, py, pz, nest))
    results.sort()
    print(nm, results[:6])

x0, A = marg(uvl[:, 0], 4)
y0, B = marg(uvl[:, 1], 4)
a_in = np.round((uvl[:, 0].astype(np.float64) * 8 - 2.5) * 256).astype(np.int64)
b_in = np.round((uvl[:, 1].astype(np.float64) * 8 - 2.5) * 256).astype(np.int64)
for (aa, bb) in ((0, 0), (64, 64), (200, 40), (256, 128)):
    sel = (a_in == aa) & (b_in == bb)
    AG = rn(A[sel] * g[sel]); BG = rn(B[sel] * g[sel]); ABG = rn(AG * B[sel])
    ix = 1 - x0[sel]; iy = 1 - y0[sel]
    print(f"a={aa} b={bb}: level-1 x0={x0[sel][0]} A1={A[sel][0]} y0={y0[sel][0]} B1={B[sel][0]}  texel(1,1) is corner ({ix[0]},{iy[0]})")
    print("   g     :", g[sel][:20].tolist())
    print("   got   :", wb[sel][:20].tolist())
    ideal = np.where(ix == 1, A[sel], 256 - A[sel]) * np.where(iy == 1, B[sel], 256 - B[sel]) * g[sel] / 65536.0
    print("   ideal :", np.round(ideal[:20], 2).tolist())
