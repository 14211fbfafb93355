"""Offline model of shared-memory wavefronts of the pass-1 depth gather for candidate lane
mappings (dev tool; uses the oracle only to get the projection tables of synthetic frames)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dtsdf_oracle as orc
import __graft_entry__ as g
pkg = g.load_package()

def tables(frame, R=32):
    h = frame["header"]; d = frame["depth"]
    pmin, pmax = orc.bbox_from_pixels(h, d)
    vox_ori, vl, trunc, _, _ = orc.grid_params(pmax, pmin, R)
    l, t, r, b = (int(v) for v in h[2:6]); bw = r - l
    f32 = np.float32
    idx = np.arange(R, dtype=np.float32)
    vx = vox_ori[0] + idx * vl; vy = vox_ori[1] + idx * vl; vz = vox_ori[2] + idx * vl
    coeff = -f32(orc.FOCAL) / vz
    px = (vx[:, None] * coeff[None, :] + f32(160)).astype(np.int64)   # [x][z]
    py = (-vy[:, None] * coeff[None, :] + f32(120)).astype(np.int64)  # [y][z]
    okx = (px >= l) & (px < r); oky = (py >= t) & (py < b)
    col = np.where(okx, px - l, -1); row = np.where(oky, py - t, -1)
    return col, row, bw

def wavefronts(addr_words):
    """addr_words: int array [32] (word addresses, -1 = zero slot)"""
    a = np.where(addr_words < 0, -1, addr_words)
    banks = a % 32
    w = 0
    for bnk in np.unique(banks):
        w = max(w, len(np.unique(a[banks == bnk])))
    return w

def sim(frame, mapping, R=32):
    col, row, bw = tables(frame, R)
    tot = 0; n = 0
    for (xs, ys, zs) in mapping(R):      # each: arrays of 32 lanes' x, y, z0 (group start)
        for j in range(4):
            z = zs + j
            c = col[xs, z]; rr = row[ys, z]
            ok = (c >= 0) & (rr >= 0)
            addr = np.where(ok, rr * bw + c, -1)
            tot += wavefronts(addr); n += 1
    return tot / n

def m3(R):      # current: 4 consecutive x  x  8 z-groups, sweep y
    lane = np.arange(32); zg = lane % 8; xs_ = lane // 8
    for xq in range(R // 4):
        for y in range(R):
            yield (4 * xq + xs_, np.full(32, y), 4 * zg)

def m3s(stride):   # x strided
    def f(R):
        lane = np.arange(32); zg = lane % 8; xs_ = lane // 8
        nq = R // 4
        for xq in range(nq):
            for y in range(R):
                yield ((xq + nq * xs_) if stride == 8 else (xq % stride + stride * xs_ + (xq // stride) * 4 * stride), np.full(32, y), 4 * zg)
    return f

def m1(R):      # 4 consecutive y x 8 z-groups, sweep x
    lane = np.arange(32); zg = lane % 8; ys_ = lane // 8
    for yq in range(R // 4):
        for x in range(R):
            yield (np.full(32, x), 4 * yq + ys_, 4 * zg)

def m1s(R):     # y strided by 8
    lane = np.arange(32); zg = lane % 8; ys_ = lane // 8
    for yq in range(R // 4):
        for x in range(R):
            yield (np.full(32, x), yq + 8 * ys_, 4 * zg)

def mz(R):      # lane = z (1 voxel per lane, reference for conflicts): 32 z of one column
    lane = np.arange(32)
    for x in range(R):
        for y in range(0, R, 4):
            # emulate 4 voxels/lane as 4 different y's
            yield (np.full(32, x), np.full(32, y), lane - np.arange(4)[0] * 0)

if __name__ == "__main__":
    frames = [pkg.synth.make_frame(0, box=(100, 60, 220, 190))] + pkg.synth.make_frames(range(200, 206))
    for name, mp in [("m3 (4 consecutive x)", m3), ("m3 x stride 8", m3s(8)), ("m3 x stride 2", m3s(2)),
                     ("m1 (4 consecutive y)", m1), ("m1 y stride 8", m1s)]:
        print("%-24s" % name, " ".join("%.2f" % sim(f, mp) for f in frames))

def sim_pitch(frame, mapping, pitch_fn, R=32):
    col, row, bw = tables(frame, R)
    P = pitch_fn(bw)
    tot = 0; n = 0
    for (xs, ys, zs) in mapping(R):
        for j in range(4):
            z = zs + j
            c = col[xs, z]; rr = row[ys, z]
            ok = (c >= 0) & (rr >= 0)
            addr = np.where(ok, rr * P + c, -1)
            tot += wavefronts(addr); n += 1
    return tot / n

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "pitch":
    frames = [pkg.synth.make_frame(0, box=(100, 60, 220, 190))] + pkg.synth.make_frames(range(200, 204))
    for res in range(0, 32, 1):
        fn = lambda bw, res=res: bw + ((res - bw) % 32)
        print("pitch = %2d mod 32:" % res, " ".join("%.2f" % sim_pitch(f, m3, fn) for f in frames))
