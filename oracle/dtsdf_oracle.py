"""CPU ORACLE (test infrastructure, not product code).

A NumPy restatement of the reference's depth-crop -> projective directional TSDF
(D-TSDF) path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker.  The product path
(``handposeestimation-with-3d-cnns_b200/``) never imports it.

Parity is PINNED: ``tests/golden/make_golden.py`` executes the unmodified
reference (``/root/reference/pre/tsdf_for.py``, ``pre/process.py`` and the
``pre/tsdf_numba.py`` kernels under numba's CUDA simulator) on seeded synthetic
frames and freezes inputs + outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them
bit-for-bit.

Arithmetic model ("f32 step-wise"): under NumPy >= 2 (NEP 50) every
intermediate of the reference's voxel loop is an ``np.float32`` scalar, i.e. each
``+ - * / sqrt`` is one IEEE-754 binary32 round-to-nearest operation with no FMA
contraction.  The vectorised functions below keep that order of operations on
float32 arrays; ``tsdf_cal_loop`` is the scalar restatement used to cross-check
them on small cases.  The point cloud of ``process.py`` is float64 (float32 array
+ np.int32 scalar promotes), restated as such.

All line citations are into /root/reference/.
"""
import numpy as np

FOCAL = 241.42       # pre/process.py:7, pre/tsdf_for.py:3, pre/tsdf_numba.py:8
CENTER_X = 160       # pre/process.py:133, pre/tsdf_numba.py:9
CENTER_Y = 120       # pre/process.py:134, pre/tsdf_numba.py:10
VOXEL_RES = 32       # pre/process.py:88 (the missing params.VOXEL_RES of pre/tsdf_numba.py:3)
TRUNC_MUL = 3        # pre/process.py:95, pre/tsdf_numba.py:146
SENTINEL = 99999.0   # pre/tsdf_numba.py:88-89

LAYOUT_CXYZ = 0      # for-loop layout tsdf_v[c, x, y, z]   (pre/process.py:196-198)
LAYOUT_CZYX = 1      # numba layout    o[c, z, y, x]        (pre/tsdf_numba.py:70-72)

_f32 = np.float32


# --------------------------------------------------------------------------- #
# a3/a4/a5: point cloud, resampling, per-axis extrema  (process.py semantics)  #
# --------------------------------------------------------------------------- #
def point_cloud(header, depth):
    """Back-projection of every crop pixel, dropping all-zero rows.

    Follows pre/process.py:30-68: ``X = ((u + left) - w/2) * d / f``,
    ``Y = -(((v + top) - h/2) * d) / f``, ``Z = -d`` in float64 (the float32
    index vector plus an np.int32 header scalar promotes to float64, the product
    with the float32 depth is then exact in float64 and the division by the
    Python float 241.42 is one float64 division).  Rows with X == Y == Z == 0 are
    dropped (:61-66), which keeps exactly the pixels with ``d != 0`` in row-major
    order.  Returns float64[n_valid, 3].
    """
    img_w, img_h, left, top, right, bottom = (int(v) for v in header[:6])
    bw, bh = right - left, bottom - top
    d = np.asarray(depth, dtype=np.float32).reshape(bh, bw).astype(np.float64)
    u = np.arange(bw, dtype=np.float64) + float(left) - (img_w / 2)
    v = np.arange(bh, dtype=np.float64) + float(top) - (img_h / 2)
    x = (u[None, :] * d) / FOCAL
    y = -(v[:, None] * d) / FOCAL
    z = -d
    pts = np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], axis=1)
    keep = np.any(pts != 0, axis=1) | np.any(np.isnan(pts), axis=1)
    return pts[keep]


def resample_indices(n_valid, point_num=6000):
    """Index vector of ``set_length`` (pre/process.py:70-84).

    Consumes the *global* ``np.random`` stream exactly like the reference: with
    fewer than ``point_num`` points all are kept and the tail is
    ``randint(0, n_valid, point_num - n_valid)`` (:75-78), otherwise
    ``randint(0, n_valid, point_num)`` with replacement (:80).
    """
    if n_valid < point_num:
        idx = np.arange(0, point_num, 1, dtype=np.int32)
        idx[n_valid:] = np.random.randint(0, n_valid, size=point_num - n_valid)
    else:
        idx = np.random.randint(0, n_valid, size=point_num)
    return idx


def set_length(hand_points, point_num=6000):
    """pre/process.py:70-84 (== cut_version/pre/process.py:52-67)."""
    return hand_points[resample_indices(hand_points.shape[0], point_num), :]


def max_min_point(pc):
    """Per-axis (max, min) rounded to float32; z ignores exact zeros.

    pre/process.py:102-120 == pre/tsdf_for.py:23-41.  Note the (max, min) order.
    """
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    z = z[z != 0]
    pmax = np.array([np.max(x), np.max(y), np.max(z)]).astype(np.float32)
    pmin = np.array([np.min(x), np.min(y), np.min(z)]).astype(np.float32)
    return pmax, pmin


# --------------------------------------------------------------------------- #
# a8: bounding box over all valid pixels  (tsdf_numba.py semantics)            #
# --------------------------------------------------------------------------- #
def bbox_from_pixels(header, depth, focal=FOCAL, cx=CENTER_X, cy=CENTER_Y):
    """min/max of the back-projected valid pixels, float32.

    Follows pre/tsdf_numba.py:83-96 per pixel and :98-116,140-141 for the
    reduction, with Python-2 integer semantics for the row (``pos / b_w`` is
    floor division in the language the file was written for):
    ``q = d/f``, ``X = q*(x-cx)``, ``Y = -q*(y-cy)``, ``Z = -d``, a pixel is
    invalid when ``abs(d) < 1`` and then contributes the +-99999 sentinels.
    Specified deviation from the reference's undefined tail behaviour
    (out-of-bounds reads in the last block, :86,138-139): the sentinels always
    take part in the reduction, as they do whenever the crop is zero-padded to a
    multiple of 128.  Returns (min_p f32[3], max_p f32[3]).
    """
    _, _, left, top, right, bottom = (int(v) for v in header[:6])
    bw, bh = right - left, bottom - top
    d = np.asarray(depth, dtype=np.float32).reshape(bh, bw)
    xs = (np.arange(bw, dtype=np.int64) + left).astype(np.float32) - _f32(cx)
    ys = (np.arange(bh, dtype=np.int64) + top).astype(np.float32) - _f32(cy)
    valid = ~(np.abs(d) < 1)
    q = d / _f32(focal)
    cam_x = q * xs[None, :]
    cam_y = (-q) * ys[:, None]
    cam_z = -d
    s = _f32(SENTINEL)
    mins, maxs = [], []
    for c in (cam_x, cam_y, cam_z):
        v = c[valid]
        mins.append(min(_f32(v.min()), s) if v.size else s)
        maxs.append(max(_f32(v.max()), -s) if v.size else -s)
    return np.array(mins, dtype=np.float32), np.array(maxs, dtype=np.float32)


# --------------------------------------------------------------------------- #
# a6: grid parameters                                                          #
# --------------------------------------------------------------------------- #
def grid_params(pmax, pmin, R=VOXEL_RES, trunc_mul=TRUNC_MUL):
    """Cube of side ``max_l`` centred on the bounding box.

    pre/process.py:91-96 == pre/tsdf_for.py:11-16 == pre/tsdf_numba.py:142-147,
    all float32: ``mid = (max+min)/2``, ``max_l = max(max-min)``,
    ``voxel_len = max_l/R``, ``trunc = voxel_len*3``,
    ``vox_ori = mid - max_l/2 + voxel_len/2`` (centre of voxel (0,0,0)).
    Returns (vox_ori f32[3], voxel_len f32, trunc f32, max_l f32, mid f32[3]).
    """
    pmax = np.asarray(pmax, dtype=np.float32)
    pmin = np.asarray(pmin, dtype=np.float32)
    mid = (pmax + pmin) / _f32(2)
    len_e = pmax - pmin
    max_l = np.max(len_e)
    voxel_len = max_l / _f32(R)
    trunc = voxel_len * _f32(trunc_mul)
    vox_ori = mid - max_l / _f32(2) + voxel_len / _f32(2)
    return vox_ori.astype(np.float32), _f32(voxel_len), _f32(trunc), _f32(max_l), mid


# --------------------------------------------------------------------------- #
# a7/a9: per-voxel projective D-TSDF                                           #
# --------------------------------------------------------------------------- #
def voxel_indices(header, vox_ori, voxel_len, R=VOXEL_RES, focal=FOCAL, cx=CENTER_X, cy=CENTER_Y):
    """Integer part of the voxel path: (pix_x, pix_y, inside) for the R^3 grid.

    pre/process.py:150-159 == pre/tsdf_numba.py:26-32,36.  Arrays are indexed
    [x, y, z]; ``pix`` uses truncation toward zero (Python ``int()``).  Voxels
    whose projection is not finite count as outside.
    """
    _, _, left, top, right, bottom = (int(v) for v in header[:6])
    k = np.arange(R, dtype=np.float32)
    vox_ori = np.asarray(vox_ori, dtype=np.float32)
    voxel_len = _f32(voxel_len)
    with np.errstate(all="ignore"):
        vx = (vox_ori[0] + k * voxel_len)[:, None, None]
        vy = (vox_ori[1] + k * voxel_len)[None, :, None]
        vz = (vox_ori[2] + k * voxel_len)[None, None, :]
        coeff = _f32(-focal) / vz
        fx = np.broadcast_to(vx * coeff + _f32(cx), (R, R, R))
        fy = np.broadcast_to(-vy * coeff + _f32(cy), (R, R, R))
        finite = np.isfinite(fx) & np.isfinite(fy) & (np.abs(fx) < 2e9) & (np.abs(fy) < 2e9)
        px = np.where(finite, fx, _f32(-1e9)).astype(np.int64)
        py = np.where(finite, fy, _f32(-1e9)).astype(np.int64)
    inside = finite & (px >= left) & (px < right) & (py >= top) & (py < bottom)
    return px, py, inside, (vx, vy, vz)


def tsdf_cal(header, depth, vox_ori, voxel_len, trunc, R=VOXEL_RES, layout=LAYOUT_CXYZ,
             focal=FOCAL, cx=CENTER_X, cy=CENTER_Y):
    """Vectorised per-voxel D-TSDF, float32[3, R, R, R].

    pre/process.py:140-198 == pre/tsdf_for.py:62-120 (layout 0, [c,x,y,z]) and
    pre/tsdf_numba.py:15-72 (layout 1, [c,z,y,x]); same arithmetic, f32 step-wise:
    project the voxel centre to one pixel (:155-157), skip outside the crop (:159)
    or where ``abs(d) < 1`` (:167), back-project that pixel (:172-175), per-axis
    ``|v-w|/trunc`` (:177-179), Euclidean truncation (:180-189), negate when the
    voxel is behind the surface (:191-194).  Skipped voxels stay 0.
    """
    _, _, left, top, right, bottom = (int(v) for v in header[:6])
    bw = right - left
    depth = np.asarray(depth, dtype=np.float32).reshape(-1)
    trunc = _f32(trunc)
    px, py, inside, (vx, vy, vz) = voxel_indices(header, vox_ori, voxel_len, R, focal, cx, cy)
    idx = np.where(inside, (py - top) * bw + (px - left), 0)
    d = depth[idx]
    valid = inside & ~(np.abs(d) < 1)
    with np.errstate(all="ignore"):
        c1 = d / _f32(focal)
        wx = ((px.astype(np.float32)) - _f32(cx)) * c1
        wy = (-((py.astype(np.float32)) - _f32(cy))) * c1
        wz = -d
        tx = np.abs(vx - wx) / trunc
        ty = np.abs(vy - wy) / trunc
        tz = np.abs(vz - wz) / trunc
        dist = np.sqrt(tx * tx + ty * ty + tz * tz)
        far = dist > 1
        one = _f32(1)
        out = []
        behind = wz > vz
        for t in (tx, ty, tz):
            t = np.where(far, one, t)
            t = np.where(one < t, one, t)           # python min(t, 1)
            t = np.where(behind, -t, t)
            out.append(np.where(valid, t, _f32(0)).astype(np.float32))
    vol = np.stack(out, axis=0)                      # [c, x, y, z]
    if layout == LAYOUT_CZYX:
        vol = np.ascontiguousarray(vol.transpose(0, 3, 2, 1))
    return vol


def tsdf_cal_loop(header, depth, vox_ori, voxel_len, trunc, R=VOXEL_RES,
                  focal=FOCAL, cx=CENTER_X, cy=CENTER_Y):
    """Scalar restatement of the voxel loop (small cases only), layout [c,x,y,z].

    Same statements as pre/process.py:140-198 on np.float32 scalars; used to
    cross-check the vectorised ``tsdf_cal``.
    """
    _, _, left, top, right, bottom = (int(v) for v in header[:6])
    bw = right - left
    depth = np.asarray(depth, dtype=np.float32).reshape(-1)
    f = _f32(focal)
    vox_ori = np.asarray(vox_ori, dtype=np.float32)
    voxel_len = _f32(voxel_len)
    trunc = _f32(trunc)
    one = _f32(1)
    out = np.zeros((3, R, R, R), dtype=np.float32)
    for x in range(R):
        vx = vox_ori[0] + _f32(x) * voxel_len
        for y in range(R):
            vy = vox_ori[1] + _f32(y) * voxel_len
            for z in range(R):
                vz = vox_ori[2] + _f32(z) * voxel_len
                coeff = -f / vz
                px = int(vx * coeff + _f32(cx))
                py = int(-vy * coeff + _f32(cy))
                if px < left or px >= right or py < top or py >= bottom:
                    continue
                d = depth[(py - top) * bw + px - left]
                if abs(d) < 1:
                    continue
                c1 = d / f
                wx = (_f32(px) - _f32(cx)) * c1
                wy = -(_f32(py) - _f32(cy)) * c1
                wz = -d
                tx = abs(vx - wx) / trunc
                ty = abs(vy - wy) / trunc
                tz = abs(vz - wz) / trunc
                if np.sqrt(tx * tx + ty * ty + tz * tz) > 1:
                    tx = ty = tz = one
                tx, ty, tz = min(tx, one), min(ty, one), min(tz, one)
                if wz > vz:
                    tx, ty, tz = -tx, -ty, -tz
                out[0, x, y, z] = tx
                out[1, x, y, z] = ty
                out[2, x, y, z] = tz
    return out


# --------------------------------------------------------------------------- #
# whole-path entries                                                           #
# --------------------------------------------------------------------------- #
def dtsdf_from_points(frame, pc, R=VOXEL_RES, layout=LAYOUT_CXYZ):
    """``tsdf_for.tsdf_f(data, point_cloud)`` / ``DataProcess.tsdf_f(pc)``.

    pre/tsdf_for.py:6-20, pre/process.py:86-100.  Returns (tsdf f32, max_l, mid_p).
    """
    depth = frame["depth"] if "depth" in frame else frame["data"]
    pmax, pmin = max_min_point(pc)
    vox_ori, voxel_len, trunc, max_l, mid = grid_params(pmax, pmin, R)
    vol = tsdf_cal(frame["header"], depth, vox_ori, voxel_len, trunc, R, layout)
    return vol, max_l, mid


def dtsdf_from_pixels(frame, R=VOXEL_RES, layout=LAYOUT_CZYX):
    """``tsdf_numba.cal_tsdf_cuda(s)`` (pre/tsdf_numba.py:119-161).

    Bounding box over all valid pixels, then the same grid parameters and voxel
    kernel.  Returns (tsdf f32[3,R,R,R] in ``layout``, max_l f32, mid_p f32[3]).
    """
    depth = frame["data"] if "data" in frame else frame["depth"]
    pmin, pmax = bbox_from_pixels(frame["header"], depth)
    vox_ori, voxel_len, trunc, max_l, mid = grid_params(pmax, pmin, R)
    vol = tsdf_cal(frame["header"], depth, vox_ori, voxel_len, trunc, R, layout)
    return vol, max_l, mid


def process_frame(frame, point_num=6000, R=VOXEL_RES):
    """``DataProcess(frame, gt, point_num).process()`` without augmentation.

    pre/process.py:13-28: point cloud -> resample (global np.random stream) ->
    grid parameters from the sampled cloud -> voxel loop.  Returns
    [point_clouds f64[point_num,3], tsdf [3,R,R,R] (c,x,y,z), max_l, mid_p].
    """
    pts = point_cloud(frame["header"], frame["depth"])
    pc = set_length(pts, point_num)
    vol, max_l, mid = dtsdf_from_points(frame, pc, R, LAYOUT_CXYZ)
    return [pc, vol, max_l, mid]


def normalize_joints(ground_truth, max_l, mid_p):
    """``(gt - mid_p) / max_l + 0.5`` per frame (pre/joint_nor.py:8-18)."""
    gt = np.asarray(ground_truth)
    shp = gt.shape
    g = gt.reshape(-1, 21, 3)
    out = np.empty(g.shape)
    for i in range(g.shape[0]):
        out[i] = (g[i] - mid_p[i]) / max_l[i] + 0.5
    return out.reshape(shp)
