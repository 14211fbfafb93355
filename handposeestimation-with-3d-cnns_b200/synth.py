"""Seeded synthetic MSRA-shaped depth frames.

An MSRA frame record is what ``read_bin`` returns in the reference
(/root/reference/pre/read_MSRA.py:155-164): ``header`` = int32[6]
``[img_w, img_h, left, top, right, bottom]`` (right/bottom exclusive) and
``depth`` = float32[bw*bh], row-major over the crop, millimetres, 0 = background.

There is no network and no MSRA data set in this environment, so every test,
fixture and benchmark uses these frames.  The generator is deterministic in
``seed`` and independent of the oracle and of the CUDA path.
"""
import numpy as np

IMG_W = 320
IMG_H = 240
# fixed-shape default used for single-number reporting:
# 120 x 130 crop = 15,600 px, about 8.8k valid pixels.
DEFAULT_BOX = (100, 60, 220, 190)


def make_frame(seed, box=None, key="depth"):
    """One synthetic frame record ``{'header': int32[6], key: float32[bw*bh]}``.

    ``box=None`` draws a ragged bounding box from the seed
    (left in [60,140), top in [30,100), bw in [80,150), bh in [80,140), clipped
    to the 320x240 image); ``box=(l, t, r, b)`` fixes it.  The hand is an
    ellipsoidal blob: inside the ellipse with semi-axes 0.4*bw, 0.45*bh the depth
    is ``400 - 60*sqrt(1 - rho^2) + N(0,1)`` mm, elsewhere 0.
    ``key`` is ``'depth'`` for process.py / tsdf_for.py and ``'data'`` for
    tsdf_numba.py (pre/tsdf_numba.py:132-133).
    """
    rng = np.random.default_rng(seed)
    if box is None:
        left = int(rng.integers(60, 140))
        top = int(rng.integers(30, 100))
        bw = int(rng.integers(80, 150))
        bh = int(rng.integers(80, 140))
        right = min(left + bw, IMG_W)
        bottom = min(top + bh, IMG_H)
    else:
        left, top, right, bottom = (int(v) for v in box)
    bw = right - left
    bh = bottom - top
    u = (np.arange(bw, dtype=np.float64) + 0.5 - bw / 2.0) / (0.4 * bw)
    v = (np.arange(bh, dtype=np.float64) + 0.5 - bh / 2.0) / (0.45 * bh)
    rho2 = v[:, None] ** 2 + u[None, :] ** 2
    inside = rho2 < 1.0
    noise = rng.standard_normal((bh, bw))
    d = 400.0 - 60.0 * np.sqrt(np.clip(1.0 - rho2, 0.0, None)) + noise
    depth = np.where(inside, d, 0.0).astype(np.float32).reshape(-1)
    header = np.array([IMG_W, IMG_H, left, top, right, bottom], dtype=np.int32)
    return {"header": header, key: depth}


def make_frames(seeds, box=None, key="depth"):
    """List of frame records for ``seeds`` (an int N means seeds 0..N-1)."""
    if isinstance(seeds, (int, np.integer)):
        seeds = range(int(seeds))
    return [make_frame(s, box=box, key=key) for s in seeds]


def make_uniform_batch(n_frames, seed0=0, box=DEFAULT_BOX):
    """Fast generator for large fixed-shape batches (bench.py).

    Returns ``(headers int32[N,6], depth float32[N, bw*bh])``.  Frame ``i`` is
    *not* bit-identical to ``make_frame(seed0 + i)`` (one RNG stream is used for
    the whole batch and the blob centre/scale jitter per frame), but has the same
    shape statistics: ellipsoidal blob around 400 mm, ~56 % valid pixels.
    """
    left, top, right, bottom = box
    bw, bh = right - left, bottom - top
    rng = np.random.default_rng(seed0)
    u = (np.arange(bw, dtype=np.float32) + 0.5 - bw / 2.0) / np.float32(0.4 * bw)
    v = (np.arange(bh, dtype=np.float32) + 0.5 - bh / 2.0) / np.float32(0.45 * bh)
    rho2 = (v[:, None] ** 2 + u[None, :] ** 2).astype(np.float32)
    inside = rho2 < 1.0
    bulge = (60.0 * np.sqrt(np.clip(1.0 - rho2, 0.0, None))).astype(np.float32)
    depth = np.empty((n_frames, bh * bw), dtype=np.float32)
    chunk = 256
    for i0 in range(0, n_frames, chunk):
        i1 = min(i0 + chunk, n_frames)
        m = i1 - i0
        base = rng.uniform(350.0, 450.0, size=(m, 1, 1)).astype(np.float32)
        scale = rng.uniform(0.8, 1.2, size=(m, 1, 1)).astype(np.float32)
        noise = rng.standard_normal((m, bh, bw), dtype=np.float32)
        d = base - scale * bulge[None] + noise
        d = np.where(inside[None], d, np.float32(0.0))
        depth[i0:i1] = d.reshape(m, -1)
    headers = np.tile(
        np.array([IMG_W, IMG_H, left, top, right, bottom], dtype=np.int32), (n_frames, 1)
    )
    return headers, depth
