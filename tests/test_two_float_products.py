"""The sampled voxeliser (csrc/dtsdf_kernels.cu: sampled_bbox_fast) keeps the extremes of u*d and v*d over the sampled
pixels as TWO floats each - hi = RN32(u d), lo = fma(u, d, -hi) - and compares them as 64-bit integers built from the
ordered bits of (hi, lo).  This checks the arithmetic it rests on, in NumPy, against the float64 products the
reference forms (pre/process.py:52-59): the pair holds the product exactly, the integer order is the order of the
float64 products (-0 below +0), and the value rebuilt from a key is the float64 product bit for bit."""
import numpy as np


def _fkey(x):
    b = x.view(np.int32).astype(np.int64)
    return np.where(b < 0, b ^ 0x7fffffff, b)                      # ordered bits of a float32 (as int64, range of int32)


def _key2(hi, lo):
    # (fkey(hi) << 32) | (unsigned)(fkey(lo) ^ 0x80000000): lexicographic (hi, lo) as one signed 64-bit integer
    return (_fkey(hi) << 32) | ((_fkey(lo) & 0xffffffff) ^ 0x80000000)


def _dkey(x):
    b = x.view(np.int64)
    return np.where(b < 0, b ^ 0x7fffffffffffffff, b)              # the kernel's dkey: ordered bits of a float64


def test_two_float_products_are_exact_and_ordered():
    rng = np.random.default_rng(31)
    n = 400000
    # u: integers and half-integers below 2^20 (crop and image coordinates), most of them image-sized; zeros included
    u = np.where(rng.random(n) < 0.9, rng.integers(-640, 641, n) * 0.5, rng.integers(-(1 << 21), 1 << 21, n) * 0.5)
    u[rng.integers(0, n, 2000)] = 0.0
    # d: any finite float32 with 2^-90 <= |d| (what the fast path admits), both signs, a share of plain depths
    bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    d = bits.view(np.float32).copy()
    plain = rng.random(n) < 0.5
    d[plain] = (rng.uniform(1.0, 2000.0, n).astype(np.float32) * np.where(rng.random(n) < 0.2, -1, 1).astype(np.float32))[plain]
    ok = np.isfinite(d) & (np.abs(d) >= np.float32(2.0 ** -90)) & (np.abs(d) < np.float32(2.0 ** 100))
    u, d = u[ok].astype(np.float32), d[ok]
    assert u.size > 300000
    exact = u.astype(np.float64) * d.astype(np.float64)           # 22 x 24 significant bits: exact in float64,
    with np.errstate(over="ignore"):                               # and what RN64(u * d) of the reference is
        hi = exact.astype(np.float32)                              # RN32(u d)
    assert np.isfinite(hi).all()
    resid = exact - hi.astype(np.float64)                          # exact (Sterbenz-like: below an ulp of hi)
    lo = resid.astype(np.float32)                                  # what fma(u, d, -hi) returns, if representable ...
    assert np.array_equal(lo.astype(np.float64), resid)            # ... and it is: the pair holds the product
    back = np.where(lo == 0, hi.astype(np.float64), hi.astype(np.float64) + lo.astype(np.float64))
    assert np.array_equal(back.view(np.int64), exact.view(np.int64))          # sign of a zero product included
    # order: integer keys of the pairs against the ordered bits of the float64 products
    k2, kd = _key2(hi, lo), _dkey(exact)
    a, b = rng.integers(0, u.size, 600000), rng.integers(0, u.size, 600000)
    assert np.array_equal(k2[a] < k2[b], kd[a] < kd[b]) and np.array_equal(k2[a] == k2[b], kd[a] == kd[b])
    # -0 sorts below +0 in both
    z = np.array([-0.0, 0.0], dtype=np.float32)
    assert _key2(z, np.zeros(2, np.float32))[0] < _key2(z, np.zeros(2, np.float32))[1]
    assert _dkey(z.astype(np.float64))[0] < _dkey(z.astype(np.float64))[1]
    # extremes picked by key are the extremes of the float64 products
    assert exact[np.argmax(k2)] == exact.max() and exact[np.argmin(k2)] == exact.min()
