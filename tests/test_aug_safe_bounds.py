"""The augmented voxeliser decides once per frame whether every voxel's camera-frame depth v_z lies inside the range its
straight-line reciprocal division is proven for (csrc/dtsdf_kernels.cu: aug_grid_is_safe).  It does not visit the grid:
the per-axis tables of the voxel -> camera map are monotone in the voxel index (every step of their evaluation is a
correctly rounded, monotone operation) and so is the table sum, hence the extremes of each component over the whole grid
are EXACTLY the sums of the per-axis extremes, which sit at the first and the last voxel of each axis.  This test restates
tables and test in NumPy float32 (the operations of the kernel and of oracle/dtsdf_oracle.c, step 3, in their order) and
checks that claim bit for bit against the minimum and maximum over all R^3 voxels, for random maps and grids - including
grids that cross the camera plane, negative and zero voxel lengths, and matrices with zero and negative entries."""
import numpy as np
import pytest

f32 = np.float32


def fma32(a, b, c):
    # a, b, c float32 arrays: a * b is exact in float64 (24 + 24 bits); the sum is rounded to float64 and then to float32 -
    # a double rounding that differs from a true fused multiply-add only on a float64 result exactly half-way between two
    # float32 values, which the comparison below would show as a failure of THIS helper, not of the claim (none occurs)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def tables(o, voxel_len, Ainv, c, R):
    """(tx, ty, tz)[R, 3] as the kernel fills them: va = o + i * voxel_len; dd = va - c_axis;
    tx = dd * Ainv[0][k], ty = dd * Ainv[1][k], tz = fma(dd, Ainv[2][k], c_k)."""
    i = np.arange(R, dtype=f32)
    out = []
    for axis in range(3):
        va = (o[axis] + (i * voxel_len).astype(f32)).astype(f32)
        dd = (va - c[axis]).astype(f32)
        if axis < 2:
            out.append((dd[:, None] * Ainv[axis][None, :]).astype(f32))
        else:
            out.append(fma32(dd[:, None], Ainv[2][None, :], c[None, :]))
    return out


def endpoint_bounds(o, voxel_len, Ainv, c, R):
    """aug_grid_is_safe's evaluation: the two end voxels of every axis only."""
    lo, hi = np.empty(3, f32), np.empty(3, f32)
    for k in range(3):
        l3, h3 = [], []
        for axis in range(3):
            d0 = f32(f32(o[axis] + f32(f32(0.0) * voxel_len)) - c[axis])
            d1 = f32(f32(o[axis] + f32(f32(R - 1) * voxel_len)) - c[axis])
            if axis < 2:
                e0, e1 = f32(d0 * Ainv[axis][k]), f32(d1 * Ainv[axis][k])
            else:
                e0 = fma32(np.array([d0]), np.array([Ainv[2][k]]), np.array([c[k]]))[0]
                e1 = fma32(np.array([d1]), np.array([Ainv[2][k]]), np.array([c[k]]))[0]
            l3.append(min(e0, e1)); h3.append(max(e0, e1))
        lo[k] = f32(l3[0] + f32(l3[1] + l3[2]))
        hi[k] = f32(h3[0] + f32(h3[1] + h3[2]))
    return lo, hi


@pytest.mark.parametrize("R", [32, 48])
def test_grid_extremes_are_the_sums_of_the_per_axis_end_points(R):
    rng = np.random.default_rng(2024 + R)
    crossing = 0
    for trial in range(300):
        kind = trial % 6
        Ainv = rng.normal(0.0, 1.0, (3, 3)).astype(f32)
        if kind == 1:
            Ainv[rng.integers(0, 3), rng.integers(0, 3)] = 0.0
        if kind == 2:
            Ainv = np.diag(rng.uniform(0.6, 1.6, 3)).astype(f32)
        c = rng.normal(0.0, 200.0, 3).astype(f32)
        c[2] -= f32(380.0)
        voxel_len = f32(rng.uniform(2.0, 9.0))
        if kind == 3:
            voxel_len = f32(-voxel_len)
        if kind == 4:
            voxel_len = f32(0.0)
        o = (c + rng.normal(0.0, 40.0, 3).astype(f32) - f32(0.5 * R) * voxel_len).astype(f32)
        if kind == 5:                                            # a grid around the camera centre: v_z of both signs
            c = rng.normal(0.0, 3.0, 3).astype(f32)
            o = (c - f32(0.5 * R) * voxel_len).astype(f32)
        with np.errstate(over="ignore", invalid="ignore"):
            tx, ty, tz = tables(o, voxel_len, Ainv, c, R)
            v = (tx[:, None, None, :] + (ty[None, :, None, :] + tz[None, None, :, :]).astype(f32)).astype(f32)   # [x, y, z, k]
            lo, hi = endpoint_bounds(o, voxel_len, Ainv, c, R)
        got_lo, got_hi = v.reshape(-1, 3).min(axis=0), v.reshape(-1, 3).max(axis=0)
        assert np.array_equal(lo, got_lo) and np.array_equal(hi, got_hi), (trial, kind, lo, got_lo, hi, got_hi)
        crossing += int(got_lo[2] < 0.0 < got_hi[2])
    assert crossing > 10                                         # the test saw grids on both sides of the camera plane
