"""The C oracle (oracle/dtsdf_oracle.c, also the CPU baseline of bench.py) against the golden
vectors of the unmodified reference and against the NumPy oracle.  Bit-exact."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import dtsdf_oracle as orc
from helpers import pack_simple, synth


def test_c_oracle_matches_for_loop_golden(forloop_golden):
    frames = [{"header": c["header"], "depth": c["depth"]} for c in forloop_golden]
    headers, offsets, packed = pack_simple(frames)
    aux = np.stack([np.concatenate([c["pmax"], c["pmin"]]) for c in forloop_golden]).astype(np.float32)
    vol, max_l, mid_p, _ = c_oracle.dtsdf_batch(headers, offsets, packed, bbox_mode=1, aux=aux, layout=0)
    for i, c in enumerate(forloop_golden):
        assert np.array_equal(vol[i], c["tsdf"])
        assert np.array_equal(np.signbit(vol[i]), np.signbit(c["tsdf"]))
        assert max_l[i] == c["max_l"] and np.array_equal(mid_p[i], c["mid_p"])


def test_c_oracle_matches_numba_golden(numba_golden):
    frames = [{"header": c["header"], "depth": c["depth"]} for c in numba_golden]
    headers, offsets, packed = pack_simple(frames)
    vol, max_l, mid_p, _ = c_oracle.dtsdf_batch(headers, offsets, packed, bbox_mode=0, layout=1)
    for i, c in enumerate(numba_golden):
        assert max_l[i] == c["max_l"] and np.array_equal(mid_p[i], c["mid_p"])
    for i, c in enumerate(numba_golden):
        if "tsdf" in c:
            assert np.array_equal(vol[i], c["tsdf"])


@pytest.mark.parametrize("R", [32, 48, 64])
@pytest.mark.parametrize("layout", [0, 1])
def test_c_oracle_matches_numpy_oracle(R, layout):
    frames = synth.make_frames(range(20, 26))
    headers, offsets, packed = pack_simple(frames)
    vol, max_l, mid_p, _ = c_oracle.dtsdf_batch(headers, offsets, packed, R=R, layout=layout, nthreads=2)
    for i, fr in enumerate(frames):
        v, ml, mp = orc.dtsdf_from_pixels(fr, R=R, layout=layout)
        assert np.array_equal(vol[i], v)
        assert np.array_equal(np.signbit(vol[i]), np.signbit(v))
        assert max_l[i] == ml and np.array_equal(mid_p[i], mp)


def test_c_oracle_grid_mode(forloop_golden):
    c = forloop_golden[0]
    vox_ori, voxel_len, trunc, _, _ = orc.grid_params(c["pmax"], c["pmin"])
    headers, offsets, packed = pack_simple([{"header": c["header"], "depth": c["depth"]}])
    aux = np.array([[vox_ori[0], vox_ori[1], vox_ori[2], voxel_len, trunc]], dtype=np.float32)
    vol, _, _, _ = c_oracle.dtsdf_batch(headers, offsets, packed, bbox_mode=2, aux=aux)
    assert np.array_equal(vol[0], c["tsdf"])


def test_identity_augmentation_specification_reduces_to_the_reference_path():
    """oracle_dtsdf_aug_frame (the specification of the augmented D-TSDF; parity with the reference
    unpinned) with A = identity, c = 0 is the pinned pixel-bbox path, bit for bit."""
    import helpers
    frames = helpers.synth.make_frames(range(40, 44))
    headers, offsets, packed = helpers.pack_simple(frames)
    ident = np.zeros((len(frames), 21), np.float32)
    ident[:, [0, 4, 8, 9, 13, 17]] = 1.0
    for layout in (0, 1):
        v, ml, mp = c_oracle.dtsdf_aug_batch(headers, offsets, packed, ident, layout=layout)
        v0, ml0, mp0, _ = c_oracle.dtsdf_batch(headers, offsets, packed, layout=layout)
        assert np.array_equal(v, v0) and np.array_equal(np.signbit(v), np.signbit(v0))
        assert np.array_equal(ml, ml0) and np.array_equal(mp, mp0)


def test_augmentation_specification_is_invariant_under_a_power_of_two_scale():
    """A size-independent property of the augmented D-TSDF's specification (oracle_dtsdf_aug_frame): with the map A
    replaced by 2A (Ainv by Ainv / 2, centre 0) every augmented point, the grid, the voxel length and the truncation
    double EXACTLY, the camera-frame voxel centres stay what they were - so the volume is the same bit for bit, max_l
    and mid_p double."""
    import helpers
    frames = helpers.synth.make_frames(range(60, 64))
    headers, offsets, packed = helpers.pack_simple(frames)
    rng = np.random.default_rng(5)
    aug = np.zeros((len(frames), 21), np.float32)
    for i in range(len(frames)):
        t = rng.uniform(-0.6, 0.6)
        A = np.diag([rng.uniform(0.7, 1.4), rng.uniform(0.7, 1.4), 1.0]) @ np.array(
            [[np.cos(t), np.sin(t), 0.0], [-np.sin(t), np.cos(t), 0.0], [0.0, 0.0, 1.0]])
        aug[i, :9] = A.astype(np.float32).reshape(-1)
        aug[i, 9:18] = np.linalg.inv(A).astype(np.float32).reshape(-1)
    aug2 = aug.copy()
    aug2[:, :9] *= np.float32(2.0)
    aug2[:, 9:18] *= np.float32(0.5)
    for layout in (0, 1):
        v, ml, mp = c_oracle.dtsdf_aug_batch(headers, offsets, packed, aug, layout=layout)
        v2, ml2, mp2 = c_oracle.dtsdf_aug_batch(headers, offsets, packed, aug2, layout=layout)
        assert np.count_nonzero(v) > 0
        assert np.array_equal(v, v2) and np.array_equal(np.signbit(v), np.signbit(v2))
        assert np.array_equal(ml2, np.float32(2.0) * ml) and np.array_equal(mp2, np.float32(2.0) * mp)
