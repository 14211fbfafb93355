#!/usr/bin/env python
"""Freeze golden vectors by EXECUTING THE UNMODIFIED REFERENCE in this container.

Run from the repo root (needs /root/reference, so it only runs in the build
container, never on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/forloop.npz   pre/process.py DataProcess.process() and
                             pre/tsdf_for.py tsdf_f() on seeded synthetic frames
  tests/golden/numba_sim.npz pre/tsdf_numba.py min_max_kernel + tsdf_kernel run
                             under numba's CUDA simulator

The reference has no tests or vectors of its own, so these are the
known answers the oracle (oracle/dtsdf_oracle.py) is pinned to.

Notes on how the numba file is driven:
  * pre/tsdf_numba.py:3 does ``from params import *`` and the reference ships no
    params.py, so a one-line module ``VOXEL_RES = 32`` is injected.
  * ``cal_tsdf_cuda`` is Python-2 host code (float block count, :132-134) and
    cannot run; the two kernels are launched directly with the launch shapes of
    :135 and :154-156.
  * ``min_max_kernel`` computes the pixel row as ``pos / b_w`` (:85), floor
    division in Python 2.  To execute the unmodified kernel source with the
    semantics it was written for, ``cuda.grid`` is wrapped so ``pos`` is an int
    subclass whose ``/`` floors.  The depth is zero-padded to a multiple of 128
    (otherwise the kernel reads out of bounds, :86).
"""
import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/pre"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "handposeestimation-with-3d-cnns_b200"))

import synth  # noqa: E402  (package-local synthetic frame generator)


def forloop_golden():
    import process      # /root/reference/pre/process.py, unmodified
    import tsdf_for     # /root/reference/pre/tsdf_for.py, unmodified

    cases = [
        # (frame seed, box, point_num, keep_pc)
        (0, synth.DEFAULT_BOX, 6000, True),
        (1, None, 6000, False),
        (2, None, 6000, False),
        (3, (130, 90, 190, 150), 6000, True),     # 60x60 crop: fewer than 6000 valid points
        (4, (0, 0, 96, 80), 6000, False),          # crop touching the image corner
        (5, (200, 120, 320, 240), 2000, False),    # crop touching the far corner, other point_num
    ]
    out = {}
    for k, (seed, box, point_num, keep_pc) in enumerate(cases):
        frame = synth.make_frame(seed, box=box)
        gt = np.zeros(63, dtype=np.float32)
        np.random.seed(1000 + seed)
        dp = process.DataProcess(frame, gt, point_num, aug=False)
        pc_all = dp.point_cloud()
        np.random.seed(1000 + seed)
        pc, tsdf, max_l, mid_p = process.DataProcess(frame, gt, point_num, aug=False).process()
        tsdf2, max_l2, mid_p2 = tsdf_for.tsdf_f(frame, pc)
        assert np.array_equal(tsdf, tsdf2) and max_l == max_l2 and np.array_equal(mid_p, mid_p2)
        pmax, pmin = tsdf_for.max_min_point(pc)
        assert tsdf.dtype == np.float64 and np.array_equal(tsdf, tsdf.astype(np.float32))
        p = "c%d_" % k
        out[p + "seed"] = np.int64(seed)
        out[p + "np_random_seed"] = np.int64(1000 + seed)
        out[p + "point_num"] = np.int64(point_num)
        out[p + "header"] = frame["header"]
        out[p + "depth"] = frame["depth"]
        out[p + "n_valid"] = np.int64(pc_all.shape[0])
        out[p + "pc_all_head"] = pc_all[:64].copy()
        out[p + "pc_all_sum"] = pc_all.sum(axis=0)
        if keep_pc:
            out[p + "pc"] = pc
        out[p + "pc_head"] = pc[:64].copy()
        out[p + "pmax"] = pmax
        out[p + "pmin"] = pmin
        out[p + "tsdf"] = tsdf.astype(np.float32)      # lossless (asserted above), [c,x,y,z]
        out[p + "max_l"] = np.float32(max_l)
        out[p + "mid_p"] = np.asarray(mid_p, dtype=np.float32)
        print("forloop case", k, "seed", seed, "n_valid", pc_all.shape[0], "max_l", max_l,
              "nonzero voxels", int(np.count_nonzero(tsdf[0])))
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "forloop.npz"), **out)


class _Py2Int(int):
    """int whose true division floors, as ``/`` did for ints in Python 2."""

    def __truediv__(self, other):
        return int(self) // other


def numba_sim_golden():
    params = types.ModuleType("params")
    params.VOXEL_RES = 32
    sys.modules["params"] = params
    import tsdf_numba   # /root/reference/pre/tsdf_numba.py, unmodified source
    from numba import cuda

    # run min_max_kernel with Python-2 integer division for `pos / b_w`
    # (the simulator swaps a FakeCUDAModule into the kernel's globals at launch)
    import numba.cuda.simulator.kernelapi as kernelapi
    sim_cuda = kernelapi.FakeCUDAModule
    orig_grid = sim_cuda.grid

    def py2_grid(self, n):
        return _Py2Int(orig_grid(self, n))

    out = {}
    mm_cases = [(10, (140, 100, 180, 136)), (11, (150, 90, 197, 131)),   # 40x36, 47x41 (ragged tail)
                (12, (96, 70, 160, 125))]                               # 64x55, off-centre
    for k, (seed, box) in enumerate(mm_cases):
        frame = synth.make_frame(seed, box=box, key="data")
        l, t, r, b = (int(v) for v in frame["header"][2:6])
        n = frame["data"].size
        nblk = (n + tsdf_numba.mm_threadsperblock - 1) // tsdf_numba.mm_threadsperblock
        padded = np.zeros(nblk * tsdf_numba.mm_threadsperblock, dtype=np.float32)
        padded[:n] = frame["data"]
        d_mm = np.zeros([nblk, 6], dtype=np.float32)
        sim_cuda.grid = py2_grid
        try:
            tsdf_numba.min_max_kernel[nblk, tsdf_numba.mm_threadsperblock](l, r, t, b, padded, d_mm)
        finally:
            sim_cuda.grid = orig_grid
        mm_p = d_mm
        min_p = np.min(mm_p[:, :3], axis=0)       # pre/tsdf_numba.py:140
        max_p = np.max(mm_p[:, 3:6], axis=0)      # :141
        mid_p = (min_p + max_p) / 2               # :142-147
        len_e = max_p - min_p
        max_l = np.max(len_e)
        voxel_len = max_l / params.VOXEL_RES
        trunc_dis = (voxel_len * 3)
        vox_ori = mid_p - max_l / 2 + voxel_len / 2
        p = "mm%d_" % k
        out[p + "header"] = frame["header"]
        out[p + "depth"] = frame["data"]
        out[p + "min_p"] = min_p
        out[p + "max_p"] = max_p
        out[p + "mid_p"] = mid_p
        out[p + "max_l"] = np.float32(max_l)
        out[p + "voxel_len"] = np.float32(voxel_len)
        out[p + "trunc"] = np.float32(trunc_dis)
        out[p + "vox_ori"] = vox_ori
        print("min_max case", k, min_p, max_p, type(voxel_len), vox_ori.dtype)
        if True:     # every case also goes through tsdf_kernel (about 12 s each in the simulator)
            tsdf = np.zeros([3, 32, 32, 32], dtype=np.float32)
            tsdf_numba.tsdf_kernel[[32, 32], tsdf_numba.threadsperblock](
                vox_ori, voxel_len, l, r, t, b, trunc_dis, frame["data"], tsdf)
            out[p + "tsdf"] = tsdf                # [c,z,y,x]
            print("tsdf_kernel case", k, "nonzero voxels", int(np.count_nonzero(tsdf[0])))
    out["n_mm_cases"] = np.int64(len(mm_cases))
    np.savez_compressed(os.path.join(HERE, "numba_sim.npz"), **out)


if __name__ == "__main__":
    forloop_golden()
    numba_sim_golden()
