"""TEST INFRASTRUCTURE.  Runs the reference's own compiled numba kernels on the GPU box.

``oracle/make_numba_ptx.py`` (build container, where /root/reference exists) lets
numba compile the UNMODIFIED ``pre/tsdf_numba.py`` kernels to PTX under
``oracle/_ref/``.  This module loads that PTX through the CUDA driver API (ctypes on
libcuda; the driver JIT-compiles numba's ``.target sm_90`` PTX for sm_100) and
launches the kernels with numba's kernel ABI, so the tests can state the tolerance
against the reference's compiled **float64** arithmetic and the
bench can time the reference's GPU path (``cal_tsdf_cuda``, pre/tsdf_numba.py:119-161)
on the same box.  Nothing here is product code; only ``tests/`` and ``bench.py``'s
baseline leg import it.

numba's ABI for an N-d array argument is the flattened struct
``(meminfo*, parent*, nitems i64, itemsize i64, data*, shape[N] i64, strides[N] i64)``;
scalars are passed by value (checked against the ``.param`` list of the PTX).
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

FOCAL = 241.42       # pre/tsdf_numba.py:8-10 (frozen into the PTX by numba)
VOXEL_RES = 32       # the params.py shim of make_numba_ptx.py
MM_THREADS = 128     # pre/tsdf_numba.py:12


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("manifest.json", "tsdf_kernel.ptx", "min_max_kernel.ptx"))


def _array_args(ptr, shape, itemsize=4):
    shape = [int(s) for s in shape]
    strides = []
    acc = itemsize
    for s in reversed(shape):
        strides.insert(0, acc)
        acc *= s
    nitems = int(np.prod(shape)) if shape else 1
    out = [ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int64(nitems), ctypes.c_int64(itemsize),
           ctypes.c_uint64(int(ptr))]
    out += [ctypes.c_int64(s) for s in shape]
    out += [ctypes.c_int64(s) for s in strides]
    return out


class NumbaPtxReference:
    """The reference's two kernels (compiled by numba, unmodified source) on one device."""

    def __init__(self, device=0):
        import torch
        if not available():
            raise RuntimeError("oracle/_ref PTX missing: run oracle/make_numba_ptx.py in the build container")
        if not torch.cuda.is_available():
            raise RuntimeError("NumbaPtxReference needs a CUDA device")
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)             # primary context current on this thread
        self.cu = ctypes.CDLL("libcuda.so.1")
        self._check(self.cu.cuInit(0), "cuInit")
        with open(os.path.join(REF_DIR, "manifest.json")) as f:
            self.manifest = json.load(f)
        self.fn = {}
        self._modules = []
        for name, info in self.manifest["kernels"].items():
            with open(os.path.join(REF_DIR, name + ".ptx"), "rb") as f:
                ptx = f.read() + b"\0"
            mod = ctypes.c_void_p()
            self._check(self.cu.cuModuleLoadData(ctypes.byref(mod), ptx), "cuModuleLoadData(%s)" % name)
            fn = ctypes.c_void_p()
            self._check(self.cu.cuModuleGetFunction(ctypes.byref(fn), mod, info["entry"].encode()),
                        "cuModuleGetFunction(%s)" % name)
            self._modules.append(mod)
            self.fn[name] = (fn, info["n_params"])

    @staticmethod
    def _check(rc, what):
        if rc != 0:
            raise RuntimeError("%s failed with CUresult %d" % (what, rc))

    def _launch(self, name, grid, block, args):
        fn, n_params = self.fn[name]
        assert len(args) == n_params, (name, len(args), n_params)
        ptrs = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
        stream = ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        self._check(self.cu.cuLaunchKernel(fn, grid[0], grid[1], 1, block, 1, 1, 0, stream, ptrs, None),
                    "cuLaunchKernel(%s)" % name)

    # --- the two kernels, argument for argument --------------------------------------------
    def tsdf_kernel(self, vox_ori, voxel_len, l, r, t, b, trunc_dis, d_depth, d_tsdf):
        """pre/tsdf_numba.py:15-72, launch shape of :154-156; tensors are CUDA float32."""
        R = VOXEL_RES
        assert d_tsdf.shape == (3, R, R, R) and d_tsdf.is_contiguous()
        args = (_array_args(vox_ori.data_ptr(), [3])
                + [ctypes.c_float(float(voxel_len))]
                + [ctypes.c_int32(int(v)) for v in (l, r, t, b)]
                + [ctypes.c_float(float(trunc_dis))]
                + _array_args(d_depth.data_ptr(), [d_depth.numel()])
                + _array_args(d_tsdf.data_ptr(), [3, R, R, R]))
        self._launch("tsdf_kernel", (R, R), 32, args)

    def min_max_kernel(self, l, r, t, b, d_depth, d_mm):
        """pre/tsdf_numba.py:75-116, launch shape of :135 (``d_depth`` padded to the grid)."""
        blocks = d_mm.shape[0]
        assert d_depth.numel() >= blocks * MM_THREADS
        args = ([ctypes.c_int32(int(v)) for v in (l, r, t, b)]
                + _array_args(d_depth.data_ptr(), [d_depth.numel()])
                + _array_args(d_mm.data_ptr(), [blocks, 6]))
        self._launch("min_max_kernel", (blocks, 1), MM_THREADS, args)

    # --- host wrapper: cal_tsdf_cuda with the two Python-3 crashes repaired -----------------
    def cal_tsdf_cuda(self, s, grid_from=None):
        """The steps of pre/tsdf_numba.py:119-161 — H2D, min_max_kernel, D2H, NumPy reduce,
        grid parameters, second H2D, device_array, tsdf_kernel, D2H — with an integer block
        count (:132 is a float under Python 3) and the depth zero-padded to the grid (the
        kernel has no bounds check, :86).  The compiled ``pos / b_w`` (:85) is Python-3 true
        division, so the y-extent differs from the Python-2 intent; pass ``grid_from=(max_l,
        mid_p)`` to voxelise on given grid parameters instead (what the parity test does).
        Returns (tsdf f32[3,R,R,R] (c,z,y,x), max_l, mid_p)."""
        torch = self.torch
        hdr = s["header"]
        l, t, r, b = (np.int32(hdr[2]), np.int32(hdr[3]), np.int32(hdr[4]), np.int32(hdr[5]))
        data = np.ascontiguousarray(s["data"], dtype=np.float32)
        blocks = (data.size + MM_THREADS - 1) // MM_THREADS
        padded = np.zeros(blocks * MM_THREADS, dtype=np.float32)
        padded[:data.size] = data
        d_depth = torch.from_numpy(padded).to(self.device)
        d_mm = torch.empty((blocks, 6), dtype=torch.float32, device=self.device)
        self.min_max_kernel(l, r, t, b, d_depth, d_mm)
        mm_p = d_mm.cpu().numpy()
        if grid_from is None:
            min_p = np.min(mm_p[:, :3], axis=0)
            max_p = np.max(mm_p[:, 3:6], axis=0)
            mid_p = (min_p + max_p) / 2
            max_l = np.max(max_p - min_p)
        else:
            max_l, mid_p = np.float32(grid_from[0]), np.asarray(grid_from[1], dtype=np.float32)
        voxel_len = max_l / VOXEL_RES
        trunc_dis = voxel_len * 3
        vox_ori = mid_p - max_l / 2 + voxel_len / 2
        assert vox_ori.dtype == np.float32 and type(trunc_dis) is np.float32
        d_depth = torch.from_numpy(padded).to(self.device)
        d_vox_ori = torch.from_numpy(np.ascontiguousarray(vox_ori)).to(self.device)
        d_tsdf = torch.empty((3, VOXEL_RES, VOXEL_RES, VOXEL_RES), dtype=torch.float32, device=self.device)
        self.tsdf_kernel(d_vox_ori, voxel_len, l, r, t, b, trunc_dis, d_depth, d_tsdf)
        tsdf = d_tsdf.cpu().numpy()
        return tsdf, max_l, mid_p
