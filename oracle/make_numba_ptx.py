#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Recipe for ``oracle/_ref``: the reference's own GPU kernels, compiled.

The reference is Python; its only native code is what numba generates from
``pre/tsdf_numba.py`` (``tsdf_kernel`` :15-72, ``min_max_kernel`` :75-116).  This
script imports that file UNMODIFIED from where it lies under /root/reference (with
the one-line ``params.py`` the reference forgot to ship), lets numba
compile both kernels to PTX for the argument types of the reference's own call
sites (:135, :154-156) and writes

    oracle/_ref/tsdf_kernel.ptx
    oracle/_ref/min_max_kernel.ptx
    oracle/_ref/manifest.json       entry names, argument lists, numba version

``oracle/_ref`` is git-ignored (no reference-derived artefact enters the history)
but travels to the GPU box, where ``oracle/numba_ptx_ref.py`` loads the PTX with the
driver API (the driver JIT-compiles numba's ``.target sm_90`` PTX for sm_100) — the
compiled-f64 arithmetic the north star pins the tolerance statement to.  No GPU is
needed here: ``cuda.compile_ptx`` only runs NVVM.

Argument types (what ``numba.typeof`` gives the reference's call-site values under
Python 3 / NumPy 2):
  tsdf_kernel(vox_ori f32[::1], voxel_len f32, l i32, r i32, t i32, b i32,
              trunc_dis f32, i f32[::1], o f32[:,:,:,::1])
  min_max_kernel(l i32, r i32, t i32, b i32, a f32[::1], o f32[:,::1])
``x * voxel_len`` is int32 * float32, which numba types as float64: that is where
the f64 arithmetic of the compiled kernel comes from.

Run from the repo root:  python oracle/make_numba_ptx.py
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/pre"


def entry_params(ptx):
    """Number of parameters of the (single) .entry of a PTX module: its list ends at the first ')'."""
    body = ptx[ptx.index(".entry"):]
    return body[:body.index(")")].count(".param .")


def main():
    if not os.path.isdir(REF):
        print("make_numba_ptx: %s absent, nothing to do" % REF)
        return 1
    os.environ.pop("NUMBA_ENABLE_CUDASIM", None)
    import numba
    from numba import cuda, float32, int32

    params = types.ModuleType("params")
    params.VOXEL_RES = 32
    sys.modules["params"] = params
    sys.path.insert(0, REF)
    import tsdf_numba  # /root/reference/pre/tsdf_numba.py, unmodified

    f1 = float32[::1]
    sigs = {
        "tsdf_kernel": (f1, float32, int32, int32, int32, int32, float32, f1, float32[:, :, :, ::1]),
        "min_max_kernel": (int32, int32, int32, int32, f1, float32[:, ::1]),
    }
    os.makedirs(OUT, exist_ok=True)
    manifest = {"numba": numba.__version__, "source": "pre/tsdf_numba.py (unmodified)",
                "VOXEL_RES": 32, "kernels": {}}
    for name, sig in sigs.items():
        pyfunc = getattr(tsdf_numba, name).py_func
        ptx, _ = cuda.compile_ptx(pyfunc, sig, cc=(9, 0), fastmath=False, opt=True)
        entry = [ln.split()[-1].rstrip("(") for ln in ptx.splitlines() if ".entry" in ln]
        assert len(entry) == 1, entry
        with open(os.path.join(OUT, name + ".ptx"), "w") as f:
            f.write(ptx)
        manifest["kernels"][name] = {"entry": entry[0], "signature": [str(t) for t in sig],
                                     "n_params": entry_params(ptx)}
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("make_numba_ptx: wrote", sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
