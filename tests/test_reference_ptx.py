"""oracle/_ref: what numba compiles from the reference's unmodified kernels (oracle/make_numba_ptx.py).
CPU-only facts about that PTX; running it is tests/test_gpu_numba_compiled.py."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def manifest():
    path = os.path.join(REF, "manifest.json")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference; __graft_entry__.build() makes it)")
    return json.load(open(path))


def _entry_params(ptx):
    body = ptx[ptx.index(".entry"):]
    return re.findall(r"\.param \.(\w+)", body[:body.index(")")])


def test_kernel_abi_matches_the_launcher(manifest):
    # oracle/numba_ptx_ref.py packs numba's array ABI by hand: (meminfo, parent, nitems, itemsize, data,
    # shape[N], strides[N]) per array, scalars by value
    ptx = open(os.path.join(REF, "tsdf_kernel.ptx")).read()
    p = _entry_params(ptx)
    assert len(p) == manifest["kernels"]["tsdf_kernel"]["n_params"] == 7 + 1 + 4 + 1 + 7 + 13
    assert p[7] == "f32" and p[8:12] == ["u32"] * 4 and p[12] == "f32"      # voxel_len, l r t b, trunc_dis
    ptx = open(os.path.join(REF, "min_max_kernel.ptx")).read()
    p = _entry_params(ptx)
    assert len(p) == manifest["kernels"]["min_max_kernel"]["n_params"] == 4 + 7 + 9
    assert p[:4] == ["u32"] * 4


def test_compiled_reference_arithmetic_is_float64(manifest):
    # int32 * float32 is typed float64 by numba: the voxel maths of tsdf_kernel runs in f64 with f32 loads
    # and stores - the reason the GPU tolerance test exists beside the bit-exact ones
    ptx = open(os.path.join(REF, "tsdf_kernel.ptx")).read()
    assert ptx.count("div.rn.f64") >= 5 and "cvt.f64.f32" in ptx and "cvt.rn.f32.f64" in ptx
    assert "st.global.f32" in ptx and "st.global.f64" not in ptx
    assert re.search(r"\.target sm_\d+", ptx)
    # min_max_kernel: `pos / b_w` compiled under Python 3 is a float64 division (pre/tsdf_numba.py:85)
    mm = open(os.path.join(REF, "min_max_kernel.ptx")).read()
    assert "div.rn.f64" in mm
