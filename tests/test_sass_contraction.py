"""What the bit-exact kernels rely on, checked on the compiler that builds them (no GPU needed):

ptxas (CUDA 12.9, sm_100a) CONTRACTS the packed forms - `mul.rn.f32x2` followed by `add.rn.f32x2` becomes one
FFMA2, although each carries an explicit rounding modifier and the build passes -fmad=false - while the scalar
`mul.rn.f32` / `add.rn.f32` pair is kept apart, as PTX guarantees.  csrc/dtsdf_kernels.cu therefore uses the packed
FP32 pipe only where no product can reach an addition (eval_pair, the augmented bounding box) and scalar
operations for every product that is then added to.  If a later toolkit stops contracting, the first assertion
fails and the restriction can go; if it starts contracting the scalar forms, the second fails and the kernels are
no longer safe."""
import os
import re
import shutil
import subprocess
import tempfile

import pytest

SRC = r"""
#include <cuda_runtime.h>
extern "C" __global__ void k_packed(const float2* a, const float2* b, const float2* c, float2* o) {
    int i = threadIdx.x;
    o[i] = __fadd2_rn(__fmul2_rn(a[i], b[i]), c[i]);
}
extern "C" __global__ void k_scalar(const float2* a, const float2* b, const float2* c, float2* o) {
    int i = threadIdx.x;
    float2 A = a[i], B = b[i], C = c[i];
    o[i] = make_float2(__fadd_rn(__fmul_rn(A.x, B.x), C.x), __fadd_rn(__fmul_rn(A.y, B.y), C.y));
}
"""


def _nvcc():
    return shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)


@pytest.mark.skipif(_nvcc() is None, reason="nvcc not found")
def test_ptxas_contracts_packed_but_not_scalar_rn_operations():
    nvcc = _nvcc()
    cuobjdump = os.path.join(os.path.dirname(nvcc), "cuobjdump")
    with tempfile.TemporaryDirectory() as d:
        src, cubin = os.path.join(d, "t.cu"), os.path.join(d, "t.cubin")
        with open(src, "w") as f:
            f.write(SRC)
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-cubin",
                        "-o", cubin, src], check=True)
        sass = subprocess.run([cuobjdump, "-sass", cubin], check=True, capture_output=True, text=True).stdout
    ops = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\w+)", line)
        if m:
            name = m.group(1)
            ops[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            ops[name].append(m.group(1).split(".")[0])
    packed, scalar = ops["k_packed"], ops["k_scalar"]
    # the packed pair is fused into one FFMA2 (one rounding instead of two)
    assert "FFMA2" in packed and "FMUL2" not in packed and "FADD2" not in packed, packed
    # the scalar pair stays two instructions, two roundings
    assert scalar.count("FMUL") == 2 and scalar.count("FADD") == 2, scalar
    assert not any(o in ("FFMA", "FFMA2", "FMUL2", "FADD2") for o in scalar), scalar
