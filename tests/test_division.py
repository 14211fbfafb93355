"""The kernel replaces x/focal and x/trunc by FMA sequences with a precomputed reciprocal
(csrc/dtsdf_kernels.cu: div_fixed).  This checks them against IEEE division on the CPU:
exhaustively for the default focal over the dividend range the kernel uses, and on a large
pseudo-random sample for variable divisors."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "division_check.c")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("div") / "division_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", SRC, "-o", exe, "-lm"], check=True)
    return exe


def test_one_correction_division_by_default_focal_is_exact(checker):
    # the kernel's range is 1 <= |d| < 2^40 (biased exponents 127..166); checked with margin
    out = subprocess.run([checker, "focal", "120", "175"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches=0" in out.stdout


def test_two_correction_division_random(checker):
    out = subprocess.run([checker, "random", "400000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches=0" in out.stdout


def test_float64_division_by_a_fixed_divisor(tmp_path):
    """point_cloud_kernel's X = (u d) / f and Y in float64: reciprocal multiplication + two residual corrections
    (ddiv_fixed) equals IEEE division on 2 x 10^8 dividends, half of them with the quotient next to a rounding
    boundary, for the default focal and 258 other divisors (tests/native/ddivision_check.c)."""
    exe = str(tmp_path / "ddivision_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", os.path.join(HERE, "native", "ddivision_check.c"),
                    "-o", exe, "-lm"], check=True)
    out = subprocess.run([exe, "200000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches=0" in out.stdout


@pytest.mark.gpu
def test_reciprocal_divisions_of_the_augmented_voxeliser_on_the_gpu(tmp_path):
    """coeff = -focal / v_z of the augmented voxeliser (csrc/dtsdf_kernels.cu: aug_div_coeff) inside the range its
    per-frame test guarantees: the reciprocal forms equal __fdiv_rn for EVERY divisor with 2^-60 <= |z| <= 2^60 -
    the five-FMA form for the default focal and 64 other dividends, the three-FMA form for the default focal.
    Runs on the GPU: the seed (MUFU.RCP) is a hardware approximation (tests/native/rcp_division_check.cu)."""
    import __graft_entry__ as g
    exe = g.RCP_CHECK
    if not os.path.exists(exe):
        exe = str(tmp_path / "rcp_division_check")
        subprocess.run([g._find("nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false",
                        os.path.join(HERE, "native", "rcp_division_check.cu"), "-o", exe], check=True)
    out = subprocess.run([exe, "64"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "five_fma mismatches=0 " in out.stdout and "three_fma mismatches=0 " in out.stdout
