"""Host expansion of the packed device->host volume format (csrc/dtsdf_decode.cpp) against a plain
restatement of the format, every output alignment, AVX-512 and baseline paths (no GPU needed)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRCS = [os.path.join(HERE, "native", "decode_check.cpp"),
        os.path.join(ROOT, "handposeestimation-with-3d-cnns_b200", "csrc", "dtsdf_decode.cpp")]


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("dec") / "decode_check")
    subprocess.run(["g++", "-O2", "-std=c++17"] + SRCS + ["-o", exe], check=True)
    return exe


@pytest.mark.parametrize("isa", ["native", "sse2"])
def test_expansion_matches_the_format(checker, isa):
    env = dict(os.environ)
    if isa == "sse2":
        env["DTSDF_DECODE_ISA"] = "sse2"
    out = subprocess.run([checker], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches=0" in out.stdout
    if isa == "sse2":
        assert "isa=sse2" in out.stdout
