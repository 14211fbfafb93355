"""CPU-side checks of bench.py's bookkeeping: the algorithmic bytes of SURVEY.md section 8d, the workload
labels, the ncu traffic scaling, and the reference arm's JSON line (run on this host's cores)."""
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_of_the_default_frame():
    hdr = np.array([[320, 240, 100, 60, 220, 190]] * 3, dtype=np.int32)
    assert bench.algorithmic_bytes(hdr, 32) == 3 * 455656           # 24 + 62,400 + 393,216 + 16
    assert bench.algorithmic_bytes(hdr[:1], 48) == 1389544 and bench.algorithmic_bytes(hdr[:1], 64) == 3208168


def test_workload_labels_follow_baseline_configs():
    mk = lambda **k: types.SimpleNamespace(**dict(dict(frames=8500, res=32, layout=0, ragged=False), **k))
    assert bench.workload_config(mk(), 1)["workload"].startswith("configs[1]")
    assert "configs[2]" in bench.workload_config(mk(frames=76375), 1)["workload"]
    assert "configs[2]" in bench.workload_config(mk(frames=9547), 8)["workload"]
    assert "resolution sweep" in bench.workload_config(mk(res=48), 1)["workload"]
    assert bench.workload_config(mk(ragged=True), 1)["workload"].startswith("ragged")


def test_ncu_traffic_is_scaled_to_the_run():
    full = bench.ncu_traffic(32, 0, 8500)
    assert full is not None and abs(full / bench.algorithmic_bytes(
        np.array([[320, 240, 100, 60, 220, 190]] * 8500, dtype=np.int32), 32) - 1.0) < 0.02   # traffic = algorithmic bytes
    assert abs(bench.ncu_traffic(32, 0, 4250) * 2 - full) < 1.0
    assert bench.ncu_traffic(32, 7, 8500) is None


def test_reference_arm_prints_one_json_line_with_all_host_threads():
    env = dict(os.environ, OMP_NUM_THREADS="1")             # what torchrun exports to its ranks
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--frames", "256"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == bench.host_threads()
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["value"] > 0
