"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the e2e byte accounting."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "HR query pixels/s" and d["unit"] == "px/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "c3" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c3" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_e2e_h2d_accounting():
    sys.path.insert(0, ROOT)
    import bench
    # one rank uploads the whole map; eight ranks upload their row tiles + the 3x3 halo (a little more in total)
    assert bench.h2d_bytes_all_ranks(339, 510, 1356, 1, 1) == 339 * 510 * 64 * 4
    total8 = bench.h2d_bytes_all_ranks(339, 510, 1356, 8, 1)
    assert 339 * 510 * 64 * 4 < total8 < 1.1 * 339 * 510 * 64 * 4
