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
    # the reference ITSELF (oracle/_ref, placed by oracle/make_ref.py at build time) where /root/reference or a previous
    # build provided it; the oracle's port only as the labelled fallback
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "src", "models", "components", "diinn.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "c3" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c3" in d["config"]["workload"]
    # the reference arm runs on OUR arm's config: the same object, whatever the arm, the precision or N
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config("c3") and set(d["config"]) == {"workload", "io_dtype", "l2"}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_sample_is_the_workload_cropped():
    sys.path.insert(0, ROOT)
    import bench
    B, h, W, hu, W_up = bench.cpu_sample_shape("c3")
    assert (B, W, W_up) == (1, 510, 2040) and hu == 4 * h and 150_000 <= B * hu * W_up <= 250_000


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="build container only")
def test_oracle_ref_recipe_copies_the_reference_unmodified():
    import hashlib
    from oracle import make_ref
    assert make_ref.make_ref()
    man = json.load(open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")))
    for rel, sha in man["sha256"].items():
        assert hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest() == sha
    assert make_ref.import_reference_decoder().__name__ == "ImplicitDecoder"
    # git-ignored (reference sources never enter the history), not gpurun-ignored (travels to the GPU box)
    assert "oracle/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()
    gi = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gi) or "oracle/_ref" not in open(gi).read()


def test_e2e_h2d_accounting():
    sys.path.insert(0, ROOT)
    import bench
    # one rank uploads the whole map; eight ranks upload their row tiles + the 3x3 halo (a little more in total)
    assert bench.h2d_bytes_all_ranks(339, 510, 1356, 1, 1) == 339 * 510 * 64 * 4
    total8 = bench.h2d_bytes_all_ranks(339, 510, 1356, 8, 1)
    assert 339 * 510 * 64 * 4 < total8 < 1.1 * 339 * 510 * 64 * 4
