"""Multi-GPU parity on real hardware (pytest -m gpu on a box with >= 2 GPUs; skipped otherwise): two ranks over NCCL run
tools/check_sharded.py -- decode_sharded (NCCL all-gather, contiguous tiles and pipelined bands) and decode_sharded_fused
(peer stores and NVSwitch multicast stores issued by the stage-B kernel itself) must assemble, on EVERY rank, an image
bit-identical to the single-GPU decode (SURVEY.md section 4 test 7), and the broadcast encoder hand-off must deliver the
same. bench.py asserts the same property inside every N>1 run (`assembled_bit_identical`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                                  reason="needs two GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_sharded_decode_is_bit_identical_on_every_rank(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "SHARDED_OK" in r.stdout and "SHARDED_MISMATCH" not in r.stdout
    assert r.stdout.count("bit-identical=True") >= 14 * world and "bit-identical=False" not in r.stdout
