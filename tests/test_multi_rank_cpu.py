"""CPU: the N > 1 path (one process per rank, gloo, world size 2): gather of the per-rank partial points
and their combination with zkp_g1_sum -- the only cross-GPU step of the hot path."""
import os
import subprocess
import sys

import pytest

from zkp_subnet_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range():
    assert [sharding.shard_range(1 << 20, r, 8) for r in (0, 7)] == [(0, 1 << 17), (7 << 17, 1 << 20)]
    with pytest.raises(ValueError):
        sharding.shard_range(10, 0, 4)


def test_two_rank_combine_gloo():
    port = 29500 + os.getpid() % 500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "helpers", "rank_combine.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "COMBINE_OK" in out.stdout
