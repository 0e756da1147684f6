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


def run_two_ranks(script: str, port_base: int):
    port = port_base + os.getpid() % 500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "helpers", script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_two_rank_combine_gloo():
    out = run_two_ranks("rank_combine.py", 29500)
    assert "COMBINE_OK" in out and "EXCHANGE_OK" in out


def test_two_rank_sharded_open_gloo():
    """Point-range-sharded commit + open of one polynomial over 2 ranks: both ranks end with the golden
    single-device commitment, evaluation and proof."""
    out = run_two_ranks("rank_sharded_open.py", 30100)
    assert out.count("SHARDED_OPEN_OK") == 2, out


def test_shard_eval_combine_host():
    """zkp_shard_eval_combine (host arithmetic) against the oracle's barycentric evaluation."""
    from oracle import bls12_381 as o
    from zkp_subnet_b200 import native
    import random
    rng = random.Random(5)
    n, G = 64, 4
    f = [rng.randrange(o.R) for _ in range(n)]
    x = rng.randrange(o.R)
    w = o.root_of_unity(n)
    parts = []
    for g in range(G):
        acc = 0
        for j in range(g * n // G, (g + 1) * n // G):
            wj = pow(w, j, o.R)
            acc = (acc + f[j] * wj * o.fr_inv((wj - x) % o.R)) % o.R
        parts.append(acc.to_bytes(32, "big"))
    y = native.shard_eval_combine(b"".join(parts), 6, x.to_bytes(32, "big"))
    assert int.from_bytes(y, "big") == o.eval_from_evals(f, x)
    with pytest.raises(native.ZkpError):
        native.shard_eval_combine((o.R).to_bytes(32, "big"), 6, x.to_bytes(32, "big"))
