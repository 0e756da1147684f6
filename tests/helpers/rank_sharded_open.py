"""Worker for tests/test_multi_rank_cpu.py (gloo, world size 2): the point-range-sharded commit + open
(zkp_subnet_b200.sharding.sharded_commit_open) with the per-rank device work replaced by the ORACLE, so that
the host-side protocol -- which bytes are gathered, zkp_shard_eval_combine, zkp_g1_sum -- is what is tested.
The result must equal the golden single-device vector B (commitment, eval, proof of TEST_POLY on 16 points)."""
import json
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bls12_381 as o  # noqa: E402
from zkp_subnet_b200 import sharding  # noqa: E402


class OracleShard:
    """What a zkp_ctx built by zkp_srs_generate_shard(tau, n, shard, log_shards) computes, in big-int Python."""

    def __init__(self, n, shard, world, tau):
        self.n, self.lo, self.hi = n, *sharding.shard_range(n, shard, world)
        self.srs = o.srs_lagrange(n, tau)[self.lo:self.hi]
        self.w = o.root_of_unity(n)

    def _vals(self, b):
        return [int.from_bytes(b[i:i + 32], "big") for i in range(0, len(b), 32)]

    def worker_commit(self, row, slice_be):
        return o.g1_compress(o.kzg_commit(self._vals(slice_be), self.srs))

    def shard_eval_partial(self, row, slice_be, x_be):
        x, acc = int.from_bytes(x_be, "big"), 0
        for k, f in enumerate(self._vals(slice_be)):
            wj = pow(self.w, self.lo + k, o.R)
            acc = (acc + f * wj * o.fr_inv((wj - x) % o.R)) % o.R
        return acc.to_bytes(32, "big")

    def shard_open_partial(self, row, slice_be, x_be, y_be):
        x, y = int.from_bytes(x_be, "big"), int.from_bytes(y_be, "big")
        q = [(f - y) * o.fr_inv((pow(self.w, self.lo + k, o.R) - x) % o.R) % o.R for k, f in enumerate(self._vals(slice_be))]
        return o.g1_compress(o.g1_msm_naive(self.srs, q))


dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
poly = [o.b64_decode(s) for s in golden["test_poly"]]
x = o.b64_decode(golden["test_point"])
ctx = OracleShard(16, rank, world, int(golden["tau_x"]))
com, y, proof = sharding.sharded_commit_open(dist, ctx, 0, b"".join(poly[ctx.lo:ctx.hi]), x, 4)
B = golden["B_eval_form"]
ok = (com.hex(), o.fr_to_b64(int.from_bytes(y, "big")), proof.hex()) == (B["commitment"], B["eval"], B["proof"])
print(f"rank {rank}: " + ("SHARDED_OPEN_OK" if ok else "SHARDED_OPEN_BAD"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
