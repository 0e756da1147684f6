"""Worker for tests/test_multi_rank_cpu.py: run under torch.distributed.run with the gloo backend.
Each rank produces the partial commitment/proof of its share of the Pianist rows with the ORACLE (CPU), the
ranks all_gather the 96-byte partials, rank 0 combines them with the product's zkp_g1_sum and checks the
aggregate against the golden vectors."""
import json
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bls12_381 as o  # noqa: E402
from zkp_subnet_b200 import native, sharding  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
rows = golden["pianist_4x16"]
lo, hi = sharding.shard_range(len(rows), rank, world)
mine_c = native.g1_sum(b"".join(bytes.fromhex(r["commitment"]) for r in rows[lo:hi]))
mine_p = native.g1_sum(b"".join(bytes.fromhex(r["proof"]) for r in rows[lo:hi]))
parts = sharding.gather_bytes(dist, mine_c + mine_p)
assert len(parts) == world and parts[rank] == mine_c + mine_p
# the uncompressed route (what bench.py uses: no square roots on the combining rank) must give the same sums
parts96 = sharding.gather_bytes(dist, sharding.expand_partials(mine_c + mine_p))
ok = True
if rank == 0:
    com, proof = sharding.combine_partials(parts)
    assert sharding.combine_expanded(parts96) == (com, proof)
    # identical rows: sum_i R_i(tau_y) = 1, so the aggregate equals the commitment / proof under the plain SRS
    ok = com.hex() == golden["B_eval_form"]["commitment"] and proof.hex() == golden["B_eval_form"]["proof"]
    print("COMBINE_OK" if ok else "COMBINE_BAD", flush=True)
# the host-to-host exchange bench.py uses for the per-step combine (POSIX shared memory; NCCL / gloo only for barriers):
# several steps in a row, rank 0 must see every rank's payload of THAT step
hx = sharding.HostExchange(rank, world, 192, "test_" + os.environ.get("MASTER_PORT", "0"))
dist.barrier()
hx.attach()
exp96 = sharding.expand_partials(mine_c + mine_p)
for step in range(1, 41):
    got = hx.gather(step, exp96[:184] + step.to_bytes(8, "little"))
    if rank == 0:
        assert [g[184:] for g in got] == [step.to_bytes(8, "little")] * world and got[rank][:184] == exp96[:184]
        if step == 40:
            ok = ok and [g[:184] for g in got] == [p[:184] for p in parts96]
    else:
        assert got is None
dist.barrier()
hx.close()
if rank == 0:
    print("EXCHANGE_OK" if ok else "EXCHANGE_BAD", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
