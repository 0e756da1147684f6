"""Host-side verification cost (pairing) per response and per batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
ctx = native.Context(0)
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, 10, 3)
x = ctx.random_point(1)
items = []
for i in range(8):
    f = ctx.random_poly(50 + i, 1 << 10)
    com, y, proof = ctx.worker_commit_open(i, f, x)
    items.append((i, proof, y, com))
t = time.perf_counter()
for _ in range(3):
    for i, p, y, c in items:
        assert ctx.worker_verify(i, p, x, y, c)
print(f"worker_verify: {(time.perf_counter() - t) / 24 * 1e3:.3f} ms per response")
for reps in (8, 32, 128):
    idx = [it[0] for it in items] * (reps // 8)
    pr = b"".join(it[1] for it in items) * (reps // 8)
    ev = b"".join(it[2] for it in items) * (reps // 8)
    cm = b"".join(it[3] for it in items) * (reps // 8)
    t = time.perf_counter()
    ok = ctx.worker_verify_batch(idx, pr, x, ev, cm)
    dt = time.perf_counter() - t
    assert all(ok)
    print(f"worker_verify_batch of {reps}: {dt * 1e3:.2f} ms = {dt / reps * 1e3:.3f} ms per response")
# two responses of a batch of 32 carry each other's (valid-looking, wrong) proof: the combined check fails and every
# response is verified on its own
bad = bytearray(pr[:48 * 32]); bad[48 * 5:48 * 6], bad[48 * 6:48 * 7] = pr[48 * 6:48 * 7], pr[48 * 5:48 * 6]
t = time.perf_counter()
ok = ctx.worker_verify_batch(idx[:32], bytes(bad), x, ev[:32 * 32], cm[:48 * 32])
dt = time.perf_counter() - t
assert [k for k, v in enumerate(ok) if not v] == [5, 6]
print(f"worker_verify_batch of 32 with two swapped proofs (fallback to single checks): {dt * 1e3:.2f} ms")
