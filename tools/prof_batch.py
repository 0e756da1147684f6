"""Small driver for ncu: one batch of K commit+open requests at 2^LOG_N through zkp_worker_commit_open_batch (one
grouped launch set), after one warm batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ctx = native.Context(0)
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 2)
ctx.prebuild_tables()
n = 1 << lg
pins = [native.PinnedBuffer(32 * n).write(ctx.random_poly_range(5, j * n, n)) for j in range(k)]
xs = b"".join(ctx.random_point(10 + j) for j in range(k))
rows = [j % 4 for j in range(k)]
for _ in range(2):
    out = ctx.worker_commit_open_batch(rows, pins, xs)
print("batch", k, "x 2^", lg, "ok", all(o[0] == 0 for o in out))
