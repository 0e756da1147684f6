import os, sys
sys.path.insert(0, ".")
from zkp_subnet_b200 import native
TX, TY = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
n = 1 << 20
ctx = native.Context(0)
ctx.srs_generate(TX, TY, 20, 0); ctx.prebuild_tables()
p = native.PinnedBuffer(32 * n).write(ctx.random_poly(7, n)); x = ctx.random_point(1)
ctx.bench_commit_open(0, p, x, 20, False)
for rep in range(2):
    tr = ctx.bench_trace(0, p, x, 3)
    print(os.environ.get("ZKP_TRACE_FLUSH", "async"), " ".join(f"{l}:{s}:{t:.2f}" for l, s, t in tr))
