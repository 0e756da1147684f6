"""Back-to-back resident commit+opens run 0.3 ms slower per step than the same step after an L2 flush; which kind of
intervening work restores the fast mode?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
TX, TY = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
n = 1 << 20
ctx = native.Context(0)
ctx.srs_generate(TX, TY, 20, 0); ctx.prebuild_tables()
aux = native.Context(0)
p = native.PinnedBuffer(32 * n).write(ctx.random_poly(7, n)); x = ctx.random_point(1)
big = native.PinnedBuffer(32 << 20).write(bytes(32 << 20))
ctx.bench_commit_open(0, p, x, 20, False)
def run(label, gap):
    rows = []
    for i in range(8):
        gap()
        ms, msk, launches, *_ = ctx.bench_commit_open(0, p, x, 1, False)
        rows.append(f"{ms:.2f}/{msk:.2f}")
    print(f"{label:34s} " + " ".join(rows), flush=True)
run("no gap", lambda: None)
run("flush on the same context", lambda: ctx.bench_flush_l2())
run("flush on ANOTHER context", lambda: aux.bench_flush_l2())
run("fft 2^20 on another context (32MiB)", lambda: aux.fft(big, True, False))
run("random_poly 2^20 (32 MiB write)", lambda: aux.random_poly(1, 1 << 20))
run("random_poly 2^23 (256 MiB write)", lambda: aux.random_poly(1, 1 << 23))
run("no gap again", lambda: None)
