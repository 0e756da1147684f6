"""Latency of one commit+open as a function of the row length: two lanes vs one grouped launch set (zkp_set_fuse),
and throughput of the batched entry as a function of the batch size.  python tools/fuse_sweep.py > profiles/r2_fuse_sweep.txt"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native

TAU_X, TAU_Y = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
print("# log_n  two_lanes_ms  fused_ms   (pinned host buffer in, results on host; mean of 50 calls)")
for lg in (8, 10, 12, 14, 16, 18, 20):
    n = 1 << lg
    with native.Context(0) as ctx:
        ctx.srs_generate(TAU_X, TAU_Y, lg, 0)
        ctx.prebuild_tables()
        p = native.PinnedBuffer(32 * n).write(ctx.random_poly(lg, n))
        x = ctx.random_point(1)
        res = {}
        for mode in (0, 1):
            ctx.set_fuse(mode)
            ref = ctx.worker_commit_open(0, p, x)
            reps = 50 if lg <= 16 else 10
            for _ in range(3):
                ctx.worker_commit_open(0, p, x)
            t0 = time.perf_counter()
            for _ in range(reps):
                out = ctx.worker_commit_open(0, p, x)
            res[mode] = (time.perf_counter() - t0) * 1e3 / reps
            assert out == ref
        print(f"{lg:6d}  {res[0]:10.3f}  {res[1]:10.3f}")
        p.close()
print("# batch entry at 2^16 (4 rows): batch size, requests/s, ms per request")
lg = 16
n = 1 << lg
with native.Context(0) as ctx:
    ctx.srs_generate(TAU_X, TAU_Y, lg, 2)
    ctx.prebuild_tables()
    pins = [native.PinnedBuffer(32 * n).write(ctx.random_poly_range(5, k * n, n)) for k in range(32)]
    xs = [ctx.random_point(10 + k) for k in range(32)]
    for k in (1, 2, 4, 8, 16, 32):
        rows = [r % 4 for r in range(k)]
        xcat = b"".join(xs[:k])
        ctx.worker_commit_open_batch(rows, pins[:k], xcat)
        reps = max(2, 64 // k)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.worker_commit_open_batch(rows, pins[:k], xcat)
        dt = (time.perf_counter() - t0) / reps
        print(f"{k:6d}  {k / dt:10.1f}  {dt * 1e3 / k:8.3f}")
