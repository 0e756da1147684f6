"""worker_commit / worker_open / commit+open latency through the C ABI (pinned input) at small sizes, for A/B builds.
python tools/lat_small.py [log_n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native

def med(f, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

for lg in [int(a) for a in sys.argv[1:]] or [12, 14, 16, 18, 20]:
    n = 1 << lg
    ctx = native.Context(0)
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    ctx.prebuild_tables()
    pin = native.PinnedBuffer(32 * n)
    pin.write(ctx.random_poly(0xB200 + 3, n))
    x = ctx.random_point(5)
    reps = 101 if lg <= 16 else 21
    for _ in range(3):
        r = (ctx.worker_commit(0, pin), ctx.worker_open(0, pin, x), ctx.worker_commit_open(0, pin, x))
    a = med(lambda: ctx.worker_commit(0, pin), reps)
    b = med(lambda: ctx.worker_open(0, pin, x), reps)
    c = med(lambda: ctx.worker_commit_open(0, pin, x), reps)
    print(f"2^{lg}: worker_commit {a[0]:.3f}/{a[1]:.3f} | worker_open {b[0]:.3f}/{b[1]:.3f} | commit_open {c[0]:.3f}/{c[1]:.3f}  (median/min ms) {r[0].hex()[:8]}", flush=True)
    ctx.close()
