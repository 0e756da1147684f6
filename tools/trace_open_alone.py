"""worker_open alone (no commitment MSM on lane 0): is the opening slow by itself or only when it shares the GPU?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
for lg in (16, 20):
    ctx = native.Context(0)
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    poly = ctx.random_poly(0xB200 + 3, 1 << lg)
    x = ctx.random_point(0xA1FA)
    pin = native.PinnedBuffer(len(poly)).write(poly)
    for name, fn in (("commit", lambda: ctx.worker_commit(0, pin)), ("open", lambda: ctx.worker_open(0, pin, x)),
                     ("commit_open", lambda: ctx.worker_commit_open(0, pin, x))):
        fn(); fn()
        t = time.perf_counter()
        for _ in range(10):
            fn()
        print(f"2^{lg} {name}: {(time.perf_counter() - t) * 100:.3f} ms per call (host wall, pinned input)")
    ctx.close()
