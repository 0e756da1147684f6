"""zkp_worker_commit_open_batch at 2^16 (32 requests per launch set; one context and two forked contexts alternating):
coset opening on / off.  python tools/batch_ab.py"""
import os, sys, time, concurrent.futures
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg, batch, reps = 16, 32, 6
n = 1 << lg
c = native.Context(0)
c.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 2)
c.prebuild_tables()
pins = [native.PinnedBuffer(32 * n).write(c.random_poly_range(0xB200 + 2, k * n, n)) for k in range(batch)]
xs = b"".join(c.random_point(100 + k) for k in range(batch))
rows = [k % 4 for k in range(batch)]
f2 = [c, c.fork()]
ref = None
for mode in (0, 1, 0, 1):
    for x in f2:
        x.set_open_coset(bool(mode))
    r = c.worker_commit_open_batch(rows, pins, xs)
    ref = ref or r
    assert r == ref
    t0 = time.perf_counter()
    for _ in range(reps):
        c.worker_commit_open_batch(rows, pins, xs)
    one = batch * reps / (time.perf_counter() - t0)
    def work(k):
        for _ in range(reps):
            o = f2[k].worker_commit_open_batch(rows, pins, xs)
        return o
    with concurrent.futures.ThreadPoolExecutor(2) as ex:
        list(ex.map(work, range(2)))
        t0 = time.perf_counter()
        outs = list(ex.map(work, range(2)))
        two = 2 * batch * reps / (time.perf_counter() - t0)
    assert outs[0] == ref and outs[1] == ref
    print(f"coset={mode}: batch of {batch} at 2^{lg}: {one:.1f} commit+open/s (one context), {two:.1f} (two contexts)", flush=True)
