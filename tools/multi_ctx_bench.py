"""Throughput of 2^16 commit+open with several contexts (one host thread each) on one GPU."""
import concurrent.futures, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
TAU = (1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF)
for nctx in ((1, 2, 4, 8, 16) if lg <= 17 else (1, 2, 3, 4)):
    ctxs = [native.Context(0) for _ in range(nctx)]
    pins = []
    for k, c in enumerate(ctxs):
        c.srs_generate(*TAU, lg, 0)
        pins.append(native.PinnedBuffer(32 << lg).write(c.random_poly(100 + k, 1 << lg)))
    x = ctxs[0].random_point(1)
    per = (64 if lg <= 17 else 24) // nctx
    def work(k):
        for _ in range(per):
            ctxs[k].worker_commit_open(0, pins[k], x)
    with concurrent.futures.ThreadPoolExecutor(nctx) as ex:
        list(ex.map(work, range(nctx)))
        t0 = time.perf_counter()
        list(ex.map(work, range(nctx)))
        dt = time.perf_counter() - t0
    print(f"2^{lg}: {nctx:2d} contexts: {per * nctx / dt:8.1f} commit+open/s", flush=True)
    for c in ctxs: c.close()
    for p in pins: p.close()
