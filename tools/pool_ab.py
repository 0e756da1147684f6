"""Host-path A/B on the GPU box: the wire codec with its persistent worker pool vs fresh threads per call
(ZKP_CODEC_POOL=0), and the reference's two-call flow / the fused call through fourier.Client at 2^LOG_N.
python tools/pool_ab.py [log_n ...]   (run once per setting of ZKP_CODEC_POOL)"""
import os, sys, time, base64, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
from zkp_subnet_b200.client import Client, encode_poly, _Helper

def med(f, reps=41):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

print("ZKP_CODEC_POOL =", os.environ.get("ZKP_CODEC_POOL", "(default: on)"), "cores:", os.cpu_count())
def spawn():
    t = threading.Thread(target=lambda: None); t.start(); t.join()
h = _Helper()
print("threading.Thread start+join (median, min ms): %.3f %.3f | persistent helper submit+wait: %.3f %.3f" % (*med(spawn, 201), *med(lambda: h.submit(lambda: None).wait(), 201)))
for lg in [int(a) for a in sys.argv[1:]] or [12, 16, 20]:
    n = 1 << lg
    ctx = native.Context(0)
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    ctx.prebuild_tables()
    cl = Client(staged_upload=False).attach(ctx, lg, 0)
    raw = ctx.random_poly(0xB200 + 3, n)
    strs = encode_poly(raw)
    xs = base64.b64encode(ctx.random_point(5)).decode().rstrip("=")
    pin = native.PinnedBuffer(32 * n)
    for _ in range(3):
        cl.worker_commit_and_open(0, strs, xs); cl.worker_commit(0, strs); cl.worker_open(0, strs, xs)
    reps = 41 if lg <= 18 else 15
    dec = med(lambda: native.wire_decode_list(strs, pin), reps)
    enc = med(lambda: native.wire_encode_list(raw), reps)
    def two():
        c = cl.worker_commit(0, strs).json()["commitment"]; o = cl.worker_open(0, strs, xs).json(); return c, o
    t2 = med(two, reps)
    tc = med(lambda: cl.worker_commit(0, strs), reps)
    to = med(lambda: (cl.worker_commit(0, strs), None)[1], 1)  # keep the polynomial resident for the next line
    topen = med(lambda: cl.worker_open(0, strs, xs), reps)
    tf = med(lambda: cl.worker_commit_and_open(0, strs, xs), reps)
    x = ctx.random_point(5)
    tabi = med(lambda: ctx.worker_commit_open(0, pin, x), reps)
    print(f"2^{lg}: decode {dec[0]:.3f}/{dec[1]:.3f} encode {enc[0]:.3f}/{enc[1]:.3f} | worker_commit {tc[0]:.3f} | worker_open (resident) {topen[0]:.3f} | "
          f"two-call {t2[0]:.3f}/{t2[1]:.3f} | fused Client {tf[0]:.3f}/{tf[1]:.3f} | C-ABI commit_open {tabi[0]:.3f}/{tabi[1]:.3f}  (median/min ms)", flush=True)
    if lg >= 17:
        orig = native.Context.stage_list
        for ch in (1 << 16, 1 << 17, 1 << 18, 1 << 19):
            native.Context.stage_list = lambda self, strs_, st, chunk=ch: orig(self, strs_, st, chunk)
            cs = Client(staged_upload=True).attach(ctx, lg, 0)
            for _ in range(3):
                cs.worker_commit_and_open(0, strs, xs)
            r1, r2 = cs.worker_commit_and_open(0, strs, xs).json(), cl.worker_commit_and_open(0, strs, xs).json()
            assert r1 == r2
            def two_s():
                c = cs.worker_commit(0, strs).json()["commitment"]; o = cs.worker_open(0, strs, xs).json(); return c, o
            a, b = med(lambda: cs.worker_commit_and_open(0, strs, xs), reps), med(two_s, reps)
            print(f"   staged upload, chunks of 2^{ch.bit_length() - 1}: fused Client {a[0]:.3f}/{a[1]:.3f} | two-call {b[0]:.3f}/{b[1]:.3f}", flush=True)
        native.Context.stage_list = orig
    ctx.close()
