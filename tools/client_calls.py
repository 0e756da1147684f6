"""Per-call wall times of fourier.Client.worker_commit_and_open / the two-call flow at 2^20 right after attach (what
bench.py times), chunked decode+upload on (default) and off."""
import os, sys, time, base64
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
from zkp_subnet_b200.client import Client, encode_poly
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = native.Context(0)
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
ctx.prebuild_tables()
strs = encode_poly(ctx.random_poly(0xB200 + 3, 1 << lg))
xs = base64.b64encode(ctx.random_point(5)).decode().rstrip("=")
for staged in (None, False, None, False):
    cl = Client(staged_upload=staged).attach(ctx, lg, 0)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); cl.worker_commit_and_open(0, strs, xs); ts.append((time.perf_counter() - t0) * 1e3)
    t2 = []
    for _ in range(6):
        t0 = time.perf_counter(); cl.worker_commit(0, strs); cl.worker_open(0, strs, xs); t2.append((time.perf_counter() - t0) * 1e3)
    print(f"staged={staged}: fused calls " + " ".join(f"{t:.2f}" for t in ts) + " | two-call " + " ".join(f"{t:.2f}" for t in t2), flush=True)
