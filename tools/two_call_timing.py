"""Per-call wall times of the reference's two-call flow through fourier.Client at 2^LOG_N (worker_commit, then
worker_open with the same list -> resident path; worker_open with an equal copy after the polynomial was dropped ->
regular path; fused worker_commit_and_open)."""
import os, sys, time, base64
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
from zkp_subnet_b200.client import Client, encode_poly
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = native.Context(0)
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
cl = Client().attach(ctx, lg, 0)
strs = encode_poly(ctx.random_poly(0xB200 + 3, 1 << lg))
xs = base64.b64encode(ctx.random_point(5)).decode().rstrip("=")
def t(f):
    t0 = time.perf_counter(); r = f(); return (time.perf_counter() - t0) * 1e3, r
cl.worker_commit_and_open(0, strs, xs)
cl.worker_open(0, strs, xs)  # allocates the second page-locked staging buffer (one-off, ~30 ms)
for rep in range(3):
    a, r1 = t(lambda: cl.worker_commit(0, strs))
    b, r2 = t(lambda: cl.worker_open(0, strs, xs))
    ctx.random_point(1); cl._slots[0].resident_n = 0
    c, r3 = t(lambda: cl.worker_open(0, strs, xs))
    d, r4 = t(lambda: cl.worker_commit_and_open(0, strs, xs))
    assert r2.json() == r3.json() and r4.json()["proof"] == r2.json()["proof"]
    print(f"2^{lg}: worker_commit {a:6.2f} ms | worker_open resident {b:6.2f} ms | worker_open regular {c:6.2f} ms | fused {d:6.2f} ms", flush=True)
