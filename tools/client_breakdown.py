"""Where a fourier.Client.worker_commit_and_open call at 2^LOG_N spends its time: list decode, the C-ABI call, the rest."""
import os, sys, time, base64
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
from zkp_subnet_b200.client import Client, encode_poly
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
ctx = native.Context(0)
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
raw = ctx.random_poly(0xB200 + 3, n)
strs = encode_poly(raw)
x = ctx.random_point(5)
xs = base64.b64encode(x).decode().rstrip("=")
pin = native.PinnedBuffer(32 * n)
cl = Client().attach(ctx, lg, 0)
cl.worker_commit_and_open(0, strs, xs)
now = time.perf_counter
for rep in range(4):
    t0 = now(); native.wire_decode_list(strs, pin); t1 = now(); ctx.worker_commit_open(0, pin, x); t2 = now()
    ctx.worker_commit_open(0, pin, x); t3 = now()
    r = cl.worker_commit_and_open(0, strs, xs); t4 = now()
    print(f"2^{lg}: decode {1e3*(t1-t0):5.2f} | C-ABI call right after the decode {1e3*(t2-t1):6.2f} | C-ABI call again {1e3*(t3-t2):6.2f} | "
          f"Client call {1e3*(t4-t3):6.2f} ms", flush=True)
