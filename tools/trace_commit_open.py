"""Stage timeline of one commit+open at 2^LOG_N (both lanes), from CUDA events between the pipeline stages."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = native.Context(0)
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
poly = ctx.random_poly(0xB200 + 3, 1 << lg)
x = ctx.random_point(0xA1FA)
for rep in range(2):
    rows = ctx.bench_trace(0, poly, x, 2)
    last = {}
    print(f"--- commit+open 2^{lg}, trace {rep}")
    for lane, stage, t in rows:
        prev = last.get(lane, 0.0)
        print(f"lane {lane:>4s} {stage:22s} done at {t:8.3f} ms  (+{t - prev:7.3f})")
        last[lane] = t
