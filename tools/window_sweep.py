"""MSM time vs window width c (fixed-base tables) at several sizes: validates the cost model in msm_window_bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
ctx = native.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [18, 19, 20, 21, 22]:
    n = 1 << lg
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    sc = ctx.random_poly(0xB200 + lg, n)
    x = ctx.random_point(3)
    ctx.set_msm_window(0)
    auto_c = ctx.msm_info(n)[0]
    res = []
    for c in range(max(12, lg - 3), min(24, lg + 2) + 1):
        ctx.set_msm_window(c)
        ms, out = ctx.bench_msm(0, sc, 4, True)
        co = ctx.bench_commit_open(0, sc, x, 3, True)[0] if lg <= 21 else float("nan")
        res.append((c, ms, co))
    ctx.set_msm_window(0)
    print(f"2^{lg} (auto c={auto_c}): " + "  ".join(f"c={c}: {ms:.3f}/{co:.2f}" for c, ms, co in res), flush=True)
