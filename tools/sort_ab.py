"""MSM time with the hand-written bucket sort vs cub::DeviceRadixSort at several sizes (picks the crossover of
msm_uses_bucket_sort).  usage: python tools/sort_ab.py 18 19 20 21 22"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
ctx = native.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [18, 20, 21, 22]:
    n = 1 << lg
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    sc = ctx.random_poly(0xB200 + lg, n)
    res = {}
    for mode in (2, 1, 0):
        ctx.set_msm_sort(mode)
        ctx.bench_msm(0, sc, 2, True)
        ms, out = ctx.bench_msm(0, sc, 6, True)
        res[mode] = (ms, out.hex()[:8])
    c, W, _ = ctx.msm_info(n)
    print(f"2^{lg} c={c} W={W}: auto {res[2][0]:8.3f} ms   bucket {res[1][0]:8.3f} ms   cub {res[0][0]:8.3f} ms   {res[1][1]} {res[0][1]}", flush=True)
