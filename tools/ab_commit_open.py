"""A/B of commit+open at 2^20 with 0 vs R batched-affine rounds, interleaved repetitions."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
ctx = native.Context(0)
lg = 20
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
sc = ctx.random_poly(0xB200 + 3, 1 << lg)
x = ctx.random_point(1)
res = {}
for rounds in (0, 2, 3):
    ctx.set_msm_affine_rounds(rounds)
    ctx.bench_commit_open(0, sc, x, 3, True)
for rep in range(5):
    for rounds in (0, 2, 3):
        ctx.set_msm_affine_rounds(rounds)
        res.setdefault(rounds, []).append(ctx.bench_commit_open(0, sc, x, 10, True)[0])
for rounds, v in res.items():
    print(f"rounds={rounds}: commit+open median {statistics.median(v):.3f} ms  min {min(v):.3f}  max {max(v):.3f}")
