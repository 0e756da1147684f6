"""Timing of the batched-affine path at 2^LOG_N for the library named by ZKP_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
tag = os.path.basename(os.environ.get("ZKP_B200_LIB", "default"))
ctx = native.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [20]:
    n = 1 << lg
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    sc = ctx.random_poly(0xB200 + 3, n)
    x = ctx.random_point(1)
    ref = None
    for rounds in (0, 3):
        ctx.set_msm_affine_rounds(rounds)
        ctx.bench_msm(0, sc, 2, True)
        ms, out = ctx.bench_msm(0, sc, 6, True)
        ctx.bench_commit_open(0, sc, x, 2, True)
        ms_co = ctx.bench_commit_open(0, sc, x, 6, True)[0]
        ref = ref or out
        print(f"{tag:12s} 2^{lg} rounds={rounds}: msm {ms:8.3f} ms  commit+open {ms_co:8.3f} ms  {'same' if out == ref else 'DIFFERENT'}", flush=True)
