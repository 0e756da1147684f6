"""Small driver for ncu: one SRS generation + a few MSMs / commit+opens at 2^LOG_N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = sys.argv[2] if len(sys.argv) > 2 else "msm"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = native.Context(0)
if os.environ.get('ZKP_AFFINE_ROUNDS'):
    ctx.set_msm_affine_rounds(int(os.environ['ZKP_AFFINE_ROUNDS']))
ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
n = 1 << lg
sc = ctx.random_poly(0xB200 + lg, n)
x = ctx.random_point(5)
if mode == "msm":
    ms, out = ctx.bench_msm(0, sc, reps, False)
    print(f"msm 2^{lg}: {ms:.3f} ms {out.hex()[:16]}")
else:
    ms, ms_k, launches, com, y, proof = ctx.bench_commit_open(0, sc, x, reps, False)
    print(f"commit+open 2^{lg}: {ms:.3f} ms acc {ms_k:.3f} ms launches {launches}")
