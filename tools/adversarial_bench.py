"""MSM time for scalar distributions that stress the slot levels (tuning builds via ZKP_B200_LIB): uniform random,
a constant polynomial (every digit of a window in ONE bucket: each level-0 slice ends in a cut run), and a
polynomial that alternates between two values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
tag = os.path.basename(os.environ.get("ZKP_B200_LIB", "default")) + ":" + os.environ.get("ZKP_SORT", "bucket")
for lg in (16, 20):
    n = 1 << lg
    ctx = native.Context(0)
    if os.environ.get('ZKP_SORT') == 'cub':
        ctx.set_msm_sort(False)
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    rnd = ctx.random_poly(0xB200 + 3, n)
    a, b = rnd[:32], rnd[32:64]
    cases = {"random": rnd, "constant": a * n, "two-valued": (a + b) * (n // 2)}
    out = []
    for name, sc in cases.items():
        ctx.bench_msm(0, sc, 2, True)
        ms, _ = ctx.bench_msm(0, sc, 8, True)
        out.append(f"{name} {ms:7.3f} ms")
    print(f"{tag:14s} 2^{lg}: " + "   ".join(out), flush=True)
    ctx.close()
