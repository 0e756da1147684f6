"""Batched-affine rounds: result invariance against the XYZZ-only path (small sizes, adversarial scalars) and timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
TAU = (1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF)
ctx = native.Context(0)
bad = 0
for lg in (2, 5, 8, 11, 13):
    n = 1 << lg
    ctx.srs_generate(*TAU, lg, 0)
    sets = {"random": ctx.random_poly(lg, n), "const": (R - 5).to_bytes(32, "big") * n, "zero": bytes(32 * n),
            "one": (1).to_bytes(32, "big") * n, "single": bytes(32 * (n - 1)) + (7).to_bytes(32, "big"),
            "rm1": (R - 1).to_bytes(32, "big") * n}
    for name, sc in sets.items():
        ctx.set_msm_affine_rounds(0)
        ref = ctx.msm_g1(0, sc)
        for rounds in (1, 2, 3, 6):
            ctx.set_msm_affine_rounds(rounds)
            got = ctx.msm_g1(0, sc)
            if got != ref:
                bad += 1
                print(f"MISMATCH lg={lg} {name} rounds={rounds}: {got.hex()[:16]} vs {ref.hex()[:16]}")
print("invariance:", "OK" if not bad else f"{bad} mismatches", flush=True)
if "--time" in sys.argv:
    for lg in (18, 20, 22):
        n = 1 << lg
        ctx.srs_generate(*TAU, lg, 0)
        sc = ctx.random_poly(0xB200 + 3, n)
        x = ctx.random_point(1)
        ref = None
        for rounds in (0, 2, 3, 4):
            ctx.set_msm_affine_rounds(rounds)
            ctx.bench_msm(0, sc, 2, True)
            ms, out = ctx.bench_msm(0, sc, 5, True)
            ms_co = ctx.bench_commit_open(0, sc, x, 4, True)[0] if lg <= 20 else float("nan")
            ref = ref or out
            print(f"2^{lg} rounds={rounds}: msm {ms:8.3f} ms (xyzz level0 {ctx.bench_last_kernel_ms():7.3f})  commit+open {ms_co:8.3f} ms  {'same' if out == ref else 'DIFFERENT'}", flush=True)
