"""One 2^21-point shard of the 2^24 MSM (what each of 8 GPUs runs) and one 2^20 commit+open: total, accumulation kernel
and the rest, for the level-0 slice target given in ZKP_L0_TARGET (environment).  Run once untimed first (cold-box effect)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
TX, TY = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
lg_local = int(sys.argv[1]) if len(sys.argv) > 1 else 21
tag = f"L0_TARGET={os.environ.get('ZKP_L0_TARGET', 'default')}"
ctx = native.Context(0)
ctx.srs_generate_shard(TX, TY, 24, 0, 0, 24 - lg_local)
n = 1 << lg_local
sc = native.PinnedBuffer(32 * n).write(ctx.random_poly_range(0xB204, 0, n))
ctx.bench_msm(0, sc, 3, True)
ms, _ = ctx.bench_msm(0, sc, 8, True)
k = ctx.bench_last_kernel_ms()
print(f"{tag} shard 2^{lg_local}: total {ms:.3f} ms  accumulate {k:.3f} ms  rest {ms - k:.3f} ms")
ctx.close()
ctx = native.Context(0)
ctx.srs_generate(TX, TY, 20, 0)
ctx.prebuild_tables()
p = native.PinnedBuffer(32 << 20).write(ctx.random_poly(7, 1 << 20))
x = ctx.random_point(1)
ctx.bench_commit_open(0, p, x, 10, False)
ms, msk, launches, *_ = ctx.bench_commit_open(0, p, x, 20, True)
print(f"{tag} commit+open 2^20: {ms:.3f} ms  accumulate-in-step {msk:.3f} ms  launches {launches}")
