import sys, time
sys.path.insert(0, ".")
from zkp_subnet_b200 import native
TX, TY = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
n = 1 << 20
with native.MultiContext([0]) as mg:
    mg.srs_generate(TX, TY, 20, 0, native.LAYOUT_ROWS)
    mg.prebuild_tables()
    ctx = mg.ctx(0)
    poly = native.PinnedBuffer(32 * n).write(ctx.random_poly_range(0xB203, 0, n))
    x = ctx.random_point(1)
    mg.pianist_commit_open([0], poly, x)
    for _ in range(3): mg.pianist_commit_open([0], poly, x, native.MGPU_RESIDENT)
    t0 = time.perf_counter()
    for _ in range(6): mg.pianist_commit_open([0], poly, x, native.MGPU_RESIDENT)
    print("wall per call ms", (time.perf_counter() - t0) / 6 * 1e3)
    ms, msk, launches, *_ = ctx.bench_commit_open(0, poly, x, 6, False)
    print("bench_commit_open (no flush) ms", ms)
    ms, msk, launches, *_ = ctx.bench_commit_open(0, poly, x, 6, True)
    print("bench_commit_open (flush) ms", ms)
