"""A/B of the NTT with the pass-2 tile fetched by per-thread loads vs ONE bulk async copy (TMA, cp.async.bulk + mbarrier).
python tools/ntt_tma_ab.py > profiles/r2_ntt_tma_ab.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
ctx = native.Context(0)
print("# log_n   loads_ms   tma_ms   (zkp_bench_ntt: CUDA events, L2 flushed, mean of 20; forward transform)")
for lg in (16, 18, 20, 22):
    n = 1 << lg
    v = ctx.random_poly(lg, n)
    res = {}
    outs = {}
    for tma in (0, 1, 0, 1):
        ctx.set_ntt_tma(bool(tma))
        outs[tma] = ctx.fft(v, True, False)
        res.setdefault(tma, []).append(ctx.bench_ntt(n, 20, False))
    assert outs[0] == outs[1], "TMA variant changes the result"
    print(f"{lg:6d}  {min(res[0]):9.4f}  {min(res[1]):9.4f}")
ctx.set_ntt_tma(False)
