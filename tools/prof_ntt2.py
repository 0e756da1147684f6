"""ncu driver: a few 2^20 NTTs with the pass-2 tile fetched by per-thread loads (argv[1] = 0) or by one bulk copy (1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
ctx = native.Context(0)
ctx.set_ntt_tma(bool(int(sys.argv[1])))
print(ctx.bench_ntt(1 << 20, 3, False))
