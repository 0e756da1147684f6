import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = native.Context(0)
print(f"ntt 2^{lg}: {ctx.bench_ntt(1 << lg, 3):.4f} ms")
