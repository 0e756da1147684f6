"""MSM / commit+open timing at small sizes for the library named by ZKP_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
tag = os.path.basename(os.environ.get("ZKP_B200_LIB", "default"))
for lg in (12, 14, 16, 17):
    ctx = native.Context(0)
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    poly = ctx.random_poly(0xB200 + 3, 1 << lg)
    x = ctx.random_point(0xA1FA)
    ms, out = ctx.bench_msm(0, poly, 20, False)
    ctx.bench_commit_open(0, poly, x, 3, False)
    ms_co = ctx.bench_commit_open(0, poly, x, 20, False)[0]
    print(f"{tag:12s} 2^{lg}: msm {ms:7.3f} ms (acc_l0 {ctx.bench_last_kernel_ms():6.3f})  commit+open {ms_co:7.3f} ms {out.hex()[:10]}", flush=True)
    ctx.close()
