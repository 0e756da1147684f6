"""Times MSM / commit+open at 2^16 and 2^20 with the library named by ZKP_B200_LIB (tuning builds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
tag = os.path.basename(os.environ.get("ZKP_B200_LIB", "default")) + ":" + os.environ.get("ZKP_SORT", "bucket")
ref = {}
for lg in (16, 20):
    ctx = native.Context(0)
    if os.environ.get('ZKP_SORT') == 'cub':
        ctx.set_msm_sort(False)
    ctx.srs_generate(1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF, lg, 0)
    poly = ctx.random_poly(0xB200 + 3, 1 << lg)
    x = ctx.random_point(0xA1FA)
    ctx.bench_msm(0, poly, 2, True)
    ms, out = ctx.bench_msm(0, poly, 8, True)
    kern = ctx.bench_last_kernel_ms()
    ctx.bench_commit_open(0, poly, x, 2, True)
    ms_co, ms_k, launches, com, y, proof = ctx.bench_commit_open(0, poly, x, 8, True)
    print(f"{tag:14s} 2^{lg}: msm {ms:7.3f} ms (acc_l0 {kern:6.3f})  commit+open {ms_co:7.3f} ms  launches {launches}  {out.hex()[:12]} {proof.hex()[:12]}", flush=True)
    ctx.close()
