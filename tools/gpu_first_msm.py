"""First GPU bring-up: MSM of small/medium sizes against the C oracle (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref
from zkp_subnet_b200 import native

TAU = 1927409816240961209460912649124
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
ctx = native.Context(0)
ok = True
for log_n in (4, 8, 12, 14):
    n = 1 << log_n
    t = time.time(); srs = ref.srs(n, TAU, "lagrange"); t_srs = time.time() - t
    ctx.srs_set_shape(log_n, 0)
    ctx.srs_import_row(0, srs)
    assert ctx.srs_export_row(0, n) == srs, "srs roundtrip"
    cases = {"random": ref.random_scalars(0xB200 + log_n, n),
             "zeros": bytes(32 * n), "ones": ref.join32([1] * n), "rm1": ref.join32([R - 1] * n),
             "single": ref.join32([0] * (n - 1) + [12345]),
             "small": ref.join32([i % 7 for i in range(n)])}
    for name, sc in cases.items():
        t = time.time(); got = ctx.msm_g1(0, sc); t_gpu = time.time() - t
        t = time.time(); exp = ref.msm(srs, sc, 16); t_cpu = time.time() - t
        good = got == exp
        ok &= good
        print(f"n=2^{log_n} {name:7s} {'OK ' if good else 'BAD'} gpu {t_gpu*1e3:.2f} ms cpu {t_cpu*1e3:.1f} ms (srs gen {t_srs:.1f}s) {got.hex()[:16]}", flush=True)
    # partial length (n not a power of two, fewer scalars than the row)
    m = n - 3
    sc = ref.random_scalars(99, m)
    good = ctx.msm_g1(0, sc) == ref.msm(srs, sc, 16); ok &= good
    print(f"n=2^{log_n}-3 partial {'OK' if good else 'BAD'}")
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
