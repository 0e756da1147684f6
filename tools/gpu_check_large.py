"""Config 4 of BASELINE.json: G1 MSM at 2^24 (SRS 1.5 GiB in HBM) -- correctness through the trapdoor identity
and timing; also 2^22.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref
from zkp_subnet_b200 import native
TAU_X, TAU_Y = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
ctx = native.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [22, 24]:
    n = 1 << lg
    t = time.time(); ctx.srs_generate(TAU_X, TAU_Y, lg, 0); t_srs = time.time() - t
    sc = ctx.random_poly(0xB200 + 4, n)
    t = time.time(); com = ctx.worker_commit(0, sc); t_first = time.time() - t   # builds the fixed-base tables
    c, W, muls = ctx.msm_info(n)
    ms, com2 = ctx.bench_msm(0, sc, 3, True)
    t = time.time(); exp = ref.g1_mul_gen(ref.fr_dot(sc, ref.lagrange_scalars(n, TAU_X))); t_or = time.time() - t
    print(f"n=2^{lg} c={c} W={W}: srs {t_srs:.2f}s, first call (tables) {t_first:.2f}s, msm {ms:.2f} ms = {n/ms/1e3:.1f} Mpts/s, "
          f"{muls/ms/1e6:.2f} G Fq-mul/s, kernel {ctx.bench_last_kernel_ms():.2f} ms; trapdoor {'OK' if com == com2 == exp else 'BAD'} (oracle {t_or:.1f}s)", flush=True)
    x = ctx.random_point(9)
    t = time.time(); com3, y, proof = ctx.worker_commit_open(0, sc, x); t_co = time.time() - t
    print(f"   commit+open {t_co*1e3:.1f} ms (host buffers), verify {ctx.worker_verify(0, proof, x, y, com3)}, ntt {ctx.bench_ntt(n, 2):.3f} ms", flush=True)
