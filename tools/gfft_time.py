"""Time of the trapdoor-free SRS derivation (zkp_srs_monomial_to_lagrange: inverse group FFTs along Y and X) and of the
compressed-point import, at a few shapes.  python tools/gfft_time.py > profiles/r2_gfft_time.txt"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
TX, TY = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
print("# log_n log_m  points  monomial_gen_s  to_lagrange_s  us_per_point  rows_equal_trapdoor_generation  import_compressed_row_ms")
for lg, lm in ((12, 4), (16, 2), (16, 4), (14, 8)):
    with native.Context(0) as ctx:
        n, rows = 1 << lg, 1 << lm
        ctx.srs_generate(TX, TY, lg, lm)
        want = [ctx.srs_export_row(i, n) for i in (0, rows - 1)]
        t0 = time.perf_counter()
        ctx.srs_generate_monomial2(TX, TY, lg, lm)
        t1 = time.perf_counter()
        ctx.srs_monomial_to_lagrange()
        t2 = time.perf_counter()
        got = [ctx.srs_export_row(i, n) for i in (0, rows - 1)]
        comp = ctx.srs_export_row_compressed(0, n)
        t3 = time.perf_counter()
        ctx.srs_import_row_compressed(0, comp)
        t4 = time.perf_counter()
        print(f"{lg:5d} {lm:5d} {n * rows:8d} {t1 - t0:12.3f} {t2 - t1:14.3f} {(t2 - t1) / (n * rows) * 1e6:12.2f} {got == want!s:>10} {(t4 - t3) * 1e3:20.2f}")
