"""Randomised differential test on the GPU: MSM / commit+open against the oracle over random sizes, window widths,
table modes, batched-affine rounds, scalar shapes, and the shapes of the request (two lanes, one grouped launch set, the
batched entry, a forked context, the in-library multi-device entries on one device).
usage: python tools/fuzz_msm.py [cases]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref
from zkp_subnet_b200 import native
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
TAU = (1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = random.Random(0xF022)
ctx = native.Context(0)
srs_cache = {}
bad = 0
for case in range(cases):
    lg = rng.choice([1, 2, 3, 5, 7, 8, 9, 10, 11, 12, 13])
    n = 1 << lg
    if lg not in srs_cache:
        ctx.srs_generate(*TAU, lg, 0)
        srs_cache[lg] = ctx.srs_export_row(0, n)
    else:
        ctx.srs_set_shape(lg, 0)
        ctx.srs_import_row(0, srs_cache[lg])
        ctx.srs_import_g2_tau(TAU[0])
    srs = srs_cache[lg]
    shape = rng.choice(["random", "small", "sparse", "equal", "near_r", "ragged"])
    m = n
    if shape == "random":
        vals = [rng.randrange(R) for _ in range(n)]
    elif shape == "small":
        vals = [rng.randrange(1 << rng.choice([1, 8, 20, 33])) for _ in range(n)]
    elif shape == "sparse":
        vals = [rng.randrange(R) if rng.random() < 0.1 else 0 for _ in range(n)]
    elif shape == "equal":
        v = rng.randrange(R); vals = [v] * n
    elif shape == "near_r":
        vals = [R - 1 - rng.randrange(4) for _ in range(n)]
    else:
        m = rng.randrange(1, n + 1); vals = [rng.randrange(R) for _ in range(m)]
    sc = ref.join32(vals)
    ctx.set_msm_mode(rng.random() < 0.7)
    ctx.set_msm_window(rng.choice([0, 0, 3, 5, 8, 11, 14]))
    ctx.set_msm_affine_rounds(rng.choice([-1, -1, 1, 2, 4]))
    got = ctx.msm_g1(0, sc)
    exp = ref.msm(srs, sc, 8)
    ok = got == exp
    if ok and m == n and lg >= 3 and rng.random() < 0.4:
        x = ref.random_scalars(case, 1)
        ctx.set_fuse(rng.choice([-1, 0, 1]))
        com, y, proof = ctx.worker_commit_open(0, sc, x)
        ctx.set_fuse(-1)
        ey, eproof = ref.open_evals(sc, x, srs, 8)
        ok = (com, y, proof) == (exp, ey, eproof) and ctx.worker_verify(0, proof, x, y, com)
        if ok and rng.random() < 0.5:
            # the same request inside a batch (with a second, random polynomial), on a forked context
            other = ref.random_scalars(1000 + case, n)
            x2 = ref.random_scalars(2000 + case, 1)
            ctx.set_msm_affine_rounds(-1)
            f = ctx.fork()
            out = f.worker_commit_open_batch([0, 0], [sc, other], x + x2)
            f.close()
            ok = out[0] == (0, com, y, proof) and out[1][0] == 0 and out[1][1] == ref.msm(srs, other, 8) and \
                (out[1][2], out[1][3]) == ref.open_evals(other, x2, srs, 8)
        if ok and rng.random() < 0.3:
            ctx.set_msm_window(0)
            with native.MultiContext([0]) as mg:
                for k, layout in enumerate((native.LAYOUT_ROWS, native.LAYOUT_POINT_RANGE)):
                    c0 = mg.ctx(0)
                    c0.srs_set_shape(lg, 0)
                    c0.srs_import_row(0, srs)
                    c0.srs_import_g2_tau(TAU[0])
                    mg.set_layout(layout, lg, 0)
                    if layout == native.LAYOUT_ROWS:
                        r = mg.pianist_commit_open([0], sc, x)
                        ok = ok and (r[0][0], r[1][0], r[2][0], r[3], r[4]) == (com, y, proof, com, proof)
                    else:
                        ok = ok and mg.commit_open(0, sc, x) == (com, y, proof) and mg.msm_g1(0, sc) == com
    if not ok:
        bad += 1
        print(f"MISMATCH case {case}: lg={lg} shape={shape} m={m}", flush=True)
print(f"fuzz: {cases} cases, {bad} mismatches")
