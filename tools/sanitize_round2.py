"""Tiny run of every kernel path added in round 2, for compute-sanitizer (memcheck / racecheck):
compute-sanitizer --tool memcheck python tools/sanitize_round2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_subnet_b200 import native
TX, TY = 1927409816240961209460912649124, 0x1234567890ABCDEF1234567890ABCDEF
lg = 8
n = 1 << lg
with native.Context(0) as ctx:
    ctx.srs_generate(TX, TY, lg, 1)
    polys = [ctx.random_poly(10 + k, n) for k in range(3)]
    xs = [ctx.random_point(k) for k in range(3)]
    single = [ctx.worker_commit_open(k % 2, polys[k], xs[k]) for k in range(3)]
    ctx.set_fuse(1)
    assert [ctx.worker_commit_open(k % 2, polys[k], xs[k]) for k in range(3)] == single
    ctx.set_fuse(-1)
    out = ctx.worker_commit_open_batch([0, 1, 0], polys, b"".join(xs))
    assert [tuple(o[1:]) for o in out] == single
    f = ctx.fork()
    assert f.worker_commit_open(1, polys[1], xs[1]) == single[1]
    f.close()
    ctx.set_poly_form(True)
    ctx.worker_commit_open(0, polys[0], xs[0])
    ctx.set_poly_form(False)
    rows = [ctx.srs_export_row(i, n) for i in range(2)]
    comp = [ctx.srs_export_row_compressed(i, n) for i in range(2)]
    ctx.srs_generate_monomial2(TX, TY, lg, 1)
    ctx.srs_monomial_to_lagrange()
    assert [ctx.srs_export_row(i, n) for i in range(2)] == rows
    ctx.srs_import_row_compressed(0, comp[0])
    assert ctx.srs_export_row(0, n) == rows[0]
    for tma in (True, False):
        ctx.set_ntt_tma(tma)
        v = ctx.random_poly(3, 1 << 18)
        assert ctx.fft(ctx.fft(v, True, False), True, True) == v
# the coset opening (rows of >= 512 elements): single requests and a batch, against the general kernels
for lg2 in (9, 11):
    n2 = 1 << lg2
    with native.Context(0) as ctx:
        ctx.srs_generate(TX, TY, lg2, 0)
        ps = [ctx.random_poly(20 + k, n2) for k in range(3)]
        x3 = [ctx.random_point(30 + k) for k in range(3)]
        a = [ctx.worker_commit_open(0, ps[k], x3[k]) for k in range(3)]
        b = ctx.worker_commit_open_batch([0, 0, 0], ps, b"".join(x3))
        ctx.set_open_coset(False)
        assert [ctx.worker_commit_open(0, ps[k], x3[k]) for k in range(3)] == a
        assert [tuple(o[1:]) for o in b] == a and ctx.worker_commit_open_batch([0, 0, 0], ps, b"".join(x3)) == b
with native.MultiContext([0]) as mg:
    mg.srs_generate(TX, TY, lg, 1, native.LAYOUT_POINT_RANGE)
    assert mg.commit_open(1, polys[1], xs[1]) == single[1]
with native.MultiContext([0]) as mg:
    mg.srs_generate(TX, TY, lg, 1, native.LAYOUT_ROWS)
    r = mg.pianist_commit_open([0, 1], polys[0] + polys[1], xs[0])
    assert r[0][0] == single[0][0]
print("sanitize run OK")
