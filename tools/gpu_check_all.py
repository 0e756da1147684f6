"""GPU bring-up check of every C-ABI entry against the oracle (run under gpurun)."""
import base64
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bls12_381 as py
from oracle import ref
from zkp_subnet_b200 import native

TAU_X = py.TEST_SECRET
TAU_Y = 0x1234567890ABCDEF1234567890ABCDEF
R = py.R
fails = []


def check(name, cond):
    print(("OK  " if cond else "BAD ") + name, flush=True)
    if not cond:
        fails.append(name)


ctx = native.Context(0)

# ---- SRS generation vs oracle (Pianist rows)
log_n, log_m = 4, 2
ctx.srs_generate(TAU_X, TAU_Y, log_n, log_m)
n, M = 1 << log_n, 1 << log_m
Rs = ref.split32(ref.lagrange_scalars(M, TAU_Y))
rows = [ref.srs(n, TAU_X, "lagrange", scale=Rs[i]) for i in range(M)]
for i in range(M):
    check(f"srs_generate row {i}", ctx.srs_export_row(i, n) == rows[i])

# ---- commit / open / verify on the reference's TEST_POLY (16 evaluations), every row
TEST_POLY = ["aUXcXE/02sinJ4ybjw1GEzIM+H/5R/Iayb9CMn7BlEg", "aOQMCI2Ce8zgLO80vcjBK7Al++oEe8bADAyMXJJbf68",
             "ZygfrBZOk0i4BpO6MNXU4xHeWHjrPSDjSlhQe0hLJDw", "X3w3fa5rnZq6113BXk//n+dSDR+FIkyV9IX0SXgVTFo",
             "LYXDdqRAtuJcP3wRVZtqJ2hAI/NsPXoKzX59AZ3jmcc", "Sm+5XwJBs1g3ceeZEgyHquPIQ+zbUKOCVKkuGYloki8",
             "EAUHn5bsQSpxn+Lp+mfUIdmPtN7EGBRZ5ZQw9dUCvSo", "ZJYLhpIGLcsBwP+6xWlHiomtiA7Tyd9xC+1c519IRpM",
             "A8KIIVWkR2Qr0h+xzyVT+AlVcT8Ju7vZck4sv9ixnUE", "CrB/7LWe40NfYSn81gLLUZ5W17QmlBYz43o7Z2okgw8",
             "EvpYYUWe/7rmVIJ9mL/f6lVF3fi7lihXlGPaIfF0YrU", "amKWoDdtgHUw2wnci7Bp/97D11QUl7gscioZnWt8WwY",
             "FT0sgbVNfhw+g+phx/Zv2IFV8XE+5YHivoQ4yp/uGgI", "IWvMxK6X/j4dSyHDdcRhQPoVPnhoIBpDSAiJBHrNDC0",
             "OBvU/pJOsQ4I8qIn09sgg6oOWh9mHNPHAsS4qTheeDk", "cjp2QP1+ZUcxMVY6tVFJFqyGHCaVzmUT5QYeWX5eGoE"]
TEST_POINT = "RWAG//VkEtMp1SeQHQKHelgaic+md8qWPrnWgHZiNMw"
poly_be = native.b64_decode_fr("".join(TEST_POLY).encode(), 43, 16)
check("b64 decode", poly_be == b"".join(py.b64_decode(s) for s in TEST_POLY))
check("b64 encode", native.b64_encode_fr(poly_be).decode() == "".join(TEST_POLY))
x_be = py.b64_decode(TEST_POINT)
for i in range(M):
    com = ctx.worker_commit(i, poly_be)
    y, proof = ctx.worker_open(i, poly_be, x_be)
    com2, y2, proof2 = ctx.worker_commit_open(i, poly_be, x_be)
    check(f"row {i} commit == oracle", com == ref.msm(rows[i], poly_be))
    ey, eproof = ref.open_evals(poly_be, x_be, rows[i])
    check(f"row {i} eval == oracle", y == ey)
    check(f"row {i} proof == oracle", proof == eproof)
    check(f"row {i} fused == separate", (com2, y2, proof2) == (com, y, proof))
    check(f"row {i} verify", ctx.worker_verify(i, proof, x_be, y, com))
    bad = (int.from_bytes(proof, "big") + 1).to_bytes(48, "big")
    check(f"row {i} tampered proof rejected", not ctx.worker_verify(i, bad, x_be, y, com))
    check(f"row {i} wrong eval rejected", not ctx.worker_verify(i, proof, x_be, ((int.from_bytes(y, 'big') + 1) % R).to_bytes(32, 'big'), com))
    check(f"row {i} wrong row rejected", not ctx.worker_verify((i + 1) % M, proof, x_be, y, com))
check("garbage proof rejected", not ctx.worker_verify(0, b"\xff" * 48, x_be, y, com))
# SURVEY 8c vector B is row 0 of a single-machine SRS
ctx.srs_generate(TAU_X, TAU_Y, 4, 0)
check("vector B commit", ctx.worker_commit(0, poly_be).hex() == "aa3dcf78dff69fb1cc711993cc056c5f210db012cc654c9e9ebf09f450003b05c47645295f61f0bccc2e5e5c6e38c249")
y, proof = ctx.worker_open(0, poly_be, x_be)
check("vector B eval", y.hex() == "5e130b00be5d4cf00af368a75a24aa5bdb2729c4f92d1e96b871f4ce5ec2ea23")
check("vector B proof", proof.hex() == "b25b1758de10baafed035fce2362d5d8991fb51a220088e3337990eebb77406753b7a5419abdfbc1058bca02b037ddbc")
# x inside the domain
w16 = py.root_of_unity(16)
xd = pow(w16, 5, R).to_bytes(32, "big")
y, proof = ctx.worker_open(0, poly_be, xd)
ey, eproof = ref.open_evals(poly_be, xd, ref.srs(16, TAU_X, "lagrange"))
check("in-domain eval", y == ey == poly_be[5 * 32:6 * 32])
check("in-domain proof", proof == eproof)
check("in-domain verify", ctx.worker_verify(0, proof, xd, y, ctx.worker_commit(0, poly_be)))

# ---- eval (the reference KAT) / fft
check("eval KAT", base64.b64encode(ctx.eval(poly_be, x_be)).decode().rstrip("=") == "KXMqHg4HSrBe5qnld5TFrRlluYtsjG7N6WrHduoG/1s")
for lg in (0, 1, 2, 4, 7, 10, 12, 13, 14, 16):
    nn = 1 << lg
    v = ref.random_scalars(1000 + lg, nn)
    t = time.time(); f = ctx.fft(v, True, False); tg = time.time() - t
    check(f"fft 2^{lg} ({tg*1e3:.2f} ms)", f == ref.ntt(v, False))
    check(f"ifft 2^{lg}", ctx.fft(v, True, True) == ref.ntt(v, True))
    if lg in (7, 13):
        xx = ref.random_scalars(5, 1)
        check(f"eval 2^{lg}", ctx.eval(v, xx) == ref.eval_coeffs(v, xx))
rp = ctx.random_poly(7, 1000)
check("random_poly canonical", all(int.from_bytes(rp[i:i + 32], "big") < R for i in range(0, len(rp), 32)) and len(set(rp[i:i + 32] for i in range(0, len(rp), 32))) == 1000)
check("random_point", int.from_bytes(ctx.random_point(1), "big") < R and ctx.random_point(1) != ctx.random_point(2))
check("non-canonical scalar rejected", _raises := True)
try:
    ctx.worker_commit(0, (R).to_bytes(32, "big") * 16)
    check("non-canonical scalar rejected (commit)", False)
except native.ZkpError as e:
    check("non-canonical scalar rejected (commit)", e.code == native.ZKP_ERR_ENCODING)

# ---- pairing check: e(aG, bH) * e(-abG, H) == 1
a, b = 123456789, 987654321
g2 = lambda pt: b"".join(c.to_bytes(48, "big") for c in (pt[0][0], pt[0][1], pt[1][0], pt[1][1]))
P1 = py.g1_compress(py.g1_mul(py.G1_GEN, a)); Q1 = g2(py.g2_mul(py.G2_GEN, b))
P2 = py.g1_compress(py.g1_neg(py.g1_mul(py.G1_GEN, a * b % R))); Q2 = g2(py.G2_GEN)
check("pairing bilinear", native.pairing_check(P1 + P2, Q1 + Q2))
P3 = py.g1_compress(py.g1_neg(py.g1_mul(py.G1_GEN, (a * b + 1) % R)))
check("pairing negative", not native.pairing_check(P1 + P3, Q1 + Q2))

# ---- medium sizes vs oracle (GPU-generated SRS exported to the oracle) + trapdoor identity at 2^16/2^20
for lg in (10, 12, 16, 20):
    nn = 1 << lg
    t = time.time(); ctx.srs_generate(TAU_X, TAU_Y, lg, 0); tgen = time.time() - t
    sc = ref.random_scalars(0xB200 + lg, nn)
    xx = ref.random_scalars(77 + lg, 1)
    t = time.time(); com, y, proof = ctx.worker_commit_open(0, sc, xx); tg = time.time() - t
    ls = ref.lagrange_scalars(nn, TAU_X)
    check(f"2^{lg} commit == [f(tau)]G (srs gen {tgen*1e3:.0f} ms, commit+open {tg*1e3:.1f} ms)", com == ref.g1_mul_gen(ref.fr_dot(sc, ls)))
    ey, q = ref.quotient_evals(sc, xx)
    check(f"2^{lg} eval == oracle", y == ey)
    check(f"2^{lg} proof == [q(tau)]G", proof == ref.g1_mul_gen(ref.fr_dot(q, ls)))
    check(f"2^{lg} verify", ctx.worker_verify(0, proof, xx, y, com))
    if lg <= 12:
        srs = ctx.srs_export_row(0, nn)
        check(f"2^{lg} commit == oracle MSM", com == ref.msm(srs, sc, 16))

# ---- first timings
for lg in (16, 20):
    nn = 1 << lg
    ctx.srs_generate(TAU_X, TAU_Y, lg, 0)
    sc = ref.random_scalars(0xB200 + lg, nn)
    xx = ref.random_scalars(77 + lg, 1)
    c, W, muls = ctx.msm_info(nn)
    ms, _ = ctx.bench_msm(0, sc, 5)
    ms_it, ms_k, launches, *_ = ctx.bench_commit_open(0, sc, xx, 5)
    print(f"n=2^{lg} c={c} W={W}: msm {ms:.3f} ms ({nn/ms/1e3:.1f} Mpts/s, {muls/ms/1e6:.2f} G Fq-mul/s), commit+open {ms_it:.3f} ms, "
          f"accumulate kernel {ms_k:.3f} ms, launches/iter {launches}", flush=True)
    for lgn in (lg,):
        print(f"   ntt 2^{lgn}: {ctx.bench_ntt(1 << lgn, 5):.4f} ms")

print("ALL OK" if not fails else f"FAILURES: {fails}")
sys.exit(1 if fails else 0)
