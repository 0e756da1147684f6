"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs -- bit-exact (integer work).  Small sizes are compared with the oracle's own MSM / NTT /
quotient; sizes the oracle's MSM cannot finish in seconds use size-independent properties: the trapdoor
identity commit == [sum_j f_j L_j(tau)]_1 (one oracle scalar product + one scalar multiplication),
proof == [sum_j q_j L_j(tau)]_1, and the pairing check of every proof."""
import pytest

from oracle import bls12_381 as o
from oracle import ref
from zkp_subnet_b200 import native

pytestmark = pytest.mark.gpu

R = o.R
TAU_X = o.TEST_SECRET
TAU_Y = 0x1234567890ABCDEF1234567890ABCDEF


def adversarial(n):
    return {
        "random": ref.random_scalars(0xB200 + n, n),
        "zeros": bytes(32 * n),
        "ones": ref.join32([1] * n),
        "r_minus_1": ref.join32([R - 1] * n),
        "single": ref.join32([0] * (n - 1) + [12345]),
        "small": ref.join32([i % 7 for i in range(n)]),
        "top_bits": ref.join32([(1 << 254) + i for i in range(n)]),
        "half_window": ref.join32([(1 << 15) | (1 << 31) | (1 << 47) for _ in range(n)]),
    }


@pytest.mark.parametrize("log_n", [0, 1, 4, 8, 12])
def test_msm_vs_oracle(gpu_ctx, log_n):
    n = 1 << log_n
    srs = ref.srs(n, TAU_X, "lagrange")
    gpu_ctx.srs_set_shape(log_n, 0)
    gpu_ctx.srs_import_row(0, srs)
    assert gpu_ctx.srs_export_row(0, n) == srs
    for name, sc in adversarial(n).items():
        assert gpu_ctx.msm_g1(0, sc) == ref.msm(srs, sc, 8), name
    if n > 4:  # ragged: fewer scalars than points
        sc = ref.random_scalars(99, n - 3)
        assert gpu_ctx.msm_g1(0, sc) == ref.msm(srs, sc, 8)


def test_msm_special_points(gpu_ctx):
    # SRS rows with repeated points, P and -P, and infinity: exercises the doubling / cancellation
    # branches of the bucket adds (SURVEY.md section 7 "adversarial")
    n = 64
    g = o.G1_GEN
    pts = [g] * 16 + [o.g1_neg(g)] * 16 + [None] * 8 + [o.g1_mul(g, k + 2) for k in range(24)]
    raw = b"".join((b"\x40" + bytes(95)) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big") for p in pts)
    gpu_ctx.srs_set_shape(6, 0)
    gpu_ctx.srs_import_row(0, raw)
    for sc in (ref.join32([5] * n), ref.join32([1] * n), ref.random_scalars(3, n), ref.join32([R - 1] * 32 + [7] * 32)):
        assert gpu_ctx.msm_g1(0, sc) == ref.msm(raw, sc, 1)


@pytest.mark.parametrize("tables", [True, False])
def test_msm_exceptional_additions_on_a_loose_accumulator(gpu_ctx, tables):
    """The level-0 accumulation adds with lazy reductions (G1Xyzz::madd_lazy): its test for equal / opposite
    operands has to work on unreduced coordinates.  Rows built so that, with every scalar equal (one bucket per
    window, entries in index order), the running sum meets a point EQUAL to itself after a few additions (doubling
    branch), later its NEGATIVE (the sum becomes infinity), and then carries on from infinity."""
    g = o.G1_GEN
    ks = [3, 5, 11]                      # acc = 19 G after three additions ...
    ks += [19]                           # ... + 19 G: doubling while the accumulator is loose -> 38 G
    ks += [7, 2]                         # 47 G
    ks += [-47, 13]                      # cancellation -> infinity, restart from infinity (end of the first 8-entry slice)
    ks += [13, 13, 26, -52, 9]           # next slice: 13 G + 13 G (doubling from an affine start), + 26 G (doubling again), - 52 G, + 9 G
    ks += [k + 100 for k in range(64 - len(ks))]
    pts = [o.g1_mul(g, k) if k > 0 else o.g1_neg(o.g1_mul(g, -k)) for k in ks]
    raw = b"".join(p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big") for p in pts)
    gpu_ctx.set_msm_mode(tables)
    try:
        gpu_ctx.srs_set_shape(6, 0)
        gpu_ctx.srs_import_row(0, raw)
        for sc in (ref.join32([1] * 64), ref.join32([0x1234567] * 64), ref.join32([R - 2] * 64),
                   ref.join32([1] * 12 + [0] * 52), ref.random_scalars(11, 64)):
            assert gpu_ctx.msm_g1(0, sc) == ref.msm(raw, sc, 1)
    finally:
        gpu_ctx.set_msm_mode(True)
        gpu_ctx.set_msm_window(0)


@pytest.mark.parametrize("rounds", [1, 2, 3, 6])
def test_batched_affine_rounds_vs_oracle(gpu_ctx, rounds):
    """Batched-affine pairwise rounds in front of the XYZZ accumulation (msm_affine.cuh): forced on at small
    sizes and checked against the oracle for every adversarial scalar set (one bucket holding every entry,
    empty lists, ragged lengths), both table modes, and an SRS with repeated points, P / -P pairs and
    infinities (the doubling, cancellation and infinity branches of the affine addition)."""
    try:
        gpu_ctx.set_msm_affine_rounds(rounds)
        for log_n in (0, 1, 3, 6, 10):
            n = 1 << log_n
            srs = ref.srs(n, TAU_X, "lagrange")
            gpu_ctx.srs_set_shape(log_n, 0)
            gpu_ctx.srs_import_row(0, srs)
            for name, sc in adversarial(n).items():
                exp = ref.msm(srs, sc, 8)
                for mode in (True, False):
                    gpu_ctx.set_msm_mode(mode)
                    assert gpu_ctx.msm_g1(0, sc) == exp, (log_n, name, mode)
            gpu_ctx.set_msm_mode(True)
            if n > 4:
                sc = ref.random_scalars(98, n - 3)
                assert gpu_ctx.msm_g1(0, sc) == ref.msm(srs, sc, 8)
        n = 64
        g = o.G1_GEN
        pts = [g] * 16 + [o.g1_neg(g)] * 16 + [None] * 8 + [o.g1_mul(g, k + 2) for k in range(24)]
        raw = b"".join((b"\x40" + bytes(95)) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big") for p in pts)
        gpu_ctx.srs_set_shape(6, 0)
        gpu_ctx.srs_import_row(0, raw)
        for sc in (ref.join32([5] * n), ref.join32([1] * n), ref.random_scalars(3, n), ref.join32([R - 1] * 32 + [7] * 32)):
            for c in (0, 3, 5):  # small windows: many entries per bucket, so all rounds have pairs to add
                gpu_ctx.set_msm_window(c)
                assert gpu_ctx.msm_g1(0, sc) == ref.msm(raw, sc, 1), (c,)
    finally:
        gpu_ctx.set_msm_affine_rounds(-1)
        gpu_ctx.set_msm_mode(True)
        gpu_ctx.set_msm_window(0)


def test_batched_affine_commit_open_full_size(gpu_ctx):
    """2^20 commit+open: default (no batched-affine rounds), 2 and 3 rounds forced on give identical bytes, and
    the proof verifies."""
    log_n = 20
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    f = gpu_ctx.random_poly(0xAFF1, n)
    x = gpu_ctx.random_point(0xAFF2)
    try:
        auto = gpu_ctx.worker_commit_open(0, f, x)
        gpu_ctx.set_msm_affine_rounds(2)
        assert gpu_ctx.worker_commit_open(0, f, x) == auto
        gpu_ctx.set_msm_affine_rounds(3)
        assert gpu_ctx.worker_commit_open(0, f, x) == auto
        assert gpu_ctx.worker_commit(0, f) == auto[0]
    finally:
        gpu_ctx.set_msm_affine_rounds(-1)
    assert gpu_ctx.worker_verify(0, auto[2], x, auto[1], auto[0])


def test_window_override_is_result_invariant(gpu_ctx):
    n = 1 << 10
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 10, 0)
    sc = ref.random_scalars(11, n)
    base = gpu_ctx.msm_g1(0, sc)
    try:
        for c in (4, 7, 11, 13, 16):
            gpu_ctx.set_msm_window(c)
            assert gpu_ctx.msm_g1(0, sc) == base, c
    finally:
        gpu_ctx.set_msm_window(0)


def test_fixed_base_tables_match_classic(gpu_ctx):
    # the per-row tables [2^(c w)] P_i (shared buckets, no window fold) and the classic per-window
    # buckets must give identical bytes, for every window width
    n = 1 << 10
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 10, 1)
    try:
        for row in (0, 1):
            srs = gpu_ctx.srs_export_row(row, n)
            for name, sc in adversarial(n).items():
                exp = ref.msm(srs, sc, 8)
                for mode in (True, False):
                    gpu_ctx.set_msm_mode(mode)
                    for c in (0, 6, 10, 12):
                        gpu_ctx.set_msm_window(c)
                        assert gpu_ctx.msm_g1(row, sc) == exp, (row, name, mode, c)
            gpu_ctx.set_msm_mode(True)
            gpu_ctx.set_msm_window(0)
            sc = ref.random_scalars(5, n - 7)  # ragged prefix with tables
            assert gpu_ctx.msm_g1(row, sc) == ref.msm(srs, sc, 8)
    finally:
        gpu_ctx.set_msm_mode(True)
        gpu_ctx.set_msm_window(0)


def test_srs_generate_matches_oracle(gpu_ctx):
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 2)
    Rs = ref.split32(ref.lagrange_scalars(4, TAU_Y))
    for i in range(4):
        assert gpu_ctx.srs_export_row(i, 16) == ref.srs(16, TAU_X, "lagrange", scale=Rs[i])
    # point-range shards tile the full row
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 6, 0)
    full = gpu_ctx.srs_export_row(0, 64)
    for s in range(4):
        gpu_ctx.srs_generate_shard(TAU_X, TAU_Y, 6, 0, s, 2)
        assert gpu_ctx.srs_export_row(0, 16) == full[s * 16 * 96:(s + 1) * 16 * 96]


def test_golden_vectors(gpu_ctx, golden):
    poly = b"".join(o.b64_decode(s) for s in golden["test_poly"])
    x = o.b64_decode(golden["test_point"])
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 0)
    B = golden["B_eval_form"]
    assert gpu_ctx.worker_commit(0, poly).hex() == B["commitment"]
    y, proof = gpu_ctx.worker_open(0, poly, x)
    assert o.fr_to_b64(int.from_bytes(y, "big")) == B["eval"] and proof.hex() == B["proof"]
    assert gpu_ctx.worker_commit_open(0, poly, x) == (bytes.fromhex(B["commitment"]), y, proof)
    assert gpu_ctx.worker_verify(0, proof, x, y, bytes.fromhex(B["commitment"]))
    D = golden["B_in_domain"]
    y, proof = gpu_ctx.worker_open(0, poly, o.b64_decode(D["x"]))
    assert o.fr_to_b64(int.from_bytes(y, "big")) == D["eval"] and proof.hex() == D["proof"]
    # coefficient-form Horner: the reference's own known answer
    assert o.fr_to_b64(int.from_bytes(gpu_ctx.eval(poly, x), "big")) == golden["test_eval"]
    assert [o.fr_to_b64(v) for v in ref.split32(gpu_ctx.fft(poly, True, False))] == golden["ntt16"]
    assert [o.fr_to_b64(v) for v in ref.split32(gpu_ctx.fft(poly, True, True))] == golden["intt16"]
    # Pianist rows (scale 6 / machines_scale 2: the reference's unit-test shape)
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 2)
    for rec in golden["pianist_4x16"]:
        i = rec["row"]
        com, y, proof = gpu_ctx.worker_commit_open(i, poly, x)
        assert (com.hex(), proof.hex(), o.fr_to_b64(int.from_bytes(y, "big"))) == (rec["commitment"], rec["proof"], rec["eval"])
        assert gpu_ctx.worker_verify(i, proof, x, y, com)
        assert not gpu_ctx.worker_verify((i + 1) % 4, proof, x, y, com)
        tampered = (int.from_bytes(proof, "big") + 1).to_bytes(48, "big")  # reference tests/test_validator.py:79-86
        assert not gpu_ctx.worker_verify(i, tampered, x, y, com)


def test_pianist_master_golden(gpu_ctx, golden):
    """Config 5 in miniature: 4 sub-polynomials on 4 SRS rows, aggregated commitment, X-opening at alpha and the
    master's Y-direction opening at beta -- bytes against the oracle's coefficient-form restatement, and the
    bivariate pairing check."""
    M = golden["pianist_master_4x16"]
    poly = [o.b64_decode(s) for s in golden["test_poly"]]
    x = o.b64_decode(golden["test_point"])
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 2)
    coms, ys, proofs = [], [], []
    for i, rec in enumerate(M["rows"]):
        fi = b"".join(poly[i:] + poly[:i])
        com, y, proof = gpu_ctx.worker_commit_open(i, fi, x)
        assert (com.hex(), o.fr_to_b64(int.from_bytes(y, "big")), proof.hex()) == (rec["commitment"], rec["eval"], rec["proof"])
        coms.append(com); ys.append(y); proofs.append(proof)
    com, pix = native.g1_sum(b"".join(coms)), native.g1_sum(b"".join(proofs))
    assert (com.hex(), pix.hex()) == (M["commitment"], M["proof_x"])
    for bk, zk, pk in (("beta", "z", "proof_y"), ("beta_in_domain", "z_in_domain", "proof_y_in_domain")):
        beta = o.b64_decode(M[bk])
        z, piy = gpu_ctx.master_open_y(b"".join(ys), beta)
        assert (o.fr_to_b64(int.from_bytes(z, "big")), piy.hex()) == (M[zk], M[pk])
        assert gpu_ctx.master_verify(com, pix, piy, x, beta, z)
        bad_z = ((int.from_bytes(z, "big") + 1) % o.R).to_bytes(32, "big")
        assert not gpu_ctx.master_verify(com, pix, piy, x, beta, bad_z)
        assert not gpu_ctx.master_verify(com, piy, pix, x, beta, z)          # proofs swapped
        assert not gpu_ctx.master_verify(coms[0], pix, piy, x, beta, z)      # a worker's commitment, not the aggregate
        assert not gpu_ctx.master_verify(com, pix, b"\xff" * 48, x, beta, z)  # malformed point: invalid, not an error
    with pytest.raises(native.ZkpError):
        gpu_ctx.master_open_y(b"".join(ys[:3]), x)  # one evaluation per worker
    # M = 1: g is constant, the Y-proof is the point at infinity
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 0)
    com, y, proof = gpu_ctx.worker_commit_open(0, b"".join(poly), x)
    z, piy = gpu_ctx.master_open_y(y, o.b64_decode(M["beta"]))
    assert z == y and piy.hex() == golden["g1_encodings"]["inf"]
    assert gpu_ctx.master_verify(com, proof, piy, x, o.b64_decode(M["beta"]), z)


@pytest.mark.parametrize("log_n", [0, 1, 3, 9, 12, 13, 15, 17])
def test_ntt_vs_oracle(gpu_ctx, log_n):
    n = 1 << log_n
    v = ref.random_scalars(1000 + log_n, n)
    f = gpu_ctx.fft(v, True, False)
    assert f == ref.ntt(v, False)
    assert gpu_ctx.fft(v, False, True) == ref.ntt(v, True)
    assert gpu_ctx.fft(f, True, True) == v  # round trip


def test_ntt_large_properties(gpu_ctx):
    # 2^20: linearity + round trip + agreement with Horner at a domain point (size-independent checks)
    n = 1 << 20
    a, b = ref.random_scalars(1, n), ref.random_scalars(2, n)
    fa, fb = gpu_ctx.fft(a), gpu_ctx.fft(b)
    assert gpu_ctx.fft(fa, True, True) == a
    s = ref.join32([(x + y) % R for x, y in zip(ref.split32(a[:32 * 64]), ref.split32(b[:32 * 64]))])
    ab = ref.join32([(x + y) % R for x, y in zip(ref.split32(a), ref.split32(b))])
    fab = gpu_ctx.fft(ab)
    assert ref.split32(fab[:32 * 64]) == [(x + y) % R for x, y in zip(ref.split32(fa[:32 * 64]), ref.split32(fb[:32 * 64]))]
    assert len(s) == 32 * 64
    w = o.root_of_unity(n)
    for k in (0, 1, 12345, n - 1):
        assert gpu_ctx.eval(a, pow(w, k, R).to_bytes(32, "big")) == fa[32 * k:32 * k + 32]


@pytest.mark.parametrize("log_n", [5, 10, 13])
def test_eval_vs_oracle(gpu_ctx, log_n):
    n = (1 << log_n) - 3  # ragged length
    c = ref.random_scalars(log_n, n)
    for seed in (1, 2):
        x = ref.random_scalars(500 + seed, 1)
        assert gpu_ctx.eval(c, x) == ref.eval_coeffs(c, x)
    assert gpu_ctx.eval(c, bytes(32)) == c[:32]


@pytest.mark.parametrize("log_n", [10, 12])
def test_commit_open_vs_oracle(gpu_ctx, log_n):
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    srs = gpu_ctx.srs_export_row(0, n)
    assert srs == ref.srs(n, TAU_X, "lagrange")
    sc = ref.random_scalars(0xB200 + log_n, n)
    x = ref.random_scalars(77, 1)
    com, y, proof = gpu_ctx.worker_commit_open(0, sc, x)
    assert com == ref.msm(srs, sc, 8)
    ey, eproof = ref.open_evals(sc, x, srs, 8)
    assert (y, proof) == (ey, eproof)
    assert gpu_ctx.worker_open(0, sc, x) == (y, proof) and gpu_ctx.worker_commit(0, sc) == com
    assert gpu_ctx.worker_verify(0, proof, x, y, com)
    # constant polynomial: quotient is zero, proof is the point at infinity
    const = ref.join32([42] * n)
    y, proof = gpu_ctx.worker_open(0, const, x)
    assert int.from_bytes(y, "big") == 42 and proof.hex() == "c0" + "00" * 47
    assert gpu_ctx.worker_verify(0, proof, x, y, gpu_ctx.worker_commit(0, const))


@pytest.mark.parametrize("log_n", [16, 20])
def test_full_size_trapdoor_identity(gpu_ctx, log_n):
    # BASELINE configs 2 and 3: sizes beyond the oracle's MSM, checked through the trapdoor
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    sc = ref.random_scalars(0xB200 + log_n, n)
    x = ref.random_scalars(77 + log_n, 1)
    com, y, proof = gpu_ctx.worker_commit_open(0, sc, x)
    ls = ref.lagrange_scalars(n, TAU_X)
    assert com == ref.g1_mul_gen(ref.fr_dot(sc, ls))
    ey, q = ref.quotient_evals(sc, x)
    assert y == ey
    assert proof == ref.g1_mul_gen(ref.fr_dot(q, ls))
    assert gpu_ctx.worker_verify(0, proof, x, y, com)
    assert not gpu_ctx.worker_verify(0, com, x, y, proof)
    # evaluation-form commitment == coefficient-form commitment: iNTT on the GPU, monomial trapdoor on the CPU
    coeffs = gpu_ctx.fft(sc, True, True)
    f_tau = ref.eval_coeffs(coeffs, TAU_X.to_bytes(32, "big"))
    assert com == ref.g1_mul_gen(f_tau)


def test_commit_open_2p24_trapdoor_identity(gpu_ctx):
    """BASELINE configs[3] and the north_star size: one SRS row of 2^24 points (1.5 GiB affine, 48 GiB of fixed-base
    tables), commitment and opening proof checked through the trapdoor (commit == [sum f_j L_j(tau)]_1,
    proof == [sum q_j L_j(tau)]_1 with q from the oracle's quotient), y against the oracle, pairing check on the
    host.  The level-0 accumulation kernel runs its longest slices here."""
    log_n = 24
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    sc = gpu_ctx.random_poly(0xB200 + 4, n)
    x = ref.random_scalars(77 + log_n, 1)
    com, y, proof = gpu_ctx.worker_commit_open(0, sc, x)
    ls = ref.lagrange_scalars(n, TAU_X)
    assert com == ref.g1_mul_gen(ref.fr_dot(sc, ls))
    assert com == gpu_ctx.worker_commit(0, sc)
    ey, q = ref.quotient_evals(sc, x)
    assert y == ey
    assert proof == ref.g1_mul_gen(ref.fr_dot(q, ls))
    assert gpu_ctx.worker_verify(0, proof, x, y, com)
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 0)  # drop the 2^24 row and its tables before the next test


def test_msm_sort_paths_and_linearity_full_size(gpu_ctx):
    """Size-independent properties at 2^20 (BASELINE configs[2]): the hand-written bucket sort, the library radix sort
    and the classic per-window buckets give the same bytes for uniform, constant and sparse scalars; the commitment is
    linear: commit(f + g) = commit(f) + commit(g) (field addition on the CPU, group addition through zkp_g1_sum)."""
    log_n = 20
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    f = gpu_ctx.random_poly(0xF00D, n)
    g = gpu_ctx.random_poly(0xBEEF, n)
    const = f[:32] * n
    sparse = b"".join(f[32 * i:32 * i + 32] if i % 97 == 0 else bytes(32) for i in range(n))
    try:
        for sc in (f, const, sparse):
            gpu_ctx.set_msm_sort(1)
            a = gpu_ctx.msm_g1(0, sc)
            gpu_ctx.set_msm_sort(0)
            assert gpu_ctx.msm_g1(0, sc) == a
        gpu_ctx.set_msm_mode(False)
        gpu_ctx.set_msm_sort(1)
        classic = gpu_ctx.msm_g1(0, f)
        gpu_ctx.set_msm_mode(True)
        cf, cg = gpu_ctx.msm_g1(0, f), gpu_ctx.msm_g1(0, g)
        assert classic == cf
        fi, gi = ref.split32(f), ref.split32(g)
        s = ref.join32([(x + y) % R for x, y in zip(fi, gi)])
        assert gpu_ctx.msm_g1(0, s) == native.g1_sum(cf + cg)
    finally:
        gpu_ctx.set_msm_sort(2)
        gpu_ctx.set_msm_mode(True)


@pytest.mark.parametrize("log_n", [4, 12, 20])
def test_commit_path_a_equals_path_b(gpu_ctx, log_n, golden):
    """BASELINE configs[2]: path A = MSM of the evaluations over the Lagrange SRS; path B = iNTT on the GPU, then
    MSM of the coefficients over the monomial SRS [tau^j]_1.  Two SRS forms, two scalar vectors, identical bytes."""
    n = 1 << log_n
    evals = b"".join(o.b64_decode(s) for s in golden["test_poly"]) if log_n == 4 else gpu_ctx.random_poly(0xAB + log_n, n)
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    path_a = gpu_ctx.worker_commit(0, evals)
    coeffs = gpu_ctx.fft(evals, True, True)
    gpu_ctx.srs_generate_monomial(TAU_X, log_n)
    path_b = gpu_ctx.msm_g1(0, coeffs)
    assert path_a == path_b
    if log_n == 4:
        assert path_a.hex() == golden["B_eval_form"]["commitment"]
        # and the monomial row itself against the oracle, plus the coefficient-form golden vector A
        assert gpu_ctx.srs_export_row(0, n) == ref.srs(n, TAU_X, "monomial")
        assert gpu_ctx.msm_g1(0, evals).hex() == golden["A_coeff_form"]["commitment"]


def test_sharded_commit_combines_to_full(gpu_ctx):
    # point-range sharding: partial commitments of the 4 shards sum to the commitment
    log_n, log_s = 12, 2
    n, S = 1 << log_n, 1 << log_s
    sc = ref.random_scalars(31337, n)
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    full = gpu_ctx.worker_commit(0, sc)
    parts = b""
    for s in range(S):
        gpu_ctx.srs_generate_shard(TAU_X, TAU_Y, log_n, 0, s, log_s)
        parts += gpu_ctx.worker_commit(0, sc[s * (n // S) * 32:(s + 1) * (n // S) * 32])
        with pytest.raises(native.ZkpError):
            gpu_ctx.worker_open(0, sc[:(n // S) * 32], bytes(31) + b"\x05")
    assert native.g1_sum(parts) == full


@pytest.mark.parametrize("log_n,log_shards", [(6, 2), (12, 3), (16, 1)])
def test_sharded_open_combines_to_full(gpu_ctx, log_n, log_shards):
    """Point-range shards (generated one after the other on this GPU): partial barycentric sums -> y, partial
    proofs -> proof; both equal the unsharded opening, bit for bit."""
    n, G = 1 << log_n, 1 << log_shards
    f = ref.random_scalars(0x5A + log_n, n)
    x = ref.random_scalars(0x5B, 1)
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    com, y, proof = gpu_ctx.worker_commit_open(0, f, x)
    per = n // G * 32
    partial_sums, coms = [], []
    for g in range(G):
        gpu_ctx.srs_generate_shard(TAU_X, TAU_Y, log_n, 0, g, log_shards)
        partial_sums.append(gpu_ctx.shard_eval_partial(0, f[g * per:(g + 1) * per], x))
        coms.append(gpu_ctx.worker_commit(0, f[g * per:(g + 1) * per]))
    assert native.shard_eval_combine(b"".join(partial_sums), log_n, x) == y
    assert native.g1_sum(b"".join(coms)) == com
    proofs = []
    for g in range(G):
        gpu_ctx.srs_generate_shard(TAU_X, TAU_Y, log_n, 0, g, log_shards)
        proofs.append(gpu_ctx.shard_open_partial(0, f[g * per:(g + 1) * per], x, y))
        with pytest.raises(native.ZkpError):  # the single-device opening is refused on a shard
            gpu_ctx.worker_open(0, f[g * per:(g + 1) * per], x)
    assert native.g1_sum(b"".join(proofs)) == proof
    # x inside this shard's part of the domain is refused explicitly
    w = o.root_of_unity(n)
    xin = pow(w, n - 1, o.R).to_bytes(32, "big")  # last shard holds index n - 1
    with pytest.raises(native.ZkpError, match="inside the domain"):
        gpu_ctx.shard_eval_partial(0, f[(G - 1) * per:], xin)


def test_errors_and_edge_inputs(gpu_ctx):
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 0)
    with pytest.raises(native.ZkpError) as e:
        gpu_ctx.worker_commit(0, R.to_bytes(32, "big") * 16)  # non-canonical scalar
    assert e.value.code == native.ZKP_ERR_ENCODING
    with pytest.raises(native.ZkpError):
        gpu_ctx.worker_commit(1, bytes(32 * 16))  # row out of range
    with pytest.raises(native.ZkpError):
        gpu_ctx.worker_commit(0, bytes(32 * 17))  # longer than the row
    with pytest.raises(native.ZkpError):
        gpu_ctx.worker_open(0, bytes(32 * 8), bytes(32))  # opening needs a full row
    with pytest.raises(native.ZkpError):
        gpu_ctx.fft(bytes(32 * 3))  # not a power of two
    assert not gpu_ctx.worker_verify(0, b"\xff" * 48, bytes(32), bytes(32), b"\xc0" + bytes(47))
    assert not gpu_ctx.worker_verify(0, b"\xc0" + bytes(47), b"\xff" * 32, bytes(32), b"\xc0" + bytes(47))
    rp = gpu_ctx.random_poly(7, 4096)
    vals = ref.split32(rp)
    assert all(v < R for v in vals) and len(set(vals)) == 4096
    assert gpu_ctx.random_poly(7, 16) == rp[:512] and gpu_ctx.random_poly(8, 16) != rp[:512]


def test_srs_file_roundtrip(gpu_ctx, tmp_path, golden):
    gpu_ctx.srs_generate(TAU_X, TAU_Y, 4, 2)
    path = str(tmp_path / "srs.bin")
    gpu_ctx.srs_save(path)
    rows = [gpu_ctx.srs_export_row(i, 16) for i in range(4)]
    other = native.Context(0)  # a second context on the same GPU (miner + validator on one host)
    try:
        other.srs_load(path)
        assert other.srs_shape() == (4, 2)
        assert [other.srs_export_row(i, 16) for i in range(4)] == rows
        poly = b"".join(o.b64_decode(s) for s in golden["test_poly"])
        x = o.b64_decode(golden["test_point"])
        rec = golden["pianist_4x16"][3]
        com, y, proof = other.worker_commit_open(3, poly, x)
        assert com.hex() == rec["commitment"] and proof.hex() == rec["proof"]
        assert other.worker_verify(3, proof, x, y, com) and gpu_ctx.worker_verify(3, proof, x, y, com)
        # [tau_y]_2 travels in the file (format version 2): the master check works on the loaded SRS
        M = golden["pianist_master_4x16"]
        assert other.master_verify(bytes.fromhex(M["commitment"]), bytes.fromhex(M["proof_x"]), bytes.fromhex(M["proof_y"]), x,
                                   o.b64_decode(M["beta"]), o.b64_decode(M["z"]))
    finally:
        other.close()


# Public constants of BLS12-381 (ZCash compressed encodings of the G1 generator, its double and its negative, and of
# the point at infinity): the same bytes appear in every implementation's own tests.
G1_GEN_HEX = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
G1_2GEN_HEX = "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e"
G1_NEG_GEN_HEX = "b7" + G1_GEN_HEX[2:]
G1_INF_HEX = "c0" + "00" * 47


@pytest.mark.parametrize("log_n", [0, 1, 4, 9, 12, 16, 20])
def test_known_answers_that_need_no_trapdoor(gpu_ctx, log_n):
    """Known answers independent of tau, of the oracle and of any implementation: the Lagrange basis sums to one, so
    for EVERY trusted setup commit(1, 1, ..., 1) is the G1 generator, commit(2, ...) its double, commit(r - 1, ...) its
    negative and commit(0, ...) the point at infinity -- compared here with the public compressed encodings; a constant
    polynomial opens to itself with the point at infinity as proof, at points inside and outside the domain, through
    both opening forms.  Two different trapdoors give the same bytes."""
    n = 1 << log_n
    ones, twos, zeros, neg = ref.join32([1] * n), ref.join32([2] * n), bytes(32 * n), ref.join32([R - 1] * n)
    w = pow(7, (R - 1) // n, R)
    xs = [ref.random_scalars(17 + log_n, 1), ref.fr_be(pow(w, 3 % n, R)), ref.fr_be(0)]
    for tau in ((TAU_X, TAU_Y), (0x1D0C0FFEE + log_n, 0xABCDEF)):
        gpu_ctx.srs_generate(tau[0], tau[1], log_n, 0)
        assert gpu_ctx.worker_commit(0, ones).hex() == G1_GEN_HEX
        assert gpu_ctx.worker_commit(0, twos).hex() == G1_2GEN_HEX
        assert gpu_ctx.worker_commit(0, neg).hex() == G1_NEG_GEN_HEX
        assert gpu_ctx.worker_commit(0, zeros).hex() == G1_INF_HEX
        for coset in (True, False):
            gpu_ctx.set_open_coset(coset)
            for x in xs:
                for poly, c, com in ((ones, 1, G1_GEN_HEX), (neg, R - 1, G1_NEG_GEN_HEX), (zeros, 0, G1_INF_HEX)):
                    got = gpu_ctx.worker_commit_open(0, poly, x)
                    assert (got[0].hex(), int.from_bytes(got[1], "big"), got[2].hex()) == (com, c, G1_INF_HEX), (log_n, coset, c)
                    assert gpu_ctx.worker_verify(0, got[2], x, got[1], got[0])
        gpu_ctx.set_open_coset(True)
