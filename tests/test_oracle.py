"""CPU: pins the oracle (Python big-int and the C restatement) against the reference's known answer,
the standard encodings and the committed golden vectors.  The oracle is the checker for every GPU
parity test, so it is checked first."""
import pytest

from oracle import bls12_381 as o
from oracle import ref

R = o.R


@pytest.fixture(scope="module")
def poly(golden):
    return [o.fr_from_b64(s) for s in golden["test_poly"]]


def test_reference_known_answer(golden, poly):
    # reference tests/test_miner.py:33-55: TEST_EVAL = Horner(TEST_POLY as coefficients, TEST_POINT), BE base64
    x = o.fr_from_b64(golden["test_point"])
    assert o.fr_to_b64(o.horner_eval(poly, x)) == golden["test_eval"]
    assert all(len(s) == 43 and o.fr_from_b64(s) < R for s in golden["test_poly"])
    pb = ref.join32(poly)
    assert ref.eval_coeffs(pb, ref.fr_be(x)) == o.b64_decode(golden["test_eval"])


def test_g1_encodings(golden):
    enc = golden["g1_encodings"]
    assert o.g1_compress(o.G1_GEN).hex() == enc["G"] == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    assert o.g1_compress(o.g1_neg(o.G1_GEN)).hex() == enc["negG"] and enc["negG"].startswith("b7f1d3a7")
    assert o.g1_compress(o.g1_mul(o.G1_GEN, 2)).hex() == enc["2G"] == "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e"
    assert o.g1_compress(None).hex() == enc["inf"] == "c0" + "00" * 47
    for k in ("G", "negG", "2G", "inf"):
        assert o.g1_compress(o.g1_decompress(bytes.fromhex(enc[k]))).hex() == enc[k]
    assert ref.g1_mul_gen(ref.fr_be(2)).hex() == enc["2G"]
    assert ref.g1_mul_gen(ref.fr_be(R - 1)).hex() == enc["negG"]
    assert ref.g1_mul_gen(ref.fr_be(0)).hex() == enc["inf"]


def test_curve_constants():
    assert o.g1_is_on_curve(o.G1_GEN) and o.g2_is_on_curve(o.G2_GEN) and o.g1_in_subgroup(o.G1_GEN)
    z = -o.Z_ABS
    assert o.R == z**4 - z**2 + 1 and o.P == (z - 1) ** 2 * o.R // 3 + z
    for k, v in {4: 0x20b1ce9140267af9dd1c0af834cec32c17beb312f20b6f7653ea61d87742bcce,
                 16: 0x2155379d12180caa88f39a78f1aeb57867a665ae1fcadc91d7118f85cd96b8ad,
                 32: 0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b}.items():
        assert o.root_of_unity(1 << k) == v


def test_survey_vectors(golden, poly):
    x = o.fr_from_b64(golden["test_point"])
    mono, lag = o.srs_monomial(16), o.srs_lagrange(16)
    A, B = golden["A_coeff_form"], golden["B_eval_form"]
    assert o.g1_compress(o.kzg_commit(poly, mono)).hex() == A["commitment"] == "80c29505f17a8421a01fa597a3691518416653c7bc32bbcdcf5e76e2d07101338065ed0510b8bef414c8a31044b7ce08"
    y, pr = o.kzg_open_coeffs(poly, x, mono)
    assert o.fr_to_b64(y) == A["eval"] == golden["test_eval"]
    assert o.g1_compress(pr).hex() == A["proof"] == "b4ab05bb9553c77f62c568ca3fa882045208d5cdabd03e11eadbce7e0b93a74884c9afaf240cc3cdab44e80c055f6c8c"
    assert o.g1_compress(o.kzg_commit(poly, lag)).hex() == B["commitment"] == "aa3dcf78dff69fb1cc711993cc056c5f210db012cc654c9e9ebf09f450003b05c47645295f61f0bccc2e5e5c6e38c249"
    y, pr = o.kzg_open_evals(poly, x, lag)
    assert hex(y) == "0x5e130b00be5d4cf00af368a75a24aa5bdb2729c4f92d1e96b871f4ce5ec2ea23" and o.fr_to_b64(y) == B["eval"]
    assert o.g1_compress(pr).hex() == B["proof"] == "b25b1758de10baafed035fce2362d5d8991fb51a220088e3337990eebb77406753b7a5419abdfbc1058bca02b037ddbc"
    # evaluation-form commitment == coefficient-form commitment of the interpolant
    assert o.kzg_commit(o.ntt(poly, inverse=True), mono) == o.kzg_commit(poly, lag)


def test_c_oracle_matches_python(golden, poly):
    pb = ref.join32(poly)
    x = o.b64_decode(golden["test_point"])
    assert [o.fr_to_b64(v) for v in ref.split32(ref.ntt(pb))] == golden["ntt16"]
    assert [o.fr_to_b64(v) for v in ref.split32(ref.ntt(pb, True))] == golden["intt16"]
    assert ref.split32(ref.ntt(pb)) == o.dft_naive(poly)
    lag = ref.srs(16, o.TEST_SECRET, "lagrange")
    mono = ref.srs(16, o.TEST_SECRET, "monomial")
    assert ref.msm(lag, pb).hex() == golden["B_eval_form"]["commitment"]
    assert ref.msm(mono, pb).hex() == golden["A_coeff_form"]["commitment"]
    y, proof = ref.open_evals(pb, x, lag)
    assert o.fr_to_b64(int.from_bytes(y, "big")) == golden["B_eval_form"]["eval"]
    assert proof.hex() == golden["B_eval_form"]["proof"]
    xd = o.b64_decode(golden["B_in_domain"]["x"])
    y, proof = ref.open_evals(pb, xd, lag)
    assert o.fr_to_b64(int.from_bytes(y, "big")) == golden["B_in_domain"]["eval"] and proof.hex() == golden["B_in_domain"]["proof"]
    assert y == pb[5 * 32:6 * 32]


def test_pianist_rows(golden, poly):
    tau_y = int(golden["tau_y"])
    Rs = ref.split32(ref.lagrange_scalars(4, tau_y))
    assert Rs == o.lagrange_at(4, tau_y) and sum(Rs) % R == 1
    pb = ref.join32(poly)
    x = o.b64_decode(golden["test_point"])
    for rec in golden["pianist_4x16"]:
        row = ref.srs(16, o.TEST_SECRET, "lagrange", scale=Rs[rec["row"]])
        assert ref.msm(row, pb).hex() == rec["commitment"]
        y, proof = ref.open_evals(pb, x, row)
        assert proof.hex() == rec["proof"] and o.fr_to_b64(int.from_bytes(y, "big")) == rec["eval"]
    # aggregated commitment of identical rows = commitment under the unscaled SRS (sum_i R_i = 1)
    agg = None
    for rec in golden["pianist_4x16"]:
        agg = o.g1_add(agg, o.g1_decompress(bytes.fromhex(rec["commitment"])))
    assert o.g1_compress(agg).hex() == golden["B_eval_form"]["commitment"]


def test_pianist_master_vectors(golden):
    """The committed master vectors (coefficient-form Y-opening over a monomial tau_Y SRS) equal the
    evaluation-form route over the row scale points [R_i(tau_Y)]_1 -- two formulations, one group element --
    and the trapdoor identity pi_Y = [(g(tau_Y) - z)/(tau_Y - beta)]_1."""
    M = golden["pianist_master_4x16"]
    tau_y = int(golden["tau_y"])
    ys = [o.fr_from_b64(r["eval"]) for r in M["rows"]]
    Rs = o.lagrange_at(4, tau_y)
    scale_pts = [o.g1_mul(o.G1_GEN, r) for r in Rs]
    for bk, zk, pk in (("beta", "z", "proof_y"), ("beta_in_domain", "z_in_domain", "proof_y_in_domain")):
        beta = o.fr_from_b64(M[bk])
        z, piy = o.master_open_y(ys, beta, tau_y)
        assert (o.fr_to_b64(z), o.g1_compress(piy).hex()) == (M[zk], M[pk])
        q, z2 = o.quotient_evals(ys, beta)
        assert z2 == z and o.g1_compress(o.g1_msm_naive(scale_pts, q)).hex() == M[pk]
        g_tau = sum(y * r for y, r in zip(ys, Rs)) % R
        assert o.g1_compress(o.g1_mul(o.G1_GEN, (g_tau - z) * o.fr_inv((tau_y - beta) % R) % R)).hex() == M[pk]
    assert o.g1_compress(o.master_aggregate([o.g1_decompress(bytes.fromhex(r["commitment"])) for r in M["rows"]])).hex() == M["commitment"]
    assert o.g1_compress(o.master_aggregate([o.g1_decompress(bytes.fromhex(r["proof"])) for r in M["rows"]])).hex() == M["proof_x"]


def test_msm_trapdoor_and_threads():
    n = 1 << 10
    srs = ref.srs(n, o.TEST_SECRET, "lagrange")
    ls = ref.lagrange_scalars(n, o.TEST_SECRET)
    for sc in (ref.random_scalars(7, n), bytes(32 * n), ref.join32([R - 1] * n), ref.join32([1] * n),
               ref.join32([0] * (n - 1) + [5]), ref.join32([(1 << 255) % R] * n)):
        a = ref.msm(srs, sc, 1)
        assert a == ref.msm(srs, sc, 4) == ref.g1_mul_gen(ref.fr_dot(sc, ls))
    assert ref.msm(srs, bytes(32 * n)).hex() == "c0" + "00" * 47


def test_pairing_and_kzg_verify(golden, poly):
    x = o.fr_from_b64(golden["test_point"])
    tau_g2 = o.g2_mul(o.G2_GEN, o.TEST_SECRET)
    assert [hex(c) for c in (tau_g2[0][0], tau_g2[0][1], tau_g2[1][0], tau_g2[1][1])] == golden["g2_tau_x"]
    B = golden["B_eval_form"]
    com = o.g1_decompress(bytes.fromhex(B["commitment"]))
    proof = o.g1_decompress(bytes.fromhex(B["proof"]))
    y = o.fr_from_b64(B["eval"])
    assert o.kzg_verify(com, proof, x, y, tau_g2)
    assert not o.kzg_verify(com, proof, x, (y + 1) % R, tau_g2)
    rec = golden["pianist_4x16"][2]
    assert o.kzg_verify(o.g1_decompress(bytes.fromhex(rec["commitment"])), o.g1_decompress(bytes.fromhex(rec["proof"])), x,
                        o.fr_from_b64(rec["eval"]), tau_g2, o.g1_decompress(bytes.fromhex(rec["scale_point"])))


def test_decompress_rejects_garbage():
    good = o.g1_compress(o.G1_GEN)
    plus_one = (int.from_bytes(good, "big") + 1).to_bytes(48, "big")  # reference tests/test_validator.py:79-86
    for bad in (plus_one, b"\xff" * 48, b"\x00" * 48, b"\xc0" + b"\x00" * 46 + b"\x01"):
        with pytest.raises(ValueError):
            o.g1_decompress(bad)


def test_known_answers_that_need_no_trapdoor():
    """The Lagrange basis sums to one: for every tau commit(1, ..., 1) is the G1 generator, commit(2, ...) its double,
    commit(0, ...) infinity -- public constants, compared with both restatements; a constant polynomial opens to itself
    with the point at infinity as proof."""
    G = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    G2x = "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e"
    INF = "c0" + "00" * 47
    for n, tau in ((16, None), (64, 0xC0FFEE)):
        lag = o.srs_lagrange(n) if tau is None else o.srs_lagrange(n, tau)
        assert o.g1_compress(o.kzg_commit([1] * n, lag)).hex() == G
        assert o.g1_compress(o.kzg_commit([2] * n, lag)).hex() == G2x
        assert o.g1_compress(o.kzg_commit([0] * n, lag)).hex() == INF
        y, pr = o.kzg_open_evals([7] * n, 123456789, lag)
        assert y == 7 and o.g1_compress(pr).hex() == INF
        srs96 = ref.srs(n, o.TEST_SECRET if tau is None else tau, "lagrange")
        assert ref.msm(srs96, ref.join32([1] * n), 2).hex() == G
        assert ref.msm(srs96, ref.join32([2] * n), 2).hex() == G2x
        y, pr = ref.open_evals(ref.join32([7] * n), ref.fr_be(123456789), srs96, 2)
        assert int.from_bytes(y, "big") == 7 and pr.hex() == INF
