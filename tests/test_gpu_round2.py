"""GPU (-m gpu): the second-round pieces, all through the C ABI and all bit-exact against the oracle or against the
plain single-request path --
  * grouped launch sets: fused commit+open (two groups) and zkp_worker_commit_open_batch (2k groups) vs k single calls;
  * the table arena (eviction, budget, logged fallback), forked contexts, the resident-upload generation;
  * coefficient-form switch; compressed SRS rows; G2 import/export;
  * the group inverse FFT: monomial SRS -> Lagrange rows without the trapdoor == zkp_srs_generate with it;
  * the `setup` CLI and the SRS file layouts Client.start() reads;
  * the in-library multi-GPU entries (zkp_mgpu_*), on however many devices the box has (they work with one).
"""
import base64
import os
import subprocess
import sys
import threading

import pytest

from oracle import bls12_381 as o
from oracle import ref
from zkp_subnet_b200 import native

pytestmark = pytest.mark.gpu

R = o.R
TAU_X = o.TEST_SECRET
TAU_Y = 0x1234567890ABCDEF1234567890ABCDEF
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_commit_open(srs, poly, x):
    com = ref.msm(srs, poly, 8)
    y, proof = ref.open_evals(poly, x, srs, 8)
    return com, y, proof


@pytest.mark.parametrize("log_n", [8, 10, 12])
def test_fused_commit_open_matches_two_lanes_and_oracle(gpu_ctx, log_n):
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 1)
    try:
        for row in (0, 1):
            srs = gpu_ctx.srs_export_row(row, n)
            cases = {"random": ref.random_scalars(log_n, n), "zeros": bytes(32 * n), "ones": ref.join32([1] * n),
                     "r_minus_1": ref.join32([R - 1] * n), "single": ref.join32([0] * (n - 1) + [77])}
            for name, poly in cases.items():
                x = ref.random_scalars(1000 + log_n, 1)
                want = oracle_commit_open(srs, poly, x)
                for mode in (1, 0, -1):
                    gpu_ctx.set_fuse(mode)
                    assert gpu_ctx.worker_commit_open(row, poly, x) == want, (row, name, mode)
            # x inside the domain, fused
            gpu_ctx.set_fuse(1)
            w = pow(7, (R - 1) // n, R)
            xd = ref.fr_be(pow(w, 5, R))
            poly = cases["random"]
            assert gpu_ctx.worker_commit_open(row, poly, xd) == oracle_commit_open(srs, poly, xd)
    finally:
        gpu_ctx.set_fuse(-1)


def test_fused_commit_open_full_size_2p16(gpu_ctx):
    log_n = 16
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    poly = gpu_ctx.random_poly(0xB200 + 2, n)
    x = gpu_ctx.random_point(5)
    try:
        gpu_ctx.set_fuse(0)
        two = gpu_ctx.worker_commit_open(0, poly, x)
        gpu_ctx.set_fuse(1)
        one = gpu_ctx.worker_commit_open(0, poly, x)
    finally:
        gpu_ctx.set_fuse(-1)
    assert one == two and gpu_ctx.worker_verify(0, one[2], x, one[1], one[0])
    # trapdoor identity: commit == [sum_j f_j L_j(tau)]_1
    assert one[0] == ref.g1_mul_gen(ref.fr_dot(poly, ref.lagrange_scalars(n, TAU_X)))
    # a constant polynomial (every digit of a window in one bucket of each group)
    const = ref.join32([123456789] * n)
    gpu_ctx.set_fuse(1)
    try:
        c1 = gpu_ctx.worker_commit_open(0, const, x)
        gpu_ctx.set_fuse(0)
        assert gpu_ctx.worker_commit_open(0, const, x) == c1
    finally:
        gpu_ctx.set_fuse(-1)


@pytest.mark.parametrize("log_n,log_m", [(8, 2), (10, 1), (12, 0)])
def test_batch_matches_single_requests(gpu_ctx, log_n, log_m):
    n, rows = 1 << log_n, 1 << log_m
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, log_m)
    for k in (1, 2, 5, 33):
        idx = [(3 * r + 1) % rows for r in range(k)]
        polys = [ref.random_scalars(100 * k + r, n) for r in range(k)]
        xs = [ref.random_scalars(7000 + 31 * k + r, 1) for r in range(k)]
        if k >= 5:
            polys[2] = bytes(32 * n)                                   # zero polynomial
            w = pow(7, (R - 1) // n, R)
            xs[3] = ref.fr_be(pow(w, 9, R))                            # evaluation point inside the domain
        single = [gpu_ctx.worker_commit_open(idx[r], polys[r], xs[r]) for r in range(k)]
        out = gpu_ctx.worker_commit_open_batch(idx, polys, b"".join(xs))
        assert [o_[0] for o_ in out] == [0] * k
        assert [tuple(o_[1:]) for o_ in out] == single, k
    # against the oracle for a small batch
    srs0 = gpu_ctx.srs_export_row(0, n)
    polys = [ref.random_scalars(5 + r, n) for r in range(3)]
    xs = [ref.random_scalars(50 + r, 1) for r in range(3)]
    out = gpu_ctx.worker_commit_open_batch([0, 0, 0], polys, b"".join(xs))
    for r in range(3):
        assert tuple(out[r][1:]) == oracle_commit_open(srs0, polys[r], xs[r])
    # one malformed request fails alone
    bad = bytearray(polys[1])
    bad[32 * 7:32 * 8] = (R + 5).to_bytes(32, "big")
    out = gpu_ctx.worker_commit_open_batch([0, 0, 0], [polys[0], bytes(bad), polys[2]], b"".join(xs))
    assert [o_[0] for o_ in out] == [0, native.ZKP_ERR_ENCODING, 0]
    assert tuple(out[0][1:]) == oracle_commit_open(srs0, polys[0], xs[0]) and out[1][1] == bytes(48)
    out = gpu_ctx.worker_commit_open_batch([0, 0], polys[:2], xs[0] + (R + 1).to_bytes(32, "big"))
    assert [o_[0] for o_ in out] == [0, native.ZKP_ERR_ENCODING]
    with pytest.raises(native.ZkpError):
        gpu_ctx.worker_commit_open_batch([rows], polys[:1], xs[0])  # worker index out of range fails the call


def test_batch_full_size_2p16_pinned(gpu_ctx):
    log_n, k = 16, 6
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 1)
    pins = [native.PinnedBuffer(32 * n).write(gpu_ctx.random_poly(0xB200 + r, n)) for r in range(k)]
    xs = [gpu_ctx.random_point(40 + r) for r in range(k)]
    idx = [r % 2 for r in range(k)]
    out = gpu_ctx.worker_commit_open_batch(idx, pins, b"".join(xs))
    for r in range(k):
        st, com, y, proof = out[r]
        assert st == 0 and (com, y, proof) == gpu_ctx.worker_commit_open(idx[r], pins[r], xs[r])
        assert gpu_ctx.worker_verify(idx[r], proof, xs[r], y, com)
    for p in pins:
        p.close()


def test_table_arena_eviction_budget_and_logged_fallback(capfd):
    log_n, log_m = 10, 3
    n = 1 << log_n
    with native.Context(0) as ctx:
        ctx.srs_generate(TAU_X, TAU_Y, log_n, log_m)
        c, W, _ = ctx.msm_info(n)
        slot_bytes = 2 * W * n * 128
        ctx.set_table_budget(3 * slot_bytes + 100)          # room for 3 of the 8 rows
        sc = ref.random_scalars(1, n)
        want = [ref.msm(ctx.srs_export_row(r, n), sc, 8) for r in range(8)]
        for r in list(range(8)) + [0, 5, 0, 7]:
            assert ctx.msm_g1(r, sc) == want[r]
        st = ctx.table_stats()
        assert st["slots"] == 3 and st["resident"] == 3 and st["builds"] >= 10 and st["evictions"] >= 7 and st["fallbacks"] == 0
        assert ctx.prebuild_tables(0, 8) == 3
        # a batch needing more tables than the arena has slots falls back to single requests -- same bytes
        polys = [ref.random_scalars(9 + r, n) for r in range(5)]
        xs = [ref.random_scalars(90 + r, 1) for r in range(5)]
        single = [ctx.worker_commit_open(r, polys[r], xs[r]) for r in range(5)]
        out = ctx.worker_commit_open_batch(list(range(5)), polys, b"".join(xs))
        assert [tuple(o_[1:]) for o_ in out] == single
    with native.Context(0) as ctx:
        ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
        ctx.set_table_budget(1000)                           # no table fits: classic path, logged, counted
        sc = ref.random_scalars(2, n)
        assert ctx.msm_g1(0, sc) == ref.msm(ctx.srs_export_row(0, n), sc, 8)
        assert ctx.table_stats()["fallbacks"] >= 1 and ctx.table_stats()["slots"] == 0
    assert "falling back to the classic" in capfd.readouterr().err


def test_forked_contexts_share_the_srs_and_run_concurrently(gpu_ctx):
    log_n = 12
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 1)
    forks = [gpu_ctx.fork() for _ in range(3)]
    try:
        polys = [ref.random_scalars(20 + k, n) for k in range(4)]
        x = ref.random_scalars(3, 1)
        want = [gpu_ctx.worker_commit_open(k % 2, polys[k], x) for k in range(4)]
        assert gpu_ctx.table_stats()["builds"] == forks[0].table_stats()["builds"]  # one arena
        got = [None] * 4
        errs = []

        def work(k, ctx):
            try:
                for _ in range(5):
                    got[k] = ctx.worker_commit_open(k % 2, polys[k], x)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=work, args=(k, ([gpu_ctx] + forks)[k])) for k in range(4)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs and got == want
        assert forks[1].fft(polys[0], True, False) == ref.ntt(polys[0], False)
    finally:
        for f in forks:
            f.close()
    # the parent is still usable after its forks are gone
    assert gpu_ctx.worker_commit_open(0, polys[0], x) == want[0]


def test_partial_point_exports_for_the_multi_process_combine(gpu_ctx):
    """zkp_last_points_uncompressed / _jacobian and their sums: what the ranks of bench.py exchange per step."""
    log_n = 8
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 1)
    outs, unc, jac = [], [], []
    for i in range(2):
        outs.append(gpu_ctx.worker_commit_open(i, ref.random_scalars(40 + i, n), ref.random_scalars(50, 1)))
        unc.append(gpu_ctx.last_points_uncompressed())
        jac.append(gpu_ctx.last_points_jacobian())
    want_c = native.g1_sum(outs[0][0] + outs[1][0])
    want_p = native.g1_sum(outs[0][2] + outs[1][2])
    assert native.g1_sum_uncompressed(unc[0][:96] + unc[1][:96]) == want_c
    assert native.g1_sum_uncompressed(unc[0][96:] + unc[1][96:]) == want_p
    cat = jac[0] + jac[1]
    assert native.g1_sum_jacobian(cat, 2, 288) == want_c and native.g1_sum_jacobian(cat[144:], 2, 288) == want_p
    assert native.g1_sum_jacobian(jac[0], 1, 288) == outs[0][0]
    bad = bytearray(cat)
    bad[5] ^= 1
    with pytest.raises(native.ZkpError):
        native.g1_sum_jacobian(bytes(bad), 2, 288)
    gpu_ctx.msm_g1(0, ref.random_scalars(41, n))        # an MSM alone: commitment slot = the result, proof slot = infinity
    j = gpu_ctx.last_points_jacobian()
    assert native.g1_sum_jacobian(j, 1, 288) == gpu_ctx.msm_g1(0, ref.random_scalars(41, n))
    assert native.g1_sum_jacobian(j[144:], 1, 144) == b"\xc0" + bytes(47)


def test_resident_open_is_bound_to_its_upload(gpu_ctx):
    log_n = 8
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    a, b = ref.random_scalars(1, n), ref.random_scalars(2, n)
    x = ref.random_scalars(3, 1)
    gpu_ctx.worker_commit(0, a)
    gen, rn = gpu_ctx.resident_generation()
    assert rn == n
    assert gpu_ctx.worker_open_resident_gen(0, n, gen, x) == gpu_ctx.worker_open(0, a, x)
    gpu_ctx.worker_commit(0, a)
    gen, _ = gpu_ctx.resident_generation()
    gpu_ctx.worker_commit(0, b)   # "another client" replaces the resident polynomial with one of the same length
    with pytest.raises(native.ZkpError) as e:
        gpu_ctx.worker_open_resident_gen(0, n, gen, x)
    assert e.value.code == native.ZKP_ERR_STATE
    gen2, _ = gpu_ctx.resident_generation()
    assert gen2 != gen and gpu_ctx.worker_open_resident_gen(0, n, gen2, x) == gpu_ctx.worker_open(0, b, x)


@pytest.mark.parametrize("log_n", [4, 10])
def test_coefficient_form_switch(gpu_ctx, log_n):
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    coeffs = ref.random_scalars(77, n)
    evals = ref.ntt(coeffs, False)
    x = ref.random_scalars(78, 1)
    want = gpu_ctx.worker_commit_open(0, evals, x)
    assert want[1] == ref.eval_coeffs(coeffs, x)
    try:
        gpu_ctx.set_poly_form(True)
        assert gpu_ctx.worker_commit(0, coeffs) == want[0]
        assert gpu_ctx.worker_open(0, coeffs, x) == want[1:]
        for mode in (0, 1):
            gpu_ctx.set_fuse(mode)
            assert gpu_ctx.worker_commit_open(0, coeffs, x) == want
        bad = bytearray(coeffs)
        bad[:32] = (R + 3).to_bytes(32, "big")
        with pytest.raises(native.ZkpError):
            gpu_ctx.worker_commit(0, bytes(bad))
    finally:
        gpu_ctx.set_poly_form(False)
        gpu_ctx.set_fuse(-1)
    assert gpu_ctx.worker_commit_open(0, evals, x) == want


@pytest.mark.parametrize("log_n,log_m", [(4, 2), (10, 1), (6, 0), (1, 3)])
def test_group_ifft_monomial_to_lagrange_equals_trapdoor_generation(gpu_ctx, log_n, log_m):
    n, rows = 1 << log_n, 1 << log_m
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, log_m)
    want_rows = [gpu_ctx.srs_export_row(i, n) for i in range(rows)]
    want_scale = [gpu_ctx.srs_export_scale_point(i) for i in range(rows)]
    want_g2 = (gpu_ctx.srs_export_g2(0), gpu_ctx.srs_export_g2(1))
    gpu_ctx.srs_generate_monomial2(TAU_X, TAU_Y, log_n, log_m)
    # the monomial rows are what they claim: row i = tau_y^i * [tau_x^j]
    if log_n == 4:
        for i in range(rows):
            assert gpu_ctx.srs_export_row(i, n) == ref.srs(n, TAU_X, "monomial", scale=pow(TAU_Y, i, R))
    gpu_ctx.srs_monomial_to_lagrange()
    assert [gpu_ctx.srs_export_row(i, n) for i in range(rows)] == want_rows
    assert [gpu_ctx.srs_export_scale_point(i) for i in range(rows)] == want_scale
    assert (gpu_ctx.srs_export_g2(0), gpu_ctx.srs_export_g2(1)) == want_g2
    if log_n >= 4:
        poly = ref.random_scalars(4, n)
        x = ref.random_scalars(5, 1)
        com, y, proof = gpu_ctx.worker_commit_open(rows - 1, poly, x)
        assert gpu_ctx.worker_verify(rows - 1, proof, x, y, com)


def test_compressed_rows_and_g2_points_roundtrip(gpu_ctx):
    log_n = 6
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 1)
    unc = [gpu_ctx.srs_export_row(i, n) for i in range(2)]
    cmp_ = [gpu_ctx.srs_export_row_compressed(i, n) for i in range(2)]
    for i in range(2):
        for j in (0, 1, n - 1):
            pt = (int.from_bytes(unc[i][96 * j:96 * j + 48], "big"), int.from_bytes(unc[i][96 * j + 48:96 * j + 96], "big"))
            assert cmp_[i][48 * j:48 * j + 48] == o.g1_compress(pt)
    g2x, g2y = gpu_ctx.srs_export_g2(0), gpu_ctx.srs_export_g2(1)
    sp = [gpu_ctx.srs_export_scale_point(i) for i in range(2)]
    with native.Context(0) as c2:
        c2.srs_set_shape(log_n, 1)
        for i in range(2):
            c2.srs_import_row_compressed(i, cmp_[i], sp[i])
        c2.srs_import_g2(0, g2x)
        c2.srs_import_g2(1, g2y)
        assert [c2.srs_export_row(i, n) for i in range(2)] == unc
        poly = ref.random_scalars(8, n)
        x = ref.random_scalars(9, 1)
        res = c2.worker_commit_open(1, poly, x)
        assert res == gpu_ctx.worker_commit_open(1, poly, x) and c2.worker_verify(1, res[2], x, res[1], res[0])
        # a compressed x with no point on the curve, and a G2 point off the curve, are refused
        bad = bytearray(cmp_[0])
        for delta in range(1, 40):
            bad[47] = (cmp_[0][47] + delta) & 0xFF
            try:
                c2.srs_import_row_compressed(0, bytes(bad), sp[0])
            except native.ZkpError as e:
                assert e.code == native.ZKP_ERR_ENCODING
                break
        else:
            pytest.fail("40 consecutive x-coordinates all had curve points")
        g2bad = bytearray(g2x)
        g2bad[191] ^= 1
        with pytest.raises(native.ZkpError):
            c2.srs_import_g2(0, bytes(g2bad))
    # the infinity point survives the compressed round trip
    gpu_ctx.srs_set_shape(2, 0)
    gpu_ctx.srs_import_row_compressed(0, b"\xc0" + bytes(47) + cmp_[0][:48 * 3])
    assert gpu_ctx.srs_export_row_compressed(0, 4) == b"\xc0" + bytes(47) + cmp_[0][:48 * 3]


@pytest.mark.parametrize("uncompressed", [True, False])
def test_setup_cli_and_client_start_from_files(tmp_path, uncompressed, golden):
    from fourier import Client
    setup, pre = str(tmp_path / "test_setup.bin"), str(tmp_path / "test_precompute.bin")
    cmd = [sys.executable, "-m", "zkp_subnet_b200.setup", "--setup-path", setup, "--precompute-path", pre, "--scale", "6",
           "--machines-scale", "2", "--generate-setup", "--generate-precompute", "--test-trapdoor"]
    if uncompressed:
        cmd += ["--uncompressed", "true"]
    subprocess.check_call(cmd + ["--overwrite"], cwd=ROOT)
    pb = 96 if uncompressed else 48
    assert os.path.getsize(setup) == pb * 64 + 384 and os.path.getsize(pre) == pb * 64 + 48 * 4
    # without --overwrite existing files are kept
    assert subprocess.call(cmd, cwd=ROOT, stderr=subprocess.DEVNULL) == 1
    rec = golden["pianist_4x16"]
    for precompute_path in (pre, str(tmp_path / "missing")):   # with the precompute file, and derived from the setup file alone
        c = Client(port=1337, bin="./prover", uncompressed=uncompressed, setup_path=setup, precompute_path=precompute_path)
        c.start(scale=6, machines_scale=2)
        try:
            assert c.srs_source == "file:raw"
            for i in range(4):
                r = c.worker_commit_and_open(i, golden["test_poly"], golden["test_point"]).json()
                assert base64.b64decode(r["commitment"]).hex() == rec[i]["commitment"]
                assert base64.b64decode(r["proof"]).hex() == rec[i]["proof"] and r["eval"] == rec[i]["eval"]
                assert c.worker_verify(i, r["proof"], golden["test_point"], r["eval"], r["commitment"]).json()["valid"]
        finally:
            c.stop()
    # precompute only from an existing setup file (the ceremony case)
    os.remove(pre)
    subprocess.check_call([sys.executable, "-m", "zkp_subnet_b200.setup", "--setup-path", setup, "--precompute-path", pre, "--scale", "6",
                           "--machines-scale", "2", "--generate-precompute"] + (["--uncompressed", "true"] if uncompressed else []), cwd=ROOT)
    c = Client(uncompressed=uncompressed, setup_path=setup, precompute_path=pre)
    c.start(scale=6, machines_scale=2)
    try:
        r = c.worker_commit(1, golden["test_poly"]).json()
        assert base64.b64decode(r["commitment"]).hex() == rec[1]["commitment"]
    finally:
        c.stop()
    # wrong scale for these files
    c = Client(uncompressed=uncompressed, setup_path=setup, precompute_path=pre)
    with pytest.raises(native.ZkpError):
        c.start(scale=8, machines_scale=2)


def test_client_refuses_to_invent_an_srs(tmp_path, monkeypatch, golden):
    from fourier import Client
    path = str(tmp_path / "setup_24_8.uncompressed")
    monkeypatch.delenv("ZKP_B200_TEST_SRS", raising=False)
    c = Client(port=1337, bin="./prover", uncompressed=True, setup_path=path, precompute_path=path + ".pre")
    with pytest.raises(native.ZkpError) as e:
        c.start(scale=6, machines_scale=2)
    assert e.value.code == native.ZKP_ERR_IO and not os.path.exists(path)
    monkeypatch.setenv("ZKP_B200_TEST_SRS", "1")
    c = Client(port=1337, bin="./prover", uncompressed="true", setup_path=path, precompute_path=path + ".pre", precompute="lazy")
    c.start(scale=6, machines_scale=2)
    try:
        assert c.srs_source == "test-trapdoor" and not os.path.exists(path)
        r = c.worker_commit(0, golden["test_poly"]).json()
        assert base64.b64decode(r["commitment"]).hex() == golden["pianist_4x16"][0]["commitment"]
    finally:
        c.stop()


def test_client_switches_pool_and_batch(golden):
    from fourier import Client
    enc = lambda v: base64.b64encode(v.to_bytes(32, "big")).decode().rstrip("=")
    rec = golden["pianist_4x16"]
    base = Client(test_srs=True, contexts=3)
    base.start(scale=6, machines_scale=2)
    try:
        # row order: worker i -> row bitrev(i)
        rev = Client(test_srs=True, row_order="bitrev")
        rev.start(scale=6, machines_scale=2)
        try:
            for i, row in enumerate([0, 2, 1, 3]):
                r = rev.worker_commit_and_open(i, golden["test_poly"], golden["test_point"]).json()
                assert base64.b64decode(r["commitment"]).hex() == rec[row]["commitment"]
                assert rev.worker_verify(i, r["proof"], golden["test_point"], r["eval"], r["commitment"]).json()["valid"]
                assert not base.worker_verify(i, r["proof"], golden["test_point"], r["eval"], r["commitment"]).json()["valid"] or i == row
        finally:
            rev.stop()
        # coefficient form: the same commitment from the coefficients of the same polynomial
        co = Client(test_srs=True, poly_form="coeffs")
        co.start(scale=6, machines_scale=2)
        try:
            coeffs = base.fft(golden["test_poly"], True, True).json()["poly"]
            r = co.worker_commit_and_open(2, coeffs, golden["test_point"]).json()
            assert base64.b64decode(r["commitment"]).hex() == rec[2]["commitment"] and r["eval"] == rec[2]["eval"]
            assert r["eval"] == co.eval(coeffs, golden["test_point"]).json()["y"]
        finally:
            co.stop()
        # batch method
        items = [{"i": i % 4, "poly": golden["test_poly"], "alpha": golden["test_point"]} for i in range(6)]
        items[4] = {"i": 1, "poly": golden["test_poly"][:-1] + ["!" * 43], "alpha": golden["test_point"]}
        res = base.worker_commit_and_open_batch(items)
        assert res.status_code == 200
        out = res.json()["results"]
        for k, it in enumerate(items):
            if k == 4:
                assert "error" in out[k]
            else:
                assert base64.b64decode(out[k]["commitment"]).hex() == rec[it["i"]]["commitment"]
                assert base64.b64decode(out[k]["proof"]).hex() == rec[it["i"]]["proof"]
        # pool: concurrent forwards from 8 threads over 3 contexts
        outs, errs = [None] * 8, []

        def work(k):
            try:
                for _ in range(4):
                    outs[k] = base.worker_commit_and_open(k % 4, golden["test_poly"], golden["test_point"]).json()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=work, args=(k,)) for k in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs
        for k in range(8):
            assert base64.b64decode(outs[k]["proof"]).hex() == rec[k % 4]["proof"]
        # the master node refuses a commitment outside the subgroup (a point of the curve's full group)
        x = 0
        while True:
            x += 1
            y2 = (x ** 3 + 4) % o.P
            y = pow(y2, (o.P + 1) // 4, o.P)
            if y * y % o.P == y2 and not o.g1_in_subgroup((x, y)):
                break
        outside = base64.b64encode(o.g1_compress((x, y))).decode()
        good = [base64.b64encode(bytes.fromhex(rec[i]["commitment"])).decode() for i in range(4)]
        assert base.master_commit(good).status_code == 200
        assert base.master_commit(good[:3] + [outside]).status_code == 400
    finally:
        base.stop()


@pytest.mark.parametrize("log_n", [12, 16, 20])
def test_ntt_bulk_copy_variant_gives_the_same_transform(gpu_ctx, log_n):
    n = 1 << log_n
    v = gpu_ctx.random_poly(0x77 + log_n, n)
    try:
        gpu_ctx.set_ntt_tma(False)
        f0, i0 = gpu_ctx.fft(v, True, False), gpu_ctx.fft(v, True, True)
        gpu_ctx.set_ntt_tma(True)
        assert gpu_ctx.fft(v, True, False) == f0 and gpu_ctx.fft(v, True, True) == i0
        assert gpu_ctx.fft(f0, True, True) == v
    finally:
        gpu_ctx.set_ntt_tma(False)
    if log_n == 12:
        assert f0 == ref.ntt(v, False)


def test_staged_upload_in_chunks_matches_plain_upload(golden):
    """Large lists go through the chunked decode + staged upload (zkp_stage_*): same answers as the plain path, the
    two-call flow still recognises the resident polynomial, a malformed element is still a 400, and a superseded upload
    is refused."""
    from fourier import Client
    from zkp_subnet_b200.client import encode_poly
    lg = 17
    n = 1 << lg
    c = Client(test_srs=True, contexts=1, staged_upload=True)
    c.start(scale=lg, machines_scale=0)
    try:
        assert n >= c.STAGE_MIN
        ctx = c._need()
        raw = ctx.random_poly(0x5146, n)
        x = ctx.random_point(3)
        xs = base64.b64encode(x).decode().rstrip("=")
        strs = encode_poly(raw)
        want = ctx.worker_commit_open(0, raw, x)
        r = c.worker_commit_and_open(0, strs, xs).json()
        assert (base64.b64decode(r["commitment"]), base64.b64decode(r["eval"] + "="), base64.b64decode(r["proof"])) == want
        assert base64.b64decode(c.worker_commit(0, strs).json()["commitment"]) == want[0]
        o1 = c.worker_open(0, strs, xs).json()          # speculative: the staged polynomial is the resident one
        o2 = c.worker_open(0, list(strs), xs).json()
        assert o1 == o2 and base64.b64decode(o1["proof"]) == want[2]
        other = encode_poly(ctx.random_poly(0x5147, n))
        o3 = c.worker_open(0, other, xs).json()         # another polynomial: staged afresh
        assert base64.b64decode(o3["proof"]) == ctx.worker_open(0, ctx.random_poly(0x5147, n), x)[1]
        bad = list(strs)
        bad[n - 7] = "!" * 43
        assert c.worker_commit(0, bad).status_code == 400
        assert c.worker_commit_and_open(0, bad, xs).status_code == 400
        assert base64.b64decode(c.worker_commit(0, strs).json()["commitment"]) == want[0]
        # the C entries directly: chunks in any order; a superseded generation is refused
        pin = native.PinnedBuffer(32 * n).write(raw)
        gen = ctx.stage_list(strs, pin, chunk=5000)
        assert pin.tobytes() == raw and ctx.worker_commit_open_resident(0, n, gen, x) == want
        assert ctx.worker_commit_resident(0, n, gen) == want[0]
        ctx.fft(raw[:32 * 16], True, False)             # anything that rewrites the staged scalars
        with pytest.raises(native.ZkpError) as e:
            ctx.worker_commit_resident(0, n, gen)
        assert e.value.code == native.ZKP_ERR_STATE
        pin.close()
    finally:
        c.stop()


def _device_sets():
    n = native.lib().zkp_device_count()
    sets = [[0]]
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    return sets


@pytest.mark.parametrize("devices", _device_sets() if native.lib().zkp_device_count() else [[0]])
def test_mgpu_point_range_matches_single_device(gpu_ctx, devices):
    log_n = 12
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 1)
    poly = ref.random_scalars(0xB200 + 3, n)
    x = ref.random_scalars(17, 1)
    want = gpu_ctx.worker_commit_open(1, poly, x)
    with native.MultiContext(devices) as mg:
        mg.srs_generate(TAU_X, TAU_Y, log_n, 1, native.LAYOUT_POINT_RANGE)
        mg.prebuild_tables()
        assert mg.msm_g1(1, poly) == want[0]
        assert mg.msm_g1(1, poly, native.MGPU_RESIDENT) == want[0]
        assert mg.msm_g1(1, poly[:32 * (n - 5)]) == gpu_ctx.msm_g1(1, poly[:32 * (n - 5)])
        assert mg.commit_open(1, poly, x) == want
        assert mg.commit_open(1, poly, x, native.MGPU_RESIDENT) == want
        assert mg.commit_open(0, poly, x) == gpu_ctx.worker_commit_open(0, poly, x)
        bad = bytearray(poly)
        bad[-32:] = (R + 9).to_bytes(32, "big")
        with pytest.raises(native.ZkpError):
            mg.commit_open(1, bytes(bad), x)
        assert mg.commit_open(1, poly, x) == want   # usable after a failed call
        with pytest.raises(native.ZkpError):
            mg.pianist_commit_open([0], poly, x)     # wrong layout
        assert mg.ctx(0).worker_verify(1, want[2], x, want[1], want[0])


@pytest.mark.parametrize("devices", _device_sets() if native.lib().zkp_device_count() else [[0]])
def test_mgpu_pianist_matches_rows_and_master_verifies(gpu_ctx, devices):
    log_n, log_m = 10, 2
    n, rows = 1 << log_n, 1 << log_m
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, log_m)
    polys = b"".join(ref.random_scalars(300 + i, n) for i in range(rows))
    alpha, beta = ref.random_scalars(31, 1), ref.random_scalars(32, 1)
    per = [gpu_ctx.worker_commit_open(i, polys[32 * n * i:32 * n * (i + 1)], alpha) for i in range(rows)]
    with native.MultiContext(devices) as mg:
        mg.srs_generate(TAU_X, TAU_Y, log_n, log_m, native.LAYOUT_ROWS)
        coms, ys, proofs, agg_c, agg_p = mg.pianist_commit_open(list(range(rows)), polys, alpha)
        assert list(zip(coms, ys, proofs)) == per
        assert agg_c == native.g1_sum(b"".join(coms)) and agg_p == native.g1_sum(b"".join(proofs))
        z, pi_y = mg.ctx(0).master_open_y(b"".join(ys), beta)
        assert mg.ctx(0).master_verify(agg_c, agg_p, pi_y, alpha, beta, z)
        tampered = bytearray(z)
        tampered[-1] ^= 1
        assert not mg.ctx(0).master_verify(agg_c, agg_p, pi_y, alpha, beta, bytes(tampered))
        if len(devices) >= rows:
            again = mg.pianist_commit_open(list(range(rows)), polys, alpha, native.MGPU_RESIDENT)
            assert again[3:] == (agg_c, agg_p)
        # a subset of the rows, in another order
        sub = mg.pianist_commit_open([2, 0], polys[32 * n * 2:32 * n * 3] + polys[:32 * n], alpha)
        assert sub[0] == [per[2][0], per[0][0]] and sub[3] == native.g1_sum(per[2][0] + per[0][0])


def test_client_over_several_devices_and_split_mode(golden):
    from fourier import Client
    ndev = native.lib().zkp_device_count()
    devs = list(range(min(ndev, 2)))
    rec = golden["pianist_4x16"]
    for mode in ("requests", "split"):
        c = Client(test_srs=True, devices=devs, multi_gpu=mode)
        c.start(scale=6, machines_scale=2)
        try:
            for i in range(4):
                r = c.worker_commit_and_open(i, golden["test_poly"], golden["test_point"]).json()
                assert base64.b64decode(r["commitment"]).hex() == rec[i]["commitment"], mode
                assert base64.b64decode(r["proof"]).hex() == rec[i]["proof"] and r["eval"] == rec[i]["eval"]
                c1 = c.worker_commit(i, golden["test_poly"]).json()["commitment"]
                o1 = c.worker_open(i, golden["test_poly"], golden["test_point"]).json()
                assert (c1, o1["proof"]) == (r["commitment"], r["proof"])
                assert c.worker_verify(i, r["proof"], golden["test_point"], r["eval"], r["commitment"]).json()["valid"]
        finally:
            c.stop()


@pytest.mark.parametrize("log_n", [9, 10, 11, 12, 14, 16, 18, 20, 22])
def test_coset_opening_equals_the_general_form(gpu_ctx, log_n):
    """Single-request opening: pass 1 on cosets with the inversion on the host (default) against the general kernels
    (one Fermat inversion per block, zkp_set_open_coset(0)) and, where the oracle is quick, against the oracle; an x
    inside the domain takes the general kernels either way; every (y, proof) passes the pairing check.  2^22 has more
    coset blocks than k_open_coset_inv has threads (several block inverses per thread)."""
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    try:
        w = pow(7, (R - 1) // n, R)
        xs = [ref.random_scalars(2000 + log_n, 1), ref.fr_be(0), ref.fr_be(R - 1), ref.fr_be(2),
              ref.fr_be(pow(w, 3, R)), ref.fr_be(1)]  # the last two lie in the domain
        polys = {"random": ref.random_scalars(77 + log_n, n)}
        if log_n <= 12:
            polys["ones"] = ref.join32([1] * n)
            polys["single"] = ref.join32([0] * (n - 1) + [5])
        srs = gpu_ctx.srs_export_row(0, n) if log_n <= 12 else None
        for name, poly in polys.items():
            com = gpu_ctx.worker_commit(0, poly)
            for x in xs:
                gpu_ctx.set_open_coset(True)
                a = gpu_ctx.worker_open(0, poly, x)
                a3 = gpu_ctx.worker_commit_open(0, poly, x)
                gpu_ctx.set_open_coset(False)
                b = gpu_ctx.worker_open(0, poly, x)
                assert a == b and a3 == (com,) + tuple(a), (name, x.hex())
                if srs is not None and name == "random":
                    assert tuple(a) == tuple(ref.open_evals(poly, x, srs, 8)), (name, x.hex())
                assert gpu_ctx.worker_verify(0, a[1], x, a[0], com)
        # a batch (one x per request): cosets + host inversions against the general kernels
        if log_n <= 16:
            k = 5
            bp = [ref.random_scalars(900 + log_n + r, n) for r in range(k)]
            bx = b"".join(ref.random_scalars(950 + log_n + r, 1) for r in range(k))
            gpu_ctx.set_open_coset(True)
            a = gpu_ctx.worker_commit_open_batch([0] * k, bp, bx)
            gpu_ctx.set_open_coset(False)
            assert a == gpu_ctx.worker_commit_open_batch([0] * k, bp, bx)
            assert all(r[0] == 0 for r in a)
    finally:
        gpu_ctx.set_open_coset(True)


@pytest.mark.parametrize("log_n", [7, 10, 12, 13, 16])
def test_rowcol_coop_equals_the_large_array_kernels(gpu_ctx, log_n):
    """Bucket reduction over a small bucket array with four lanes per share and two warps per sum (default for a single
    request up to the mainnet row size) against the kernels used for large arrays: same bytes for random, constant,
    sparse and r - 1 scalars, with fixed-base tables (one bucket set) and without (one per window), and against the
    oracle where it is quick."""
    n = 1 << log_n
    gpu_ctx.srs_generate(TAU_X, TAU_Y, log_n, 0)
    cases = {"random": ref.random_scalars(400 + log_n, n), "constant": ref.join32([0xABCDEF + (1 << 222)] * n),
             "single": ref.join32([0] * (n - 1) + [9]), "r_minus_1": ref.join32([R - 1] * n), "zeros": bytes(32 * n)}
    srs = gpu_ctx.srs_export_row(0, n) if log_n <= 12 else None
    x = ref.random_scalars(41 + log_n, 1)
    try:
        for tables in (True, False):
            gpu_ctx.set_msm_mode(tables)
            for name, poly in cases.items():
                # lone MSMs take the cooperative kernels (a two-lane commit+open keeps the large-array ones)
                gpu_ctx.set_rowcol_coop(True)
                a = (gpu_ctx.worker_commit(0, poly),) + tuple(gpu_ctx.worker_open(0, poly, x))
                gpu_ctx.set_rowcol_coop(False)
                assert (gpu_ctx.worker_commit(0, poly),) + tuple(gpu_ctx.worker_open(0, poly, x)) == a, (name, tables)
                assert gpu_ctx.worker_commit_open(0, poly, x) == a, (name, tables)
                if srs is not None and tables:
                    assert a == oracle_commit_open(srs, poly, x), name
    finally:
        gpu_ctx.set_rowcol_coop(True)
        gpu_ctx.set_msm_mode(True)
