"""GPU (-m gpu): the reference's own integration tests re-stated against the drop-in `fourier.Client`
(reference tests/test_miner.py:84-121 and tests/test_validator.py:59-163), with the bittensor chain mocks
removed (the prover is what is under test; the reference never mocks it either)."""
import base64
import threading

import pytest

from fourier import Client
from oracle import bls12_381 as o
from zkp_subnet_b200.miner import Miner
from zkp_subnet_b200.protocol import Prove
from zkp_subnet_b200.validator import Validator

pytestmark = pytest.mark.gpu

TEST_SCALE, TEST_MACHINES_SCALE = 6, 2  # reference tests/conftest.py:26-27
TEST_MACHINE_COUNT = 2


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("srs") / "test_setup.compressed")
    # no SRS file exists at `path`: the throw-away test SRS has to be asked for explicitly, and nothing is written
    c = Client(port=1337, bin="./test_prover", uncompressed=False, setup_path=path, precompute_path=path + ".pre", test_srs=True)
    c.start(scale=TEST_SCALE, machines_scale=TEST_MACHINES_SCALE)
    import os
    assert c.srs_source == "test-trapdoor" and not os.path.exists(path)
    yield c
    c.stop()


@pytest.mark.parametrize("include_point", [True, False])
def test_miner_forward(client, golden, include_point):
    miner = Miner(client)
    synapse = Prove(index=0, poly=golden["test_poly"], alpha=golden["test_point"], eval=golden["test_eval"])
    with client.worker_commit(i=synapse.index, poly=synapse.poly) as resp:
        assert resp.status_code == 200
        commitment = resp.json().get("commitment")
    with client.worker_open(i=synapse.index, poly=synapse.poly, x=synapse.alpha) as resp:
        assert resp.status_code == 200
        eval, proof = resp.json().get("eval"), resp.json().get("proof")
    with client.worker_verify(i=synapse.index, proof=proof, alpha=synapse.alpha, eval=eval, commitment=commitment) as resp:
        assert resp.status_code == 200
        assert resp.json().get("valid")
    # bytes pinned by the golden vectors (row 0 of the 4x16 Pianist SRS)
    rec = golden["pianist_4x16"][0]
    assert base64.b64decode(commitment).hex() == rec["commitment"] and base64.b64decode(proof).hex() == rec["proof"]
    assert eval == rec["eval"] and len(commitment) == 64 and len(eval) == 43
    if not include_point:
        synapse.alpha = None
    ret = miner.forward(synapse)
    if include_point:
        assert ret.commitment == commitment and ret.proof == proof and ret.eval == eval
        assert ret.poly == [] and ret.alpha is None and ret.index == 0
        # the non-fused path (two RPCs, as the reference's rpc_commit_and_open) gives the same bytes
        assert Miner(client, fused=False).forward(synapse).proof == proof
    else:
        assert ret is synapse and ret.commitment is None  # exception swallowed, unfilled synapse returned


def make_proofs(validator):
    challenge = validator.generate_challenge(TEST_MACHINE_COUNT)
    responses = []
    for i in range(TEST_MACHINE_COUNT):
        with validator.client.worker_commit(i, challenge.polys[i]) as resp:
            commitment = resp.json().get("commitment")
        with validator.client.worker_open(i, challenge.polys[i], challenge.alpha) as resp:
            eval, proof = resp.json().get("eval"), resp.json().get("proof")
        responses.append(Prove(index=i, poly=[], alpha=None, eval=eval, commitment=commitment, proof=proof))
    for r in responses:
        with validator.client.worker_verify(r.index, r.proof, challenge.alpha, challenge.evals[r.index], r.commitment) as resp:
            assert resp.status_code == 200 and resp.json().get("valid")
    return challenge, responses


@pytest.mark.parametrize("missing_info,too_late,invalid_proof,half_time,expected_value", [
    (False, False, False, False, [1.0, 1.0]),
    (True, False, False, False, [0.0, 1.0]),
    (False, True, False, False, [0.0, 1.0]),
    (False, False, True, False, [0.0, 1.0]),
    (False, False, False, True, [0.5, 1.0]),
])
def test_reward(client, missing_info, too_late, invalid_proof, half_time, expected_value):
    def change_proof(proof):
        raw = base64.b64decode(proof)
        plus_one = int.from_bytes(raw, "big") + 1 % 2 ** (len(raw) * 8)
        return base64.b64encode(plus_one.to_bytes(len(raw), "big")).decode()

    validator = Validator(client)
    challenge, responses = make_proofs(validator)
    # the validator's eval (iNTT + Horner) equals the miner's barycentric eval
    assert [r.eval for r in responses] == challenge.evals[:TEST_MACHINE_COUNT]
    times = [0.0, 0.0]
    timeout = 10.0
    if missing_info:
        responses[0].commitment = None
    if too_late:
        times[0] = 11.0
    if invalid_proof:
        responses[0].proof = change_proof(responses[0].proof)
    if half_time:
        times[0] = 5.0
    assert validator.get_rewards(challenge, responses, times, timeout) == expected_value


def test_worker_verify_batch_matches_individual(client, golden):
    """Batched verification (one random linear combination for the whole challenge) gives, item by item, the
    answer of worker_verify: all-valid batches, and batches where some items are wrong in different ways (the
    fallback must single out exactly those)."""
    alpha = golden["test_point"]
    good = []
    for rec in golden["pianist_4x16"]:
        good.append({"i": rec["row"], "proof": base64.b64encode(bytes.fromhex(rec["proof"])).decode(), "eval": rec["eval"],
                     "commitment": base64.b64encode(bytes.fromhex(rec["commitment"])).decode()})

    def individual(items):
        out = []
        for it in items:
            with client.worker_verify(it["i"], it["proof"], alpha, it["eval"], it["commitment"]) as r:
                assert r.status_code == 200
                out.append(r.json()["valid"])
        return out

    def batch(items):
        with client.worker_verify_batch(items, alpha) as r:
            assert r.status_code == 200
            return r.json()["valid"]

    assert batch(good) == individual(good) == [True] * 4
    assert batch(good[:1]) == [True] and batch([]) == []
    bad = [dict(it) for it in good] + [dict(good[0])]
    bad[1]["eval"] = good[0]["eval"][:-1] + ("A" if good[0]["eval"][-1] != "A" else "E")   # wrong evaluation
    bad[2]["commitment"] = good[3]["commitment"]                                           # someone else's commitment
    bad[3]["proof"] = "!!not base64!!"                                                     # garbage
    got = batch(bad)
    assert got == individual(bad) and got[0] is True and got[4] is True and got[1:4] == [False, False, False]
    wrong_row = [dict(good[0], i=1)]                                                       # right bytes, wrong SRS row
    assert batch(wrong_row + good[1:]) == [False, True, True, True]
    with client.worker_verify_batch(good, "not a field element") as r:
        assert r.status_code == 200 and r.json()["valid"] == [False] * 4


def test_batched_challenge_equals_reference_flow(client):
    """generate_challenge through the one-call barycentric entry equals the reference's per-row
    inverse-fft + Horner flow (neurons/validator.py:106-120) on the same polynomial and point."""
    with client.random_poly() as r:
        rows = r.json()["poly"]
    with client.random_point() as r:
        alpha = r.json()["point"]
    v = Validator(client, batched=False)
    per_row = [v.rpc_eval(v.rpc_fft(row, left=True, inverse=True), alpha) for row in rows]
    with client.challenge_evals(rows, alpha) as r:
        assert r.status_code == 200 and r.json()["evals"] == per_row
    # alpha inside the domain: the evaluation is the stored value itself
    w = o.root_of_unity(len(rows[0]))
    xin = o.fr_to_b64(pow(w, 3, o.R))
    with client.challenge_evals(rows, xin) as r:
        assert r.json()["evals"] == [row[3] for row in rows]
    with client.challenge_evals([rows[0], rows[1][:-1]], alpha) as r:
        assert r.status_code == 400
    ch = Validator(client).generate_challenge(TEST_MACHINE_COUNT)
    assert len(ch.evals) == TEST_MACHINE_COUNT and all(len(e) == 43 for e in ch.evals)


def test_master_node_through_client(client, golden):
    """Client.master_commit / master_open / master_verify (Pianist master node) against the golden vectors: the
    four workers commit and open different rotations of TEST_POLY, the master aggregates and opens in Y."""
    M = golden["pianist_master_4x16"]
    poly, alpha = golden["test_poly"], golden["test_point"]
    coms, evals, proofs = [], [], []
    for i, rec in enumerate(M["rows"]):
        with client.worker_commit_and_open(i, poly[i:] + poly[:i], alpha) as r:
            assert r.status_code == 200
            j = r.json()
        assert base64.b64decode(j["commitment"]).hex() == rec["commitment"] and j["eval"] == rec["eval"]
        coms.append(j["commitment"]); evals.append(j["eval"]); proofs.append(j["proof"])
    with client.master_commit(coms) as r:
        com = r.json()["commitment"]
    assert base64.b64decode(com).hex() == M["commitment"]
    with client.master_open(evals, proofs, M["beta"]) as r:
        assert r.status_code == 200
        j = r.json()
    assert j["eval"] == M["z"] and base64.b64decode(j["proof_x"]).hex() == M["proof_x"]
    assert base64.b64decode(j["proof_y"]).hex() == M["proof_y"]
    with client.master_verify(j["proof_x"], j["proof_y"], alpha, M["beta"], j["eval"], com) as r:
        assert r.status_code == 200 and r.json()["valid"] is True
    with client.master_verify(j["proof_y"], j["proof_x"], alpha, M["beta"], j["eval"], com) as r:
        assert r.json()["valid"] is False
    with client.master_verify("garbage", j["proof_y"], alpha, M["beta"], j["eval"], com) as r:
        assert r.status_code == 200 and r.json()["valid"] is False
    with client.master_open(evals[:3], proofs[:3], M["beta"]) as r:
        assert r.status_code == 400  # one evaluation per worker


def test_pinned_staging_and_trace(client, golden):
    """A polynomial handed over in a page-locked buffer (zkp_host_alloc) gives the same bytes as one in ordinary
    memory; the stage trace entry reports both lanes."""
    from zkp_subnet_b200 import native
    ctx = client._need()
    raw = b"".join(o.b64_decode(s) for s in golden["test_poly"])
    x = o.b64_decode(golden["test_point"])
    pin = native.PinnedBuffer(1 << 12).write(raw)
    assert len(pin) == len(raw) and pin.tobytes() == raw
    assert ctx.worker_commit_open(0, pin, x) == ctx.worker_commit_open(0, raw, x)
    assert ctx.worker_commit(0, pin) == ctx.worker_commit(0, raw) and ctx.fft(pin) == ctx.fft(raw)
    rows = ctx.bench_trace(0, raw, x, 1)
    stages = {(lane, stage) for lane, stage, _ in rows}
    assert ("0", "accumulate_l0") in stages and ("1", "open_field_kernels") in stages and rows[-1][0] == "host"
    pin.close()
    with pytest.raises(ValueError):
        native.PinnedBuffer(64).write(b"x" * 65)


def test_challenge_shape_and_wire_format(client):
    with client.random_poly() as r:
        poly = r.json()["poly"]
    assert len(poly) == 4 and all(len(row) == 16 for row in poly)
    assert all(len(s) == 43 and o.fr_from_b64(s) < o.R for row in poly for s in row)
    with client.random_point() as r:
        assert o.fr_from_b64(r.json()["point"]) < o.R
    with client.fft(poly[0], left=True, inverse=True) as r:
        coeffs = r.json()["poly"]
    with client.fft(coeffs, left=True, inverse=False) as r:
        assert r.json()["poly"] == poly[0]
    # error convention: non-200 rather than an exception, except verify which is 200/false
    with client.worker_commit(0, ["not base64!"] * 16) as r:
        assert r.status_code != 200
    with client.worker_commit(9, poly[0]) as r:
        assert r.status_code != 200
    with client.worker_verify(0, "AAAA", "AAAA", "AAAA", "AAAA") as r:
        assert r.status_code == 200 and r.json()["valid"] is False


def test_concurrent_forwards(client, golden):
    # the axon calls forward from several threads: results must not interfere
    miner = Miner(client)
    rec = golden["pianist_4x16"]
    out = {}

    def work(i):
        s = Prove(index=i, poly=golden["test_poly"], alpha=golden["test_point"])
        out[i] = miner.forward(s)

    threads = [threading.Thread(target=work, args=(i % 4,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(4):
        assert base64.b64decode(out[i].commitment).hex() == rec[i]["commitment"]
        assert base64.b64decode(out[i].proof).hex() == rec[i]["proof"]


def test_worker_open_after_worker_commit_reuses_the_resident_polynomial(client, golden):
    """The reference's two-call flow (neurons/miner.py:56-61) through the shim: worker_open after worker_commit of the
    same list takes the speculative path (resident polynomial, no second upload) and must give the bytes of a plain
    worker_open; a different list of the same length, a list mutated in place, and a call that drops the resident
    polynomial in between must all be noticed."""
    import random
    rng = random.Random(7)
    n = 1 << (TEST_SCALE - TEST_MACHINES_SCALE)
    enc = lambda v: base64.b64encode(v.to_bytes(32, "big")).decode().rstrip("=")
    poly_a = [enc(rng.randrange(o.R)) for _ in range(n)]
    poly_b = [enc(rng.randrange(o.R)) for _ in range(n)]
    x = enc(rng.randrange(o.R))
    fresh = Client(test_srs=True)
    fresh.start(scale=TEST_SCALE, machines_scale=TEST_MACHINES_SCALE)
    try:
        expect = {}
        for name, p in (("a", poly_a), ("b", poly_b)):
            expect[name] = fresh.worker_open(1, list(p), x).json()      # never preceded by a call on the same list
            fresh.fft(poly_a, True, False)                              # drops the resident polynomial
        assert client.worker_commit(1, poly_a).status_code == 200
        assert client._free.queue[-1].resident_n == n  # the slot that served the call (the pool is last-in first-out)
        assert client.worker_open(1, poly_a, x).json() == expect["a"]   # speculative result accepted
        assert client.worker_open(1, list(poly_a), x).json() == expect["a"]   # an equal COPY of the list as well
        assert client.worker_open(1, poly_b, x).json() == expect["b"]   # same length, other polynomial: rejected, redone
        assert client.worker_open(1, poly_b, x).json() == expect["b"]   # ... and poly_b is the resident one now
        poly_b2 = list(poly_b)
        poly_b2[5] = poly_a[5]                                          # "mutated in place"
        want = fresh.worker_open(1, poly_b2, x).json()
        assert client.worker_open(1, poly_b2, x).json() == want != expect["b"]
        assert client.worker_commit(1, poly_a).status_code == 200
        assert client.fft(poly_b, True, False).status_code == 200       # another call stages scalars on the device
        assert client.worker_open(1, poly_a, x).json() == expect["a"]
        # a malformed element is an error response whichever path is taken
        assert client.worker_commit(1, poly_a).status_code == 200
        bad = list(poly_a)
        bad[3] = "!" * 43
        assert client.worker_open(1, bad, x).status_code == 400
        assert client.worker_open(1, poly_a, x).json() == expect["a"]
    finally:
        fresh.stop()
