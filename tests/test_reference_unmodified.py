"""The UNMODIFIED reference classes and the reference's OWN test functions, driven against the drop-in `fourier.Client`.

What runs here is reference code, imported from /root/reference with zero edits:
  base/protocol.py          Prove
  neurons/miner.py          Miner.rpc_commit / rpc_open / rpc_commit_and_open / forward
  neurons/validator.py      Challenge, Validator.rpc_*, generate_challenge, reward, get_rewards
  tests/test_miner.py       test_miner_forward (both parametrisations), TEST_SYNAPSE
  tests/test_validator.py   test_reward (all five parametrisations), make_proofs, change_proof
with `bittensor` replaced by tests/stubs/bittensor (no chain: the neuron objects are created with __new__ and given the
client, which is all those methods use) and `fourier` resolved to this repository's shim.

Two backends behind the same Client class:
  "oracle"  (runs here, no GPU)  tests/helpers/oracle_client.py -- the product's Client with its device context swapped
            for the CPU oracle; checks the drop-in SURFACE (signatures, Response objects, wire format, error
            conventions) against the reference's own code and pins the bytes to tests/golden/vectors.json;
  "cuda"    (-m gpu)             the real thing; needs /root/reference AND a GPU, so it is skipped on the GPU box, which
            has no reference tree.  The link between the two is the golden file: tests/test_gpu_client.py pins the CUDA
            path to the same vectors through the mirrors zkp_subnet_b200/{miner,validator}.py, whose behaviour is
            compared with the reference classes below.
"""
import base64
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
STUBS = os.path.join(ROOT, "tests", "stubs")

if not os.path.isdir(os.path.join(REFERENCE, "neurons")):
    pytest.skip("the reference tree is not present on this machine", allow_module_level=True)


@pytest.fixture(scope="module")
def refmods():
    """import the reference's modules (and its two test modules) unmodified"""
    saved = list(sys.path)
    sys.path.insert(0, STUBS)
    sys.path.insert(1, REFERENCE)
    try:
        import bittensor
        assert bittensor.__file__.startswith(STUBS)
        from base.protocol import Prove
        from neurons.miner import Miner
        from neurons.validator import Challenge, Validator
        import fourier
        assert os.path.dirname(os.path.dirname(fourier.__file__)) == ROOT  # this repository's shim, not the Rust crate's binding
        mods = {}
        for name in ("test_miner", "test_validator"):
            spec = importlib.util.spec_from_file_location("reference_" + name, os.path.join(REFERENCE, "tests", name + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods[name] = m
        yield {"Prove": Prove, "Miner": Miner, "Validator": Validator, "Challenge": Challenge, **mods}
    finally:
        sys.path[:] = saved


def make_client(backend):
    if backend == "oracle":
        # loaded by path: with the reference tree on sys.path the name `tests` belongs to the reference's own package
        spec = importlib.util.spec_from_file_location("zkp_oracle_client", os.path.join(ROOT, "tests", "helpers", "oracle_client.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        OracleClient = mod.OracleClient
        c = OracleClient(port=1337, bin="./test_prover", uncompressed=False, setup_path="test_setup.compressed",
                         precompute_path="test_precompute.compressed", seed=99)
    else:
        from fourier import Client
        c = Client(port=1337, bin="./test_prover", uncompressed=False, setup_path="test_setup.compressed",
                   precompute_path="test_precompute.compressed", test_srs=True, seed=99)
    c.start(scale=6, machines_scale=2)  # reference tests/conftest.py:26-27
    return c


BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=BACKENDS)
def client(request):
    c = make_client(request.param)
    yield c
    c.stop()


def test_reference_test_miner_forward(refmods, client, golden):
    """reference tests/test_miner.py:84-121, called as is with a Miner whose client is ours"""
    tm = refmods["test_miner"]
    assert tm.TEST_POLY == golden["test_poly"] and tm.TEST_POINT == golden["test_point"]
    miner = refmods["Miner"].__new__(refmods["Miner"])
    miner.client = client
    # what the reference's assertions compare, pinned to the golden vectors first (row 0 of the 4 x 16 Pianist SRS)
    out = miner.forward(refmods["Prove"](index=0, poly=tm.TEST_POLY, alpha=tm.TEST_POINT, eval=tm.TEST_EVAL))
    rec = golden["pianist_4x16"][0]
    assert base64.b64decode(out.commitment).hex() == rec["commitment"] and base64.b64decode(out.proof).hex() == rec["proof"]
    assert out.eval == rec["eval"] and out.poly == [] and out.alpha is None
    # the reference's own test, both parametrisations, in the reference's order (the second mutates TEST_SYNAPSE)
    tm.test_miner_forward(miner, True)
    tm.test_miner_forward(miner, False)
    # ... and our mirror of the miner behaves like the reference class on the same client
    from zkp_subnet_b200.miner import Miner as Mirror
    from zkp_subnet_b200.protocol import Prove as MirrorProve
    for fused in (True, False):
        got = Mirror(client, fused=fused).forward(MirrorProve(index=0, poly=tm.TEST_POLY, alpha=tm.TEST_POINT, eval=tm.TEST_EVAL))
        assert (got.commitment, got.eval, got.proof, got.poly, got.alpha) == (out.commitment, out.eval, out.proof, [], None)
    unfilled = MirrorProve(index=0, poly=tm.TEST_POLY, alpha=None)
    assert Mirror(client).forward(unfilled) is unfilled and unfilled.commitment is None


@pytest.mark.parametrize("missing_info,too_late,invalid_proof,half_time,expected_value", [
    (False, False, False, False, [1.0, 1.0]),
    (True, False, False, False, [0.0, 1.0]),
    (False, True, False, False, [0.0, 1.0]),
    (False, False, True, False, [0.0, 1.0]),
    (False, False, False, True, [0.5, 1.0]),
])
def test_reference_test_reward(refmods, client, missing_info, too_late, invalid_proof, half_time, expected_value):
    """reference tests/test_validator.py:59-121 (its parametrisation restated above), called as is"""
    validator = refmods["Validator"].__new__(refmods["Validator"])
    validator.client = client
    refmods["test_validator"].test_reward(validator, missing_info, too_late, invalid_proof, half_time, expected_value)


def test_reference_generate_challenge_matches_mirror_and_batched_entry(refmods, client):
    """Validator.generate_challenge (reference neurons/validator.py:106-120: per row an inverse fft and an eval) gives
    the evaluations the one-call entry gives, and the mirror's rewards equal the reference's."""
    from zkp_subnet_b200.validator import Validator as Mirror
    validator = refmods["Validator"].__new__(refmods["Validator"])
    validator.client = client
    challenge, responses, is_valid = refmods["test_validator"].make_proofs(validator)
    assert is_valid == [True, True]
    with client.challenge_evals(challenge.polys[:2], challenge.alpha) as r:
        assert r.status_code == 200 and r.json()["evals"] == challenge.evals
    for resp in responses:
        resp.dendrite.process_time = 2.5
    ref_rewards = [float(v) for v in validator.get_rewards(challenge, responses, 10.0)]
    mirror = Mirror(client)
    assert mirror.get_rewards(challenge, responses, [2.5, 2.5], 10.0) == ref_rewards == [0.75, 0.75]
    # the wire type round-trips through JSON the way bittensor ships it
    Prove = refmods["Prove"]
    s = challenge.to_synapse(1)
    again = Prove.model_validate_json(s.model_dump_json())
    assert (again.index, again.poly, again.alpha, again.eval) == (1, challenge.polys[1], challenge.alpha, challenge.evals[1])
