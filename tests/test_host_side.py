"""CPU: host-side parts of the product library that need no device -- the base64 wire codec, the G1 sum
used to combine per-GPU partial points, and the pairing behind worker_verify -- against the oracle."""
import base64
import os

import pytest

from oracle import bls12_381 as o
from zkp_subnet_b200 import native
from zkp_subnet_b200.client import Response, decode_poly, encode_poly


def test_b64_codec(golden):
    strs = golden["test_poly"]
    raw = native.b64_decode_fr("".join(strs).encode(), 43, len(strs))
    assert raw == b"".join(o.b64_decode(s) for s in strs)
    assert native.b64_encode_fr(raw).decode() == "".join(strs)
    assert decode_poly(strs) == raw and encode_poly(raw) == strs
    padded = [s + "=" for s in strs]
    assert decode_poly(padded) == raw
    assert decode_poly([]) == b""
    with pytest.raises(native.ZkpError):
        native.b64_decode_fr(b"!" * 43, 43, 1)
    with pytest.raises(native.ZkpError):  # non-zero trailing bits
        native.b64_decode_fr(("A" * 42 + "B").encode(), 43, 1)
    for v in (0, 1, o.R - 1, 2**256 - 1):
        b = v.to_bytes(32, "big")
        assert native.b64_encode_fr(b).decode() == base64.b64encode(b).decode().rstrip("=")


def test_wire_list_codec(golden):
    """csrc/wire_py.cpp: List[str] <-> bytes through the CPython API, against Python's own base64."""
    import os
    import random
    rng = random.Random(0xB200)
    vals = [rng.randrange(o.R) for _ in range(20000)] + [0, 1, o.R - 1]
    raw = b"".join(v.to_bytes(32, "big") for v in vals)
    strs = native.wire_encode_list(raw)
    assert isinstance(strs, list) and len(strs) == len(vals)
    assert strs[:50] == [base64.b64encode(v.to_bytes(32, "big")).decode().rstrip("=") for v in vals[:50]]
    assert native.wire_decode_list(strs) == raw
    assert native.wire_decode_list(tuple(strs)) == raw
    assert native.wire_decode_list([s.encode() for s in strs[:100]]) == raw[:3200]      # bytes items
    assert native.wire_decode_list([s + "=" for s in strs[:100]]) == raw[:3200]         # padded form
    assert native.wire_decode_list(iter(strs[:10])) == raw[:320]                        # any iterable
    for bad in ("*" + strs[7][1:], strs[7][:42], strs[7] + "A", strs[7][:42] + "B", "é" * 43, 5, None):
        lst = list(strs[:64])
        lst[7] = bad
        with pytest.raises(ValueError, match="element 7"):
            native.wire_decode_list(lst)
    # a bad element far into a long list is reported with its own index (threads race for the minimum)
    lst = list(strs)
    lst[15001] = "!" * 43
    lst[19000] = "!" * 43
    with pytest.raises(ValueError, match="element 15001"):
        native.wire_decode_list(lst)
    assert decode_poly(golden["test_poly"]) == b"".join(o.b64_decode(s) for s in golden["test_poly"])
    # decode + compare in one pass (the speculative worker_open of the Client shim): same list, one changed element,
    # a changed element behind a subclassed str (serial fallback path), malformed element
    import ctypes
    w = native.wire()
    n = len(strs)
    out, ref, same = ctypes.create_string_buffer(32 * n), ctypes.create_string_buffer(raw, 32 * n), ctypes.c_int(-1)
    call = lambda lst: w.zkp_wire_decode_list_cmp(lst, ctypes.addressof(out), 32 * n, ctypes.addressof(ref), ctypes.byref(same))
    assert call(strs) == n and same.value == 1 and out.raw == raw
    lst = list(strs)
    lst[12345] = strs[12344]
    assert call(lst) == n and same.value == 0 and out.raw[32 * 12345:32 * 12346] == raw[32 * 12344:32 * 12345]

    class S(str):
        pass
    lst = list(strs)
    lst[3] = S(strs[3])
    assert call(lst) == n and same.value == 1
    lst[3] = S(strs[4])
    assert call(lst) == n and same.value == 0
    lst[3] = "?" * 43
    assert call(lst) == -1 - 3


def test_g1_sum(golden):
    pts = [bytes.fromhex(r["commitment"]) for r in golden["pianist_4x16"]]
    assert native.g1_sum(b"".join(pts)).hex() == golden["B_eval_form"]["commitment"]
    assert native.g1_sum(b"").hex() == "c0" + "00" * 47
    g, ng = bytes.fromhex(golden["g1_encodings"]["G"]), bytes.fromhex(golden["g1_encodings"]["negG"])
    assert native.g1_sum(g + ng).hex() == golden["g1_encodings"]["inf"]
    assert native.g1_sum(g + g).hex() == golden["g1_encodings"]["2G"]
    with pytest.raises(native.ZkpError):
        native.g1_sum(b"\xff" * 48)
    # uncompressed route: same sums without square roots on the combining side
    exp = b"".join(native.g1_uncompress(p) for p in pts)
    assert len(exp) == 96 * len(pts)
    assert native.g1_sum_uncompressed(exp).hex() == golden["B_eval_form"]["commitment"]
    gx, gy = o.G1_GEN
    assert native.g1_uncompress(g) == gx.to_bytes(48, "big") + gy.to_bytes(48, "big")
    inf96 = native.g1_uncompress(bytes.fromhex(golden["g1_encodings"]["inf"]))
    assert inf96[0] == 0x40 and native.g1_sum_uncompressed(inf96 + exp[:96]) == pts[0]
    with pytest.raises(native.ZkpError):  # off-curve
        native.g1_sum_uncompressed(exp[:95] + bytes([exp[95] ^ 1]))


def test_g1_sum_checked_many_points_and_first_bad_index(golden):
    """zkp_g1_sum_checked (points received from other parties: decompression + subgroup check per point, spread over the
    host threads): k*G for k = 1..80 sums to (80*81/2)*G; the error names the FIRST point that fails, whichever thread met
    a bad one first; a curve point outside the prime-order subgroup is refused."""
    from oracle import ref
    pts = [ref.g1_mul_gen(ref.fr_be(k)) for k in range(1, 81)]
    assert native.g1_sum_checked(b"".join(pts)) == ref.g1_mul_gen(ref.fr_be(80 * 81 // 2))
    assert native.g1_sum_checked(b"".join(pts)) == native.g1_sum(b"".join(pts))
    bad = list(pts)
    bad[70] = b"\xff" * 48
    bad[23] = b"\xff" * 48
    with pytest.raises(native.ZkpError, match="point 23 "):
        native.g1_sum_checked(b"".join(bad))
    # x = 4 lies on the curve E(Fq) but not in G1 (y^2 = 68 has a root; cofactor != 1): find its compressed form
    for x in range(2, 40):
        y2 = (x ** 3 + 4) % o.P
        y = pow(y2, (o.P + 1) // 4, o.P)
        if y * y % o.P == y2 and not o.g1_in_subgroup((x, y)):
            enc = bytearray(x.to_bytes(48, "big"))
            enc[0] |= 0x80 | (0x20 if y > o.P - y else 0)
            with pytest.raises(native.ZkpError, match="point 5 "):
                native.g1_sum_checked(b"".join(pts[:5]) + bytes(enc) + b"".join(pts[5:]))
            break
    else:  # pragma: no cover
        raise AssertionError("no small curve point outside the subgroup found")


def _g2(pt):
    return b"".join(c.to_bytes(48, "big") for c in (pt[0][0], pt[0][1], pt[1][0], pt[1][1]))


def test_pairing_check_bilinear():
    a, b = 0x1234567, 0x89ABCDEF01
    P1 = o.g1_compress(o.g1_mul(o.G1_GEN, a))
    Q1 = _g2(o.g2_mul(o.G2_GEN, b))
    P2 = o.g1_compress(o.g1_neg(o.g1_mul(o.G1_GEN, a * b % o.R)))
    assert native.pairing_check(P1 + P2, Q1 + _g2(o.G2_GEN))
    P3 = o.g1_compress(o.g1_neg(o.g1_mul(o.G1_GEN, (a * b + 1) % o.R)))
    assert not native.pairing_check(P1 + P3, Q1 + _g2(o.G2_GEN))
    # infinity contributes 1
    assert native.pairing_check(o.g1_compress(None), Q1)


def test_pairing_check_kzg_equation(golden):
    # e(C - [y]_1, g2) * e(-pi, [tau - x]_2) == 1 on the golden evaluation-form vector
    B = golden["B_eval_form"]
    x = o.fr_from_b64(golden["test_point"])
    y = o.fr_from_b64(B["eval"])
    com = o.g1_decompress(bytes.fromhex(B["commitment"]))
    proof = o.g1_decompress(bytes.fromhex(B["proof"]))
    lhs = o.g1_add(com, o.g1_neg(o.g1_mul(o.G1_GEN, y)))
    q2 = o.g2_add(o.g2_mul(o.G2_GEN, o.TEST_SECRET), o.g2_neg(o.g2_mul(o.G2_GEN, x)))
    assert native.pairing_check(o.g1_compress(lhs) + o.g1_compress(o.g1_neg(proof)), _g2(o.G2_GEN) + _g2(q2))
    lhs_bad = o.g1_add(com, o.g1_neg(o.g1_mul(o.G1_GEN, y + 1)))
    assert not native.pairing_check(o.g1_compress(lhs_bad) + o.g1_compress(o.g1_neg(proof)), _g2(o.G2_GEN) + _g2(q2))


def test_response_contract():
    with Response(200, {"commitment": "x"}) as r:
        assert r.status_code == 200 and r.json().get("commitment") == "x"


def test_wire_decoder_every_byte_at_every_position():
    """The list decoder (AVX2 on x86-64, table-driven otherwise) against Python's base64 for every byte value at
    every one of the 43 positions: accepted exactly when the byte is in the standard alphabet (and, in the last
    position, carries no trailing bits), and then decoded to the same 32 bytes."""
    import random
    alphabet = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"
    rng = random.Random(43)
    base = base64.b64encode(bytes(rng.randrange(256) for _ in range(32)))[:43]
    checked = 0
    for pos in range(43):
        for c in range(256):
            s = base[:pos] + bytes([c]) + base[pos + 1:]
            ok = c in alphabet and (pos != 42 or alphabet.index(c) % 4 == 0)
            for item in (s, s + b"="):
                if ok:
                    assert native.wire_decode_list([item]) == base64.b64decode(s + b"="), (pos, c)
                else:
                    with pytest.raises(ValueError):
                        native.wire_decode_list([item])
            if c < 128:  # the same through a str
                if ok:
                    assert native.wire_decode_list([s.decode()]) == base64.b64decode(s + b"=")
                else:
                    with pytest.raises(ValueError):
                        native.wire_decode_list([s.decode()])
            checked += 1
    assert checked == 43 * 256
    # random valid elements in bulk
    raw = bytes(rng.randrange(256) for _ in range(32 * 5000))
    strs = [base64.b64encode(raw[32 * i:32 * i + 32]).decode().rstrip("=") for i in range(5000)]
    assert native.wire_decode_list(strs) == raw


# ---- round 2: the shim's start-up rules and the SRS file recognition need no GPU
def test_client_refuses_a_missing_srs_before_touching_the_gpu(tmp_path, monkeypatch):
    from zkp_subnet_b200 import native
    from zkp_subnet_b200.client import Client, _bitrev, _truthy
    monkeypatch.delenv("ZKP_B200_TEST_SRS", raising=False)
    c = Client(port=1337, bin="./prover", uncompressed="true", setup_path=str(tmp_path / "setup_24_8.uncompressed"),
               precompute_path=str(tmp_path / "precompute_24_8.uncompressed"))
    assert c.uncompressed is True
    with pytest.raises(native.ZkpError) as e:
        c.start(scale=24, machines_scale=8)
    assert e.value.code == native.ZKP_ERR_IO and "ZKP_B200_TEST_SRS" in str(e.value)
    assert not os.listdir(tmp_path)  # nothing was written
    with pytest.raises(ValueError):
        Client(poly_form="wavelets")
    with pytest.raises(ValueError):
        Client().start(scale=4, machines_scale=5)
    assert [_bitrev(i, 3) for i in range(8)] == [0, 4, 2, 6, 1, 5, 3, 7]
    assert _truthy("true") and _truthy(True) and not _truthy("false") and not _truthy("0") and not _truthy("")


def test_srs_file_recognition(tmp_path):
    from zkp_subnet_b200 import native, srsfile
    scale, ms = 6, 2
    count = 1 << scale
    setup, pre = tmp_path / "setup", tmp_path / "pre"
    assert srsfile.find_source(str(setup), str(pre), True, scale, ms) is None
    setup.write_bytes(bytes(96 * count + 384))
    src = srsfile.find_source(str(setup), str(pre), False, scale, ms)  # the size decides, the flag only breaks ties
    assert (src.kind, src.point_bytes, src.precompute_path, src.log_n, src.log_m) == ("raw", 96, None, 4, 2)
    pre.write_bytes(bytes(48 * count + 48 * 4))
    src = srsfile.find_source(str(setup), str(pre), True, scale, ms)
    assert (src.precompute_path, src.pre_point_bytes) == (str(pre), 48)
    setup.write_bytes(bytes(100))
    with pytest.raises(native.ZkpError) as e:
        srsfile.find_source(str(setup), None, True, scale, ms)
    assert "neither" in str(e.value)
    setup.write_bytes(b"ZKPB200S" + bytes(40))
    assert srsfile.find_source(str(setup), None, True, scale, ms).kind == "native"


def test_setup_cli_argument_checks(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [sys.executable, "-m", "zkp_subnet_b200.setup", "--setup-path", str(tmp_path / "s"), "--precompute-path", str(tmp_path / "p"),
            "--scale", "6", "--machines-scale", "2"]
    r = subprocess.run(base, cwd=root, capture_output=True, text=True)
    assert r.returncode == 1 and "nothing to do" in r.stderr
    (tmp_path / "s").write_bytes(b"x")
    r = subprocess.run(base + ["--generate-setup"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 1 and "--overwrite" in r.stderr and (tmp_path / "s").read_bytes() == b"x"
    r = subprocess.run(base[:-4] + ["--scale", "2", "--machines-scale", "3", "--generate-setup"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 2


def _pool_poly(n, seed):
    import random
    rng = random.Random(seed)
    raw = b"".join(rng.randrange(o.R).to_bytes(32, "big") for _ in range(n))
    return raw, encode_poly(raw)


def test_codec_pool_concurrent_callers_and_fork():
    """The codec's persistent worker pool (csrc/codec.hpp WorkerPool): repeated calls reuse it, callers that find it busy
    fall back to fresh threads, and a fork()ed child (which has no workers) builds its own."""
    import threading
    n = 1 << 15  # enough elements for several codec threads
    polys = [_pool_poly(n, s) for s in range(4)]
    for raw, strs in polys:  # sequential reuse
        for _ in range(3):
            assert decode_poly(strs) == raw
    # encode + decode from several Python threads at once (the decoder holds the GIL, the encoder's fill phase does
    # too: this exercises re-entry after a busy pool more than true overlap, which libzkp_b200's callers provide)
    errs = []

    def worker(k):
        raw, strs = polys[k]
        try:
            for _ in range(10):
                assert decode_poly(strs) == raw
                assert native.wire_encode_list(raw) == strs
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs
    # fork: the child must not wait for workers that only exist in the parent
    pid = os.fork()
    if pid == 0:
        ok = False
        try:
            raw, strs = polys[0]
            ok = decode_poly(strs) == raw and native.wire_encode_list(raw) == strs
        finally:
            os._exit(0 if ok else 1)
    import time
    deadline = time.time() + 60
    while True:
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            break
        if time.time() > deadline:  # pragma: no cover
            os.kill(pid, 9)
            os.waitpid(pid, 0)
            raise AssertionError("forked child hung in the codec pool")
        time.sleep(0.01)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
    # and the parent's pool still works
    assert decode_poly(polys[1][1]) == polys[1][0]
