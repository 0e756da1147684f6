"""CPU: the 32-bit-limb device algorithms (ff.cuh even/odd Montgomery product, g1.cuh XYZZ formulas incl.
doubling / inverse / infinity branches) compiled for the host with an emulated carry flag and checked
against Python big integers.  The same source is what nvcc compiles for sm_100a."""
import os
import subprocess

import pytest

from oracle import bls12_381 as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(name, tmp_path):
    exe = tmp_path / name
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-x", "c++", "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", name + ".cpp")])
    return str(exe)


def test_field_arithmetic(tmp_path):
    out = subprocess.check_output([build("ff_host_test", tmp_path), "400"]).decode()
    mods = {"fq": (o.P, 1 << 384), "fr": (o.R, 1 << 256)}
    cur, checked = {}, 0
    for line in out.splitlines():
        k, v = line.split()
        if k == "field":
            m, Rm = mods[v]
            Ri = pow(Rm, -1, m)
            continue
        v = int(v, 16)
        cur[k] = v
        a, b = cur.get("a"), cur.get("b")
        exp = {"mul": lambda: a * b * Ri % m, "add": lambda: (a + b) % m, "sub": lambda: (a - b) % m, "neg": lambda: (-a) % m,
               "sqr": lambda: a * a * Ri % m, "mul2": lambda: (a * b * Ri + b * (a * a * Ri % m) * Ri) % m, "tom": lambda: a * Rm % m, "fromm": lambda: a * Ri % m,
               "inv": lambda: (pow(a * Ri % m, -1, m) * Rm % m if a else 0)}
        if k in exp:
            assert exp[k]() == v, (k, hex(a), hex(b))
            checked += 1
        else:
            assert v < m
    assert checked > 5000


def test_g1_xyzz_formulas(tmp_path):
    out = subprocess.check_output([build("g1_host_test", tmp_path)]).decode()
    n = 0
    for line in out.splitlines():
        f = line.split()
        k = int(f[0])
        exp = o.g1_mul(o.G1_GEN, k) if k else None
        got = None if f[1] == "inf" else (int(f[1], 16), int(f[2], 16))
        assert got == exp, k
        n += 1
    assert n >= 17


def test_fq_inverse_binary_gcd(tmp_path):
    """csrc/fq_inv.cuh (chunked binary GCD, the division of the batched-affine bucket accumulation) on random,
    2^k, p - 2^k, small and short inputs: x * inv(x) == 1 (both in Montgomery form), canonical output, inv(0) = 0."""
    out = subprocess.check_output([build("inv_host_test", tmp_path), "3000"]).decode().split("\n")
    Rm = (1 << 384) % o.P
    n = 0
    for i in range(0, len(out) - 1, 2):
        x, v = int(out[i].split()[1], 16), int(out[i + 1].split()[1], 16)
        assert v < o.P
        if x == 0:
            assert v == 0
        else:
            assert x * v % o.P == Rm * Rm % o.P, hex(x)
        n += 1
    assert n == 3001


def test_host_pairing_fast_paths(tmp_path):
    """csrc/host: complex Fq12 squaring, cyclotomic squaring and the endomorphism subgroup check agree with their
    plain definitions (random Fq12 elements; curve points inside and outside the prime-order subgroup)."""
    out = subprocess.check_output([build("pairing_host_test", tmp_path)]).decode().splitlines()
    assert len(out) == 5 and all(line.split()[1] == "ok" for line in out), out


def test_lazy_field_helpers_and_madd_lazy(tmp_path):
    """The lazily reduced Fq helpers on WIDE operands (up to the ranges tools/lazy_bounds.py allows, edge values
    included) against exact integer formulas -- an overflow of the even/odd accumulators would show up as a wrong
    integer, not just a wrong residue -- and G1Xyzz::madd_lazy beside the fully reduced madd (doubling and
    cancellation hit while the accumulator is loose; coordinate ranges checked after every step)."""
    import random
    exe = build("lazy_host_test", tmp_path)
    p, M = o.P, 1 << 384
    ninv = (-pow(p, -1, M)) % M
    rng = random.Random(0xB200)

    def wide(k_num, k_den=1):  # random value below k p, with edge values mixed in
        top = p * k_num // k_den
        c = rng.randrange(8)
        if c == 0:
            return max(0, min(top - 1, p * rng.randrange(0, k_num // k_den + 1) + rng.randrange(-2, 3)))
        if c == 1:
            return top - 1 - rng.randrange(3)
        if c == 2:
            return min(top - 1, int("".join(rng.choice(["00000000", "ffffffff"]) for _ in range(12)), 16))
        return rng.randrange(top)

    cases = []
    for _ in range(300):
        cases.append(("A", wide(16, 5), wide(16, 5)))       # both below 3.2 p: mul, sqr, add, mul2
        cases.append(("B", wide(44, 5), wide(49, 5)))       # multiplicand below 8.8 p, multiplier below 9.8 p: mul
        cases.append(("C", wide(2), wide(2)))               # differences with + 2p
        cases.append(("D", wide(2), wide(6)))               # sub_fix
    inp = "".join(f"{a:096x} {b:096x}\n" for _, a, b in cases).encode()
    out = subprocess.run([exe], input=inp, stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    recs, cur = [], {}
    for line in out:
        if not line:
            continue
        k, v = line.split()
        if k == "a" and cur:
            recs.append(cur)
            cur = {}
        cur[k] = int(v, 16)
    recs.append(cur)
    assert len(recs) == len(cases)

    def redc(t):
        return (t + (t * ninv % M) * p) >> 384

    n = 0
    for (kind, a, b), r in zip(cases, recs):
        assert (r["a"], r["b"]) == (a, b)
        if kind in "AB":
            assert r["mul"] == redc(a * b) < M, (kind, hex(a), hex(b))
            n += 1
        if kind == "A":
            assert r["sqr"] == redc(a * a)
            assert r["add"] == a + b
            t = redc(2 * a * b)
            assert r["mul2"] == (t - p if t >= p else t)
            n += 3
        if kind == "C":
            assert r["subp2"] == a - b + 2 * p and r["rsubp2"] == 2 * p - b
            n += 2
        if kind == "D":
            v = r["subfix"]
            assert (v - (a - b)) % p == 0 and 0 <= v < max(a + 1, p + (p >> 24)), (hex(a), hex(b), hex(v))
            if a >= b:
                assert v == a - b
            n += 1
    assert n > 2000
    res = subprocess.run([exe, "group"], stdout=subprocess.PIPE).stdout.decode().split()
    assert res[0] == "ok" and int(res[1]) >= 400, res


def test_lazy_range_proof_script():
    """tools/lazy_bounds.py: the coordinate box of madd_lazy maps into itself and no accumulator of the even/odd
    product rows can overflow (exact rationals; the script asserts, this test runs it)."""
    out = subprocess.run(["python", os.path.join(ROOT, "tools", "lazy_bounds.py")], stdout=subprocess.PIPE, check=True,
                         timeout=120).stdout.decode()
    assert "invariant box" in out and "X < 1.99" in out, out


def test_host_inverse_binary_gcd_equals_fermat(tmp_path):
    """csrc/host/field64.hpp F64::inverse (binary extended Euclid, what g1_compress and the opening's 1/(x^n - 1) use)
    against the Fermat form, Fq and Fr: single-bit values, 0, 1, p - 1, small and random values."""
    out = subprocess.check_output([build("inv64_host_test", tmp_path)]).decode().splitlines()
    assert len(out) == 2 and all(line.split()[1] == "ok" and int(line.split()[2]) > 3000 for line in out), out


def test_host_mont_asm_equals_portable_product(tmp_path):
    """csrc/host/mont_asm.hpp (mulx / adcx / adox Montgomery products, Fq and Fr) against the portable product of
    field64.hpp on 400 000 random and edge operand pairs per field; skipped by the binary on a CPU without BMI2 + ADX."""
    out = subprocess.check_output([build("mont_asm_host_test", tmp_path)]).decode().splitlines()
    assert len(out) == 2, out
    for line in out:
        name, verdict, count = line.split()
        assert verdict in ("ok", "skipped"), out
        if verdict == "ok":
            assert int(count) > 300000
