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
    assert len(out) == 3 and all(line.split()[1] == "ok" for line in out), out
