"""
ORACLE (test infrastructure, NOT product code) -- BLS12-381 / KZG in plain Python big integers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product path (zkp_subnet_b200 + libzkp_b200.so) must never import it.

Parity status: **parity unpinned at the commitment/proof byte level**.  The reference's prover is
the external, un-vendored Rust crate `fourier` (requirements.txt:3, Makefile:24-28 of the
reference); its source is not on this box.  What this file is pinned against:
  * the reference's own de-facto known answer TEST_POLY / TEST_POINT / TEST_EVAL
    (reference tests/test_miner.py:33-55)  -> `horner_eval`, big-endian base64 Fr wire format;
  * the standard ZCash compressed-G1 encodings of G, -G, 2G, infinity;
  * the self-derived commitment/proof vectors A and B of SURVEY.md section 8c;
  * bilinearity / trapdoor identities (tests/test_oracle.py).

Everything here is deliberately the slowest, most obviously-correct formulation: affine group
law, schoolbook DFT for tiny n, generic polynomial-ring Fq12.
"""
from __future__ import annotations

import base64
from typing import List, Optional, Sequence, Tuple

# ------------------------------------------------------------------------------------------------
# constants (p, r prime, r = z^4 - z^2 + 1, z = -0xd201000000010000)
# ------------------------------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Z_ABS = 0xD201000000010000  # |z|, z negative
B_COEFF = 4
G1X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2X = (
    0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
    0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
)
G2Y = (
    0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
    0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
)
FR_GENERATOR = 7  # multiplicative generator of Fr*; omega_n = 7^((r-1)/n)
TEST_SECRET = 1927409816240961209460912649124  # public test trapdoor tau

G1Point = Optional[Tuple[int, int]]  # None = infinity


# ------------------------------------------------------------------------------------------------
# Fr wire codec: 32-byte big-endian, standard base64, unpadded on output (reference
# tests/test_miner.py:33-55 use 43-char strings; tests/test_validator.py:81 b64decode()s outputs)
# ------------------------------------------------------------------------------------------------
def b64_decode(s: str) -> bytes:
    return base64.b64decode(s + "=" * (-len(s) % 4))


def fr_from_b64(s: str) -> int:
    raw = b64_decode(s)
    if len(raw) != 32:
        raise ValueError("Fr must be 32 bytes")
    v = int.from_bytes(raw, "big")
    if v >= R:
        raise ValueError("non-canonical Fr")
    return v


def fr_to_b64(v: int) -> str:
    return base64.b64encode((v % R).to_bytes(32, "big")).decode().rstrip("=")


def g1_to_b64(pt: G1Point) -> str:
    return base64.b64encode(g1_compress(pt)).decode()


def g1_from_b64(s: str) -> G1Point:
    return g1_decompress(b64_decode(s))


# ------------------------------------------------------------------------------------------------
# Fr helpers
# ------------------------------------------------------------------------------------------------
def fr_inv(a: int) -> int:
    return pow(a % R, R - 2, R)


def root_of_unity(n: int) -> int:
    assert n & (n - 1) == 0 and (R - 1) % n == 0
    return pow(FR_GENERATOR, (R - 1) // n, R)


def horner_eval(coeffs: Sequence[int], x: int) -> int:
    """Client.eval: coefficient-form Horner (pinned by reference tests/test_miner.py:33-55)."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def ntt(vals: Sequence[int], inverse: bool = False) -> List[int]:
    """out[i] = sum_j in[j] w^(ij); inverse uses w^-1 and 1/n.  Natural order in and out.
    Client.fft(poly, left, inverse) of the reference (neurons/validator.py:58-65) [convention:
    SURVEY.md section 8c item 4]."""
    n = len(vals)
    if n == 1:
        return [vals[0] % R]
    w = root_of_unity(n)
    if inverse:
        w = fr_inv(w)
    out = _fft_rec(list(vals), w)
    if inverse:
        ninv = fr_inv(n)
        out = [v * ninv % R for v in out]
    return out


def _fft_rec(a: List[int], w: int) -> List[int]:
    n = len(a)
    if n == 1:
        return a
    even = _fft_rec(a[0::2], w * w % R)
    odd = _fft_rec(a[1::2], w * w % R)
    out = [0] * n
    t = 1
    for i in range(n // 2):
        v = t * odd[i] % R
        out[i] = (even[i] + v) % R
        out[i + n // 2] = (even[i] - v) % R
        t = t * w % R
    return out


def dft_naive(vals: Sequence[int], inverse: bool = False) -> List[int]:
    n = len(vals)
    w = root_of_unity(n)
    if inverse:
        w = fr_inv(w)
    out = [sum(vals[j] * pow(w, i * j, R) for j in range(n)) % R for i in range(n)]
    if inverse:
        ninv = fr_inv(n)
        out = [v * ninv % R for v in out]
    return out


def quotient_coeffs(coeffs: Sequence[int], x: int) -> Tuple[List[int], int]:
    """(f(X) - f(x)) / (X - x) by synthetic division; returns (q coeffs, y)."""
    n = len(coeffs)
    q = [0] * max(n - 1, 0)
    acc = 0
    for k in range(n - 1, 0, -1):
        acc = (coeffs[k] + acc * x) % R
        q[k - 1] = acc
    y = (coeffs[0] + acc * x) % R if n else 0
    return q, y


def lagrange_at(n: int, tau: int) -> List[int]:
    """L_j(tau) for the size-n domain (natural order): (tau^n-1) w^j / (n (tau - w^j))."""
    w = root_of_unity(n)
    zn = (pow(tau, n, R) - 1) % R
    ninv = fr_inv(n)
    out = []
    wj = 1
    for _ in range(n):
        d = (tau - wj) % R
        if d == 0:
            raise ValueError("tau in domain")
        out.append(zn * wj % R * ninv % R * fr_inv(d) % R)
        wj = wj * w % R
    return out


# ------------------------------------------------------------------------------------------------
# G1 (affine, y^2 = x^3 + 4 over Fq)
# ------------------------------------------------------------------------------------------------
G1_GEN: G1Point = (G1X, G1Y)


def g1_is_on_curve(pt: G1Point) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def g1_neg(pt: G1Point) -> G1Point:
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a: G1Point, b: G1Point) -> G1Point:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_mul(pt: G1Point, k: int) -> G1Point:
    k %= R
    acc: G1Point = None
    add = pt
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


def g1_msm_naive(points: Sequence[G1Point], scalars: Sequence[int]) -> G1Point:
    acc: G1Point = None
    for pt, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


def g1_in_subgroup(pt: G1Point) -> bool:
    if pt is None:
        return True
    # full-order check by plain scalar multiplication (k taken mod nothing here)
    acc: G1Point = None
    add = pt
    k = R
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc is None


def g1_compress(pt: G1Point) -> bytes:
    """ZCash / IETF compressed G1: 48 B big-endian x; bit7 compressed, bit6 infinity,
    bit5 set iff y > (p-1)/2."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > (P - 1) // 2:
        out[0] |= 0x20
    return bytes(out)


def g1_decompress(raw: bytes, check_subgroup: bool = True) -> G1Point:
    if len(raw) != 48:
        raise ValueError("compressed G1 must be 48 bytes")
    flags = raw[0] >> 5
    if not flags & 4:
        raise ValueError("uncompressed flag")
    x = int.from_bytes(bytes([raw[0] & 0x1F]) + raw[1:], "big")
    if flags & 2:
        if x != 0 or flags & 1:
            raise ValueError("bad infinity encoding")
        return None
    if x >= P:
        raise ValueError("x not canonical")
    y2 = (x * x * x + B_COEFF) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise ValueError("x not on curve")
    if (y > (P - 1) // 2) != bool(flags & 1):
        y = P - y
    pt = (x, y)
    if check_subgroup and not g1_in_subgroup(pt):
        raise ValueError("not in G1 subgroup")
    return pt


# ------------------------------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1), tuples (c0, c1)
# ------------------------------------------------------------------------------------------------
Fq2 = Tuple[int, int]


def f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_scalar(a: Fq2, k: int) -> Fq2:
    return (a[0] * k % P, a[1] * k % P)


def f2_inv(a: Fq2) -> Fq2:
    d = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def f2_neg(a: Fq2) -> Fq2:
    return ((-a[0]) % P, (-a[1]) % P)


B2: Fq2 = (4, 4)  # twist: y^2 = x^3 + 4(1+u)
G2Point = Optional[Tuple[Fq2, Fq2]]
G2_GEN: G2Point = (G2X, G2Y)


def g2_is_on_curve(pt: G2Point) -> bool:
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def g2_neg(pt: G2Point) -> G2Point:
    if pt is None:
        return None
    return (pt[0], f2_neg(pt[1]))


def g2_add(a: G2Point, b: G2Point) -> G2Point:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(pt: G2Point, k: int) -> G2Point:
    k %= R
    acc: G2Point = None
    add = pt
    while k:
        if k & 1:
            acc = g2_add(acc, add)
        add = g2_add(add, add)
        k >>= 1
    return acc


# ------------------------------------------------------------------------------------------------
# Fq12 = Fq[w]/(w^12 - 2 w^6 + 2)   (so w^6 = 1 + u); 12-coefficient lists
# ------------------------------------------------------------------------------------------------
def f12_one() -> List[int]:
    return [1] + [0] * 11


def f12_mul(a: Sequence[int], b: Sequence[int]) -> List[int]:
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                if bj:
                    t[i + j] += ai * bj
    # w^12 = 2 w^6 - 2
    for k in range(22, 11, -1):
        v = t[k]
        if v:
            t[k - 6] += 2 * v
            t[k - 12] -= 2 * v
    return [v % P for v in t[:12]]


def f12_pow(a: Sequence[int], e: int) -> List[int]:
    out = f12_one()
    base = list(a)
    while e:
        if e & 1:
            out = f12_mul(out, base)
        base = f12_mul(base, base)
        e >>= 1
    return out


def _f2_to_f12_at(c: Fq2, deg: int, out: List[int]) -> None:
    """add c * w^deg (c in Fq2, c0 + c1 u = (c0 - c1) + c1 w^6) into out; deg < 6."""
    out[deg] = (out[deg] + c[0] - c[1]) % P
    out[deg + 6] = (out[deg + 6] + c[1]) % P


def _line(lam: Fq2, xr: Fq2, yr: Fq2, p: Tuple[int, int]) -> List[int]:
    """Line through psi(R) with twist-slope lam, evaluated at P in G1 and scaled by w^3
    (w^3 lies in a proper subfield, so the factor dies in the final exponentiation):
        (yR - lam xR) + (lam xP) w^2 - yP w^3."""
    out = [0] * 12
    _f2_to_f12_at(f2_sub(yr, f2_mul(lam, xr)), 0, out)
    _f2_to_f12_at(f2_scalar(lam, p[0]), 2, out)
    out[3] = (out[3] - p[1]) % P
    return out


def miller_loop(p: G1Point, q: G2Point) -> List[int]:
    """f_{|z|,Q}(P) with Q on the twist, affine doubling/addition (sign of z ignored: the
    result is the inverse pairing, which is still bilinear and non-degenerate)."""
    if p is None or q is None:
        return f12_one()
    f = f12_one()
    rx, ry = q
    qx, qy = q
    for i in range(Z_ABS.bit_length() - 2, -1, -1):
        lam = f2_mul(f2_scalar(f2_mul(rx, rx), 3), f2_inv(f2_scalar(ry, 2)))
        f = f12_mul(f12_mul(f, f), _line(lam, rx, ry, p))
        nx = f2_sub(f2_mul(lam, lam), f2_scalar(rx, 2))
        ny = f2_sub(f2_mul(lam, f2_sub(rx, nx)), ry)
        rx, ry = nx, ny
        if (Z_ABS >> i) & 1:
            lam = f2_mul(f2_sub(qy, ry), f2_inv(f2_sub(qx, rx)))
            f = f12_mul(f, _line(lam, rx, ry, p))
            nx = f2_sub(f2_sub(f2_mul(lam, lam), rx), qx)
            ny = f2_sub(f2_mul(lam, f2_sub(rx, nx)), ry)
            rx, ry = nx, ny
    return f


FINAL_EXP = (P**12 - 1) // R


def final_exponentiation(f: Sequence[int]) -> List[int]:
    return f12_pow(f, FINAL_EXP)


def pairing(p: G1Point, q: G2Point) -> List[int]:
    return final_exponentiation(miller_loop(p, q))


def pairing_product_is_one(pairs: Sequence[Tuple[G1Point, G2Point]]) -> bool:
    f = f12_one()
    for p, q in pairs:
        f = f12_mul(f, miller_loop(p, q))
    return final_exponentiation(f) == f12_one()


# ------------------------------------------------------------------------------------------------
# KZG and the Pianist worker variant
# ------------------------------------------------------------------------------------------------
def srs_monomial(n: int, tau: int = TEST_SECRET) -> List[G1Point]:
    out = []
    t = 1
    for _ in range(n):
        out.append(g1_mul(G1_GEN, t))
        t = t * tau % R
    return out


def srs_lagrange(n: int, tau: int = TEST_SECRET, scale: int = 1) -> List[G1Point]:
    """[scale * L_j(tau)]_1, natural order.  `scale` = R_i(tau_Y) for Pianist row i."""
    return [g1_mul(G1_GEN, lj * scale % R) for lj in lagrange_at(n, tau)]


def kzg_commit(scalars: Sequence[int], srs: Sequence[G1Point]) -> G1Point:
    return g1_msm_naive(srs[: len(scalars)], scalars)


def kzg_open_coeffs(coeffs: Sequence[int], x: int, srs_mono: Sequence[G1Point]):
    q, y = quotient_coeffs(coeffs, x)
    return y, g1_msm_naive(srs_mono[: len(q)], q)


def eval_from_evals(evals: Sequence[int], x: int) -> int:
    return horner_eval(ntt(evals, inverse=True), x)


def quotient_evals(evals: Sequence[int], x: int) -> Tuple[List[int], int]:
    """Evaluation-form quotient q_j = (f_j - y)/(w^j - x); handles x in the domain
    (q_m = -sum_{j!=m} q_j w^(j-m))."""
    n = len(evals)
    w = root_of_unity(n)
    y = eval_from_evals(evals, x)
    q = [0] * n
    hit = -1
    wj = 1
    for j in range(n):
        d = (wj - x) % R
        if d == 0:
            hit = j
        else:
            q[j] = (evals[j] - y) * fr_inv(d) % R
        wj = wj * w % R
    if hit >= 0:
        acc = 0
        for j in range(n):
            if j != hit:
                acc = (acc + q[j] * pow(w, (j - hit) % n, R)) % R
        q[hit] = (-acc) % R
    return q, y


def kzg_open_evals(evals: Sequence[int], x: int, srs_lag: Sequence[G1Point]):
    q, y = quotient_evals(evals, x)
    return y, g1_msm_naive(srs_lag[: len(q)], q)


def kzg_verify(commitment: G1Point, proof: G1Point, x: int, y: int,
               tau_g2: G2Point, scale_g1: G1Point = G1_GEN) -> bool:
    """e(C - y*S, [1]_2) == e(pi, [tau - x]_2) with S = [1]_1 (plain KZG) or
    S = [R_i(tau_Y)]_1 (Pianist worker i)."""
    lhs = g1_add(commitment, g1_neg(g1_mul(scale_g1, y)))
    rhs_g2 = g2_add(tau_g2, g2_neg(g2_mul(G2_GEN, x)))
    return pairing_product_is_one([(lhs, g2_neg(G2_GEN)), (proof, rhs_g2)])


# ------------------------------------------------------------------------------------------------
# Pianist master node (eprint 2023/1271 section 3; the reference's roadmap item, README.md:38 and the
# "not yet implemented" note at neurons/validator.py:198).  f(X, Y) = sum_i R_i(Y) f_i(X):
#   com = sum_i com_i,  pi_X = sum_i pi_i,  g(Y) = f(alpha, Y) has evaluations y_i = f_i(alpha) on the
#   size-M domain,  z = g(beta),  pi_Y = [(g(tau_Y) - z)/(tau_Y - beta)]_1,
#   check  e(com - [z]_1, g2) == e(pi_X, [tau_X - alpha]_2) * e(pi_Y, [tau_Y - beta]_2).
# The Y-direction opening is restated here in COEFFICIENT form (interpolate, Horner, synthetic division,
# monomial SRS in tau_Y) -- deliberately not the evaluation-form route the product takes.
# ------------------------------------------------------------------------------------------------
def master_aggregate(points: Sequence[G1Point]) -> G1Point:
    acc = None
    for p in points:
        acc = g1_add(acc, p)
    return acc


def master_open_y(worker_evals: Sequence[int], beta: int, tau_y: int):
    coeffs = ntt(list(worker_evals), inverse=True) if len(worker_evals) > 1 else list(worker_evals)
    q, z = quotient_coeffs(coeffs, beta)
    mono_y = srs_monomial(max(len(q), 1), tau_y)
    return z, (g1_msm_naive(mono_y[: len(q)], q) if q else None)


def master_verify(com: G1Point, pi_x: G1Point, pi_y: G1Point, alpha: int, beta: int, z: int,
                  tau_x_g2: G2Point, tau_y_g2: G2Point) -> bool:
    lhs = g1_add(com, g1_neg(g1_mul(G1_GEN, z)))
    qx = g2_add(tau_x_g2, g2_neg(g2_mul(G2_GEN, alpha)))
    qy = g2_add(tau_y_g2, g2_neg(g2_mul(G2_GEN, beta)))
    return pairing_product_is_one([(lhs, g2_neg(G2_GEN)), (pi_x, qx), (pi_y, qy)])
