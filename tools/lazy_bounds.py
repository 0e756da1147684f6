#!/usr/bin/env python3
"""Range analysis of the lazily reduced mixed addition G1Xyzz::madd_lazy (csrc/g1.cuh).

Values are 12-limb integers in [0, 2^384); a bound k means "value < k p".  The Montgomery product without its final
conditional subtraction returns (a b + m p) / 2^384 < a b / 2^384 + p, i.e. bound rho ka kb + 1 with
rho = p / 2^384 (0.1016); the accumulators of the even/odd CIOS rows hold t_prev + a b_i + m_i p < 2^32 (a + p), which
has to stay below 2^416: a + p < 2^384 for a product, 2a + p < 2^384 for the dedicated squaring (its rows multiply by
limbs of 2a), a + c + p < 2^384 for the fused a b + c d.  The script iterates the bounds of one addition to their
fixed point and checks every such constraint with exact rationals."""
from fractions import Fraction as F

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
RHO = F(P, 1 << 384)
CAP = F(1 << 384, P)  # 9.84: values must stay below CAP p


def mul(ka, kb):
    assert ka + 1 < CAP, ("multiplicand too wide", float(ka))
    out = RHO * ka * kb + 1
    assert out < CAP
    return out


def sqr(ka):
    assert 2 * ka + 1 < CAP, ("squaring operand too wide", float(ka))
    return RHO * ka * ka + 1


def mul2(ka, kb, kc, kd):
    assert ka + kc + 1 < CAP
    out = RHO * (ka * kb + kc * kd) + 1
    assert out < CAP
    return out


def step(kx, ky, kzz, kzzz):
    """bounds of (X3, Y3, ZZ3, ZZZ3) given bounds of the accumulator; the table point is canonical (< p)"""
    u2 = mul(1, kzz)
    s2 = mul(1, kzzz)
    assert kx <= 2 and ky <= 2, "the differences below add 2p"
    p_ = u2 + 2          # u2 - X + 2p, in (0, .)
    r_ = s2 + 2
    assert p_ < 4 and r_ < 4, "zero test compares with p, 2p, 3p only"
    pp = sqr(p_)
    ppp = mul(p_, pp)
    q = mul(kx, pp)
    rr = sqr(r_)
    s = ppp + 2 * q
    assert s < CAP
    # X3 = rr - s, corrected by k p with k = (2^32 - top limb) / top limb of p + 1 when negative: result < p (1 + 2^-24)
    pt = P >> 352
    k_max = int(s * P) // (pt << 352) + 2
    assert k_max <= 7, k_max
    x3 = max(rr, 1 + F(1, 1 << 24))
    assert x3 <= 2
    d = q + 2            # q - X3 + 2p
    ny = F(2)            # 2p - Y
    y3_raw = mul2(r_, d, ppp, ny)
    assert y3_raw < 3    # one conditional subtraction of p
    y3 = max(F(1), y3_raw - 1)
    zz3 = mul(kzz, pp)
    zzz3 = mul(kzzz, ppp)
    return x3, y3, zz3, zzz3, dict(P=p_, R=r_, PP=pp, PPP=ppp, Q=q, RR=rr, S=s, Y3raw=y3_raw, kmax=k_max)


if __name__ == "__main__":
    def up(v):  # round up to a multiple of 1/1000 (keeps the rationals small; bounds only get looser)
        return F(-((-v * 1000) // 1), 1000)

    b = (F(1), F(1), F(1), F(1))  # a fresh accumulator is a canonical affine point with ZZ = ZZZ = 1
    for it in range(200):
        x3, y3, zz3, zzz3, info = step(*b)
        nb = (max(b[0], up(x3)), max(b[1], up(y3)), max(b[2], up(zz3)), max(b[3], up(zzz3)))
        if nb == b:
            break
        b = nb
    # the step is monotone in every bound: confirm that the box maps into itself
    box = b
    x3, y3, zz3, zzz3, info = step(*box)
    assert x3 <= box[0] and y3 <= box[1] and zz3 <= box[2] and zzz3 <= box[3], "not a fixed box"
    print("invariant box (units of p): X < %.3f  Y < %.3f  ZZ < %.3f  ZZZ < %.3f" % tuple(map(float, box)))
    print("intermediates:", {k: round(float(v), 4) for k, v in info.items()})
    print("capacity 2^384 / p = %.4f" % float(CAP))
