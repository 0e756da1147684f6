"""Python model of the chunked binary-GCD inversion used by the batched-affine bucket accumulation
(csrc/fq_inv.cuh): same word sizes, same approximations, same fixed number of chunks.  Used to validate the
algorithm and to derive the final correction constant.  Algorithm: T. Pornin, "Optimized Binary GCD for Modular
Inversion" (eprint 2020/972), with 30 inner iterations per chunk on 64-bit approximations (30 low + 34 top bits)."""
import random

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
K = 30          # inner iterations per chunk
CHUNKS = 27     # 810 >= 2*381 - 1 iterations
MASK64 = (1 << 64) - 1
PINV32 = (-pow(P, -1, 1 << 32)) % (1 << 32)


def approx(a, b):
    n = max(a.bit_length(), b.bit_length())
    if n <= 64:
        return a, b
    lo = (1 << K) - 1
    return (a & lo) | ((a >> (n - 34)) << K), (b & lo) | ((b >> (n - 34)) << K)


def inner(ah, bh):
    f0, g0, f1, g1 = 1, 0, 0, 1
    for _ in range(K):
        if ah & 1:
            if ah < bh:
                ah, bh, f0, f1, g0, g1 = bh, ah, f1, f0, g1, g0
            ah -= bh
            f0 -= f1
            g0 -= g1
        ah >>= 1
        f1 <<= 1
        g1 <<= 1
    assert all(-(1 << 31) <= v < (1 << 31) for v in (f0, g0, f1, g1))
    return f0, g0, f1, g1


def redc32(t):
    """(t + m p) / 2^32 brought to [0, p), t signed with |t| < 2^31 p"""
    m = ((t & 0xFFFFFFFF) * PINV32) & 0xFFFFFFFF
    r = (t + m * P) >> 32
    assert (t + m * P) & 0xFFFFFFFF == 0
    if r < 0:
        r += P
    if r >= P:
        r -= P
    assert 0 <= r < P
    return r


def inv_raw(x):
    """returns v with x^-1 = v * 4^CHUNKS (mod p); x in [1, p)"""
    a, b, u, v = x, P, 1, 0
    for _ in range(CHUNKS):
        f0, g0, f1, g1 = inner(*approx(a, b))
        na, nb = f0 * a + g0 * b, f1 * a + g1 * b
        assert na % (1 << K) == 0 and nb % (1 << K) == 0
        na >>= K
        nb >>= K
        if na < 0:
            na, f0, g0 = -na, -f0, -g0
        if nb < 0:
            nb, f1, g1 = -nb, -f1, -g1
        a, b = na, nb
        assert a < (1 << 384) and b < (1 << 384)
        u, v = redc32(f0 * u + g0 * v), redc32(f1 * u + g1 * v)
    assert a == 0 and b == 1, (a, b)
    return v


CORR = pow(4, CHUNKS, P)

if __name__ == "__main__":
    rng = random.Random(1)
    special = [1, 2, 3, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 380, (1 << 380) - 1, (1 << 64) - 1, 1 << 64, 1 << 32, 5, 7]
    special += [pow(2, k, P) for k in range(0, 400, 7)] + [(P - pow(2, k, P)) % P for k in range(1, 400, 11)]
    n = 0
    for x in special + [rng.randrange(1, P) for _ in range(20000)] + [rng.randrange(1, 1 << rng.randrange(1, 381)) for _ in range(5000)]:
        if x % P == 0:
            continue
        v = inv_raw(x % P)
        assert v * CORR % P * x % P == 1, hex(x)
        n += 1
    print("ok", n, "inputs; correction 4^CHUNKS =", hex(CORR))
    R = 1 << 384
    print("C = 4^T R^3 mod p =", hex(CORR * pow(R, 3, P) % P))
