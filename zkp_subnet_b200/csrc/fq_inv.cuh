// Fq inversion by a chunked binary GCD (T. Pornin, "Optimized Binary GCD for Modular Inversion",
// eprint 2020/972) -- the one division of the batched-affine bucket accumulation (msm_affine.cuh).
//
// Why not Fermat: x^(p-2) is ~460 dependent Fq products (~140k multiply-pipe cycles and 0.5 ms of latency for
// a lone warp).  The binary GCD needs 2*381 - 1 halving steps; doing 30 of them at a time on 64-bit
// approximations of (a, b) (30 low bits + 34 top bits) turns them into ~26 rounds of
//   * 30 cheap steps on two 64-bit words and four 32-bit coefficients f0, g0, f1, g1,
//   * one exact update  (a, b) <- (f0 a + g0 b, f1 a + g1 b) / 2^30  (48 32x32 multiply-accumulates),
//   * one update of the cofactors mod p, each divided by 2^32 Montgomery-style (72 MACs),
// i.e. ~3k MACs plus ~35k simple ALU instructions per inversion: a few dozen Fq-product equivalents, paid once per
// batch of bucket additions (Montgomery's trick), not once per addition.
//
// Invariant: a = x u 4^t, b = x v 4^t (mod p) after t rounds (the exact update divides by 2^30, the cofactor
// update by 2^32).  At the end a = 0, b = gcd = 1, so x^-1 = v 4^t; with x given in Montgomery form (x R) and the
// result wanted in Montgomery form, out = montmul(v, R^3) * 4^t (2t modular doublings).  Rounds beyond
// convergence are harmless (f0 = 1, g1 = 2^30), so a warp runs until all its lanes have a = 0.
// The model tools/bingcd_model.py (same word sizes, checked against pow(x, -1, p) on 25k inputs including the
// 2^k and p - 2^k families that need the most rounds) fixes the bound: never more than 26 rounds.
//
// __host__ __device__, no inline asm: the same code is unit-tested on the CPU (tests/host/inv_host_test.cpp).
#pragma once
#include "ff.cuh"

namespace zkp {
namespace inv {

constexpr int N = 12;           // 32-bit limbs of Fq
constexpr int STEPS = 30;       // inner steps per round
constexpr int MAX_ROUNDS = 28;  // model: <= 26 observed; bound (2*381 - 1)/30 = 25.4
constexpr uint32_t LOW_MASK = (1u << STEPS) - 1;

// R^3 mod p (R = 2^384), little-endian limbs: montmul(v, R^3) = v R^2
ZKP_HD uint32_t r_cubed(int i) {
    constexpr uint32_t t[12] = {0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au, 0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u};
    return t[i];
}

ZKP_HD uint32_t clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

// 64-bit approximations of a and b: exact if both fit in 64 bits, else 30 low bits + the 34 bits at the top of
// the longer of the two
ZKP_HD void approximate(const uint32_t* a, const uint32_t* b, uint64_t& ah, uint64_t& bh) {
    uint32_t a2 = a[N - 1], a1 = a[N - 2], a0 = a[N - 3];
    uint32_t b2 = b[N - 1], b1 = b[N - 2], b0 = b[N - 3];
#pragma unroll
    for (int i = N - 4; i >= 0; i--) {
        bool shift = (a2 | b2) == 0;
        a2 = shift ? a1 : a2; a1 = shift ? a0 : a1; a0 = shift ? a[i] : a0;
        b2 = shift ? b1 : b2; b1 = shift ? b0 : b1; b0 = shift ? b[i] : b0;
    }
    if ((a2 | b2) == 0) {  // both below 2^64 (the window has slid down to limbs 2, 1, 0)
        ah = ((uint64_t)a1 << 32) | a0;
        bh = ((uint64_t)b1 << 32) | b0;
        return;
    }
    const uint32_t s = clz32(a2 | b2);  // 0..31
    uint64_t wa = ((uint64_t)a2 << 32) | a1, wb = ((uint64_t)b2 << 32) | b1;
    if (s) {
        wa = (wa << s) | (a0 >> (32 - s));
        wb = (wb << s) | (b0 >> (32 - s));
    }
    ah = ((wa >> STEPS) << STEPS) | (a[0] & LOW_MASK);
    bh = ((wb >> STEPS) << STEPS) | (b[0] & LOW_MASK);
}

// 30 binary-GCD steps on the approximations; the exact (a, b) must then be updated with the returned matrix
ZKP_HD void inner_steps(uint64_t ah, uint64_t bh, int32_t& f0, int32_t& g0, int32_t& f1, int32_t& g1) {
    f0 = 1; g0 = 0; f1 = 0; g1 = 1;
#pragma unroll 2
    for (int i = 0; i < STEPS; i++) {
        const bool odd = ah & 1;
        const bool sw = odd && ah < bh;
        const uint64_t ta = sw ? bh : ah, tb = sw ? ah : bh;
        const int32_t tf0 = sw ? f1 : f0, tf1 = sw ? f0 : f1, tg0 = sw ? g1 : g0, tg1 = sw ? g0 : g1;
        ah = odd ? ta - tb : ta;
        bh = tb;
        f0 = odd ? tf0 - tf1 : tf0;
        g0 = odd ? tg0 - tg1 : tg0;
        ah >>= 1;
        f1 = tf1 * 2;
        g1 = tg1 * 2;
    }
}

// r = |f x + g y| / 2^30 (exact division); returns true if f x + g y was negative
ZKP_HD bool lin_comb_shift(uint32_t* r, const uint32_t* x, const uint32_t* y, int32_t f, int32_t g) {
    uint32_t t[N + 1];
    int64_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        acc += (int64_t)f * (int64_t)x[i] + (int64_t)g * (int64_t)y[i];  // |.| <= 2^62 + carry
        t[i] = (uint32_t)acc;
        acc >>= 32;  // arithmetic
    }
    t[N] = (uint32_t)acc;
    const bool negative = acc < 0;
    // two's-complement negate the 13-limb value when negative
    uint32_t m = negative ? 0xffffffffu : 0u, carry = negative ? 1u : 0u;
#pragma unroll
    for (int i = 0; i <= N; i++) {
        uint64_t s = (uint64_t)(t[i] ^ m) + carry;
        t[i] = (uint32_t)s;
        carry = (uint32_t)(s >> 32);
    }
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = (t[i] >> STEPS) | (t[i + 1] << (32 - STEPS));
    return negative;
}

// r = (f u + g v) / 2^32 mod p, r in [0, p); u, v in [0, p); |f| + |g| <= 2^30
ZKP_HD void cofactor_update(uint32_t* r, const uint32_t* u, const uint32_t* v, int32_t f, int32_t g) {
    // negative coefficient: f u = |f| (p - u) mod p, so that everything below is unsigned
    const uint32_t af = (uint32_t)(f < 0 ? -(int64_t)f : (int64_t)f), ag = (uint32_t)(g < 0 ? -(int64_t)g : (int64_t)g);
    uint32_t uu[N], vv[N];
    {
        uint32_t bu = 0, bv = 0, zu = 0, zv = 0;
#pragma unroll
        for (int i = 0; i < N; i++) { zu |= u[i]; zv |= v[i]; }
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint64_t du = (uint64_t)FqParams::mod(i) - u[i] - bu, dv = (uint64_t)FqParams::mod(i) - v[i] - bv;
            bu = (uint32_t)(du >> 63);
            bv = (uint32_t)(dv >> 63);
            uu[i] = (f < 0 && zu) ? (uint32_t)du : u[i];
            vv[i] = (g < 0 && zv) ? (uint32_t)dv : v[i];
        }
    }
    uint32_t t[N + 2];
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        acc += (uint64_t)af * uu[i] + (uint64_t)ag * vv[i];  // < 2^62 + 2^62 + carry
        t[i] = (uint32_t)acc;
        acc >>= 32;
    }
    t[N] = (uint32_t)acc;
    t[N + 1] = 0;
    const uint32_t m = t[0] * FqParams::INV;
    acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        acc += (uint64_t)m * FqParams::mod(i) + t[i];
        t[i] = (uint32_t)acc;
        acc >>= 32;
    }
    acc += t[N];
    t[N] = (uint32_t)acc;
    t[N + 1] = (uint32_t)(acc >> 32);  // (t + m p) / 2^32 = t[1..13) < 1.25 p, so t[13] == 0
    // one conditional subtraction of p
    uint32_t d[N], borrow = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint64_t s = (uint64_t)t[i + 1] - FqParams::mod(i) - borrow;
        d[i] = (uint32_t)s;
        borrow = (uint32_t)(s >> 63);
    }
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = borrow ? t[i + 1] : d[i];
}

ZKP_HD bool is_zero12(const uint32_t* a) {
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < N; i++) z |= a[i];
    return z == 0;
}

}  // namespace inv

// x^-1 for x != 0, both in Montgomery form.  For x == 0 the result is 0.
ZKP_HD Fq fq_inverse(const Fq& x) {
    using namespace inv;
    uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = x.v[i]; b[i] = FqParams::mod(i); u[i] = 0; v[i] = 0; }
    u[0] = 1;
    int rounds = 0;
#pragma unroll 1
    for (; rounds < MAX_ROUNDS; rounds++) {
        bool done = is_zero12(a);
#if defined(__CUDA_ARCH__)
        if (__all_sync(__activemask(), done)) break;  // warp-uniform round count -> one constant for all lanes
#else
        if (done) break;
#endif
        uint64_t ah, bh;
        approximate(a, b, ah, bh);
        int32_t f0, g0, f1, g1;
        inner_steps(ah, bh, f0, g0, f1, g1);
        uint32_t na[N], nb[N], nu[N], nv[N];
        const bool sa = lin_comb_shift(na, a, b, f0, g0);
        const bool sb = lin_comb_shift(nb, a, b, f1, g1);
        if (sa) { f0 = -f0; g0 = -g0; }
        if (sb) { f1 = -f1; g1 = -g1; }
        cofactor_update(nu, u, v, f0, g0);
        cofactor_update(nv, u, v, f1, g1);
#pragma unroll
        for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
    }
    Fq vv, c;
#pragma unroll
    for (int i = 0; i < N; i++) { vv.v[i] = v[i]; c.v[i] = r_cubed(i); }
    Fq out = vv * c;  // v R^2
#pragma unroll 1
    for (int k = 0; k < 2 * rounds; k++) out = out + out;  // * 4^rounds
    // x == 0 never converges to b == 1 (b stays p): report 0
    bool unit = b[0] == 1;
#pragma unroll
    for (int i = 1; i < N; i++) unit = unit && b[i] == 0;
    return unit ? out : Fq::zero();
}

}  // namespace zkp
