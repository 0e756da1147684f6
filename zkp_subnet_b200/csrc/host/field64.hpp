// Host-side (CPU) Montgomery arithmetic on 64-bit limbs for the parts of the path that stay on the
// host by design: the final fold of the W window sums of an MSM, affine conversion + ZCash
// compression of the single result point, G1/G2 decompression and the pairing check of
// worker_verify (reference neurons/validator.py:77-86 -> fourier Client.worker_verify).
// Same Montgomery domain as the device code (R = 2^384 for Fq, 2^256 for Fr), so limbs can be
// memcpy'd between the two.
#pragma once
#include <stdint.h>
#include <string.h>
#include "../field_params.h"
#include "mont_asm.hpp"
#if defined(__x86_64__) && defined(__GNUC__)
#include <x86intrin.h>
#endif

namespace zkp {
namespace host {

typedef unsigned __int128 u128;

template <class P>
struct F64 {
    static constexpr int N = P::N64;
    uint64_t v[N];

    static F64 zero() { F64 r; memset(r.v, 0, sizeof(r.v)); return r; }
    static F64 one() { F64 r; memcpy(r.v, P::ONE64, sizeof(r.v)); return r; }
    static F64 r2() { F64 r; memcpy(r.v, P::R264, sizeof(r.v)); return r; }
    static F64 from_u64(uint64_t x) { F64 r = zero(); r.v[0] = x; return r.to_mont(); }

    bool is_zero() const { uint64_t a = 0; for (int i = 0; i < N; i++) a |= v[i]; return a == 0; }
    bool operator==(const F64& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const F64& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > P::MOD64[i]) return true;
            if (a[i] < P::MOD64[i]) return false;
        }
        return true;
    }
    static void sub_mod_raw(uint64_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a[i] - P::MOD64[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    // Additions are branchless (the comparison with the modulus is an unpredictable branch otherwise: a third of the time
    // of an Fq12 product went there).  The moduli leave the top bit of the top limb clear, so a + b never carries out.
    F64 operator+(const F64& o) const {
        F64 r;
#if defined(__x86_64__) && defined(__GNUC__)
        unsigned long long t[N], d[N];
        unsigned char c = 0, b = 0;
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) c = _addcarry_u64(c, v[i], o.v[i], &t[i]);
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) b = _subborrow_u64(b, t[i], P::MOD64[i], &d[i]);
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) r.v[i] = b ? t[i] : d[i];  // borrow: t < p (cmov)
#else
        uint64_t t[N], d[N];
        uint64_t c = 0;
        for (int i = 0; i < N; i++) { u128 s = (u128)v[i] + o.v[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        uint64_t b = 0;
        for (int i = 0; i < N; i++) { u128 e = (u128)t[i] - P::MOD64[i] - b; d[i] = (uint64_t)e; b = (uint64_t)(e >> 64) & 1; }
        const uint64_t keep = (uint64_t)0 - b;  // all ones: t < p, keep t
        for (int i = 0; i < N; i++) r.v[i] = (t[i] & keep) | (d[i] & ~keep);
#endif
        return r;
    }
    F64 operator-(const F64& o) const {
        F64 r;
#if defined(__x86_64__) && defined(__GNUC__)
        unsigned long long t[N], d[N];
        unsigned char b = 0, c = 0;
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) b = _subborrow_u64(b, v[i], o.v[i], &t[i]);
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) c = _addcarry_u64(c, t[i], P::MOD64[i], &d[i]);
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) r.v[i] = b ? d[i] : t[i];  // borrow: add the modulus back
#else
        uint64_t t[N];
        uint64_t b = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)v[i] - o.v[i] - b; t[i] = (uint64_t)d; b = (uint64_t)(d >> 64) & 1; }
        const uint64_t mask = (uint64_t)0 - b;
        uint64_t c = 0;
        for (int i = 0; i < N; i++) { u128 s = (u128)t[i] + (P::MOD64[i] & mask) + c; r.v[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
#endif
        return r;
    }
    F64 neg() const { return zero() - *this; }
    F64 dbl() const { return *this + *this; }
    // {modulus limbs, -p^-1 mod 2^64}: the constants block of the assembly products (mont_asm.hpp)
    static const uint64_t* asm_consts() {
        static const struct Block { uint64_t w[N + 1]; Block() { memcpy(w, P::MOD64, 8 * N); w[N] = P::INV64; } } b;
        return b.w;
    }
    F64 operator*(const F64& o) const {
#ifdef ZKP_HOST_MONT_ASM
        if (have_mulx_adx()) {
            F64 r;
            if (N == 6) { zkp_host_mont_mul_6(r.v, v, o.v, asm_consts()); return r; }
            if (N == 4) { zkp_host_mont_mul_4(r.v, v, o.v, asm_consts()); return r; }
        }
#endif
        return mul_portable(o);
    }
    F64 mul_portable(const F64& o) const {
        uint64_t t[N + 2];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < N; i++) {
            u128 c = 0;
            for (int j = 0; j < N; j++) { c += (u128)v[j] * o.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P::INV64;
            c = ((u128)m * P::MOD64[0] + t[0]) >> 64;
            for (int j = 1; j < N; j++) { c += (u128)m * P::MOD64[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        F64 r;
        memcpy(r.v, t, sizeof(r.v));
        if (t[N] || geq_mod(r.v)) sub_mod_raw(r.v);
        return r;
    }
    F64 sqr() const { return *this * *this; }
    F64 to_mont() const { return *this * r2(); }
    F64 from_mont() const { F64 o = zero(); o.v[0] = 1; return *this * o; }

    // exponent: little-endian 64-bit limbs
    F64 pow(const uint64_t* e, int elimbs) const {
        F64 acc = one();
        for (int i = elimbs * 64 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
        }
        return acc;
    }
    F64 inverse_fermat() const {  // a^(p-2); 0 -> 0
        uint64_t e[N];
        uint64_t borrow = 2;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)P::MOD64[i] - borrow;
            e[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        return pow(e, N);
    }
    // 0 -> 0.  Binary extended Euclid on the limbs (the values are public: no constant-time requirement): at most
    // 2 log2(p) halvings and log2(p) subtractions of N-limb integers, ~7x faster than the 1.5 log2(p) Montgomery products
    // of the Fermat form -- and the inversion inside g1_compress is what a request waits for after the last kernel of
    // its proof MSM has finished.  Works on the Montgomery representative v = aR: w = v^-1 = a^-1 R^-1 as an integer,
    // then two Montgomery products by R^2 give a^-1 R.
    F64 inverse() const {
        if (is_zero()) return zero();
        uint64_t u[N], w[N], x1[N], x2[N];
        memcpy(u, v, sizeof(u));
        memcpy(w, P::MOD64, sizeof(w));
        memset(x1, 0, sizeof(x1));
        memset(x2, 0, sizeof(x2));
        x1[0] = 1;
        auto shr1 = [](uint64_t* a, uint64_t top) {
            for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
            a[N - 1] = (a[N - 1] >> 1) | (top << 63);
        };
        auto halve_mod = [&](uint64_t* x) {  // x / 2 mod p, x < p
            uint64_t c = 0;
            if (x[0] & 1) {
                for (int i = 0; i < N; i++) { u128 t = (u128)x[i] + P::MOD64[i] + c; x[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
            }
            shr1(x, c);
        };
        auto geq = [](const uint64_t* a, const uint64_t* b) {
            for (int i = N - 1; i >= 0; i--) {
                if (a[i] > b[i]) return true;
                if (a[i] < b[i]) return false;
            }
            return true;
        };
        auto sub = [](uint64_t* a, const uint64_t* b) -> uint64_t {  // a -= b, returns the borrow
            uint64_t br = 0;
            for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - b[i] - br; a[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
            return br;
        };
        auto sub_mod = [&](uint64_t* a, const uint64_t* b) {  // a = a - b mod p, both < p
            if (sub(a, b)) {
                uint64_t c = 0;
                for (int i = 0; i < N; i++) { u128 t = (u128)a[i] + P::MOD64[i] + c; a[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
            }
        };
        auto is_one = [](const uint64_t* a) {
            uint64_t r = a[0] ^ 1;
            for (int i = 1; i < N; i++) r |= a[i];
            return r == 0;
        };
        auto is_nil = [](const uint64_t* a) {
            uint64_t r = 0;
            for (int i = 0; i < N; i++) r |= a[i];
            return r == 0;
        };
        while (!is_one(u) && !is_one(w)) {
            while (!(u[0] & 1)) { shr1(u, 0); halve_mod(x1); }
            while (!(w[0] & 1)) { shr1(w, 0); halve_mod(x2); }
            if (geq(u, w)) { sub(u, w); sub_mod(x1, x2); }
            else { sub(w, u); sub_mod(x2, x1); }
            // gcd(v, p) != 1: only a non-canonical representative of 0 (v = p) gets here; answer like inverse(0)
            if (is_nil(u) || is_nil(w)) return zero();
        }
        F64 r;
        memcpy(r.v, is_one(u) ? x1 : x2, sizeof(r.v));
        return (r * r2()) * r2();
    }

    // canonical big-endian bytes (N*8) <-> Montgomery form.  from_be returns false if >= modulus.
    static bool from_be(F64& out, const uint8_t* be) {
        F64 c;
        for (int i = 0; i < N; i++) {
            uint64_t w = 0;
            for (int k = 0; k < 8; k++) w = (w << 8) | be[(N - 1 - i) * 8 + k];
            c.v[i] = w;
        }
        if (geq_mod(c.v)) return false;
        out = c.to_mont();
        return true;
    }
    void to_be(uint8_t* be) const {
        F64 c = from_mont();
        for (int i = 0; i < N; i++)
            for (int k = 0; k < 8; k++) be[(N - 1 - i) * 8 + k] = (uint8_t)(c.v[i] >> (56 - 8 * k));
    }
    // canonical value compare: this > (p-1)/2 ?   (this in Montgomery form)
    bool lexicographically_largest() const {
        F64 c = from_mont();
        F64 n = neg().from_mont();
        for (int i = N - 1; i >= 0; i--) {
            if (c.v[i] > n.v[i]) return true;
            if (c.v[i] < n.v[i]) return false;
        }
        return false;
    }
};

using Fq64 = F64<FqParams>;
using Fr64 = F64<FrParams>;

}  // namespace host
}  // namespace zkp
