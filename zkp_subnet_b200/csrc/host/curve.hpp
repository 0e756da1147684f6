// Host-side BLS12-381 group arithmetic (G1 over Fq, G2 over Fq2) in Jacobian coordinates, plus the
// ZCash / IETF point encodings.  Used for: the W-window fold that finishes every MSM, compression of
// result points, decompression + subgroup checks of proofs/commitments in worker_verify, and the
// SRS tooling.  Generic over the coordinate field F (needs + - * sqr dbl neg is_zero zero one).
#pragma once
#include "field64.hpp"

namespace zkp {
namespace host {

// ------------------------------------------------------------------------------------------ Fq2
struct Fq2 {
    Fq64 c0, c1;  // c0 + c1 u, u^2 = -1
    static Fq2 zero() { return {Fq64::zero(), Fq64::zero()}; }
    static Fq2 one() { return {Fq64::one(), Fq64::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fq2& o) const { return !(*this == o); }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fq2 conj() const { return {c0, c1.neg()}; }
    Fq2 operator*(const Fq2& o) const {
        Fq64 a = c0 * o.c0, b = c1 * o.c1;
        Fq64 c = (c0 + c1) * (o.c0 + o.c1);
        return {a - b, c - a - b};
    }
    Fq2 sqr() const {
        Fq64 a = (c0 + c1) * (c0 - c1);
        Fq64 b = (c0 * c1).dbl();
        return {a, b};
    }
    Fq2 mul_fq(const Fq64& k) const { return {c0 * k, c1 * k}; }
    Fq2 mul_by_nonresidue() const { return {c0 - c1, c0 + c1}; }  // * (1 + u)
    Fq2 inverse() const {
        Fq64 d = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * d, (c1 * d).neg()};
    }
};

// ------------------------------------------------------------------------------------------ Jacobian
template <class F>
struct Jac {
    F x, y, z;  // z == 0 -> infinity
    static Jac infinity() { return {F::zero(), F::one(), F::zero()}; }
    bool is_inf() const { return z.is_zero(); }
    static Jac from_affine(const F& ax, const F& ay) { return {ax, ay, F::one()}; }

    Jac dbl() const {  // dbl-2009-l (a = 0)
        if (is_inf()) return *this;
        F a = x.sqr(), b = y.sqr(), c = b.sqr();
        F d = ((x + b).sqr() - a - c).dbl();
        F e = a.dbl() + a;
        F f = e.sqr();
        Jac r;
        r.x = f - d.dbl();
        r.y = e * (d - r.x) - c.dbl().dbl().dbl();
        r.z = (y * z).dbl();
        return r;
    }
    Jac add(const Jac& o) const {  // add-2007-bl without the 2x scaling tricks
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        F z1z1 = z.sqr(), z2z2 = o.z.sqr();
        F u1 = x * z2z2, u2 = o.x * z1z1;
        F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
        F h = u2 - u1, rr = s2 - s1;
        if (h.is_zero()) {
            if (rr.is_zero()) return dbl();
            return infinity();
        }
        F hh = h.sqr(), hhh = hh * h, v = u1 * hh;
        Jac r;
        r.x = rr.sqr() - hhh - v.dbl();
        r.y = rr * (v - r.x) - s1 * hhh;
        r.z = z * o.z * h;
        return r;
    }
    Jac neg() const { return {x, y.neg(), z}; }
    // scalar: canonical little-endian 64-bit limbs
    Jac mul(const uint64_t* k, int limbs) const {
        Jac acc = infinity();
        for (int i = limbs * 64 - 1; i >= 0; i--) {
            acc = acc.dbl();
            if ((k[i >> 6] >> (i & 63)) & 1) acc = acc.add(*this);
        }
        return acc;
    }
    // the same with 4-bit windows: 15 precomputed multiples, then 4 doublings + at most one addition per window --
    // a fifth fewer field products than double-and-add for a dense scalar (the 255-bit scalars of worker_verify)
    Jac mul_w4(const uint64_t* k, int limbs) const {
        Jac tab[16];
        tab[0] = infinity();
        tab[1] = *this;
        tab[2] = dbl();
        for (int i = 3; i < 16; i++) tab[i] = tab[i - 1].add(*this);
        Jac acc = infinity();
        for (int i = limbs * 16 - 1; i >= 0; i--) {
            if (!acc.is_inf()) acc = acc.dbl().dbl().dbl().dbl();
            const unsigned d = (unsigned)(k[i >> 4] >> (4 * (i & 15))) & 15u;
            if (d) acc = acc.add(tab[d]);
        }
        return acc;
    }
    // returns false for infinity
    bool to_affine(F& ax, F& ay) const {
        if (is_inf()) return false;
        F zi = z.inverse(), zi2 = zi.sqr();
        ax = x * zi2;
        ay = y * zi2 * zi;
        return true;
    }
    bool equals(const Jac& o) const {
        if (is_inf() || o.is_inf()) return is_inf() && o.is_inf();
        F z1z1 = z.sqr(), z2z2 = o.z.sqr();
        return x * z2z2 == o.x * z1z1 && y * o.z * z2z2 == o.y * z * z1z1;
    }
};

using G1J = Jac<Fq64>;
using G2J = Jac<Fq2>;

static const uint64_t FR_MOD64[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                     0x73eda753299d7d48ull};

inline Fq64 fq_b4() { Fq64 r; memcpy(r.v, FqParams::B464, sizeof(r.v)); return r; }
inline G1J g1_generator() {
    Fq64 x, y;
    memcpy(x.v, FqParams::GX64, sizeof(x.v));
    memcpy(y.v, FqParams::GY64, sizeof(y.v));
    return G1J::from_affine(x, y);
}
inline bool g1_on_curve(const Fq64& x, const Fq64& y) { return y.sqr() == x.sqr() * x + fq_b4(); }
inline bool g1_in_subgroup_slow(const G1J& p) { return p.mul(FR_MOD64, 4).is_inf(); }
// Subgroup membership through the GLV endomorphism phi(x, y) = (beta x, y) (M. Scott, eprint 2021/1130 section 6;
// correctness: eprint 2022/352): P on the curve lies in G1 iff phi(P) = -[z^2] P, z the BLS parameter.  Two
// multiplications by the 64-bit, weight-6 |z| instead of one by the 255-bit r (~2.8x fewer group operations).
// The constant and the sign convention were checked against the big-int oracle on the generator, on a curve point
// outside the subgroup and on its cofactor-cleared image; g1_in_subgroup_slow stays for the host tests.
static const uint64_t BLS_Z_ABS = 0xd201000000010000ull;
inline bool g1_in_subgroup(const G1J& p) {
    if (p.is_inf()) return true;
    static const uint64_t BETA_MONT[6] = {0x30f1361b798a64e8ull, 0xf3b8ddab7ece5a2aull, 0x16a8ca3ac61577f7ull,
                                          0xc26a2ff874fd029bull, 0x3636b76660701c6eull, 0x051ba4ab241b6160ull};
    Fq64 beta;
    memcpy(beta.v, BETA_MONT, sizeof(BETA_MONT));
    const uint64_t z[1] = {BLS_Z_ABS};
    G1J z2p = p.mul(z, 1).mul(z, 1);           // [z^2] P (the two sign flips of z < 0 cancel)
    G1J phi = {p.x * beta, p.y, p.z};          // Jacobian (X, Y, Z) -> x = X/Z^2 scales by beta
    return phi.equals(z2p.neg());
}

// ZCash compressed G1 (48 B): bit7 compressed, bit6 infinity, bit5 y lexicographically largest.
inline void g1_compress(uint8_t out[48], const G1J& p) {
    Fq64 x, y;
    if (!p.to_affine(x, y)) { memset(out, 0, 48); out[0] = 0xc0; return; }
    x.to_be(out);
    out[0] |= 0x80;
    if (y.lexicographically_largest()) out[0] |= 0x20;
}
// returns false on any malformed / off-curve / wrong-subgroup input
inline bool g1_decompress(G1J& out, const uint8_t in[48], bool check_subgroup = true) {
    uint8_t flags = in[0] >> 5;
    if (!(flags & 4)) return false;
    uint8_t tmp[48];
    memcpy(tmp, in, 48);
    tmp[0] &= 0x1f;
    if (flags & 2) {
        if (flags & 1) return false;
        for (int i = 0; i < 48; i++) if (tmp[i]) return false;
        out = G1J::infinity();
        return true;
    }
    Fq64 x;
    if (!Fq64::from_be(x, tmp)) return false;
    Fq64 y2 = x.sqr() * x + fq_b4();
    // p = 3 mod 4: sqrt = y2^((p+1)/4)
    uint64_t e[6];
    {
        // (p + 1) / 4
        u128 c = 1;
        uint64_t t[6];
        for (int i = 0; i < 6; i++) { c += FqParams::MOD64[i]; t[i] = (uint64_t)c; c >>= 64; }
        for (int i = 0; i < 6; i++) e[i] = (t[i] >> 2) | (i + 1 < 6 ? t[i + 1] << 62 : 0);
    }
    Fq64 y = y2.pow(e, 6);
    if (y.sqr() != y2) return false;
    if (y.lexicographically_largest() != bool(flags & 1)) y = y.neg();
    out = G1J::from_affine(x, y);
    if (check_subgroup && !g1_in_subgroup(out)) return false;
    return true;
}
// ZCash uncompressed G1 (96 B)
inline void g1_serialize96(uint8_t out[96], const G1J& p) {
    Fq64 x, y;
    if (!p.to_affine(x, y)) { memset(out, 0, 96); out[0] = 0x40; return; }
    x.to_be(out);
    y.to_be(out + 48);
}
inline bool g1_deserialize96(G1J& out, const uint8_t in[96], bool check = true) {
    if (in[0] & 0x80) return false;
    if (in[0] & 0x40) { out = G1J::infinity(); return true; }
    Fq64 x, y;
    if (!Fq64::from_be(x, in) || !Fq64::from_be(y, in + 48)) return false;
    if (check && !g1_on_curve(x, y)) return false;
    out = G1J::from_affine(x, y);
    return true;
}

// G2 generator (twist y^2 = x^3 + 4(1+u)), canonical big-endian hex from SURVEY.md section 8c
inline Fq64 fq_from_hex(const char* hex) {
    uint8_t be[48];
    for (int i = 0; i < 48; i++) {
        auto nib = [](char c) -> int { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
        be[i] = (uint8_t)(nib(hex[2 * i]) << 4 | nib(hex[2 * i + 1]));
    }
    Fq64 r;
    Fq64::from_be(r, be);
    return r;
}
inline G2J g2_generator() {
    Fq2 x = {fq_from_hex("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"),
             fq_from_hex("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e")};
    Fq2 y = {fq_from_hex("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801"),
             fq_from_hex("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be")};
    return G2J::from_affine(x, y);
}

}  // namespace host
}  // namespace zkp
