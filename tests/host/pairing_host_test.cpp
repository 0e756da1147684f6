// Host-side exerciser for csrc/host/pairing.hpp and curve.hpp: the fast paths (complex Fq12 squaring, Granger-Scott
// cyclotomic squaring, endomorphism subgroup check) against their plain definitions.  Prints one "name ok|BAD" line
// per check for tests/test_host_limbs.py.
#include <cstdio>
#include "../../zkp_subnet_b200/csrc/host/pairing.hpp"
using namespace zkp::host;
static uint64_t s = 88172645463325252ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main() {
    int bad_sqr = 0, bad_cyc = 0;
    for (int t = 0; t < 8; t++) {
        Fq12 f = Fq12::one();
        for (int i = 0; i < 6; i++) {
            Fq2& c = f.coeff(i);
            for (int k = 0; k < 6; k++) { c.c0.v[k] = rnd(); c.c1.v[k] = rnd(); }
            c.c0.v[5] &= 0x0fffffffffffffffull; c.c1.v[5] &= 0x0fffffffffffffffull;
        }
        if (!(f.sqr() == f * f)) bad_sqr++;
        Fq12 c = f.conj() * f.inverse();          // easy part of the final exponentiation: lands in the
        c = frobenius(frobenius(c)) * c;          // cyclotomic subgroup
        if (!(c.cyclotomic_sqr() == c * c)) bad_cyc++;
        if (!(c.cyclotomic_sqr().cyclotomic_sqr() == (c * c) * (c * c))) bad_cyc++;
    }
    // sparse product by a Miller-loop line against the general product by line_eval's element
    int bad_line = 0;
    for (int t = 0; t < 16; t++) {
        Fq12 f = Fq12::one();
        for (int i = 0; i < 6; i++) {
            Fq2& c = f.coeff(i);
            for (int k = 0; k < 6; k++) { c.c0.v[k] = rnd(); c.c1.v[k] = rnd(); }
            c.c0.v[5] &= 0x0fffffffffffffffull; c.c1.v[5] &= 0x0fffffffffffffffull;
        }
        G2Lines::Line l;
        Fq64 px, py;
        for (int k = 0; k < 6; k++) { l.lambda.c0.v[k] = rnd(); l.lambda.c1.v[k] = rnd(); l.c.c0.v[k] = rnd(); l.c.c1.v[k] = rnd(); px.v[k] = rnd(); py.v[k] = rnd(); }
        l.lambda.c0.v[5] &= 0x0fffffffffffffffull; l.lambda.c1.v[5] &= 0x0fffffffffffffffull; l.c.c0.v[5] &= 0x0fffffffffffffffull;
        l.c.c1.v[5] &= 0x0fffffffffffffffull; px.v[5] &= 0x0fffffffffffffffull; py.v[5] &= 0x0fffffffffffffffull;
        if (t == 0) py = Fq64::zero();
        if (t == 1) l.c = Fq2::zero();
        if (!(f.mul_by_line(l.c, l.lambda.mul_fq(px), py.neg()) == f * line_eval(l, px, py))) bad_line++;
    }
    printf("fq12_complex_sqr %s\n", bad_sqr ? "BAD" : "ok");
    printf("mul_by_line %s\n", bad_line ? "BAD" : "ok");
    printf("cyclotomic_sqr %s\n", bad_cyc ? "BAD" : "ok");
    int bad_sub = 0, in = 0, out = 0;
    uint64_t e[6];
    { u128 c = 1; uint64_t tt[6];
      for (int i = 0; i < 6; i++) { c += zkp::FqParams::MOD64[i]; tt[i] = (uint64_t)c; c >>= 64; }
      for (int i = 0; i < 6; i++) e[i] = (tt[i] >> 2) | (i + 1 < 6 ? tt[i + 1] << 62 : 0); }
    G1J g = g1_generator();
    for (int t = 0; t < 60; t++) {
        Fq64 x = Fq64::from_u64(5 + t), y2 = x.sqr() * x + fq_b4(), y = y2.pow(e, 6);
        if (y.sqr() == y2) {                      // a curve point, usually outside the prime-order subgroup
            G1J p = G1J::from_affine(x, y);
            bool a = g1_in_subgroup(p), b = g1_in_subgroup_slow(p);
            if (a != b) bad_sub++;
            (b ? in : out)++;
        }
        uint64_t k[2] = {rnd(), rnd()};
        G1J q = g.mul(k, 2);                      // a subgroup point with Z != 1
        if (!g1_in_subgroup(q)) bad_sub++;
        in++;
    }
    if (!g1_in_subgroup(G1J::infinity())) bad_sub++;
    // windowed scalar multiplication against double-and-add: dense, sparse, short, zero and all-ones scalars
    int bad_w4 = 0;
    for (int t = 0; t < 40; t++) {
        uint64_t k[4] = {rnd(), rnd(), rnd(), rnd() >> 2};
        if (t == 0) k[0] = k[1] = k[2] = k[3] = 0;
        if (t == 1) { k[0] = 1; k[1] = k[2] = k[3] = 0; }
        if (t == 2) { k[0] = k[1] = k[2] = ~0ull; k[3] = ~0ull >> 2; }
        if (t == 3) { k[0] = 0; k[1] = 0; k[2] = 0; k[3] = 1ull << 60; }
        if (t == 4) { k[0] = 0xf0f0f0f0f0f0f0f0ull; k[1] = 0x0f0f0f0f0f0f0f0full; k[2] = 0x1000000000000001ull; k[3] = 0; }
        uint64_t kk[2] = {rnd(), rnd()};
        G1J q = g.mul(kk, 2);
        if (!q.mul_w4(k, 4).equals(q.mul(k, 4))) bad_w4++;
        if (!q.mul_w4(k, 2).equals(q.mul(k, 2))) bad_w4++;
        if (!G1J::infinity().mul_w4(k, 4).is_inf()) bad_w4++;
    }
    printf("mul_w4 %s\n", bad_w4 ? "BAD" : "ok");
    printf("subgroup_endomorphism %s (%d inside, %d outside)\n", (bad_sub || !out) ? "BAD" : "ok", in, out);
    return 0;
}
