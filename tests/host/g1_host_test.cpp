// Host-side exerciser for g1.cuh (XYZZ formulas incl. exceptional cases); prints affine results
// as hex for the Python checker in tests/test_host_limbs.py.
#include <cstdio>
#include "../../zkp_subnet_b200/csrc/g1.cuh"
using namespace zkp;
static void pr(const char* tag, const G1Xyzz& p) {
    if (p.is_inf()) { printf("%s inf\n", tag); return; }
    Fq x = (p.x * p.zz.inverse()).from_mont(), y = (p.y * p.zzz.inverse()).from_mont();
    printf("%s ", tag);
    for (int i = 11; i >= 0; i--) printf("%08x", x.v[i]);
    printf(" ");
    for (int i = 11; i >= 0; i--) printf("%08x", y.v[i]);
    printf("\n");
}
int main() {
    G1Affine g;
    for (int i = 0; i < 12; i++) { g.x.v[i] = FqParams::GX[i]; g.y.v[i] = FqParams::GY[i]; }
    G1Xyzz acc = G1Xyzz::infinity();
    acc.madd(g, 0); pr("1", acc);            // inf + G
    acc.madd(g, 0); pr("2", acc);            // G + G  (doubling branch)
    acc.madd(g, 0); pr("3", acc);
    G1Xyzz three = acc;
    for (int i = 4; i <= 10; i++) acc.madd(g, 0);
    pr("10", acc);
    G1Xyzz t = acc; t.add(acc); pr("20", t);  // add with equal operands (doubling branch)
    t.add(three); pr("23", t);
    t = three; t.add(acc); pr("13", t);
    t = acc.dbl(); pr("20", t);
    t = acc; t.madd(g, 1); pr("9", t);        // subtract G
    t = G1Xyzz::from_affine(g, 0); t.madd(g, 1); pr("0", t);   // G - G
    t = three; G1Xyzz n3 = three; n3.y = n3.y.neg(); t.add(n3); pr("0", t);
    t = G1Xyzz::infinity(); t.add(three); pr("3", t);
    t = three; t.add(G1Xyzz::infinity()); pr("3", t);
    G1Affine inf; inf.x = Fq::zero(); inf.y = Fq::zero();
    t = three; t.madd(inf, 0); pr("3", t);
    // non-trivial ZZ on both sides: (2G as xyzz with zz != 1) + (3G)
    G1Xyzz two = G1Xyzz::dbl_affine(g.x, g.y); t = two; t.add(three); pr("5", t);
    t = two; t.add(two); pr("4", t);
    t = two.dbl(); pr("4", t);
    return 0;
}
