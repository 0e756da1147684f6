// Host-side exerciser for the lazily reduced Fq helpers (ff.cuh mul_lazy / sqr_lazy / mul2 on wide operands, sub_p2,
// rsub_p2, add_raw, sub_fix) and for G1Xyzz::madd_lazy (g1.cuh).
//   field mode:  reads "a b" pairs (96 hex digits each, any 384-bit values) from stdin and prints the raw results;
//                tests/test_host_limbs.py compares them with exact big-integer formulas.
//   group mode:  runs chains of lazy additions beside the fully reduced madd, exceptional cases included, checks the
//                coordinate ranges claimed in tools/lazy_bounds.py, prints "ok <count>" or the first failure.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../zkp_subnet_b200/csrc/g1.cuh"
using namespace zkp;

static uint32_t KP[FQ_KP_ROWS * 12];
static void pr(const char* tag, const Fq& a) {
    printf("%s ", tag);
    for (int i = 11; i >= 0; i--) printf("%08x", a.v[i]);
    printf("\n");
}
static bool rd(const char* h, Fq& a) {
    if (strlen(h) != 96) return false;
    for (int i = 0; i < 12; i++) {
        char buf[9];
        memcpy(buf, h + 8 * (11 - i), 8);
        buf[8] = 0;
        a.v[i] = (uint32_t)strtoul(buf, nullptr, 16);
    }
    return true;
}
// a < k p ?  (k < 8)
static bool below(const Fq& a, uint32_t k) {
    for (int i = 11; i >= 0; i--) {
        uint32_t m = KP[k * 12 + i];
        if (a.v[i] != m) return a.v[i] < m;
    }
    return false;
}
static bool same_point(const G1Xyzz& a, const G1Xyzz& b) {  // both canonical
    if (a.is_inf() || b.is_inf()) return a.is_inf() && b.is_inf();
    return a.x * b.zz == b.x * a.zz && a.y * b.zzz == b.y * a.zzz;
}
static G1Affine to_affine(const G1Xyzz& p) {
    G1Affine r;
    r.x = p.x * p.zz.inverse();
    r.y = p.y * p.zzz.inverse();
    return r;
}

static int group_mode() {
    G1Affine g;
    for (int i = 0; i < 12; i++) { g.x.v[i] = FqParams::GX[i]; g.y.v[i] = FqParams::GY[i]; }
    // a few affine multiples of G: 1, 2, 3, 5, 8, 13, ...
    G1Affine pts[10];
    {
        G1Xyzz a = G1Xyzz::from_affine(g, 0), b = G1Xyzz::dbl_affine(g.x, g.y);
        for (int i = 0; i < 10; i++) {
            pts[i] = to_affine(a);
            G1Xyzz c = a;
            c.add(b);
            a = b;
            b = c;
        }
    }
    int checks = 0;
    uint64_t s = 0x243F6A8885A308D3ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    auto check = [&](const G1Xyzz& lazy, const G1Xyzz& ref, const char* what, int step) -> bool {
        if (!lazy.is_inf() && !(below(lazy.x, 2) && below(lazy.y, 2) && below(lazy.zz, 2) && below(lazy.zzz, 2))) {
            printf("range violated in %s at step %d\n", what, step);
            return false;
        }
        G1Xyzz n = lazy;
        n.normalize();
        if (!n.is_inf() && !(below(n.x, 1) && below(n.y, 1) && below(n.zz, 1) && below(n.zzz, 1))) {
            printf("normalize not canonical in %s at step %d\n", what, step);
            return false;
        }
        if (!same_point(n, ref)) {
            printf("mismatch in %s at step %d\n", what, step);
            return false;
        }
        checks++;
        return true;
    };
    // long random chains from infinity, with the exceptional cases injected while the accumulator is loose
    for (int chain = 0; chain < 6; chain++) {
        G1Xyzz lazy = G1Xyzz::infinity(), ref = G1Xyzz::infinity();
        for (int step = 0; step < 60; step++) {
            G1Affine p = pts[rnd() % 10];
            if (rnd() & 1) p.y = p.y.neg();
            int special = step % 20;
            if (special == 7 && !ref.is_inf()) p = to_affine(ref);                                  // P + P while loose
            if (special == 13 && !ref.is_inf()) { p = to_affine(ref); p.y = p.y.neg(); }            // P + (-P)
            lazy.madd_lazy(p.x, p.y, KP);
            ref.madd(p, 0);
            if (!check(lazy, ref, "chain", chain * 100 + step)) return 1;
        }
    }
    // the same point over and over (2G, 3G, ... : doubling first, then ordinary additions)
    {
        G1Xyzz lazy = G1Xyzz::infinity(), ref = G1Xyzz::infinity();
        for (int step = 0; step < 40; step++) {
            lazy.madd_lazy(g.x, g.y, KP);
            ref.madd(g, 0);
            if (!check(lazy, ref, "repeat", step)) return 1;
        }
    }
    printf("ok %d\n", checks);
    return 0;
}

int main(int argc, char** argv) {
    for (uint32_t k = 0; k < FQ_KP_ROWS; k++)
        for (int i = 0; i < 12; i++) KP[k * 12 + i] = fq_kp_limb(k, i);
    if (argc > 1 && !strcmp(argv[1], "group")) return group_mode();
    char ha[256], hb[256];
    while (scanf("%255s %255s", ha, hb) == 2) {
        Fq a, b;
        if (!rd(ha, a) || !rd(hb, b)) return 2;
        pr("a", a); pr("b", b);
        pr("mul", Fq::mul_lazy(a, b));
        pr("sqr", a.sqr_lazy());
        pr("add", Fq::add_raw(a, b));
        pr("subp2", Fq::sub_p2(a, b));
        pr("rsubp2", b.rsub_p2());
        pr("subfix", Fq::sub_fix(a, b, KP));
        pr("mul2", Fq::mul2(a, b, b, a));
    }
    return 0;
}
