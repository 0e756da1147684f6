// Host-side optimal-ate pairing on BLS12-381 for worker_verify
// (reference neurons/validator.py:77-86,168-170 -> fourier Client.worker_verify).
// The verification equation  e(C - y*S_i, g2) == e(pi, [tau - alpha]_2)  is rearranged to
//     e(C - y*S_i + alpha*pi, g2) * e(-pi, [tau]_2) == 1
// so that both G2 arguments are fixed per SRS: their Miller-loop line coefficients are computed once
// (affine, at SRS load) and every verification costs two line-evaluation loops that share one
// accumulator squaring chain plus one final exponentiation.
// Tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - (1+u)), Fq12 = Fq6[w]/(w^2 - v).
#pragma once
#include <vector>

#include "curve.hpp"

namespace zkp {
namespace host {

struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    Fq6 operator*(const Fq6& o) const {
        Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
        Fq2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_by_nonresidue() + t0;
        Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_by_nonresidue();
        Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
        return {r0, r1, r2};
    }
    Fq6 mul_by_v() const { return {c2.mul_by_nonresidue(), c0, c1}; }
    Fq6 inverse() const {
        Fq2 t0 = c0.sqr() - (c1 * c2).mul_by_nonresidue();
        Fq2 t1 = c2.sqr().mul_by_nonresidue() - c0 * c1;
        Fq2 t2 = c1.sqr() - c0 * c2;
        Fq2 d = (c0 * t0 + (c2 * t1 + c1 * t2).mul_by_nonresidue()).inverse();
        return {t0 * d, t1 * d, t2 * d};
    }
};

struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq12 operator*(const Fq12& o) const {
        Fq6 t0 = c0 * o.c0, t1 = c1 * o.c1;
        Fq6 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1;
        return {t0 + t1.mul_by_v(), r1};
    }
    // (c0 + c1 w)^2 = (c0^2 + v c1^2) + 2 c0 c1 w with two Fq6 products: (c0 + c1)(c0 + v c1) = c0^2 + v c1^2 + (1 + v) c0 c1
    Fq12 sqr() const {
        Fq6 ab = c0 * c1;
        Fq6 t = (c0 + c1) * (c0 + c1.mul_by_v());
        return {t - ab - ab.mul_by_v(), ab + ab};
    }
    // Squaring in the cyclotomic subgroup (after the easy part of the final exponentiation): Granger-Scott, three
    // Fq4 squarings = 9 Fq2 squarings instead of 12 Fq2 products.
    static void fq4_square(const Fq2& a, const Fq2& b, Fq2& o0, Fq2& o1) {
        Fq2 t0 = a.sqr(), t1 = b.sqr();
        o0 = t1.mul_by_nonresidue() + t0;
        o1 = (a + b).sqr() - t0 - t1;
    }
    Fq12 cyclotomic_sqr() const {
        Fq2 z0 = c0.c0, z4 = c0.c1, z3 = c0.c2, z2 = c1.c0, z1 = c1.c1, z5 = c1.c2;
        Fq2 t0, t1, t2, t3;
        fq4_square(z0, z1, t0, t1);
        z0 = (t0 - z0).dbl() + t0;
        z1 = (t1 + z1).dbl() + t1;
        fq4_square(z2, z3, t0, t1);
        fq4_square(z4, z5, t2, t3);
        z4 = (t0 - z4).dbl() + t0;
        z5 = (t1 + z5).dbl() + t1;
        t0 = t3.mul_by_nonresidue();
        z2 = (t0 + z2).dbl() + t0;
        z3 = (t2 - z3).dbl() + t2;
        return {{z0, z4, z3}, {z2, z1, z5}};
    }
    // this * ((a + b v) + (c v) w) with c in Fq: the shape of a Miller-loop line (line_eval below), 12 Fq2-product
    // equivalents instead of the 18 of a general product
    static Fq6 fq6_mul_by_01(const Fq6& x, const Fq2& a, const Fq2& b) {
        Fq2 xa = x.c0 * a, xb = x.c1 * b;
        Fq2 r1 = (x.c0 + x.c1) * (a + b) - xa - xb;
        return {xa + (x.c2 * b).mul_by_nonresidue(), r1, xb + x.c2 * a};
    }
    Fq12 mul_by_line(const Fq2& a, const Fq2& b, const Fq64& c) const {
        Fq6 t0 = fq6_mul_by_01(c0, a, b);
        Fq6 t1 = {c1.c2.mul_fq(c).mul_by_nonresidue(), c1.c0.mul_fq(c), c1.c1.mul_fq(c)};  // c1 * (c v)
        Fq2 bc = {b.c0 + c, b.c1};
        Fq6 r1 = fq6_mul_by_01(c0 + c1, a, bc) - t0 - t1;
        return {t0 + t1.mul_by_v(), r1};
    }
    Fq12 conj() const { return {c0, c1.neg()}; }
    Fq12 inverse() const {
        Fq6 d = (c0 * c0 - (c1 * c1).mul_by_v()).inverse();
        return {c0 * d, (c1 * d).neg()};
    }
    // coefficient of w^i (i = 0..5) as a reference
    Fq2& coeff(int i) {
        Fq6& h = (i & 1) ? c1 : c0;
        int k = i >> 1;
        return k == 0 ? h.c0 : (k == 1 ? h.c1 : h.c2);
    }
};

struct PairingConsts {
    Fq2 gamma[6];  // (1+u)^(i (p-1)/6)
    PairingConsts() {
        // (p - 1) / 6
        uint64_t e[6];
        u128 rem = 0;
        uint64_t pm1[6];
        memcpy(pm1, FqParams::MOD64, sizeof(pm1));
        pm1[0] -= 1;
        for (int i = 5; i >= 0; i--) {
            u128 cur = (rem << 64) | pm1[i];
            e[i] = (uint64_t)(cur / 6);
            rem = cur % 6;
        }
        Fq2 xi = {Fq64::one(), Fq64::one()};
        Fq2 g = Fq2::one();
        for (int i = 6 * 64 - 1; i >= 0; i--) {
            g = g.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) g = g * xi;
        }
        gamma[0] = Fq2::one();
        for (int i = 1; i < 6; i++) gamma[i] = gamma[i - 1] * g;
    }
};
inline const PairingConsts& pairing_consts() {
    static const PairingConsts c;
    return c;
}

inline Fq12 frobenius(const Fq12& f) {
    const PairingConsts& pc = pairing_consts();
    Fq12 r = f;
    for (int i = 0; i < 6; i++) r.coeff(i) = r.coeff(i).conj() * pc.gamma[i];
    return r;
}

static const uint64_t BLS_X_ABS = 0xd201000000010000ull;  // |z|, z < 0

// f^z for f in the cyclotomic subgroup (inverse = conjugate), z = -|z|
inline Fq12 exp_by_z(const Fq12& f) {
    Fq12 acc = Fq12::one();
    for (int i = 63; i >= 0; i--) {
        acc = acc.cyclotomic_sqr();
        if ((BLS_X_ABS >> i) & 1) acc = acc * f;
    }
    return acc.conj();
}

// f^(3 (p^12 - 1)/r): easy part, then hard part as (z-1)^2 (z+p) (z^2+p^2-1) + 3.
// gcd(3, r) = 1, so the result is 1 iff the reduced pairing is 1.
inline Fq12 final_exponentiation(const Fq12& f) {
    Fq12 t = f.conj() * f.inverse();          // f^(p^6 - 1)
    t = frobenius(frobenius(t)) * t;          // ^(p^2 + 1)
    Fq12 a = exp_by_z(t) * t.conj();          // t^(z-1)
    Fq12 b = exp_by_z(a) * a.conj();          // t^((z-1)^2)
    Fq12 c = exp_by_z(b) * frobenius(b);      // ^(z+p)
    Fq12 d = exp_by_z(exp_by_z(c)) * frobenius(frobenius(c)) * c.conj();  // ^(z^2+p^2-1)
    return d * t.cyclotomic_sqr() * t;        // * t^3
}

// Line coefficients of the Miller loop for a fixed Q on the twist (affine arithmetic).
struct G2Lines {
    struct Line { Fq2 lambda, c; };  // evaluated at P: c + (lambda xP) w^2 - yP w^3
    std::vector<Line> lines;
    bool infinity = true;
};
inline G2Lines g2_precompute(const G2J& q) {
    G2Lines out;
    Fq2 qx, qy;
    if (!q.to_affine(qx, qy)) return out;
    out.infinity = false;
    Fq2 rx = qx, ry = qy;
    for (int i = 62; i >= 0; i--) {
        Fq2 xx = rx.sqr();
        Fq2 lam = (xx.dbl() + xx) * ry.dbl().inverse();
        out.lines.push_back({lam, ry - lam * rx});
        Fq2 nx = lam.sqr() - rx.dbl();
        Fq2 ny = lam * (rx - nx) - ry;
        rx = nx; ry = ny;
        if ((BLS_X_ABS >> i) & 1) {
            Fq2 l2 = (qy - ry) * (qx - rx).inverse();
            out.lines.push_back({l2, ry - l2 * rx});
            Fq2 mx = l2.sqr() - rx - qx;
            Fq2 my = l2 * (rx - mx) - ry;
            rx = mx; ry = my;
        }
    }
    return out;
}

inline Fq12 line_eval(const G2Lines::Line& l, const Fq64& px, const Fq64& py) {
    Fq12 r = {Fq6::zero(), Fq6::zero()};
    r.c0.c0 = l.c;                          // w^0
    r.c0.c1 = l.lambda.mul_fq(px);          // w^2
    r.c1.c1 = {py.neg(), Fq64::zero()};     // w^3
    return r;
}

// prod_k e(P_k, Q_k) == 1 ?   P_k affine G1 (skip flag for infinity), Q_k as precomputed lines
struct G1AffineHost { Fq64 x, y; bool inf; };
inline bool pairing_product_is_one(const std::vector<G1AffineHost>& ps, const std::vector<const G2Lines*>& qs) {
    Fq12 f = Fq12::one();
    std::vector<size_t> cursor(ps.size(), 0);
    for (int i = 62; i >= 0; i--) {
        f = f.sqr();
        for (size_t k = 0; k < ps.size(); k++) {
            if (ps[k].inf || qs[k]->infinity) continue;
            const Fq64 ny = ps[k].y.neg();
            const G2Lines::Line* l = &qs[k]->lines[cursor[k]++];
            f = f.mul_by_line(l->c, l->lambda.mul_fq(ps[k].x), ny);
            if ((BLS_X_ABS >> i) & 1) {
                l = &qs[k]->lines[cursor[k]++];
                f = f.mul_by_line(l->c, l->lambda.mul_fq(ps[k].x), ny);
            }
        }
    }
    return final_exponentiation(f) == Fq12::one();
}

inline G1AffineHost g1_affine_host(const G1J& p) {
    G1AffineHost a;
    a.inf = !p.to_affine(a.x, a.y);
    return a;
}

inline bool g2_on_curve(const Fq2& x, const Fq2& y) {
    Fq2 b = {fq_b4(), fq_b4()};  // 4 (1 + u)
    return y.sqr() == x.sqr() * x + b;
}

}  // namespace host
}  // namespace zkp
