// BLS12-381 G1 (y^2 = x^3 + 4 over Fq) point arithmetic for the MSM kernels.
//
// Buckets live in extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// mixed add 8M+2S, full add 12M+2S, doubling of an affine point 3M+3S... (EFD madd-2008-s,
// add-2008-s, mdbl-2008-s-1).  Every exceptional case (operand at infinity, P + P, P + (-P)) is
// handled exactly -- the contract with the reference is bit-exact output, not "overwhelmingly
// likely correct".  Infinity is ZZ == 0; an all-zero XYZZ record (fresh cudaMemset) is infinity.
// Affine SRS points use (0, 0) as the infinity marker ((0,0) is not on the curve).
//
// __host__ __device__ like ff.cuh, so tests/host exercises these formulas on the CPU.
#pragma once
#include "ff.cuh"

namespace zkp {

struct G1Affine {
    Fq x, y;  // Montgomery form
    ZKP_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

struct G1Xyzz {
    Fq x, y, zz, zzz;

    static ZKP_HD G1Xyzz infinity() {
        G1Xyzz r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    ZKP_HD bool is_inf() const { return zz.is_zero(); }

    // from an affine point, optionally negated
    static ZKP_HD G1Xyzz from_affine(const G1Affine& p, uint32_t negate) {
        G1Xyzz r;
        if (p.is_inf()) return infinity();
        r.x = p.x;
        r.y = p.y.cneg(negate);
        r.zz = Fq::one();
        r.zzz = Fq::one();
        return r;
    }

    // 2 * (affine p)   (mdbl-2008-s-1: 3M + 3S... here 2S + 3M + small)
    static ZKP_HD G1Xyzz dbl_affine(const Fq& px, const Fq& py) {
        G1Xyzz r;
        Fq u = py.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = px * v;
        Fq xx = px.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = Fq::mul2(m, s - r.x, w, py.neg());  // m (s - x3) - w y with one reduction
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // 2 * this   (dbl-2008-s-1, a = 0)
    ZKP_HD G1Xyzz dbl() const {
        if (is_inf()) return *this;
        G1Xyzz r;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq xx = x.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = Fq::mul2(m, s - r.x, w, y.neg());
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += (px, +-py) affine, not infinity   (madd-2008-s: 8M + 2S)
    ZKP_HD void madd(const Fq& px, const Fq& py_in, uint32_t negate) {
        Fq py = py_in.cneg(negate);
        if (is_inf()) {
            x = px; y = py; zz = Fq::one(); zzz = Fq::one();
            return;
        }
        Fq u2 = px * zz;
        Fq s2 = py * zzz;
        Fq p = u2 - x;
        Fq r = s2 - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(px, py);  // P + P
            else *this = infinity();                      // P + (-P)
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = x * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = Fq::mul2(r, q - x3, ppp, y.neg());  // r (q - x3) - y ppp: two products, one Montgomery reduction
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }
    ZKP_HD void madd(const G1Affine& p, uint32_t negate) {
        if (p.is_inf()) return;
        madd(p.x, p.y, negate);
    }

    // The same addition with LAZY reductions, for the bucket-accumulation kernel (msm.cuh k_accumulate<level0>), where
    // one accumulator takes a run of additions and only the end of the run is stored.  Coordinates are plain
    // 384-bit integers congruent to the true values: X < 2p, Y < 1.42p, ZZ < 1.26p, ZZZ < 1.2p (a box that the
    // step maps into itself: tools/lazy_bounds.py checks it and every overflow condition with exact rationals).
    //  * the seven single products skip their final conditional subtraction (result < ab/2^384 + p);
    //  * the differences that feed products are a - b + 2p, unconditionally;
    //  * X3 = R^2 - (PPP + 2Q) is the one value corrected on both sides (Fq::sub_fix: + k p with k read off the
    //    top limb), Y3 keeps the single conditional subtraction of the fused product.
    // About 170 add/sub/select instructions per addition instead of 460; the multiply count is unchanged.
    // (px, py) canonical and not infinity.  kp = limbs of k p for k < 8 (shared memory on the device).
    // normalize() returns canonical coordinates; infinity stays the all-zero record throughout.
    ZKP_HD void madd_lazy(const Fq& px, const Fq& py, const uint32_t* kp) {
        const int st = madd_lazy_core(px, py, kp);
        if (st) madd_lazy_rare(st, px, py);
    }
    // The arithmetic of madd_lazy with the exceptional cases only DETECTED: returns 0 when the sum is in *this,
    // 1 when the operands were equal (the sum is 2 (px, py)), 2 when they were opposite (the sum is infinity); in
    // both cases *this is garbage and madd_lazy_rare finishes the job.  Keeping the test out of the way lets the
    // products start before it resolves and lets the caller reload (px, py) in the rare case instead of holding
    // them in registers through the whole addition.
    ZKP_HD int madd_lazy_core(const Fq& px, const Fq& py, const uint32_t* kp) {
        if (is_inf()) {
            x = px; y = py; zz = Fq::one(); zzz = Fq::one();
            return 0;
        }
        Fq u2 = Fq::mul_lazy(px, zz);
        Fq s2 = Fq::mul_lazy(py, zzz);
        Fq p = Fq::sub_p2(u2, x);   // in (0, 3.13p)
        Fq r = Fq::sub_p2(s2, y);   // in (0, 3.13p)
        const bool p_zero = p.is_zero_mod_p_lt4();
        Fq pp = p.sqr_lazy();
        Fq ppp = Fq::mul_lazy(p, pp);
        Fq q = Fq::mul_lazy(x, pp);
        Fq s = Fq::add_raw(Fq::add_raw(ppp, q), q);  // PPP + 2Q < 4.45p
        Fq x3 = Fq::sub_fix(r.sqr_lazy(), s, kp);
        y = Fq::mul2(r, Fq::sub_p2(q, x3), ppp, y.rsub_p2());  // < 2.42p before, < 1.42p after its conditional subtraction
        x = x3;
        zz = Fq::mul_lazy(zz, pp);
        zzz = Fq::mul_lazy(zzz, ppp);
        if (p_zero) return r.is_zero_mod_p_lt4() ? 1 : 2;
        return 0;
    }
    ZKP_HD void madd_lazy_rare(int st, const Fq& px, const Fq& py) {
        if (st == 1) *this = dbl_affine(px, py);  // P + P
        else *this = infinity();                  // P + (-P)
    }
    ZKP_HD void normalize() {
        x = x.reduce_once(); y = y.reduce_once(); zz = zz.reduce_once(); zzz = zzz.reduce_once();
    }

    // this += o   (add-2008-s: 12M + 2S)
    ZKP_HD void add(const G1Xyzz& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        Fq u1 = x * o.zz;
        Fq u2 = o.x * zz;
        Fq s1 = y * o.zzz;
        Fq s2 = o.y * zzz;
        Fq p = u2 - u1;
        Fq r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = infinity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = u1 * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = Fq::mul2(r, q - x3, ppp, s1.neg());
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }
};

}  // namespace zkp
