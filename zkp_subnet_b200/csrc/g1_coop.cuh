// Four-lane cooperative XYZZ addition for the latency-bound stages of the MSM (slot levels, row/column finish,
// bit planes).  There a warp has a handful of additions to do, one after the other, and a lone warp needs ~17 us
// for the 13 dependent field products of one addition.  Here four consecutive lanes share one addition: every
// lane holds the SAME two operands, each computes a quarter of the products, and the intermediate values travel
// by warp shuffles (12 per field element).  Dependent products per addition: 4.5 instead of 13.
//
//   round 1   L0: u1 = x1 zz2      L1: u2 = x2 zz1      L2: s1 = y1 zzz2     L3: s2 = y2 zzz1
//   round 2   L0: pp = P^2         L1: zz12 = zz1 zz2   L2: rr = R^2         L3: zzz12 = zzz1 zzz2
//   round 3   L0: ppp = P pp       L1: q = u1 pp        L2: zz3 = zz12 pp    L3: zzz3 = zzz12 ppp (after round 3 of L0)
//   round 4   L1: x3 = rr - ppp - 2q                    L2: y3 = R (q - x3) - s1 ppp   (one fused two-product multiply)
// with P = u2 - u1, R = s2 - s1.  Exceptional cases (an operand at infinity, P = 0: doubling or cancellation) are
// detected by all four lanes and handled by the scalar formulas on every lane (each has both operands).
// All lanes of the group must call together (no divergence inside a group); `gmask` = the group's four lane bits.
#pragma once
#include "g1.cuh"

namespace zkp {

__device__ __forceinline__ Fq shfl_fq(uint32_t mask, const Fq& v, int src_lane) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = __shfl_sync(mask, v.v[i], src_lane);
    return r;
}

// per-lane operand choice: every lane then executes the SAME multiply instruction stream (a divergent
// if/else over four different products would be serialised by the warp and gain nothing)
__device__ __forceinline__ Fq sel4(int gl, const Fq& v0, const Fq& v1, const Fq& v2, const Fq& v3) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t lo = gl & 1 ? v1.v[i] : v0.v[i], hi = gl & 1 ? v3.v[i] : v2.v[i];
        r.v[i] = gl & 2 ? hi : lo;
    }
    return r;
}

// a += b; a and b identical on the four lanes of the group before, a identical on all four after
__device__ __forceinline__ void coop_add4(G1Xyzz& a, const G1Xyzz& b, uint32_t gmask, int gbase, int gl) {
    if (b.is_inf()) return;           // uniform inside the group: every lane holds the same operands
    if (a.is_inf()) { a = b; return; }
    // round 1: u1 = x1 zz2 | u2 = x2 zz1 | s1 = y1 zzz2 | s2 = y2 zzz1
    const Fq t1 = sel4(gl, a.x, b.x, a.y, b.y) * sel4(gl, b.zz, a.zz, b.zzz, a.zzz);
    const Fq partner = shfl_fq(gmask, t1, gbase + (gl ^ 1));
    // lanes 0,1 know (u1, u2); lanes 2,3 know (s1, s2): P = u2 - u1 on lanes 0,1, R = s2 - s1 on lanes 2,3
    const Fq d = (gl & 1) ? t1 - partner : partner - t1;
    const Fq P = shfl_fq(gmask, d, gbase);
    const Fq R = shfl_fq(gmask, d, gbase + 2);
    if (P.is_zero()) {  // doubling or P + (-P): rare, every lane runs the scalar formulas on its own copy
        a.add(b);
        return;
    }
    // round 2: pp = P^2 | zz12 = zz1 zz2 | rr = R^2 | zzz12 = zzz1 zzz2
    const Fq t2 = sel4(gl, P, a.zz, R, a.zzz) * sel4(gl, P, b.zz, R, b.zzz);
    const Fq pp = shfl_fq(gmask, t2, gbase);
    const Fq zz12 = shfl_fq(gmask, t2, gbase + 1);
    const Fq rr = shfl_fq(gmask, t2, gbase + 2);
    // round 3: ppp = P pp | q = u1 pp | zz3 = zz12 pp | (lane 3: a product nobody reads)
    const Fq u1 = gl == 1 ? partner : t1;  // valid on lanes 0 and 1
    const Fq t3 = sel4(gl, P, u1, zz12, zz12) * pp;
    const Fq ppp = shfl_fq(gmask, t3, gbase);
    const Fq q = shfl_fq(gmask, t3, gbase + 1);
    const Fq zz3 = shfl_fq(gmask, t3, gbase + 2);
    // round 4: every lane forms x3 (three subtractions: cheaper than another shuffle round); lane 2 computes
    // y3 = R (q - x3) - s1 ppp, lane 3 zzz3 = zzz12 ppp, both as one fused two-product multiply (lanes 0, 1 idle along)
    const Fq x3 = rr - ppp - q.dbl();
    const Fq s1neg = t1.neg();  // meaningful on lane 2 (t1 = s1)
    const Fq zero = Fq::zero();
    const Fq t4 = Fq::mul2(sel4(gl, zero, zero, R, t2), sel4(gl, zero, zero, q - x3, ppp), sel4(gl, zero, zero, ppp, zero),
                           sel4(gl, zero, zero, s1neg, zero));
    a.x = x3;
    a.y = shfl_fq(gmask, t4, gbase + 2);
    a.zz = zz3;
    a.zzz = shfl_fq(gmask, t4, gbase + 3);
}

}  // namespace zkp
