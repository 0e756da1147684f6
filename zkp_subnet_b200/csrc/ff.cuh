// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a (Fq: 12 limbs, Fr: 8 limbs).
//
// Multiplication is an interleaved (CIOS-order) Montgomery product that keeps TWO accumulators,
// one aligned on even limb positions and one on odd positions, so that every 32x32->64 product
// a[j]*b_i lands on an aligned (lo,hi) register pair: each mad.lo.cc/madc.hi.cc pair becomes one
// IMAD.WIDE.U32 in SASS and the carry chains of the two accumulators are independent (ILP 2).
// Work per product: 2*N^2 + N wide multiply-accumulates (300 for Fq, 136 for Fr) -- the figure the
// roofline accounting in DESIGN.md uses.
//
// Values are kept fully reduced in [0, p).  All functions are __host__ __device__; the host
// instantiation runs on the emulated carry flag of ptx_chain.cuh and exists for CPU unit tests.
#pragma once
#include "field_params.h"
#include "ptx_chain.cuh"
#include "mont_chains.cuh"

namespace zkp {

// limb i of k p (Fq), k < 8: the correction table of Fp::sub_fix and the constants of is_zero_mod_p_lt4
ZKP_PARAM_HD constexpr uint32_t fq_kp_limb(uint32_t k, int i) {
    uint64_t carry = 0;
    uint32_t out = 0;
    for (int j = 0; j <= i; j++) {
        uint64_t t = (uint64_t)FqParams::mod(j) * k + carry;
        out = (uint32_t)t;
        carry = t >> 32;
    }
    return out;
}
constexpr int FQ_KP_ROWS = 8;

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    // ---------------------------------------------------------------- constants
    static ZKP_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    static ZKP_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one(i);
        return r;
    }
    static ZKP_HD Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
        return r;
    }

    // ---------------------------------------------------------------- predicates
    ZKP_HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i];
        return acc == 0;
    }
    friend ZKP_HD bool operator==(const Fp& a, const Fp& b) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= a.v[i] ^ b.v[i];
        return acc == 0;
    }
    friend ZKP_HD bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }

    // ---------------------------------------------------------------- add / sub / neg
    // (carry chains live in the generated mont_chains.cuh, one asm statement per chain)
    friend ZKP_HD Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        if constexpr (N == 12) chains::fq_add(r.v, a.v, b.v);
        else chains::fr_add(r.v, a.v, b.v);
        return r;
    }
    friend ZKP_HD Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        if constexpr (N == 12) chains::fq_sub(r.v, a.v, b.v);
        else chains::fr_sub(r.v, a.v, b.v);
        return r;
    }
    ZKP_HD Fp neg() const { return zero() - *this; }  // 0 - a = p - a, and 0 for a == 0
    // conditional negate (flag != 0 -> -a)
    ZKP_HD Fp cneg(uint32_t flag) const {
        Fp n = neg();
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = flag ? n.v[i] : v[i];
        return r;
    }
    ZKP_HD Fp dbl() const { return *this + *this; }

    // ---------------------------------------------------------------- Montgomery product
    // Row i adds a*b_i and m_i*p to the pair of accumulators and divides by 2^32; the accumulator
    // aligned on even limb positions and the one aligned on odd positions swap roles every row.
    friend ZKP_HD Fp operator*(const Fp& a, const Fp& b) { return mul_t<true>(a, b); }
    // REDUCE = false leaves out the final conditional subtraction: the result is (a b + m p) / 2^(32 N) < a b / 2^(32 N) + p
    // (Fq only; operand ranges in tools/lazy_bounds.py)
    template <bool REDUCE>
    static ZKP_HD Fp mul_t(const Fp& a, const Fp& b) {
        uint32_t even[N], odd[N];
        Fp r;
        if constexpr (N == 12) {
#if defined(__CUDA_ARCH__) && defined(ZKP_MONT_ROWS_PER_ITER)
            // Uniform rolled form: ALL rows are generic rows on zero-initialised accumulators, ZKP_MONT_ROWS_PER_ITER
            // (even, divides 12) rows per trip.  Every trip ends by physically moving the loop-carried
            // accumulator and b registers (~26 MOV / IMAD.MOV per trip, part of them on the multiply pipe --
            // ncu, profiles/), so fewer, longer trips trade code size for moves and loop back-edges.
            constexpr int RPI = ZKP_MONT_ROWS_PER_ITER;
            static_assert(N % RPI == 0 && RPI % 2 == 0, "rows per iteration must be even and divide N");
#pragma unroll
            for (int k = 0; k < N; k++) even[k] = odd[k] = 0;
            uint32_t bb[N];
#pragma unroll
            for (int k = 0; k < N; k++) bb[k] = b.v[k];
#pragma unroll 1
            for (int i = 0; i < N; i += RPI) {
#pragma unroll
                for (int k = 0; k < RPI; k += 2) {
                    chains::fq_row(even, odd, a.v, bb[k]);
                    chains::fq_row(odd, even, a.v, bb[k + 1]);
                }
#pragma unroll
                for (int k = 0; k + RPI < N; k++) bb[k] = bb[k + RPI];
            }
#elif defined(__CUDA_ARCH__) && !defined(ZKP_UNROLL_MONT_ROWS)
            // All 12 rows as a ROLLED loop of row pairs (a generic row on zero accumulators is the first
            // row): the bucket-accumulation kernel inlines ten of these products and was instruction-fetch
            // bound when fully unrolled (ncu: stall_no_instruction 1.7 per issue, profiles/).  b's limbs
            // rotate down two places per trip; the moves issue on slots the 4-cycle IMAD.WIDE pipe leaves idle.
            // (Measured on B200: rolling ALL rows, first pair included, costs 14% of raw product throughput
            // -- no overlap across the loop back-edge -- so the first pair stays outside and the loop body is
            // unrolled by two pairs.)
            chains::fq_row_first(even, odd, a.v, b.v[0]);
            chains::fq_row(odd, even, a.v, b.v[1]);
            uint32_t bb[N - 2];
#pragma unroll
            for (int k = 0; k < N - 2; k++) bb[k] = b.v[k + 2];
#pragma unroll 1
            for (int i = 2; i < N; i += 2) {
                chains::fq_row(even, odd, a.v, bb[0]);
                chains::fq_row(odd, even, a.v, bb[1]);
#pragma unroll
                for (int k = 0; k < N - 4; k++) bb[k] = bb[k + 2];
            }
#else
            chains::fq_row_first(even, odd, a.v, b.v[0]);
            chains::fq_row(odd, even, a.v, b.v[1]);
#pragma unroll
            for (int i = 2; i < N; i += 2) {
                chains::fq_row(even, odd, a.v, b.v[i]);
                chains::fq_row(odd, even, a.v, b.v[i + 1]);
            }
#endif
            chains::fq_merge(r.v, odd, even);
            if (REDUCE) chains::fq_reduce_once(r.v, 0);
        } else {
            chains::fr_row_first(even, odd, a.v, b.v[0]);
            chains::fr_row(odd, even, a.v, b.v[1]);
#pragma unroll
            for (int i = 2; i < N; i += 2) {
                chains::fr_row(even, odd, a.v, b.v[i]);
                chains::fr_row(odd, even, a.v, b.v[i + 1]);
            }
            chains::fr_merge(r.v, odd, even);
            chains::fr_reduce_once(r.v, 0);
        }
        return r;
    }
    // Dedicated squaring (Fq): row i multiplies a_i with (a_i, 2a_{i+1}, .., 2a_{N-1}) only -- N(N+1)/2 = 78
    // products instead of 144, plus the same 156 of the reduction: 234 wide multiply-accumulates instead of 300.
    // 2a < 2^382 fits the 12 limbs.  Fully unrolled (every row has its own shape).
    ZKP_HD Fp sqr() const { return sqr_t<true>(); }
    template <bool REDUCE>
    ZKP_HD Fp sqr_t() const {
        if constexpr (N == 12) {
            uint32_t even[N], odd[N], d[N];
#pragma unroll
            for (int k = 0; k < N; k++) even[k] = odd[k] = 0;
            d[0] = v[0] << 1;
#pragma unroll
            for (int k = 1; k < N; k++) d[k] = (v[k] << 1) | (v[k - 1] >> 31);
            chains::fq_sqr_row_0(even, odd, v, d);
            chains::fq_sqr_row_1(odd, even, v, d);
            chains::fq_sqr_row_2(even, odd, v, d);
            chains::fq_sqr_row_3(odd, even, v, d);
            chains::fq_sqr_row_4(even, odd, v, d);
            chains::fq_sqr_row_5(odd, even, v, d);
            chains::fq_sqr_row_6(even, odd, v, d);
            chains::fq_sqr_row_7(odd, even, v, d);
            chains::fq_sqr_row_8(even, odd, v, d);
            chains::fq_sqr_row_9(odd, even, v, d);
            chains::fq_sqr_row_10(even, odd, v, d);
            chains::fq_sqr_row_11(odd, even, v, d);
            Fp r;
            chains::fq_merge(r.v, odd, even);
            if (REDUCE) chains::fq_reduce_once(r.v, 0);
            return r;
        } else {
            return *this * *this;
        }
    }
    // a*b + c*d with ONE Montgomery reduction (Fq): every row adds a*b_i and c*d_i before its reduction step,
    // 2*144 + 156 = 444 wide multiply-accumulates instead of 600 for two products, and no separate addition.
    // Inputs in [0, p); the accumulator stays below a + c + p < 3p and the result below 1.2p before the final
    // conditional subtraction.
    static ZKP_HD Fp mul2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return mul2_t<true>(a, b, c, d); }
    template <bool REDUCE>
    static ZKP_HD Fp mul2_t(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        if constexpr (N == 12) {
            uint32_t even[N], odd[N];
#pragma unroll
            for (int k = 0; k < N; k++) even[k] = odd[k] = 0;
#if defined(__CUDA_ARCH__)
            uint32_t bb[N], dd[N];
#pragma unroll
            for (int k = 0; k < N; k++) { bb[k] = b.v[k]; dd[k] = d.v[k]; }
#pragma unroll 1
            for (int i = 0; i < N; i += 2) {
                chains::fq_row2(even, odd, a.v, bb[0], c.v, dd[0]);
                chains::fq_row2(odd, even, a.v, bb[1], c.v, dd[1]);
#pragma unroll
                for (int k = 0; k < N - 2; k++) { bb[k] = bb[k + 2]; dd[k] = dd[k + 2]; }
            }
#else
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                chains::fq_row2(even, odd, a.v, b.v[i], c.v, d.v[i]);
                chains::fq_row2(odd, even, a.v, b.v[i + 1], c.v, d.v[i + 1]);
            }
#endif
            Fp r;
            chains::fq_merge(r.v, odd, even);
            if (REDUCE) chains::fq_reduce_once(r.v, 0);
            return r;
        } else {
            return a * b + c * d;
        }
    }

    // ---------------------------------------------------------------- lazily reduced helpers (Fq)
    // Plain 384-bit integers, not residues in [0, p): used by G1Xyzz::madd_lazy only, which keeps every value below
    // 2^384 (tools/lazy_bounds.py) and hands canonical coordinates back through reduce_once().
    static ZKP_HD Fp mul_lazy(const Fp& a, const Fp& b) { return mul_t<false>(a, b); }
    ZKP_HD Fp sqr_lazy() const { return sqr_t<false>(); }
    static ZKP_HD Fp add_raw(const Fp& a, const Fp& b) {
        static_assert(N == 12, "Fq only");
        Fp r;
        chains::fq_add_raw(r.v, a.v, b.v);
        return r;
    }
    // a - b + 2p (b < 2p)
    static ZKP_HD Fp sub_p2(const Fp& a, const Fp& b) {
        static_assert(N == 12, "Fq only");
        Fp r;
        chains::fq_sub_p2(r.v, a.v, b.v);
        return r;
    }
    // 2p - a (a <= 2p)
    ZKP_HD Fp rsub_p2() const {
        static_assert(N == 12, "Fq only");
        Fp r;
        chains::fq_rsub_p2(r.v, v);
        return r;
    }
    // a - b brought back into [0, max(a, p (1 + 2^-24))): when the difference is negative, k p is added with
    // k = (2^32 - top limb) / (top limb of p) + 1, the smallest multiple that is certain to make it positive.
    // kp = table of the limbs of k p, k < 8 (fq_kp_limb); |a - b| < 6 p.
    static ZKP_HD Fp sub_fix(const Fp& a, const Fp& b, const uint32_t* kp) {
        static_assert(N == 12, "Fq only");
        Fp r;
        uint32_t bw;
        chains::fq_sub_borrow(r.v, a.v, b.v, &bw);
        const uint32_t k = bw ? (0u - r.v[N - 1]) / P::mod(N - 1) + 1u : 0u;
        Fp c;
#pragma unroll
        for (int i = 0; i < N; i++) c.v[i] = kp[k * N + i];
        return add_raw(r, c);
    }
    // one conditional subtraction of p: [0, 2p) -> [0, p)
    ZKP_HD Fp reduce_once() const {
        static_assert(N == 12, "Fq only");
        Fp r = *this;
        chains::fq_reduce_once(r.v, 0);
        return r;
    }
    // value in (0, 4p): is it a multiple of p?  (limb 0 of p, 2p, 3p first: almost never equal)
    ZKP_HD bool is_zero_mod_p_lt4() const {
        bool hit = false;
#pragma unroll
        for (uint32_t k = 1; k <= 3; k++) {
            if (v[0] == fq_kp_limb(k, 0)) {
                uint32_t acc = 0;
#pragma unroll
                for (int i = 0; i < N; i++) acc |= v[i] ^ fq_kp_limb(k, i);
                hit |= acc == 0;
            }
        }
        return hit;
    }

    // ---------------------------------------------------------------- Montgomery domain
    ZKP_HD Fp to_mont() const { return *this * r2(); }
    ZKP_HD Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }

    // x^e for a little-endian limb exponent (host/test helper; the device paths that need an
    // inverse use batch inversion with one Fermat exponentiation per thread)
    template <int EN>
    ZKP_HD Fp pow_limbs(const uint32_t* e) const {
        Fp acc = one();
        for (int i = EN * 32 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * *this;
        }
        return acc;
    }
    ZKP_HD Fp inverse() const {
        uint32_t e[N];
        // p - 2
        e[0] = ptx::sub_cc(P::mod(0), 2);
#pragma unroll
        for (int i = 1; i < N - 1; i++) e[i] = ptx::subc_cc(P::mod(i), 0);
        e[N - 1] = ptx::subc(P::mod(N - 1), 0);
        return pow_limbs<N>(e);
    }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace zkp
