// EXPERIMENT (not part of the product library; see tools/microbench_ll.cu and DESIGN.md section 7): a Montgomery product
// with no carry between instructions -- every 32 x 32 -> 64 product split into halves that are added to 64-bit COLUMN
// sums, the reduction on the columns, carries resolved once at the end -- meant to cut the LATENCY of a product in the
// latency-bound kernels of the MSM tail.  Same limbs, same Montgomery form, same canonical result as Fp::operator*
// (checked on the host, 400 000 operand pairs per field).  MEASURED ON B200: it is slower.  A lone warp issues roughly one
// instruction every 2-4 cycles whatever the dependencies, so what a lone warp pays for is the instruction COUNT, and the
// throughput product of ff.cuh (~330 instructions, 2028 cycles per product at 1-4 warps per SM) beats this one (~950
// instructions, 3538 cycles).  The only way to a lower-latency point addition is fewer instructions per LANE, i.e. one
// product spread over several lanes.
#pragma once
#include "../zkp_subnet_b200/csrc/ff.cuh"

namespace zkp {

// columns t[0 .. 2N) of a * b (+ c * d): t[k] = sum of the 32-bit halves of the limb products that land on limb k
template <int N>
ZKP_HD void ll_columns(uint64_t* t, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint64_t p = (uint64_t)a[i] * b[j];
            t[i + j] += (uint32_t)p;
            t[i + j + 1] += p >> 32;
        }
    }
}
// Montgomery reduction of the column sums (each < 2^40 on entry), canonical result
template <class P>
ZKP_HD Fp<P> ll_reduce(uint64_t* t) {
    constexpr int N = P::N;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t m = (uint32_t)t[i] * P::INV;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint64_t p = (uint64_t)m * P::mod(j);
            t[i + j] += (uint32_t)p;
            t[i + j + 1] += p >> 32;
        }
        t[i + 1] += t[i] >> 32;  // the low word of t[i] is zero now
    }
    // carries of the upper half, then one conditional subtraction (value < 2p for canonical operands)
    Fp<P> r;
    uint64_t c = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const uint64_t s = t[N + k] + c;
        r.v[k] = (uint32_t)s;
        c = s >> 32;
    }
    // c (the word above the top limb) is 0: the moduli leave their top bits clear and the value is < 2p
    uint32_t d[N];
    uint32_t borrow = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const uint64_t s = (uint64_t)r.v[k] - P::mod(k) - borrow;
        d[k] = (uint32_t)s;
        borrow = (uint32_t)(s >> 63);
    }
#pragma unroll
    for (int k = 0; k < N; k++) r.v[k] = borrow ? r.v[k] : d[k];
    return r;
}
template <class P>
ZKP_HD Fp<P> mul_ll(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    uint64_t t[2 * N + 1];
#pragma unroll
    for (int k = 0; k < 2 * N + 1; k++) t[k] = 0;
    ll_columns<N>(t, a.v, b.v);
    return ll_reduce<P>(t);
}
// a b + c d with one reduction (operands canonical: the sum of the two products is < 2 p^2 < p 2^(32 N), so the result is
// < 2p before the conditional subtraction exactly as for one product)
template <class P>
ZKP_HD Fp<P> mul2_ll(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    constexpr int N = P::N;
    uint64_t t[2 * N + 1];
#pragma unroll
    for (int k = 0; k < 2 * N + 1; k++) t[k] = 0;
    ll_columns<N>(t, a.v, b.v);
    ll_columns<N>(t, c.v, d.v);
    return ll_reduce<P>(t);
}

}  // namespace zkp
