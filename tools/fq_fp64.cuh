// EXPERIMENT (not on the product path): an Fq Montgomery product on the FP64 pipe.
//
// B200 issues 60-63 DFMA/clk/SM -- twice the IMAD.WIDE rate -- and a 1:1 mix of DFMA and IMAD.WIDE chains
// sustains 23 pairs/clk/SM, i.e. the two pipes overlap by ~80% (tools/microbench.cu).  The integer product of
// ff.cuh saturates the IMAD pipe and leaves the FP64 pipe idle, so a second, floating-point formulation of the
// same Montgomery product could run beside it.  This header is that formulation, written to measure it:
//
//   * representation: 8 limbs of 48 bits, each held exactly in a double (8 x 48 = 384 bits, same R = 2^384 as the
//     integer code, so the two representations are re-limbings of the same Montgomery value);
//   * one limb product x*y < 2^96 is split exactly into hi = floor(xy / 2^48) 2^48 and lo = xy - hi with two
//     round-toward-zero FMAs against an accumulator biased by 2^100 (ulp 2^48 there): H' = fma_rz(x, y, H) adds
//     the truncated product to the column's high accumulator, lo = fma_rz(x, y, -(H' - H)) is exact;
//   * a column receives at most 16 low parts (< 2^48 each) and the high accumulator at most 16 truncated
//     products (< 2^96 each, biased at 2^100): every partial sum is an integer below 2^53 resp. inside
//     [2^100, 2^101) and therefore exact;
//   * Montgomery reduction limb by limb: m = (column mod 2^48) * (-p^-1) mod 2^48, add m * p.
// Cost: 128 limb products x 4 FP64 operations + ~150 for the reduction bookkeeping and normalisation.
#pragma once
#include "../zkp_subnet_b200/csrc/ff.cuh"

namespace zkp {
namespace fp64 {

struct FqD { double v[8]; };

__device__ __forceinline__ double p48(int i) {
    constexpr double t[8] = {281474976688811.0, 194974335351294.0, 270634993844222.0, 113459389855408.0,
                             83034393350847.0, 73992301405303.0, 253550359455670.0, 28591897852287.0};
    return t[i];
}
constexpr double PPRIME48 = 281462091612157.0;  // -p^-1 mod 2^48
constexpr double BIAS = 1267650600228229401496703205376.0;  // 2^100
constexpr double TWO48 = 281474976710656.0;
constexpr double INV48 = 1.0 / 281474976710656.0;
constexpr double TWO52 = 4503599627370496.0;

// 12 x 32-bit limbs -> 8 x 48-bit limbs in doubles (and back); values < 2^384
__device__ __forceinline__ FqD to_fp(const Fq& a) {
    FqD r;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const int w = 3 * (k / 2);  // three 32-bit words hold two 48-bit limbs
        uint64_t lo = (uint64_t)a.v[w] | ((uint64_t)(a.v[w + 1] & 0xffffu) << 32);
        uint64_t hi = (uint64_t)(a.v[w + 1] >> 16) | ((uint64_t)a.v[w + 2] << 16);
        r.v[k] = __longlong_as_double(0x4330000000000000ll | (long long)lo) - TWO52;
        r.v[k + 1] = __longlong_as_double(0x4330000000000000ll | (long long)hi) - TWO52;
    }
    return r;
}
__device__ __forceinline__ Fq from_fp(const FqD& a) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const int w = 3 * (k / 2);
        uint64_t lo = (uint64_t)__double_as_longlong(a.v[k] + TWO52) & 0xffffffffffffull;
        uint64_t hi = (uint64_t)__double_as_longlong(a.v[k + 1] + TWO52) & 0xffffffffffffull;
        r.v[w] = (uint32_t)lo;
        r.v[w + 1] = (uint32_t)(lo >> 32) | ((uint32_t)hi << 16);
        r.v[w + 2] = (uint32_t)(hi >> 16);
    }
    return r;
}

// accumulate x*y into column accumulators (H biased high part, L low part)
__device__ __forceinline__ void mac(double x, double y, double& H, double& L) {
    const double hn = __fma_rz(x, y, H);
    const double hs = hn - H;                 // floor(xy / 2^48) 2^48, exact
    L += __fma_rz(x, y, -hs);                 // xy - hs in [0, 2^48), exact
    H = hn;
}
__device__ __forceinline__ double floor48(double v) {  // floor(v / 2^48) for 0 <= v < 2^100
    return __fma_rz(v, INV48, TWO52) - TWO52;
}

// Montgomery product a b / 2^384 mod p; inputs in [0, p) (limbs < 2^48), output canonical
__device__ __forceinline__ FqD mul(const FqD& a, const FqD& b) {
    double H[15], L[16];
#pragma unroll
    for (int k = 0; k < 15; k++) H[k] = BIAS;
#pragma unroll
    for (int k = 0; k < 16; k++) L[k] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) mac(a.v[i], b.v[j], H[i + j], L[i + j]);
    double carry = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double v = L[i] + carry;
        if (i) v = __fma_rn(H[i - 1] - BIAS, INV48, v);
        const double vl = __fma_rn(floor48(v), -TWO48, v);        // v mod 2^48
        const double pn = __fma_rz(vl, PPRIME48, BIAS);
        const double m = __fma_rz(vl, PPRIME48, -(pn - BIAS));    // (vl * p') mod 2^48
#pragma unroll
        for (int j = 0; j < 8; j++) mac(m, p48(j), H[i + j], L[i + j]);
        double v2 = L[i] + carry;
        if (i) v2 = __fma_rn(H[i - 1] - BIAS, INV48, v2);
        carry = v2 * INV48;                                       // column i is now 0 mod 2^48
    }
    // columns 8..15 -> 48-bit limbs
    FqD r;
#pragma unroll
    for (int k = 8; k < 16; k++) {
        double v = __fma_rn(H[k - 1] - BIAS, INV48, L[k] + carry);
        carry = floor48(v);
        r.v[k - 8] = __fma_rn(carry, -TWO48, v);
    }
    // result < 2p: subtract p if >= p
    double d[8], borrow = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double t = r.v[k] - p48(k) - borrow;
        borrow = t < 0.0 ? 1.0 : 0.0;
        d[k] = t < 0.0 ? t + TWO48 : t;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = borrow != 0.0 ? r.v[k] : d[k];
    return r;
}

}  // namespace fp64
}  // namespace zkp
