// SRS tooling without a trapdoor (replaces the reference's `prover setup --generate-precompute`, tests/conftest.py:50-65,
// and the loaders behind --setup_path / --precompute_path / --uncompressed, Makefile:30-48,64-74):
//   * G1 points in the ZCash COMPRESSED wire format decoded / encoded on the device (a square root per point: 2^24 of
//     them are minutes on a host core and a fraction of a second here);
//   * the group inverse FFT that turns a monomial SRS [tau_x^j tau_y^i]_1 into the Pianist Lagrange rows
//     U[i][j] = [R_i(tau_y) L_j(tau_x)]_1 -- the only way to obtain them from a ceremony SRS, whose tau nobody knows:
//        L_j(tau) = (1/n) sum_k w^(-jk) tau^k,
//     i.e. an inverse DFT over the exponent index with group elements as values, first along Y (size M, every column),
//     then along X (size n, every row).  [R_i(tau_y)]_1, the row scale points, are column 0 after the Y pass.
// Radix-2 decimation in frequency on XYZZ points: (a, b) -> (a + b, [w^-k](a - b)); one 255-bit scalar
// multiplication per butterfly (stages whose twiddle is 1 skip it), bit-reversal + [1/n] scaling in the last pass.
// Bound by the multiply pipe like everything else here: (n/2)(log n - 1) + n scalar multiplications of ~380 point
// operations each per transform.  A one-off job (measured 2.3-3.6 us per point: 40-60 s for the 2^24-point mainnet SRS), not on
// the request path.
#pragma once
#include "g1.cuh"
#include "kzg.cuh"
#include "srs.cuh"

namespace zkp {

__device__ __forceinline__ bool fq_lex_largest(const Fq& y_canon) {
    // y > (p - 1) / 2  <=>  y > p >> 1 (p odd)
    for (int i = 11; i >= 0; i--) {
        const uint32_t h = (FqParams::mod(i) >> 1) | (i < 11 ? FqParams::mod(i + 1) << 31 : 0u);
        if (y_canon.v[i] > h) return true;
        if (y_canon.v[i] < h) return false;
    }
    return false;
}

// 48-byte compressed points -> Montgomery affine; *bad |= 1 on a malformed encoding or an x with no point
__global__ void __launch_bounds__(128)
k_points_from_be48(const uint8_t* __restrict__ in, size_t n, G1Affine* __restrict__ out, uint32_t* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in + i * 48);
    Fq x;
#pragma unroll
    for (int k = 0; k < 12; k++) x.v[11 - k] = __byte_perm(w[k], 0, 0x0123);
    const uint32_t flags = x.v[11] >> 29;
    x.v[11] &= 0x1fffffffu;
    G1Affine p;
    p.x = Fq::zero();
    p.y = Fq::zero();
    if (!(flags & 4)) {
        atomicOr(bad, 1u);
    } else if (flags & 2) {
        if ((flags & 1) || !x.is_zero()) atomicOr(bad, 1u);
    } else if (!fq_lt_mod(x)) {
        atomicOr(bad, 1u);
    } else {
        const Fq xm = x.to_mont();
        const Fq y2 = xm.sqr() * xm + fq_b4();
        // p = 3 mod 4: sqrt = y2^((p + 1) / 4)
        uint32_t e[12], t[12];
#pragma unroll
        for (int k = 0; k < 12; k++) t[k] = FqParams::mod(k);
        t[0] += 1;  // p ends in ...aaab: no carry
#pragma unroll
        for (int k = 0; k < 12; k++) e[k] = (t[k] >> 2) | (k < 11 ? t[k + 1] << 30 : 0u);
        Fq y = y2.pow_limbs<12>(e);
        if (y.sqr() != y2) {
            atomicOr(bad, 1u);
        } else {
            if (fq_lex_largest(y.from_mont()) != ((flags & 1) != 0)) y = y.neg();
            p.x = xm;
            p.y = y;
        }
    }
    out[i] = p;
}

__global__ void k_points_to_be48(const G1Affine* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = in[i];
    uint32_t* w = reinterpret_cast<uint32_t*>(out + i * 48);
    if (p.is_inf()) {
        for (int k = 0; k < 12; k++) w[k] = 0;
        w[0] = 0x000000c0u;  // byte 0 = 0xc0
        return;
    }
    Fq x = p.x.from_mont(), y = p.y.from_mont();
    x.v[11] |= 0x80000000u | (fq_lex_largest(y) ? 0x20000000u : 0u);
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = __byte_perm(x.v[11 - k], 0, 0x0123);
}

// ---- group FFT ------------------------------------------------------------------------------------------------------
__global__ void k_gfft_load(const G1Affine* __restrict__ in, size_t n, G1Xyzz* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = G1Xyzz::from_affine(in[i], 0);
}

__global__ void k_gather_stride(const G1Xyzz* __restrict__ in, size_t stride, size_t count, G1Xyzz* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = in[i * stride];
}

// [s] p for a canonical little-endian scalar (MSB-first double-and-add; exact for every input)
__device__ __noinline__ G1Xyzz xyzz_scalar_mul(const G1Xyzz& p, const Fr& s) {
    G1Xyzz acc = G1Xyzz::infinity();
    int i = 254;
    while (i >= 0 && !((s.v[i >> 5] >> (i & 31)) & 1)) i--;
    for (; i >= 0; i--) {
        acc = acc.dbl();
        if ((s.v[i >> 5] >> (i & 31)) & 1) acc.add(p);
    }
    return acc;
}

// One decimation-in-frequency stage of `batch` inverse transforms of length len = 2^log_len: element e of transform t
// lives at data[t * tstride + e * estride].  Butterfly (j, j + half) inside blocks of 2 * half:
//     a' = a + b,   b' = [w_len^(-(j mod half) * (len / (2 half)))] (a - b)
// wt[k] = w_len^(2^k).  One thread per butterfly.
__global__ void __launch_bounds__(128)
k_gfft_stage(G1Xyzz* __restrict__ data, uint32_t log_len, uint32_t log_half, size_t batch, size_t tstride, size_t estride,
             const Fr* __restrict__ wt) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = (size_t)1 << (log_len - 1);
    if (tid >= batch * per) return;
    const size_t t = tid >> (log_len - 1), b = tid & (per - 1);
    const size_t half = (size_t)1 << log_half;
    const size_t j = b & (half - 1), blk = b >> log_half;
    const size_t i0 = (blk << (log_half + 1)) + j, i1 = i0 + half;
    G1Xyzz* pa = data + t * tstride + i0 * estride;
    G1Xyzz* pb = data + t * tstride + i1 * estride;
    G1Xyzz a = *pa, bb = *pb;
    G1Xyzz sum = a;
    sum.add(bb);
    G1Xyzz nb = bb;
    nb.y = nb.y.neg();  // -(X, Y, ZZ, ZZZ) = (X, -Y, ZZ, ZZZ); infinity (ZZ = 0) stays infinity
    G1Xyzz diff = a;
    diff.add(nb);
    const uint64_t len = 1ull << log_len;
    const uint64_t k = (uint64_t)j << (log_len - 1 - log_half);  // exponent of w^-1
    if (k) {
        const Fr tw = pow_from_table(wt, len - k).from_mont();   // w^(len - k) = w^(-k)
        diff = xyzz_scalar_mul(diff, tw);
    }
    *pa = sum;
    *pb = diff;
}

// last pass: out[t][bitrev(e)] = [scale] in[t][e]   (scale = 1 / len, canonical)
__global__ void __launch_bounds__(128)
k_gfft_finish(const G1Xyzz* __restrict__ in, G1Xyzz* __restrict__ out, uint32_t log_len, size_t batch, size_t tstride, size_t estride,
              Fr scale_canon) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t len = (size_t)1 << log_len;
    if (tid >= batch * len) return;
    const size_t t = tid >> log_len, e = tid & (len - 1);
    const size_t r = log_len ? (size_t)(__brevll((unsigned long long)e) >> (64 - log_len)) : 0;
    G1Xyzz p = in[t * tstride + e * estride];
    out[t * tstride + r * estride] = xyzz_scalar_mul(p, scale_canon);
}

}  // namespace zkp
