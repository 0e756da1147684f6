// SRS kernels: import/export of G1 points in the ZCash 96-byte uncompressed wire format and on-device
// generation of Pianist Lagrange rows U[i][j] = [R_i(tau_y) L_j(tau_x)]_1 from a public test trapdoor
// (replaces `prover setup --generate-setup --generate-precompute`, reference tests/conftest.py:50-65).
#pragma once
#include "g1.cuh"
#include "kzg.cuh"

namespace zkp {

__device__ __forceinline__ bool fq_lt_mod(const Fq& a) {
    // a < p ?  (canonical limbs)
    for (int i = 11; i >= 0; i--) {
        uint32_t m = FqParams::mod(i);
        if (a.v[i] < m) return true;
        if (a.v[i] > m) return false;
    }
    return false;
}

__device__ __forceinline__ Fq fq_b4() {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = FqParams::b4(i);
    return r;
}

// 96 big-endian bytes (x || y, flag bits in the top of byte 0) -> Montgomery affine; *bad |= 1 on
// malformed or off-curve input
__global__ void k_points_from_be96(const uint8_t* __restrict__ in, size_t n, G1Affine* __restrict__ out,
                                   uint32_t* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in + i * 96);
    Fq x, y;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        x.v[11 - k] = __byte_perm(w[k], 0, 0x0123);
        y.v[11 - k] = __byte_perm(w[12 + k], 0, 0x0123);
    }
    uint32_t flags = x.v[11] >> 29;
    x.v[11] &= 0x1fffffffu;
    G1Affine p;
    if (flags & 4) {  // compressed flag in an uncompressed slot
        atomicOr(bad, 1u);
        p.x = Fq::zero(); p.y = Fq::zero();
    } else if (flags & 2) {
        p.x = Fq::zero(); p.y = Fq::zero();
    } else {
        bool ok = fq_lt_mod(x) && fq_lt_mod(y);
        p.x = x.to_mont();
        p.y = y.to_mont();
        Fq lhs = p.y.sqr();
        Fq rhs = p.x.sqr() * p.x + fq_b4();
        if (!ok || lhs != rhs) atomicOr(bad, 1u);
    }
    out[i] = p;
}

__global__ void k_points_to_be96(const G1Affine* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = in[i];
    uint32_t* w = reinterpret_cast<uint32_t*>(out + i * 96);
    if (p.is_inf()) {
        for (int k = 0; k < 24; k++) w[k] = 0;
        w[0] = 0x00000040u;  // byte 0 = 0x40
        return;
    }
    Fq x = p.x.from_mont(), y = p.y.from_mont();
#pragma unroll
    for (int k = 0; k < 12; k++) {
        w[k] = __byte_perm(x.v[11 - k], 0, 0x0123);
        w[12 + k] = __byte_perm(y.v[11 - k], 0, 0x0123);
    }
}


// ---- generation from a public test trapdoor ------------------------------------------------------
// out[j] = coef * w^j * inv_d[j]   (Lagrange basis values: coef = -(tau^n - 1)/n * R_i(tau_y),
// inv_d[j] = 1/(w^j - tau))
__global__ void k_lagrange_scalars(const Fr* __restrict__ inv_d, uint32_t n, const Fr* __restrict__ wt, Fr coef,
                                   Fr* __restrict__ out, uint64_t j0) {
    constexpr uint32_t E = 8;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * E;
    if (lo >= n) return;
    Fr a = pow_from_table(wt, j0 + lo) * coef;
    const Fr w = load_fr(wt);
    for (uint32_t i = 0; i < E && lo + i < n; i++) {
        store_fr(out + lo + i, a * load_fr(inv_d + lo + i));
        a = a * w;
    }
}

// out[j] = tau^j (the scalars of the monomial SRS [tau^j]_1); tt[k] = tau^(2^k)
// (times `scale`: row i of the bivariate monomial SRS uses scale = tau_y^i)
__global__ void k_power_scalars(const Fr* __restrict__ tt, uint32_t n, Fr* __restrict__ out, Fr scale) {
    constexpr uint32_t E = 8;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * E;
    if (lo >= n) return;
    Fr a = pow_from_table(tt, lo) * scale;
    const Fr tau = load_fr(tt);
    for (uint32_t i = 0; i < E && lo + i < n; i++) {
        store_fr(out + lo + i, a);
        a = a * tau;
    }
}

// [s]G for Montgomery-form scalars with the fixed-base table tab[w*256 + d] = [d * 256^w]G (affine)
__global__ void __launch_bounds__(128)
k_fixed_base_mul(const Fr* __restrict__ scalars, size_t n, const G1Affine* __restrict__ tab, G1Xyzz* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = load_fr(scalars + i).from_mont();
    G1Xyzz acc = G1Xyzz::infinity();
    for (int w = 0; w < 32; w++) {
        uint32_t d = (s.v[w >> 2] >> ((w & 3) * 8)) & 0xff;
        if (d) {
            G1Affine p = tab[w * 256 + d];
            acc.madd(p.x, p.y, 0);
        }
    }
    out[i] = acc;
}

// fixed-base table step: out[i] = [2^c] in[i]
__global__ void __launch_bounds__(128)
k_table_next(const char* __restrict__ in /* TABLE_STRIDE-byte records */, size_t n, uint32_t c, G1Xyzz* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = *reinterpret_cast<const G1Affine*>(in + i * TABLE_STRIDE);
    G1Xyzz acc = G1Xyzz::infinity();
    if (!p.is_inf()) {
        acc = G1Xyzz::dbl_affine(p.x, p.y);
        for (uint32_t d = 1; d < c; d++) acc = acc.dbl();
    }
    out[i] = acc;
}

// out[i] = -in[i] (the negated half of a fixed-base table; the infinity marker (0, 0) maps to itself)
__global__ void k_negate_points(const char* __restrict__ in, size_t n, char* __restrict__ out) {  // TABLE_STRIDE-byte records
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = *reinterpret_cast<const G1Affine*>(in + i * TABLE_STRIDE);
    p.y = p.y.neg();
    *reinterpret_cast<G1Affine*>(out + i * TABLE_STRIDE) = p;
}
// 96-byte points -> TABLE_STRIDE-byte records (a slice of a fixed-base table)
__global__ void k_pad_points(const G1Affine* __restrict__ in, size_t n, char* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    *reinterpret_cast<G1Affine*>(out + i * TABLE_STRIDE) = in[i];
}

// XYZZ -> affine with one inversion per run of E points (prefix products parked in scratch)
__global__ void __launch_bounds__(128)
k_xyzz_to_affine(const G1Xyzz* __restrict__ in, size_t n, uint32_t E, Fq* __restrict__ scratch, G1Affine* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * E;
    if (lo >= n) return;
    uint32_t cnt = n - lo < E ? (uint32_t)(n - lo) : E;
    Fq run = Fq::one();
    for (uint32_t i = 0; i < cnt; i++) {
        Fq z = in[lo + i].zzz;
        if (z.is_zero()) z = Fq::one();
        run = run * z;
        scratch[lo + i] = run;
    }
    Fq u = run.inverse();
    for (int i = (int)cnt - 1; i >= 0; i--) {
        G1Xyzz p = in[lo + i];
        G1Affine a;
        if (p.zz.is_zero()) {
            a.x = Fq::zero(); a.y = Fq::zero();
        } else {
            Fq zzz_inv = i > 0 ? u * scratch[lo + i - 1] : u;
            u = u * p.zzz;
            Fq t2 = p.zz * zzz_inv;  // ZZ / ZZZ ; its square is 1 / ZZ
            a.x = p.x * t2.sqr();
            a.y = p.y * zzz_inv;
        }
        out[lo + i] = a;
    }
}

}  // namespace zkp
