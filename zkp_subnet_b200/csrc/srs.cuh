// SRS kernels: import/export of G1 points in the ZCash 96-byte uncompressed wire format and on-device
// generation of Pianist Lagrange rows U[i][j] = [R_i(tau_y) L_j(tau_x)]_1 from a public test trapdoor
// (replaces `prover setup --generate-setup --generate-precompute`, reference tests/conftest.py:50-65).
#pragma once
#include "g1.cuh"

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

}  // namespace zkp
