// Fr NTT / iNTT over the natural-order radix-2 domain w_n = 7^((r-1)/n): out[k] = sum_i in[i] w^(ik),
// natural order in and out -- the transform behind Client.fft (reference neurons/validator.py:58-65)
// and the evaluation-form <-> coefficient-form bridge of the commit path.
//
// Two-pass ("four-step") decomposition n = n1 * n2 with every sub-transform done entirely in shared
// memory by one CTA (radix-2 DIT stages on a bit-reversed tile, up to 4096 elements = 128 KB):
//   pass 1: for each i2, size-n1 transform over i1 of x[i1*n2 + i2], times w_n^(i2*k1)  -> Y[k1*n2 + i2]
//   pass 2: for each k1, size-n2 transform over i2 of Y[k1*n2 + i2]                    -> X[k1 + n1*k2]
// Each element is read and written exactly once per pass; a 32-byte Fr is exactly one DRAM sector, so
// the strided tile accesses move no extra bytes.  Shared memory holds the tile as two 16-byte planes
// so that consecutive butterflies hit consecutive banks.
// One table per size: tw[e] = w_n^e for e < n/2; inverse twiddles are -tw[n/2 - e].
#pragma once
#include "kzg.cuh"

namespace zkp {

constexpr int NTT_MAX_THREADS = 512;  // launched with tile/8 threads: 4 butterflies per thread and stage
constexpr uint32_t NTT_MAX_TILE_LOG = 12;  // 4096 elements * 32 B = 128 KB dynamic shared memory

// tw[e] = w^e, e < half; wt[k] = w^(2^k)
__global__ void k_build_twiddles(Fr* __restrict__ tw, uint32_t half, const Fr* __restrict__ wt) {
    constexpr uint32_t E = 16;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * E;
    if (lo >= half) return;
    Fr a = pow_from_table(wt, lo);
    const Fr w = load_fr(wt);
    for (uint32_t i = 0; i < E && lo + i < half; i++) {
        store_fr(tw + lo + i, a);
        a = a * w;
    }
}

__device__ __forceinline__ Fr ntt_twiddle(const Fr* __restrict__ tw, uint32_t e, uint32_t half_n, int inverse) {
    // e in [0, n): w^e (forward) or w^-e (inverse) from the half table
    if (inverse) e = e ? 2 * half_n - e : 0;
    if (e < half_n) return load_fr(tw + e);
    return load_fr(tw + (e - half_n)).neg();
}

struct NttPass {
    uint32_t log_m;       // sub-transform size m = 2^log_m
    uint32_t log_cols;    // columns per CTA
    uint32_t ncols;       // total columns
    uint64_t in_rs, in_cs, out_rs, out_cs;
    uint32_t load_rows_fast;  // 1: consecutive threads walk rows (in_rs == 1), 0: walk columns
    uint32_t log_n;       // full transform size
    uint32_t twiddle;     // 1: multiply output (k, c) by w_n^(c*k)
    uint32_t inverse;
    uint32_t scale;       // 1: multiply output by n_inv
};

// tw: w_n^e (e < n/2) for the inter-pass twiddle; tw_sub: w_m^e (e < m/2), contiguous, for the butterflies
// (a 16 KB table that stays in L1 instead of one 128-byte line per twiddle of the big table)
__global__ void __launch_bounds__(NTT_MAX_THREADS)
k_ntt_pass(const Fr* __restrict__ in, Fr* __restrict__ out, const Fr* __restrict__ tw, const Fr* __restrict__ tw_sub,
           NttPass p, Fr n_inv) {
    const uint32_t NTT_THREADS = blockDim.x;
    extern __shared__ uint4 smem[];
    const uint32_t m = 1u << p.log_m, cols = 1u << p.log_cols, tile = m * cols;
    uint4* lo = smem;
    uint4* hi = smem + tile;
    const uint32_t c0 = blockIdx.x * cols;
    const uint32_t half_n = 1u << (p.log_n - 1);

    // load, rows bit-reversed
    for (uint32_t idx = threadIdx.x; idx < tile; idx += NTT_THREADS) {
        uint32_t r, c;
        if (p.load_rows_fast) { c = idx >> p.log_m; r = idx & (m - 1); }
        else { r = idx >> p.log_cols; c = idx & (cols - 1); }
        const uint4* src = reinterpret_cast<const uint4*>(in + (uint64_t)r * p.in_rs + (uint64_t)(c0 + c) * p.in_cs);
        uint32_t pos = p.log_m ? (__brev(r) >> (32 - p.log_m)) : 0;
        lo[c * m + pos] = src[0];
        hi[c * m + pos] = src[1];
    }
    __syncthreads();

    // radix-2 DIT stages
    const uint32_t half_m = m >> 1;
    for (uint32_t s = 1; s <= p.log_m; s++) {
        const uint32_t half = 1u << (s - 1);
        for (uint32_t b = threadIdx.x; b < tile / 2; b += NTT_THREADS) {
            uint32_t c = b >> (p.log_m - 1), j = b & (m / 2 - 1);
            uint32_t jj = j & (half - 1);
            uint32_t pos = ((j >> (s - 1)) << s) + jj;
            uint32_t i0 = c * m + pos, i1 = i0 + half;
            uint4 a0 = lo[i0], a1 = hi[i0], b0 = lo[i1], b1 = hi[i1];
            Fr u, v;
            u.v[0] = a0.x; u.v[1] = a0.y; u.v[2] = a0.z; u.v[3] = a0.w; u.v[4] = a1.x; u.v[5] = a1.y; u.v[6] = a1.z; u.v[7] = a1.w;
            v.v[0] = b0.x; v.v[1] = b0.y; v.v[2] = b0.z; v.v[3] = b0.w; v.v[4] = b1.x; v.v[5] = b1.y; v.v[6] = b1.z; v.v[7] = b1.w;
            if (jj) v = v * ntt_twiddle(tw_sub, jj << (p.log_m - s), half_m, p.inverse);
            Fr x = u + v, y = u - v;
            lo[i0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
            hi[i0] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
            lo[i1] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
            hi[i1] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
        }
        __syncthreads();
    }

    // store (consecutive threads walk columns: out_cs == 1 in both passes), fused twiddle / scaling
    for (uint32_t idx = threadIdx.x; idx < tile; idx += NTT_THREADS) {
        uint32_t k = idx >> p.log_cols, c = idx & (cols - 1);
        uint4 a0 = lo[c * m + k], a1 = hi[c * m + k];
        Fr x;
        x.v[0] = a0.x; x.v[1] = a0.y; x.v[2] = a0.z; x.v[3] = a0.w; x.v[4] = a1.x; x.v[5] = a1.y; x.v[6] = a1.z; x.v[7] = a1.w;
        if (p.twiddle) {
            uint32_t e = (uint32_t)(((uint64_t)(c0 + c) * k) & ((1ull << p.log_n) - 1));
            if (e) x = x * ntt_twiddle(tw, e, half_n, p.inverse);
        }
        if (p.scale) x = x * n_inv;
        store_fr(out + (uint64_t)k * p.out_rs + (uint64_t)(c0 + c) * p.out_cs, x);
    }
}

}  // namespace zkp
