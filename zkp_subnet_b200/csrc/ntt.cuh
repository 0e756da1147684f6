// Fr NTT / iNTT over the natural-order radix-2 domain w_n = 7^((r-1)/n): out[k] = sum_i in[i] w^(ik),
// natural order in and out -- the transform behind Client.fft (reference neurons/validator.py:58-65)
// and the evaluation-form <-> coefficient-form bridge of the commit path.
//
// Two-pass ("four-step") decomposition n = n1 * n2 with every sub-transform done entirely in shared
// memory by one CTA (radix-4 DIT stage pairs in registers on a bit-reversed tile, up to 4096 elements = 128 KB,
// butterfly twiddles staged in shared memory):
//   pass 1: for each i2, size-n1 transform over i1 of x[i1*n2 + i2], times w_n^(i2*k1)  -> Y[k1*n2 + i2]
//   pass 2: for each k1, size-n2 transform over i2 of Y[k1*n2 + i2]                    -> X[k1 + n1*k2]
// Each element is read and written exactly once per pass; a 32-byte Fr is exactly one DRAM sector, so
// the strided tile accesses move no extra bytes.  Shared memory holds the tile as two 16-byte planes
// so that consecutive butterflies hit consecutive banks.
// One table per size: tw[e] = w_n^e for e < n/2; inverse twiddles are -tw[n/2 - e].
#pragma once
#include "kzg.cuh"

namespace zkp {

constexpr int NTT_MAX_THREADS = 512;  // launched with tile/4 threads: one radix-4 group per thread and stage pair
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
    uint32_t tma;         // 1: the tile is ONE contiguous run in global memory (pass 2, one column per CTA) and is brought
                          // in by a single bulk copy (cp.async.bulk -> mbarrier) into a staging area, then permuted into
                          // the bit-reversed two-plane layout.  Experiment (zkp_set_ntt_tma): DESIGN.md section 4.
};

// ---- TMA (bulk async copy) helpers: one elected thread arms the barrier with the byte count and issues the copy; every
// thread waits on the barrier's phase.  SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64 (B200_PROFILING.md).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

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

    if (p.tma) {
        // one bulk copy of the whole (contiguous) tile into the staging area behind the twiddles, then the permutation
        __shared__ __align__(8) uint64_t bar;
        uint4* stage = smem + 2 * tile + m;
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(stage, in + (uint64_t)c0 * p.in_cs, tile * 32u, &bar);
        mbar_wait(&bar, 0);
        for (uint32_t r = threadIdx.x; r < tile; r += NTT_THREADS) {
            const uint32_t pos = p.log_m ? (__brev(r) >> (32 - p.log_m)) : 0;
            lo[pos] = stage[2 * r];
            hi[pos] = stage[2 * r + 1];
        }
    } else
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

    // butterfly twiddles w_m^e (e < m/2) staged in shared memory once per CTA: a butterfly reads its twiddle
    // with shared-memory latency instead of waiting for L2 (ncu: long-scoreboard was the top stall)
    uint4* twl = smem + 2 * tile;
    uint4* twh = twl + (m >> 1);
    for (uint32_t e = threadIdx.x; e < (m >> 1); e += NTT_THREADS) {
        const uint4* src = reinterpret_cast<const uint4*>(tw_sub + e);
        twl[e] = src[0];
        twh[e] = src[1];
    }
    __syncthreads();
    const uint32_t half_m = m >> 1;
    auto twiddle = [&](uint32_t e) -> Fr {  // w_m^e (forward) or w_m^-e (inverse), e < m
        if (p.inverse) e = e ? m - e : 0;
        const bool negate = e >= half_m;
        if (negate) e -= half_m;
        uint4 a0 = twl[e], a1 = twh[e];
        Fr t;
        t.v[0] = a0.x; t.v[1] = a0.y; t.v[2] = a0.z; t.v[3] = a0.w; t.v[4] = a1.x; t.v[5] = a1.y; t.v[6] = a1.z; t.v[7] = a1.w;
        return negate ? t.neg() : t;
    };
    auto ld = [&](uint32_t i) -> Fr {
        uint4 a0 = lo[i], a1 = hi[i];
        Fr t;
        t.v[0] = a0.x; t.v[1] = a0.y; t.v[2] = a0.z; t.v[3] = a0.w; t.v[4] = a1.x; t.v[5] = a1.y; t.v[6] = a1.z; t.v[7] = a1.w;
        return t;
    };
    auto st = [&](uint32_t i, const Fr& t) {
        lo[i] = make_uint4(t.v[0], t.v[1], t.v[2], t.v[3]);
        hi[i] = make_uint4(t.v[4], t.v[5], t.v[6], t.v[7]);
    };

    // radix-4 DIT: stages s and s + 1 on four elements held in registers (same four twiddle products as two
    // radix-2 stages, half the shared-memory traffic and barriers)
    uint32_t s = 1;
    for (; s + 1 <= p.log_m; s += 2) {
        const uint32_t h = 1u << (s - 1);
        for (uint32_t g = threadIdx.x; g < tile / 4; g += NTT_THREADS) {
            const uint32_t c = g >> (p.log_m - 2), gi = g & ((m >> 2) - 1);
            const uint32_t jj = gi & (h - 1), blk = gi >> (s - 1);
            const uint32_t i0 = c * m + (blk << (s + 1)) + jj, i1 = i0 + h, i2 = i1 + h, i3 = i2 + h;
            Fr x0 = ld(i0), x1 = ld(i1), x2 = ld(i2), x3 = ld(i3);
            if (jj) {
                const Fr ta = twiddle(jj << (p.log_m - s));
                x1 = x1 * ta;
                x3 = x3 * ta;
            }
            Fr u0 = x0 + x1, u1 = x0 - x1, u2 = x2 + x3, u3 = x2 - x3;
            if (jj) u2 = u2 * twiddle(jj << (p.log_m - s - 1));
            u3 = u3 * twiddle((jj << (p.log_m - s - 1)) + (m >> 2));
            st(i0, u0 + u2);
            st(i2, u0 - u2);
            st(i1, u1 + u3);
            st(i3, u1 - u3);
        }
        __syncthreads();
    }
    // odd log_m: one radix-2 stage left
    for (; s <= p.log_m; s++) {
        const uint32_t half = 1u << (s - 1);
        for (uint32_t b = threadIdx.x; b < tile / 2; b += NTT_THREADS) {
            uint32_t c = b >> (p.log_m - 1), j = b & (m / 2 - 1);
            uint32_t jj = j & (half - 1);
            uint32_t pos = ((j >> (s - 1)) << s) + jj;
            uint32_t i0 = c * m + pos, i1 = i0 + half;
            Fr u = ld(i0), v = ld(i1);
            if (jj) v = v * twiddle(jj << (p.log_m - s));
            st(i0, u + v);
            st(i1, u - v);
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
