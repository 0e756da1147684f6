// Roofline denominators measured on the device the benchmark runs on: issue rate of IMAD.WIDE.U32
// (the 32x32+64->64 multiply-accumulate every Fq/Fr product is made of) and of a dependent chain of Fq
// Montgomery products.  Used by bench.py for roofline.peak; see DESIGN.md section "Roofline".
#pragma once
#include "ff.cuh"

namespace zkp {

constexpr int PEAK_ITERS = 4096;

__global__ void k_peak_imad_wide(uint32_t* out, uint32_t m) {
    uint64_t acc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) acc[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = (uint64_t)(uint32_t)acc[k] * m + acc[k];
    }
    uint64_t s = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

// fully unrolled Montgomery product (the fastest known form on this pipe; the kernels use a partly rolled
// one to stay inside the instruction cache), so that the measured ceiling does not depend on that choice
__device__ __forceinline__ Fq fq_mul_unrolled(const Fq& a, const Fq& b) {
    uint32_t even[12], odd[12];
    Fq r;
    chains::fq_row_first(even, odd, a.v, b.v[0]);
    chains::fq_row(odd, even, a.v, b.v[1]);
#pragma unroll
    for (int i = 2; i < 12; i += 2) {
        chains::fq_row(even, odd, a.v, b.v[i]);
        chains::fq_row(odd, even, a.v, b.v[i + 1]);
    }
    chains::fq_merge(r.v, odd, even);
    chains::fq_reduce_once(r.v, 0);
    return r;
}
__global__ void k_peak_fq_mul(Fq* out, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = Fq::one(), y = Fq::r2();
    x.v[0] += tid;
    for (int it = 0; it < iters; it++) x = fq_mul_unrolled(x, y);
    out[tid] = x;
}

}  // namespace zkp
