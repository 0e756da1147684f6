// Integer-pipe microbenchmarks for B200 (sm_100a): measures the issue rate of IMAD, IMAD.WIDE.U32,
// IMAD.HI, IADD3 and of the Fq / Fr Montgomery products built from them.  The IMAD.WIDE figure is the
// roofline denominator for the MSM.  Build: make -C tools microbench ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../zkp_subnet_b200/csrc/ff.cuh"

using namespace zkp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

template <int ILP>
__global__ void k_imad(uint32_t* out, uint32_t m, uint32_t c) {
    uint32_t acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x + k;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[k] = acc[k] * m + c;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_imad_hi(uint32_t* out, uint32_t m, uint32_t c) {
    uint32_t acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x * 0x9e3779b9u + k;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[k] = __umulhi(acc[k], m) + c;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_imad_wide(uint32_t* out, uint32_t m) {
    uint64_t acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[k] = (uint64_t)(uint32_t)acc[k] * m + acc[k];
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <int ILP>
__global__ void k_iadd3(uint32_t* out, uint32_t m, uint32_t c) {
    uint32_t acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x + k;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[k] = (acc[k] + m) ^ c;  // IADD3 + LOP3 on the alu pipe
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// IMAD.WIDE interleaved 1:1 with alu-pipe work: do the two pipes overlap?
template <int ILP>
__global__ void k_wide_plus_alu(uint32_t* out, uint32_t m, uint32_t c) {
    uint64_t acc[ILP];
    uint32_t alu[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) { acc[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k; alu[k] = threadIdx.x + k; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            acc[k] = (uint64_t)(uint32_t)acc[k] * m + acc[k];
            alu[k] = (alu[k] + m) ^ c;
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k] ^ alu[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <class F, int ILP>
__global__ void k_fmul(F* out, const F* in, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    F x[ILP], y = in[1];
#pragma unroll
    for (int k = 0; k < ILP; k++) { x[k] = in[0]; x[k].v[0] += tid + k; x[k].v[F::N - 1] &= 0x0fffffffu; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = x[k] * y;
    }
    F s = x[0];
#pragma unroll
    for (int k = 1; k < ILP; k++) s = s + x[k];
    out[tid] = s;
}

template <class F>
__global__ void k_faddsub(F* out, const F* in, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    F x = in[0], y = in[1];
    x.v[0] += tid;
    for (int it = 0; it < iters; it++) { x = x + y; y = y - x; }
    out[tid] = x + y;
}

// One Fq product followed by ADDS dependent Fq additions (~40 alu-pipe instructions each): is alu work hidden
// under the multiply pipe at the occupancy of the real kernels (12 warps/SM)?
template <int ADDS>
__global__ void k_fmul_adds(Fq* out, const Fq* in, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = in[0], y = in[1], z = in[1];
    x.v[0] += tid; x.v[11] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) {
        x = x * y;
#pragma unroll
        for (int k = 0; k < ADDS; k++) z = z + x;
    }
    out[tid] = x + z;
}
// FP64 pipe: dependent DFMA chains, alone and interleaved 1:1 with IMAD.WIDE chains
template <int ILP>
__global__ void k_dfma(double* out, double m, double c) {
    double acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[k] = fma(acc[k], m, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dfma_plus_wide(double* out, double m, double c, uint32_t mi) {
    double acc[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) { acc[k] = threadIdx.x * 1e-3 + k; w[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            acc[k] = fma(acc[k], m, c);
            w[k] = (uint64_t)(uint32_t)w[k] * mi + w[k];
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += acc[k] + (double)(w[k] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

int main(int argc, char** argv) {
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    uint32_t* out;
    CK(cudaMalloc(&out, sizeof(uint32_t) * sms * 64 * 1024));
    Fq* fq_out; Fq* fq_in;
    CK(cudaMalloc(&fq_out, sizeof(Fq) * sms * 16 * 1024));
    CK(cudaMalloc(&fq_in, sizeof(Fq) * 2));
    Fq hin[2];
    for (int i = 0; i < 12; i++) { hin[0].v[i] = FqParams::GX[i]; hin[1].v[i] = FqParams::GY[i]; }
    CK(cudaMemcpy(fq_in, hin, sizeof(hin), cudaMemcpyHostToDevice));
    Timer t;

    auto report = [&](const char* name, int warps_per_sm, double ops, float ms) {
        printf("{\"kernel\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gops_per_s\": %.1f, \"ops_per_clk_per_sm_at_max\": %.2f}\n",
               name, warps_per_sm, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
        fflush(stdout);
    };

    for (int wps : {8, 16, 32, 64}) {
        int threads = 256, blocks = sms * wps * 32 / threads;
        double nthreads = (double)threads * blocks;
#define RUN(NAME, OPS_PER_THREAD, ...)                                  \
        for (int rep = 0; rep < 3; rep++) {                             \
            t.start(); __VA_ARGS__; float ms = t.stop(); CK(cudaGetLastError());        \
            if (rep == 2) report(NAME, wps, nthreads * (OPS_PER_THREAD), ms);           \
        }
        RUN("imad32_ilp8", 8.0 * ITERS, k_imad<8><<<blocks, threads>>>(out, 0x9e3779b1u, 12345u))
        RUN("imad_hi_ilp8", 8.0 * ITERS, k_imad_hi<8><<<blocks, threads>>>(out, 0x9e3779b1u, 12345u))
        RUN("imad_wide_ilp8", 8.0 * ITERS, k_imad_wide<8><<<blocks, threads>>>(out, 0x9e3779b1u))
        RUN("imad_wide_ilp4", 4.0 * ITERS, k_imad_wide<4><<<blocks, threads>>>(out, 0x9e3779b1u))
        RUN("iadd3_lop3_ilp8", 16.0 * ITERS, k_iadd3<8><<<blocks, threads>>>(out, 0x9e3779b1u, 12345u))
        RUN("wide_plus_2alu_ilp8", 8.0 * ITERS, k_wide_plus_alu<8><<<blocks, threads>>>(out, 0x9e3779b1u, 12345u))
    }
    for (int wps : {4, 8, 12, 16, 24, 32}) {
        int threads = 128, blocks = sms * wps * 32 / threads;
        double nthreads = (double)threads * blocks;
        int iters = 512;
        RUN("fq_mul_ilp1", 1.0 * iters, k_fmul<Fq, 1><<<blocks, threads>>>(fq_out, fq_in, iters))
        RUN("fq_mul_ilp2", 2.0 * iters, k_fmul<Fq, 2><<<blocks, threads>>>(fq_out, fq_in, iters))
        RUN("fr_mul_ilp1", 1.0 * iters, k_fmul<Fr, 1><<<blocks, threads>>>((Fr*)fq_out, (Fr*)fq_in, iters))
        RUN("fr_mul_ilp2", 2.0 * iters, k_fmul<Fr, 2><<<blocks, threads>>>((Fr*)fq_out, (Fr*)fq_in, iters))
        RUN("fq_addsub", 2.0 * iters, k_faddsub<Fq><<<blocks, threads>>>(fq_out, fq_in, iters))
    }
    for (int wps : {12, 32}) {
        int threads = 128, blocks = sms * wps * 32 / threads;
        double nthreads = (double)threads * blocks;
        int iters = 512;
        RUN("fq_mul_plus_0_adds", 1.0 * iters, k_fmul_adds<0><<<blocks, threads>>>(fq_out, fq_in, iters))
        RUN("fq_mul_plus_1_adds", 1.0 * iters, k_fmul_adds<1><<<blocks, threads>>>(fq_out, fq_in, iters))
        RUN("fq_mul_plus_3_adds", 1.0 * iters, k_fmul_adds<3><<<blocks, threads>>>(fq_out, fq_in, iters))
        RUN("fq_mul_plus_6_adds", 1.0 * iters, k_fmul_adds<6><<<blocks, threads>>>(fq_out, fq_in, iters))
    }
    for (int wps : {16, 64}) {
        int threads = 256, blocks = sms * wps * 32 / threads;
        double nthreads = (double)threads * blocks;
        RUN("dfma_ilp8", 8.0 * ITERS, k_dfma<8><<<blocks, threads>>>((double*)out, 1.0000001, 1e-9))
        RUN("dfma_plus_wide_ilp4(pairs)", 4.0 * ITERS, k_dfma_plus_wide<4><<<blocks, threads>>>((double*)out, 1.0000001, 1e-9, 0x9e3779b1u))
    }
    CK(cudaDeviceSynchronize());
    // correctness spot check of the device Fq product against the host emulation of the same code
    {
        k_fmul<Fq, 1><<<1, 32>>>(fq_out, fq_in, 3);
        Fq got[32];
        CK(cudaMemcpy(got, fq_out, sizeof(got), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int tid = 0; tid < 32; tid++) {
            Fq x = hin[0], y = hin[1];
            x.v[0] += tid; x.v[11] &= 0x0fffffffu;
            for (int it = 0; it < 3; it++) x = x * y;
            if (x != got[tid]) bad++;
        }
        printf("{\"check\": \"fq_mul device==host-emulation\", \"bad\": %d}\n", bad);
    }
    return 0;
}
