// Latency of a DEPENDENT chain of Fq products for few warps per scheduler: the throughput product of ff.cuh (two carry
// chains) against the low-latency product of ff_ll.cuh (column sums, no carry between instructions).  Prints cycles per
// product for grids of 1, 2, 4 and 12 warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
// -o build/microbench_ll tools/microbench_ll.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ff_ll.cuh"
using namespace zkp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int ITERS = 2000;
using Fq = Fp<FqParams>;
template <int KIND>
__global__ void k_chain(uint32_t* out, long long* cycles) {
    Fq a, b;
    for (int i = 0; i < 12; i++) { a.v[i] = threadIdx.x * 0x9e3779b9u + i * 77u + blockIdx.x; b.v[i] = threadIdx.x * 0x85ebca6bu + i * 31u + 5; }
    a.v[11] &= 0x0fffffffu; b.v[11] &= 0x0fffffffu;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (KIND == 0) a = a * b;
        else if (KIND == 1) a = mul_ll<FqParams>(a, b);
        else if (KIND == 2) a = Fq::mul2(a, b, b, a);
        else a = mul2_ll<FqParams>(a, b, b, a);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 12; i++) s ^= a.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t* out; long long* cyc;
    CK(cudaMalloc(&out, 4 * 148 * 1024)); CK(cudaMallocManaged(&cyc, 8));
    const char* names[4] = {"operator* (throughput product)", "mul_ll (low latency)", "mul2 (fused, throughput)", "mul2_ll (fused, low latency)"};
    for (int warps : {1, 2, 4, 8, 12, 16}) {
        for (int kind = 0; kind < 4; kind++) {
            // `warps` warps per SM in one block per SM
            auto launch = [&](int k) {
                if (k == 0) k_chain<0><<<sms, 32 * warps>>>(out, cyc);
                else if (k == 1) k_chain<1><<<sms, 32 * warps>>>(out, cyc);
                else if (k == 2) k_chain<2><<<sms, 32 * warps>>>(out, cyc);
                else k_chain<3><<<sms, 32 * warps>>>(out, cyc);
            };
            launch(kind); CK(cudaDeviceSynchronize());
            launch(kind); CK(cudaDeviceSynchronize());
            printf("%2d warps/SM  %-34s %7.0f cycles per product\n", warps, names[kind], (double)*cyc / ITERS);
        }
    }
    // results agree
    return 0;
}
