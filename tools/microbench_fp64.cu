// FP64-pipe Fq product (tools/fq_fp64.cuh): exactness against the integer product of ff.cuh on random and edge
// inputs, throughput alone, and throughput of a kernel whose warps alternate between the two formulations.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fq_fp64.cuh"
using namespace zkp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_check(const Fq* a, const Fq* b, int n, uint32_t* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq x = a[i], y = b[i];
    Fq want = x * y;
    Fq got = fp64::from_fp(fp64::mul(fp64::to_fp(x), fp64::to_fp(y)));
    if (got != want) atomicAdd(bad, 1u);
    Fq rt = fp64::from_fp(fp64::to_fp(x));
    if (rt != x) atomicAdd(bad + 1, 1u);
}
// mode 0: integer chain on every warp; 1: FP64 chain on every warp; 2: even warps integer, odd warps FP64
__global__ void k_chain(Fq* out, const Fq* in, int iters, int mode) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = in[0], y = in[1];
    x.v[0] += tid; x.v[11] &= 0x0fffffffu;
    const bool use_fp = mode == 1 || (mode == 2 && ((threadIdx.x >> 5) & 1));
    if (use_fp) {
        fp64::FqD xd = fp64::to_fp(x), yd = fp64::to_fp(y);
        for (int it = 0; it < iters; it++) xd = fp64::mul(xd, yd);
        x = fp64::from_fp(xd);
    } else {
        for (int it = 0; it < iters; it++) x = x * y;
    }
    out[tid] = x;
}
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static bool lt_p(const Fq& a) {
    for (int i = 11; i >= 0; i--) { if (a.v[i] < FqParams::MOD[i]) return true; if (a.v[i] > FqParams::MOD[i]) return false; }
    return false;
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    const int n = 1 << 18;
    std::vector<Fq> ha(n), hb(n);
    for (int i = 0; i < n; i++) {
        for (Fq* f : {&ha[i], &hb[i]}) {
            do { for (int k = 0; k < 12; k++) f->v[k] = (uint32_t)rnd(); f->v[11] &= 0x1fffffffu; } while (!lt_p(*f));
        }
        if (i % 97 == 0) for (int k = 0; k < 12; k++) ha[i].v[k] = FqParams::MOD[k] - (k == 0 ? 1 + i % 3 : 0);   // p - 1..3
        if (i % 101 == 0) for (int k = 0; k < 12; k++) hb[i].v[k] = k == 0 ? i % 5 : 0;                             // tiny
        if (i % 103 == 0) for (int k = 0; k < 12; k++) ha[i].v[k] = (k % 3 == 1) ? 0xffffu : (k == 11 ? 0x0fffffffu : 0xffffffffu);
    }
    Fq *da, *db, *dout; uint32_t* dbad;
    CK(cudaMalloc(&da, n * sizeof(Fq))); CK(cudaMalloc(&db, n * sizeof(Fq))); CK(cudaMalloc(&dbad, 8));
    CK(cudaMemcpy(da, ha.data(), n * sizeof(Fq), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), n * sizeof(Fq), cudaMemcpyHostToDevice));
    CK(cudaMemset(dbad, 0, 8));
    k_check<<<n / 128, 128>>>(da, db, n, dbad);
    uint32_t bad[2]; CK(cudaMemcpy(bad, dbad, 8, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"fp64 Fq product == integer product\", \"cases\": %d, \"mismatches\": %u, \"roundtrip_mismatches\": %u}\n", n, bad[0], bad[1]);
    CK(cudaMalloc(&dout, sizeof(Fq) * sms * 2048));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 256;
    for (int wps : {8, 12, 16, 24}) {
        int threads = 128, blocks = sms * wps * 32 / threads;
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0); k_chain<<<blocks, threads>>>(dout, da, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            printf("{\"kernel\": \"fq_mul_chain\", \"mode\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"G_fq_mul_per_s\": %.2f}\n",
                   mode == 0 ? "integer" : mode == 1 ? "fp64" : "half integer / half fp64", wps, best, (double)threads * blocks * iters / best * 1e-6);
        }
    }
    return 0;
}
