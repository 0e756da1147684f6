// Fr-side kernels of the KZG path: wire codec (32-byte big-endian <-> Montgomery limbs), the
// evaluation-form opening  y = f(x),  q_j = (f_j - y)/(w^j - x)  behind worker_open
// (reference neurons/miner.py:47-54), coefficient-form Horner behind Client.eval
// (reference neurons/validator.py:97-104; pinned by tests/test_miner.py:33-55) and the challenge RNG.
//
// Opening math (barycentric form on the natural-order domain {w^j}):
//     f(x) = (x^n - 1)/n * sum_j f_j w^j / (x - w^j)            x outside the domain
//     q_j  = (f_j - y) / (w^j - x)
//     x = w^m:  y = f_m,  q_m = - sum_{j != m} q_j w^(j-m)
// The n inversions use Montgomery's trick per thread over a run of E consecutive elements
// (prefix products parked in the output buffer, one Fermat inversion per run).
#pragma once
#include <cuda_runtime.h>

#include "ff.cuh"

namespace zkp {

constexpr uint32_t HIT_NONE = 0xffffffffu;

// Per-request device scalars of the opening: one SM_BYTES record per request (offsets in bytes).  The opening kernels
// are batch-aware: blockIdx.y selects the request -- polynomial blockIdx.y of n elements, record blockIdx.y, evaluation
// point xs[blockIdx.y] -- and a single request is simply gridDim.y == 1.
constexpr size_t SM_BAD = 0, SM_HIT = 4, SM_Y = 32, SM_S1 = 64, SM_S2 = 96, SM_EVAL = 128, SM_BYTES = 256;
constexpr uint32_t SM_STRIDE_U32 = SM_BYTES / 4, SM_STRIDE_FR = SM_BYTES / 32;

__device__ __forceinline__ Fr load_fr(const Fr* p) {
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4 a = s[0], b = s[1];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_fr(Fr* p, const Fr& r) {
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    d[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ bool fr_lt_mod(const Fr& a) {
    for (int i = 7; i >= 0; i--) {
        uint32_t m = FrParams::mod(i);
        if (a.v[i] < m) return true;
        if (a.v[i] > m) return false;
    }
    return false;
}
__device__ __forceinline__ Fr fr_bswap_load(const uint32_t* w) {
    const uint4* s = reinterpret_cast<const uint4*>(w);
    uint4 a = s[0], b = s[1];
    Fr r;
    r.v[7] = __byte_perm(a.x, 0, 0x0123); r.v[6] = __byte_perm(a.y, 0, 0x0123);
    r.v[5] = __byte_perm(a.z, 0, 0x0123); r.v[4] = __byte_perm(a.w, 0, 0x0123);
    r.v[3] = __byte_perm(b.x, 0, 0x0123); r.v[2] = __byte_perm(b.y, 0, 0x0123);
    r.v[1] = __byte_perm(b.z, 0, 0x0123); r.v[0] = __byte_perm(b.w, 0, 0x0123);
    return r;
}
__device__ __forceinline__ void fr_bswap_store(uint32_t* w, const Fr& r) {
    uint4* d = reinterpret_cast<uint4*>(w);
    d[0] = make_uint4(__byte_perm(r.v[7], 0, 0x0123), __byte_perm(r.v[6], 0, 0x0123),
                      __byte_perm(r.v[5], 0, 0x0123), __byte_perm(r.v[4], 0, 0x0123));
    d[1] = make_uint4(__byte_perm(r.v[3], 0, 0x0123), __byte_perm(r.v[2], 0, 0x0123),
                      __byte_perm(r.v[1], 0, 0x0123), __byte_perm(r.v[0], 0, 0x0123));
}

// 32-byte big-endian canonical -> Montgomery limbs; *bad |= 1 if any value >= r
__global__ void k_fr_from_be(const uint32_t* __restrict__ in, size_t n, Fr* __restrict__ out, uint32_t* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i += (size_t)blockIdx.y * n;
    Fr c = fr_bswap_load(in + i * 8);
    if (!fr_lt_mod(c)) atomicOr(bad + blockIdx.y * SM_STRIDE_U32, 1u);
    store_fr(out + i, c.to_mont());
}
__global__ void k_fr_to_be(const Fr* __restrict__ in, size_t n, uint32_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_bswap_store(out + i * 8, load_fr(in + i).from_mont());
}

// one field element per request record (Montgomery, at byte offset SM_Y) -> big-endian bytes at SM_EVAL
__global__ void k_fr_to_be_records(uint8_t* __restrict__ records, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint8_t* rec = records + (size_t)i * SM_BYTES;
    fr_bswap_store(reinterpret_cast<uint32_t*>(rec + SM_EVAL), load_fr(reinterpret_cast<const Fr*>(rec + SM_Y)).from_mont());
}

// w^e from the table wt[k] = w^(2^k)
__device__ __forceinline__ Fr pow_from_table(const Fr* __restrict__ wt, uint64_t e) {
    Fr acc = Fr::one();
    for (int k = 0; e; k++, e >>= 1)
        if (e & 1) acc = acc * load_fr(wt + k);
    return acc;
}

// block-wide sum of one Fr per thread -> partial[blockIdx.x] (blockDim.x <= 256, power of two)
__device__ __forceinline__ void block_sum_store(const Fr& v, Fr* __restrict__ partial) {
    __shared__ Fr sh[256];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) store_fr(partial + blockIdx.x, sh[0]);
}
// single block: out[0] = sum of count partials
__global__ void k_fr_reduce(const Fr* __restrict__ partial, uint32_t count, Fr* __restrict__ out) {
    partial += (size_t)blockIdx.y * count;
    Fr acc = Fr::zero();
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) acc = acc + load_fr(partial + i);
    block_sum_store(acc, out + blockIdx.y * SM_STRIDE_FR);
}

// ---- opening, pass 1: inv_d[j] = 1/(w^j - x), partial sums of f_j w^j / (w^j - x).
// Thread t owns elements [t*E, (t+1)*E).  wt[k] = w^(2^k), w_inv = w^-1.
// Montgomery's trick on two levels: prefix products inside the thread (parked in inv_d), then an inclusive
// prefix and suffix product scan over the block's 128 thread products in shared memory, and ONE Fermat inversion
// per block (a lone thread needs ~0.2 ms for the 255 dependent squarings -- with one inversion per thread that
// latency, times a handful of resident warps, was 1 ms of the opening's critical path at 2^20).
__global__ void __launch_bounds__(128)
k_open_pass1(const Fr* __restrict__ f, uint32_t n, uint32_t E, Fr x, const Fr* __restrict__ wt, Fr w_inv,
             Fr* __restrict__ inv_d, Fr* __restrict__ partial, uint32_t* __restrict__ hit, uint64_t j0,
             const Fr* __restrict__ xs = nullptr) {
    __shared__ Fr pre[128], suf[128];
    __shared__ Fr total_inv;
    const uint32_t tid = threadIdx.x;
    if (xs) x = load_fr(xs + blockIdx.y);
    if (f) f += (size_t)blockIdx.y * n;
    inv_d += (size_t)blockIdx.y * n;
    partial += (size_t)blockIdx.y * gridDim.x;
    hit += blockIdx.y * SM_STRIDE_U32;
    uint32_t t = blockIdx.x * blockDim.x + tid;
    uint64_t lo = (uint64_t)t * E;
    const uint32_t cnt = lo < n ? (n - lo < E ? (uint32_t)(n - lo) : E) : 0;
    Fr s1 = Fr::zero();
    const Fr w = load_fr(wt);
    Fr a = Fr::one(), run = Fr::one();
    if (cnt) {
        a = pow_from_table(wt, j0 + lo);  // w^(j0 + lo): element lo of a shard that starts at domain index j0
        // forward: prefix products of d_j parked in inv_d
        for (uint32_t i = 0; i < cnt; i++) {
            Fr d = a - x;
            if (d.is_zero()) {
                atomicMin(hit, (uint32_t)(lo + i));
                d = Fr::one();
            }
            run = run * d;
            store_fr(inv_d + lo + i, run);
            if (i + 1 < cnt) a = a * w;
        }
    }
    // block level: pre[k] = run_0 .. run_k, suf[k] = run_k .. run_127 (Hillis-Steele, 7 steps each)
    pre[tid] = run;
    suf[tid] = run;
    __syncthreads();
    for (uint32_t s = 1; s < 128; s <<= 1) {
        Fr p = pre[tid], q = suf[tid];
        if (tid >= s) p = pre[tid - s] * p;
        if (tid + s < 128) q = q * suf[tid + s];
        __syncthreads();
        pre[tid] = p;
        suf[tid] = q;
        __syncthreads();
    }
    if (tid == 0) total_inv = pre[127].inverse();
    __syncthreads();
    if (cnt) {
        // 1 / run_tid = (product of the other threads' runs) / (product of all runs)
        Fr u = total_inv;
        if (tid > 0) u = u * pre[tid - 1];
        if (tid < 127) u = u * suf[tid + 1];
        // backward: a currently holds w^(lo+cnt-1)
        for (int i = (int)cnt - 1; i >= 0; i--) {
            Fr d = a - x;
            bool zero = d.is_zero();
            if (zero) d = Fr::one();
            Fr inv = i > 0 ? u * load_fr(inv_d + lo + i - 1) : u;
            u = u * d;
            store_fr(inv_d + lo + i, inv);
            if (f && !zero) s1 = s1 + load_fr(f + lo + i) * a * inv;
            if (i > 0) a = a * w_inv;
        }
    }
    block_sum_store(s1, partial);
}

// ---- opening, pass 1 WITHOUT an inversion on the device (single request, x outside the domain).
// The Fermat inversion of k_open_pass1 is ~0.2 ms of pure latency (255 dependent squarings on one lane) -- nothing
// at 2^20 next to two MSMs, a tenth of a whole worker_open at the mainnet row size (2^16).  It disappears when the
// blocks own COSETS instead of contiguous runs: with m = 128 E elements per block and S = n / m blocks, block b takes
// j = b + k S, k < m, i.e. the points w^b z^k with z = w^S of order m, and
//        prod_k (w^b z^k - x) = x^m - w^(b m)                      (m even)
// is a closed form; the S block products multiply to x^n - 1, whose inverse the HOST computes (it has x; ~15 us of
// 64-bit limb arithmetic while the device is busy with the previous kernels).  k_open_coset_inv turns that single
// inverse into the S inverses 1/(x^m - w^(b m)) by Montgomery's trick (one block: runs of consecutive b per thread,
// prefix and suffix scans over the thread products), and pass 1 proper runs with no inversion at all.
// Elements of a block are strided by S in memory: 32-byte accesses, one sector each, and the blocks of the grid --
// all resident at once -- walk neighbouring sectors at the same time.
constexpr uint32_t COSET_INV_THREADS = 512;
__global__ void __launch_bounds__(COSET_INV_THREADS)
k_open_coset_inv(Fr x_m, Fr total_inv, const Fr* __restrict__ wt, uint32_t log_m, uint32_t S, Fr* __restrict__ inv_blocks,
                 const Fr* __restrict__ host_vals = nullptr) {
    __shared__ Fr pre[COSET_INV_THREADS], suf[COSET_INV_THREADS];
    if (host_vals) {  // a batch: request blockIdx.y has its own x (host_vals: x^m, 1/(x^n - 1), (x^n - 1)/n per request)
        x_m = load_fr(host_vals + 3 * blockIdx.y);
        total_inv = load_fr(host_vals + 3 * blockIdx.y + 1);
        inv_blocks += (size_t)blockIdx.y * S;
    }
    const uint32_t tid = threadIdx.x, T = blockDim.x;  // T = min(S, 512), a power of two; R = S / T values per thread
    const uint32_t R = S / T, lo = tid * R;
    const Fr g = load_fr(wt + log_m);  // w^m
    Fr a = pow_from_table(wt, (uint64_t)lo << log_m), run = Fr::one();
    for (uint32_t i = 0; i < R; i++) {
        run = run * (x_m - a);
        store_fr(inv_blocks + lo + i, run);
        if (i + 1 < R) a = a * g;
    }
    pre[tid] = run;
    suf[tid] = run;
    __syncthreads();
    for (uint32_t s = 1; s < T; s <<= 1) {
        Fr p = pre[tid], q = suf[tid];
        if (tid >= s) p = pre[tid - s] * p;
        if (tid + s < T) q = q * suf[tid + s];
        __syncthreads();
        pre[tid] = p;
        suf[tid] = q;
        __syncthreads();
    }
    Fr u = total_inv;
    if (tid > 0) u = u * pre[tid - 1];
    if (tid + 1 < T) u = u * suf[tid + 1];
    if (R > 1) {
        // w^-m = w^(n - m) with n = S m
        const Fr g_inv = pow_from_table(wt, ((uint64_t)S - 1) << log_m);
        for (int i = (int)R - 1; i >= 0; i--) {
            const Fr d = x_m - a;
            const Fr inv = i > 0 ? u * load_fr(inv_blocks + lo + i - 1) : u;
            u = u * d;
            store_fr(inv_blocks + lo + i, inv);
            if (i > 0) a = a * g_inv;
        }
    } else {
        store_fr(inv_blocks + lo, u);
    }
}
// grid = S blocks of 128 threads, thread t of block b owns k in [t E, (t + 1) E): element j = b + k S.  g_inv = w^-S.
__global__ void __launch_bounds__(128)
k_open_pass1_coset(const Fr* __restrict__ f, uint32_t E, uint32_t log_S, Fr x, const Fr* __restrict__ wt, Fr g_inv,
                   const Fr* __restrict__ inv_blocks, Fr* __restrict__ inv_d, Fr* __restrict__ partial,
                   const Fr* __restrict__ xs = nullptr) {
    __shared__ Fr pre[128], suf[128];
    if (xs) {  // a batch: request blockIdx.y
        const size_t n = (size_t)(128 * E) << log_S;
        x = load_fr(xs + blockIdx.y);
        f += blockIdx.y * n;
        inv_d += blockIdx.y * n;
        inv_blocks += (size_t)blockIdx.y << log_S;
        partial += (size_t)blockIdx.y << log_S;
    }
    const uint32_t tid = threadIdx.x, b = blockIdx.x, k0 = tid * E;
    auto at = [&](uint32_t k) -> size_t { return (size_t)b + ((size_t)k << log_S); };
    const Fr g = load_fr(wt + log_S);  // w^S
    Fr a = pow_from_table(wt, at(k0)), run = Fr::one();
    for (uint32_t i = 0; i < E; i++) {
        run = run * (a - x);  // never zero: the host has checked x^n != 1
        store_fr(inv_d + at(k0 + i), run);
        if (i + 1 < E) a = a * g;
    }
    pre[tid] = run;
    suf[tid] = run;
    __syncthreads();
    for (uint32_t s = 1; s < 128; s <<= 1) {
        Fr p = pre[tid], q = suf[tid];
        if (tid >= s) p = pre[tid - s] * p;
        if (tid + s < 128) q = q * suf[tid + s];
        __syncthreads();
        pre[tid] = p;
        suf[tid] = q;
        __syncthreads();
    }
    Fr u = load_fr(inv_blocks + b);
    if (tid > 0) u = u * pre[tid - 1];
    if (tid < 127) u = u * suf[tid + 1];
    Fr s1 = Fr::zero();
    for (int i = (int)E - 1; i >= 0; i--) {
        const Fr d = a - x;
        const Fr inv = i > 0 ? u * load_fr(inv_d + at(k0 + i - 1)) : u;
        u = u * d;
        store_fr(inv_d + at(k0 + i), inv);
        s1 = s1 + load_fr(f + at(k0 + i)) * a * inv;
        if (i > 0) a = a * g_inv;
    }
    block_sum_store(s1, partial);
}
// y = -zn * (sum of the partials), zn = (x^n - 1)/n from the host.  One block.
__global__ void __launch_bounds__(256)
k_open_reduce_y(const Fr* __restrict__ partial, uint32_t count, Fr zn, Fr* __restrict__ s1_out, Fr* __restrict__ y,
                const Fr* __restrict__ host_vals = nullptr) {
    __shared__ Fr sh[256];
    if (host_vals) {  // a batch: request blockIdx.y, outputs in its record
        zn = load_fr(host_vals + 3 * blockIdx.y + 2);
        partial += (size_t)blockIdx.y * count;
        s1_out += blockIdx.y * SM_STRIDE_FR;
        y += blockIdx.y * SM_STRIDE_FR;
    }
    Fr acc = Fr::zero();
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) acc = acc + load_fr(partial + i);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        store_fr(s1_out, sh[0]);
        store_fr(y, (zn * sh[0]).neg());
    }
}

// y = f(x):  hit -> f[hit], else -(x^n - 1)/n * S1.   Single thread.
__global__ void k_open_y(const Fr* __restrict__ f, uint32_t log_n, Fr x, Fr n_inv, const Fr* __restrict__ s1,
                         const uint32_t* __restrict__ hit, Fr* __restrict__ y, const Fr* __restrict__ xs = nullptr,
                         uint32_t n = 0) {
    if (threadIdx.x || blockIdx.x) return;
    if (xs) x = load_fr(xs + blockIdx.y);
    f += (size_t)blockIdx.y * n;
    s1 += blockIdx.y * SM_STRIDE_FR;
    hit += blockIdx.y * SM_STRIDE_U32;
    y += blockIdx.y * SM_STRIDE_FR;
    if (*hit != HIT_NONE) {
        store_fr(y, load_fr(f + *hit));
        return;
    }
    Fr xn = x;
    for (uint32_t k = 0; k < log_n; k++) xn = xn.sqr();
    Fr zn = (xn - Fr::one()) * n_inv;
    store_fr(y, (zn * load_fr(s1)).neg());
}

// ---- opening, pass 2: q_j = (f_j - y) * inv_d_j   (Montgomery form; the MSM's digit kernel converts)
__global__ void k_open_pass2(const Fr* __restrict__ f, const Fr* __restrict__ inv_d, uint32_t n, const Fr* __restrict__ y,
                             Fr* __restrict__ q) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t o = (size_t)blockIdx.y * n + j;
    Fr yy = load_fr(y + blockIdx.y * SM_STRIDE_FR);
    store_fr(q + o, (load_fr(f + o) - yy) * load_fr(inv_d + o));
}

// ---- x = w^m: partial sums of q_j w^(j-m), j != m; then q_m = -sum
__global__ void k_open_fix_partial(const Fr* __restrict__ q, uint32_t n, const Fr* __restrict__ wt,
                                   const uint32_t* __restrict__ hit, Fr* __restrict__ partial) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    Fr v = Fr::zero();
    uint32_t m = hit[blockIdx.y * SM_STRIDE_U32];
    q += (size_t)blockIdx.y * n;
    if (m != HIT_NONE && j < n && j != m) v = load_fr(q + j) * pow_from_table(wt, (uint64_t)((j + n - m) & (n - 1)));
    block_sum_store(v, partial + (size_t)blockIdx.y * gridDim.x);
}
__global__ void k_open_fix_apply(Fr* __restrict__ q, const uint32_t* __restrict__ hit, const Fr* __restrict__ s2, uint32_t n = 0) {
    if (threadIdx.x || blockIdx.x) return;
    const uint32_t m = hit[blockIdx.y * SM_STRIDE_U32];
    if (m != HIT_NONE) store_fr(q + (size_t)blockIdx.y * n + m, load_fr(s2 + blockIdx.y * SM_STRIDE_FR).neg());
}

// ---- batched barycentric evaluation (validator challenge: every row of the random bivariate polynomial at the
// same alpha).  k_bary_weights: wd[j] = w^j / (w^j - x) from inv_d (in place); k_bary_rows: block (bx, row)
// sums f[row][j] * wd[j] over its share -> partial[row * gridDim.x + bx].
__global__ void k_bary_weights(Fr* __restrict__ inv_d, uint32_t n, const Fr* __restrict__ wt) {
    constexpr uint32_t E = 16;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * E;
    if (lo >= n) return;
    Fr a = pow_from_table(wt, lo);
    const Fr w = load_fr(wt);
    for (uint32_t i = 0; i < E && lo + i < n; i++) {
        store_fr(inv_d + lo + i, load_fr(inv_d + lo + i) * a);
        a = a * w;
    }
}
__global__ void __launch_bounds__(256)
k_bary_rows(const Fr* __restrict__ f, const Fr* __restrict__ wd, uint32_t n, uint32_t per_block, Fr* __restrict__ partial) {
    const uint32_t row = blockIdx.y;
    const Fr* fr = f + (size_t)row * n;
    const uint32_t lo = blockIdx.x * per_block, hi = lo + per_block < n ? lo + per_block : n;
    Fr acc = Fr::zero();
    for (uint32_t j = lo + threadIdx.x; j < hi; j += blockDim.x) acc = acc + load_fr(fr + j) * load_fr(wd + j);
    __shared__ Fr sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) store_fr(partial + (size_t)row * gridDim.x + blockIdx.x, sh[0]);
}
// y[row] = hit ? f[row][hit] : -(zn) * sum of the row's partials       (one block per row)
__global__ void __launch_bounds__(32)
k_bary_finish(const Fr* __restrict__ f, uint32_t n, const Fr* __restrict__ partial, uint32_t parts, Fr zn,
              const uint32_t* __restrict__ hit, uint32_t* __restrict__ out_be) {
    const uint32_t row = blockIdx.x;
    if (threadIdx.x) return;
    Fr y;
    if (*hit != HIT_NONE) {
        y = load_fr(f + (size_t)row * n + *hit);
    } else {
        Fr acc = Fr::zero();
        for (uint32_t k = 0; k < parts; k++) acc = acc + load_fr(partial + (size_t)row * parts + k);
        y = (zn * acc).neg();
    }
    fr_bswap_store(out_be + (size_t)row * 8, y.from_mont());
}

// ---- Horner on coefficient form: thread t evaluates its run of E coefficients and scales by x^(tE)
// xt[k] = x^(2^k)
__global__ void __launch_bounds__(128)
k_eval_partial(const Fr* __restrict__ c, uint32_t n, uint32_t E, Fr x, const Fr* __restrict__ xt, Fr* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * E;
    Fr acc = Fr::zero();
    if (lo < n) {
        uint32_t cnt = n - lo < E ? (uint32_t)(n - lo) : E;
        for (int i = (int)cnt - 1; i >= 0; i--) acc = acc * x + load_fr(c + lo + i);
        acc = acc * pow_from_table(xt, lo);
    }
    block_sum_store(acc, partial);
}

// ---- challenge RNG: counter-based SplitMix64, 255-bit candidates, rejection until < r
__global__ void k_random_fr(uint64_t seed, size_t n, uint32_t* __restrict__ out_be, uint64_t index0 = 0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_be += i * 8;
    i += index0;  // element i of the stream `seed`, wherever it is written
    Fr v;
    for (uint64_t attempt = 0;; attempt++) {
        uint64_t s = seed + 0x9e3779b97f4a7c15ull * (4 * (i * 64 + attempt) + 1);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            s += 0x9e3779b97f4a7c15ull;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            z ^= z >> 31;
            v.v[2 * k] = (uint32_t)z;
            v.v[2 * k + 1] = (uint32_t)(z >> 32);
        }
        v.v[7] &= 0x7fffffffu;
        if (fr_lt_mod(v)) break;
    }
    fr_bswap_store(out_be, v);
}

}  // namespace zkp
