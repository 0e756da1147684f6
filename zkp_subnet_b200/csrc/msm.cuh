// G1 Pippenger MSM for sm_100a -- the hot loop behind worker_commit / worker_open
// (reference neurons/miner.py:38-54 -> fourier Client.worker_commit / worker_open).
//
// Pipeline (all on one stream, no host sync until the W window sums are read back):
//   1. k_decompose      scalars (32 B, canonical) -> W signed c-bit digits each; emits
//                       key = window * 2^(c-1) + |digit| - 1, val = point index | sign << 31
//                       (zero digits get the DISCARD key and sort to the end).
//   2. radix sort       (cub::DeviceRadixSort over the key bits actually used) -> bucket order.
//   3. k_accumulate     BALANCED bucket accumulation: thread t owns the fixed-length slice
//                       [t*L, (t+1)*L) of the sorted entries, whatever buckets it spans, so every
//                       thread performs the same number of mixed additions regardless of the
//                       scalar distribution (no long-bucket stragglers, no warp divergence on
//                       bucket length).  Runs of equal keys that lie inside the slice are complete
//                       buckets and are written straight to the bucket array; the (at most two)
//                       runs cut by a slice boundary go to a "slot" list, which is itself a sorted
//                       (key, point) sequence and is reduced by the same kernel (XYZZ + XYZZ
//                       instead of XYZZ + affine) level by level until one thread sees it all.
//                       Deterministic: no atomics anywhere.
//   4. k_bucket_reduce  sum_b b * B[w][b] by chunked running sums, recursively on the chunk sums
//                       (each level multiplies its chunk sums by the chunk size with doublings, so
//                       all levels feed one pool of plain addends); k_sum_segments tree-sums the pool.
//   5. host             fold the W window sums (c doublings each), to affine, compress.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "g1.cuh"

namespace zkp {

constexpr uint32_t KEY_NONE = 0xffffffffu;   // slot of a thread that had no entries
constexpr uint32_t KEY_EMPTY_FLAG = 0x80000000u;  // slot carries a key but no point

struct MsmPlan {
    uint32_t n = 0;        // points
    uint32_t c = 0;        // window bits
    uint32_t W = 0;        // windows
    uint32_t B = 0;        // buckets per window = 2^(c-1)
    uint32_t key_bits = 0; // radix-sort bits
    uint32_t discard = 0;  // key of zero digits
    size_t N = 0;          // entries = n * W
    // accumulation levels: level 0 consumes entries, level k>0 consumes the slots of level k-1
    struct Level { size_t items; uint32_t L; size_t threads; };
    std::vector<Level> levels;
    // reduction levels
    struct RLevel { uint32_t n_in; uint32_t m; uint32_t chunks; };
    std::vector<RLevel> rlevels;
    uint32_t pool_per_window = 0;
    // tail: P[j] = sum of the inputs whose weight has bit j set (tail_n <= REDUCE_TAIL_MAX inputs)
    uint32_t tail_n = 0, tail_bits = 0, out_per_window = 0;
    bool tail_one_based = false;
};
constexpr uint32_t REDUCE_TAIL_MAX = 1024;

inline uint32_t msm_window_bits(uint32_t n) {
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    // measured sweet spots: buckets ~ n/32 per window
    int c = (int)lg - 4;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    return (uint32_t)c;
}

inline MsmPlan msm_make_plan(uint32_t n, int sm_count, uint32_t c_override = 0) {
    MsmPlan p;
    p.n = n;
    p.c = c_override ? c_override : msm_window_bits(n);
    p.W = 255 / p.c + 1;
    p.B = 1u << (p.c - 1);
    p.discard = p.W * p.B;
    p.key_bits = 1;
    while ((1ull << p.key_bits) <= p.discard) p.key_bits++;
    p.N = (size_t)n * p.W;
    // level 0: slice length chosen so that the grid is a whole number of waves of resident threads
    const size_t resident = (size_t)sm_count * 384;
    size_t waves = (p.N + resident * 32 - 1) / (resident * 32);
    if (waves < 1) waves = 1;
    uint32_t L0 = (uint32_t)((p.N + waves * resident - 1) / (waves * resident));
    if (L0 < 8) L0 = 8;
    size_t items = p.N;
    uint32_t L = L0;
    for (int lvl = 0;; lvl++) {
        size_t shift = lvl ? 1 : 0;  // slot levels slice on odd indices (see k_accumulate)
        size_t threads = items > shift ? (items - shift + L - 1) / L : 1;
        p.levels.push_back({items, L, threads});
        if (threads <= 1) break;
        items = threads * 2;
        L = lvl == 0 ? 16 : 32;
    }
    // reduction plan: chunked running sums while the level is large, bit-plane sums for the tail
    uint32_t n_in = p.B;
    while (n_in > REDUCE_TAIL_MAX) {
        uint32_t m = 8;
        uint32_t chunks = (n_in + m - 1) / m;
        p.rlevels.push_back({n_in, m, chunks});
        p.pool_per_window += chunks;
        n_in = chunks;
    }
    p.tail_n = n_in;
    p.tail_one_based = p.rlevels.empty();
    uint32_t max_weight = p.tail_one_based ? n_in : n_in - 1;
    p.tail_bits = 0;
    while ((1u << p.tail_bits) <= max_weight) p.tail_bits++;
    p.out_per_window = 1 + p.tail_bits;
    return p;
}

// ------------------------------------------------------------------------------------------------
// 1. signed-digit decomposition
// ------------------------------------------------------------------------------------------------
// scalars: 8 x u32 per scalar.  fmt: SCALAR_BE = the 32-byte big-endian wire format (reference
// base/protocol.py:35-40 poly strings after base64 decoding), SCALAR_LE = canonical little-endian limbs,
// SCALAR_MONT = little-endian Montgomery limbs (output of the opening kernels).  Non-canonical
// inputs (>= r) set *bad.
enum { SCALAR_LE = 0, SCALAR_BE = 1, SCALAR_MONT = 2 };
__global__ void k_decompose(const uint32_t* __restrict__ scalars, uint32_t n, uint32_t c, uint32_t W,
                            uint32_t B, uint32_t discard, int fmt, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ vals, uint32_t* __restrict__ bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[9];
    const uint4* src = reinterpret_cast<const uint4*>(scalars + (size_t)i * 8);
    uint4 a = src[0], b = src[1];
    if (fmt == SCALAR_BE) {
        s[7] = __byte_perm(a.x, 0, 0x0123); s[6] = __byte_perm(a.y, 0, 0x0123);
        s[5] = __byte_perm(a.z, 0, 0x0123); s[4] = __byte_perm(a.w, 0, 0x0123);
        s[3] = __byte_perm(b.x, 0, 0x0123); s[2] = __byte_perm(b.y, 0, 0x0123);
        s[1] = __byte_perm(b.z, 0, 0x0123); s[0] = __byte_perm(b.w, 0, 0x0123);
    } else {
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    }
    if (fmt == SCALAR_MONT) {
        Fr m;
#pragma unroll
        for (int k = 0; k < 8; k++) m.v[k] = s[k];
        m = m.from_mont();
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = m.v[k];
    } else {
        bool lt = false, decided = false;
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            uint32_t mk = FrParams::mod(k);
            if (!decided && s[k] != mk) { decided = true; lt = s[k] < mk; }
        }
        if (!lt) atomicOr(bad, 1u);
    }
    s[8] = 0;
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
    for (uint32_t w = 0; w < W; w++) {
        uint32_t bit = w * c, limb = bit >> 5, off = bit & 31;
        uint32_t raw = 0;
        if (limb < 8) {
            uint64_t two = ((uint64_t)s[limb + 1] << 32) | s[limb];
            raw = (uint32_t)(two >> off) & mask;
        }
        raw += carry;
        uint32_t neg = raw > half;
        uint32_t mag = neg ? (1u << c) - raw : raw;
        carry = neg;
        size_t o = (size_t)w * n + i;
        keys[o] = mag ? w * B + mag - 1 : discard;
        vals[o] = i | (neg << 31);
    }
}

// ------------------------------------------------------------------------------------------------
// 3. balanced bucket accumulation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_xyzz(G1Xyzz* dst, const G1Xyzz& p) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint32_t* s = p.x.v;  // x,y,zz,zzz are contiguous
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
__device__ __forceinline__ G1Xyzz load_xyzz(const G1Xyzz* src) {
    G1Xyzz p;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint32_t* d = p.x.v;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint4 t = s[i];
        d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
    }
    return p;
}
__device__ __forceinline__ G1Affine load_affine(const G1Affine* src) {
    G1Affine p;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint32_t* d = p.x.v;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint4 t = __ldg(s + i);
        d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
    }
    return p;
}

// LEVEL0: items are (key, point index|sign) entries, points gathered from the affine SRS row.
// else  : items are (key|flags, XYZZ) slots written by the previous level.
template <bool LEVEL0>
__global__ void __launch_bounds__(128, LEVEL0 ? 3 : 1)
k_accumulate(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
             const G1Affine* __restrict__ points, const G1Xyzz* __restrict__ slots_in, size_t items,
             uint32_t L, uint32_t discard, G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ slot_keys,
             G1Xyzz* __restrict__ slot_pts, int last_level) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // Slot lists hold (head_k, tail_k) pairs and the runs worth merging join tail_k with head_{k+1}
    // (odd index, even index): slices of the slot levels therefore start on ODD indices so that a
    // slice boundary falls between head_k and tail_k, never inside such a pair.
    const size_t shift = LEVEL0 ? 0 : 1;
    size_t start = t * L + (t ? shift : 0);
    if (start >= items) return;
    size_t end = (t + 1) * L + shift < items ? (t + 1) * L + shift : items;

    auto key_at = [&](size_t i) -> uint32_t {
        uint32_t k = keys[i];
        if (!LEVEL0 && k != KEY_NONE) k &= ~KEY_EMPTY_FLAG;
        return k;
    };
    const uint32_t prev_key = start > 0 ? key_at(start - 1) : KEY_NONE;
    const uint32_t next_key = end < items ? key_at(end) : KEY_NONE;

    uint32_t cur = KEY_NONE;      // key of the run being accumulated
    uint32_t first_key = KEY_NONE, last_key = KEY_NONE;
    bool is_first_run = true;     // the run being accumulated is the first of this slice
    bool have_head = false, have_tail = false;
    G1Xyzz acc = G1Xyzz::infinity();

    // a finished run goes to the bucket array if it is wholly inside this slice, else to a slot
    auto flush = [&](bool continues_after) {
        bool starts_before = is_first_run && cur == prev_key;
        if (last_level || (!starts_before && !continues_after)) {
            if (!acc.is_inf()) store_xyzz(buckets + cur, acc);
        } else if (is_first_run) {
            store_xyzz(slot_pts + 2 * t, acc);
            have_head = true;
        } else {
            store_xyzz(slot_pts + 2 * t + 1, acc);
            have_tail = true;
        }
    };

    for (size_t i = start; i < end; i++) {
        uint32_t raw = keys[i];
        uint32_t k = raw;
        if (LEVEL0) {
            if (k >= discard) break;  // zero digits sort last: nothing further in this slice
        } else {
            if (k == KEY_NONE) break;  // slots of threads beyond the valid range sort last
            k &= ~KEY_EMPTY_FLAG;
        }
        if (k != cur) {
            if (cur != KEY_NONE) {
                flush(false);
                is_first_run = false;
            } else {
                first_key = k;
            }
            cur = k;
            acc = G1Xyzz::infinity();
        }
        last_key = k;
        if (LEVEL0) {
            uint32_t v = vals[i];
            G1Affine p = load_affine(points + (v & 0x7fffffffu));
            acc.madd(p, v >> 31);
        } else if (!(raw & KEY_EMPTY_FLAG)) {
            G1Xyzz p = load_xyzz(slots_in + i);
            acc.add(p);
        }
    }
    if (cur != KEY_NONE) flush(next_key == cur);
    if (!last_level) {
        // every slice with at least one item publishes both keys so the next level can detect
        // run boundaries by looking at adjacent slots only
        slot_keys[2 * t] = first_key == KEY_NONE ? KEY_NONE : (have_head ? first_key : (first_key | KEY_EMPTY_FLAG));
        slot_keys[2 * t + 1] = last_key == KEY_NONE ? KEY_NONE : (have_tail ? last_key : (last_key | KEY_EMPTY_FLAG));
    }
}

// ------------------------------------------------------------------------------------------------
// 4. bucket reduction
// ------------------------------------------------------------------------------------------------
// One thread per (window, chunk of m inputs).  one_based: input j carries weight j+1 (bucket array),
// else weight j (chunk sums of the previous level).  Writes T = sum_i weight_local(i) * X[i] to the
// pool and, unless this is the last level, m * S = m * sum_i X[i] (log2 m doublings) as next input.
__global__ void __launch_bounds__(128, 1)
k_bucket_reduce(const G1Xyzz* __restrict__ in, uint32_t n_in, uint32_t in_stride, uint32_t m, uint32_t log_m,
                uint32_t chunks, int one_based, G1Xyzz* __restrict__ next, uint32_t next_stride,
                G1Xyzz* __restrict__ pool, uint32_t pool_stride, uint32_t pool_off, uint32_t W) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * W) return;
    uint32_t w = t / chunks, k = t % chunks;
    const G1Xyzz* x = in + (size_t)w * in_stride + (size_t)k * m;
    uint32_t cnt = n_in - k * m < m ? n_in - k * m : m;
    G1Xyzz running = G1Xyzz::infinity(), sum = G1Xyzz::infinity();
    for (int i = (int)cnt - 1; i >= 0; i--) {
        G1Xyzz p = load_xyzz(x + i);
        running.add(p);
        if (one_based || i > 0) sum.add(running);
    }
    store_xyzz(pool + (size_t)w * pool_stride + pool_off + k, sum);
    if (next) {
        for (uint32_t d = 0; d < log_m; d++) running = running.dbl();
        store_xyzz(next + (size_t)w * next_stride + k, running);
    }
}

// Tail of the reduction: block (j, w) computes P[w][j] = sum of in[w][k] over the k whose weight
// (k + one_based) has bit j set; n_in <= REDUCE_TAIL_MAX.  The host finishes with a Horner pass
// over the bits (sum_j 2^j P_j).
constexpr int TAIL_THREADS = 128;
__global__ void __launch_bounds__(TAIL_THREADS)
k_bit_sums(const G1Xyzz* __restrict__ in, uint32_t n_in, uint32_t in_stride, int one_based, G1Xyzz* __restrict__ out,
           uint32_t out_stride, uint32_t out_off) {
    __shared__ G1Xyzz sh[TAIL_THREADS];
    uint32_t j = blockIdx.x, w = blockIdx.y;
    G1Xyzz acc = G1Xyzz::infinity();
    for (uint32_t k = threadIdx.x; k < n_in; k += TAIL_THREADS) {
        if (((k + (one_based ? 1u : 0u)) >> j) & 1) {
            G1Xyzz p = load_xyzz(in + (size_t)w * in_stride + k);
            acc.add(p);
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = TAIL_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            G1Xyzz a = sh[threadIdx.x];
            a.add(sh[threadIdx.x + s]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(out + (size_t)w * out_stride + out_off + j, sh[0]);
}

// Tree-sum: block (part, w) adds in[w*in_stride + part*PART .. +PART) (clipped to count) into
// out[w*out_stride + part].
constexpr int SUM_THREADS = 128;
constexpr int SUM_PART = 512;
__global__ void __launch_bounds__(SUM_THREADS)
k_sum_segments(const G1Xyzz* __restrict__ in, uint32_t count, uint32_t in_stride, G1Xyzz* __restrict__ out,
               uint32_t out_stride) {
    __shared__ G1Xyzz sh[SUM_THREADS];
    uint32_t part = blockIdx.x, w = blockIdx.y;
    uint32_t lo = part * SUM_PART;
    uint32_t hi = lo + SUM_PART < count ? lo + SUM_PART : count;
    G1Xyzz acc = G1Xyzz::infinity();
    for (uint32_t i = lo + threadIdx.x; i < hi; i += SUM_THREADS) {
        G1Xyzz p = load_xyzz(in + (size_t)w * in_stride + i);
        acc.add(p);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SUM_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            G1Xyzz a = sh[threadIdx.x];
            a.add(sh[threadIdx.x + s]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(out + (size_t)w * out_stride + part, sh[0]);
}

}  // namespace zkp
