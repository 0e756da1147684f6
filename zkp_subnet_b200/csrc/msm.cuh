// G1 Pippenger MSM for sm_100a -- the hot loop behind worker_commit / worker_open
// (reference neurons/miner.py:38-54 -> fourier Client.worker_commit / worker_open).
//
// Pipeline (all on one stream, no host sync until the W window sums are read back):
//   1. k_decompose      scalars (32 B, canonical) -> W signed c-bit digits each; an entry is
//                       key = window * 2^(c-1) + |digit| - 1 and val = point index | sign << 31, or, with
//                       fixed-base tables, key = |digit| - 1 and val = index into the table half of that sign
//                       (zero digits get the DISCARD key and end up last).
//   2. bucket sort      the entries GROUPED by key, in key order (the order inside a bucket is irrelevant): a
//                       counting sort written for this pipeline -- k_decompose<COUNT> histograms the keys,
//                       k_sort_scan_chunks / k_sort_scan_sums turn the histogram into start positions,
//                       k_decompose<SCATTER> recomputes the digits and places every (key, val) pair with one
//                       warp-aggregated atomicAdd on its key's cursor (section 2).  cub::DeviceRadixSort over the
//                       key bits actually used stays selectable (zkp_set_msm_sort) and is used above 2^27 entries
//                       and in front of the batched-affine rounds.
//  (2b. optional        rounds of batched-affine pairwise additions inside every bucket, msm_affine.cuh; off by default)
//   3. k_accumulate     BALANCED bucket accumulation: thread t owns the fixed-length slice
//                       [t*L, (t+1)*L) of the sorted entries, whatever buckets it spans, so every
//                       thread performs the same number of mixed additions regardless of the
//                       scalar distribution (no long-bucket stragglers, no warp divergence on
//                       bucket length).  Runs of equal keys that lie inside the slice are complete
//                       buckets and are written straight to the bucket array; the (at most two)
//                       runs cut by a slice boundary go to a "slot" list, which is itself a sorted
//                       (key, point) sequence and is reduced by the same kernel (XYZZ + XYZZ
//                       instead of XYZZ + affine) level by level until one thread sees it all; small slot
//                       levels use four lanes per slice that share every addition (g1_coop.cuh).
//                       No atomics: every bucket and every slot has exactly one writer.  (The bucket sort of step 2
//                       does use atomics, so the ORDER of the additions inside a bucket varies from run to run; the
//                       sums are exact group elements, so the output bytes do not.)
//   4. reduction        sum_b (b+1) B[b] with b = hi * 2^cl + lo: k_rowcol_partial / k_rowcol_finish form the row
//                       sums R_hi and column sums C_lo (2 additions per bucket: one thread per interleaved
//                       share of a sum, then one warp per sum), k_bit_sums the bit planes P_j = sum of the R
//                       (resp. C) whose weight has bit j set; the host finishes with two Horner passes.  No long
//                       sequential chains: a lone warp needs ~17 us per point addition on this machine, so depth,
//                       not work, is what the tail of an MSM costs.
//   5. host             Horner over the bit planes, window fold (classic mode only), to affine, compress.
//
// Fixed-base mode (plan.precomp): the SRS row is resident, so the context keeps [+-2^(c w)] P_i for every
// digit position w (2 W x the row in HBM).  All digits then feed ONE set of 2^(c-1) buckets: no window fold,
// a 16x smaller reduction, and room for a wider window (c = log2 n), i.e. ~19% fewer bucket additions.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#include "g1_coop.cuh"

namespace zkp {

constexpr uint32_t TABLE_STRIDE = 128;       // bytes per record of a fixed-base table (96 used, padded to one L2 line)
constexpr uint32_t KEY_NONE = 0xffffffffu;   // slot of a thread that had no entries
constexpr uint32_t KEY_EMPTY_FLAG = 0x80000000u;  // slot carries a key but no point

struct MsmPlan {
    uint32_t n = 0;        // points
    uint32_t c = 0;        // window bits
    uint32_t W = 0;        // digit positions per scalar
    uint32_t Wb = 0;       // bucket windows: W (classic) or one per GROUP (fixed-base tables: all digit positions of a
                           // group share one bucket set; a group is one MSM of a fused launch set, see MsmGroups)
    uint32_t groups = 1;   // MSMs sharing this launch set (same n, same c; > 1 only with fixed-base tables)
    uint32_t B = 0;        // buckets per bucket window = 2^(c-1)
    uint32_t key_bits = 0; // key bits (library radix-sort path)
    uint32_t discard = 0;  // key of zero digits
    bool precomp = false;  // vals index the table [2^(c w)] P_i at w * win_stride + i
    uint32_t win_stride = 0;
    uint32_t neg_offset = 0; // precomp: the table holds a second, negated half at +neg_offset (sign folded into the index)
    size_t N = 0;          // entries = groups * n * W
    // accumulation levels: level 0 consumes entries, level k>0 consumes the slots of level k-1
    struct Level { size_t items; uint32_t L; size_t threads; };
    std::vector<Level> levels;
    // reduction: B > REDUCE_DIRECT_MAX -> rows x cols split, else bit planes straight from the buckets
    bool rowcol = false;
    uint32_t log_rows = 0, log_cols = 0;
    uint32_t bits_c = 0, bits_r = 0;  // bit planes of the column / row weights
    uint32_t out_per_window = 0;      // bits_c + bits_r records per bucket window
    // batched-affine pre-reduction (msm_affine.cuh): rounds of pairwise additions before the XYZZ accumulation;
    // bound[r] = upper bound on the list length after r rounds (bound[0] = N), acc_items = bound[affine_rounds]
    uint32_t affine_rounds = 0;
    size_t bound[8] = {0};
    size_t acc_items = 0;
};
constexpr uint32_t AFFINE_MAX_ROUNDS = 6;
constexpr uint32_t REDUCE_DIRECT_MAX = 1024;
// tunables (compile-time so that experiments are separate builds): resident CTAs per SM the level-0
// accumulation kernel is compiled for, and the slice length of the first slot level
#ifndef ZKP_ACC_MIN_BLOCKS
#define ZKP_ACC_MIN_BLOCKS 3
#endif
#ifndef ZKP_ACC_THREADS
#define ZKP_ACC_THREADS 128
#endif
#ifndef ZKP_SLOT_L1
#define ZKP_SLOT_L1 8
#endif
#ifndef ZKP_FUSE_MAX_LOG
// A SINGLE commit+open runs as one grouped launch set (capi_rest.cuh commit_open_fused) up to this row length.  Measured
// on B200 (tools/fuse_sweep.py, profiles/r2_fuse_sweep.txt): the two-lane form already hides the tail of the first MSM
// under the accumulation of the second, while the fused form puts the opening's field kernels in front of the shared
// sort -- 1.77 vs 1.53 ms at 2^16 -- so a lone request never fuses by default; grouping pays for BATCHES
// (zkp_worker_commit_open_batch), where 2k tails collapse into one.  zkp_set_fuse(ctx, 1) forces it.
#define ZKP_FUSE_MAX_LOG 0
#endif
#ifndef ZKP_SLOT_LN
#define ZKP_SLOT_LN 8   // slice length of the slot levels after the first
#endif

inline uint32_t ilog2_floor(uint32_t n) {
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    return lg;
}
inline uint32_t msm_window_bits(uint32_t n, bool precomp) {
    int lg = (int)ilog2_floor(n);
    // classic: buckets ~ n/32 per window.  fixed-base: one shared bucket set of ~n/2, ~2W entries per bucket
    int c = precomp ? lg : lg - 4;
    if (c < 4) c = 4;
    if (c > (precomp ? 22 : 16)) c = precomp ? 22 : 16;
    if (precomp && lg >= 18) {
        // Large fixed-base MSMs: W = floor(255/c) + 1 only changes at a few widths (13 for c = 20 and 21, 12 for
        // 22 and 23), so the width is chosen by cost (tools/window_sweep.py, B200): n W additions at 0.36 ns, plus
        // a reduction that is latency-bound (~1.1 ms) up to 2^19 buckets and 2.1 ns per bucket beyond.  E.g.
        // n = 2^21 takes c = 20, not 21 (same W, half the buckets: 11.96 vs 13.09 ms); 2^18 and 2^19 measured
        // best at c = 16.
        if (lg <= 19) return 16;
        double best = 1e300;
        for (int cc = lg - 3; cc <= 24; cc++) {
            double tail = 2.1 * (double)(1ull << (cc - 1));
            if (tail < 1.1e6) tail = 1.1e6;
            double cost = 0.36 * (double)n * (255 / cc + 1) + tail;
            if (cost < best * 0.999) { best = cost; c = cc; }
        }
    }
    return (uint32_t)c;
}

inline MsmPlan msm_make_plan(uint32_t n, int sm_count, uint32_t c, bool precomp, uint32_t win_stride, uint32_t affine_rounds = 0,
                             uint32_t groups = 1) {
    MsmPlan p;
    p.n = n;
    p.c = c;
    p.W = 255 / p.c + 1;
    p.precomp = precomp;
    p.win_stride = win_stride;
    p.groups = precomp ? groups : 1;
    p.Wb = precomp ? p.groups : p.W;
    p.B = 1u << (p.c - 1);
    p.discard = p.Wb * p.B;
    p.key_bits = 1;
    while ((1ull << p.key_bits) <= p.discard) p.key_bits++;
    p.N = (size_t)n * p.W * p.groups;
    p.affine_rounds = affine_rounds > AFFINE_MAX_ROUNDS ? AFFINE_MAX_ROUNDS : affine_rounds;
    p.bound[0] = p.N;
    for (uint32_t r = 0; r < p.affine_rounds; r++) p.bound[r + 1] = (p.bound[r] + p.discard + 1) / 2;  // sum ceil(len/2) <= (N + buckets)/2
    p.acc_items = p.bound[p.affine_rounds];
    // level 0: slice length chosen so that the grid is a whole number of waves of resident threads
    const size_t resident = (size_t)sm_count * ZKP_ACC_THREADS * ZKP_ACC_MIN_BLOCKS;
    // ~6 waves: long slices mean few slice-boundary partials for the slot levels, and because every
    // thread does the same work the last wave is as full as the first
#ifndef ZKP_L0_TARGET
#define ZKP_L0_TARGET 48
#endif
    static const size_t l0_target = [] {
        const char* e = getenv("ZKP_L0_TARGET");  // tuning runs only (tools/shard_tail.py)
        size_t v = e ? (size_t)strtoull(e, nullptr, 10) : 0;
        return v ? v : (size_t)ZKP_L0_TARGET;
    }();
    size_t waves = (p.acc_items + resident * l0_target - 1) / (resident * l0_target);
    if (waves < 1) waves = 1;
    uint32_t L0 = (uint32_t)((p.acc_items + waves * resident - 1) / (waves * resident));
#ifndef ZKP_MIN_L0
#define ZKP_MIN_L0 8
#endif
    if (L0 < ZKP_MIN_L0) L0 = ZKP_MIN_L0;
    size_t items = p.acc_items;
    uint32_t L = L0;
    for (int lvl = 0;; lvl++) {
        size_t shift = lvl ? 1 : 0;  // slot levels slice on odd indices (see k_accumulate)
        size_t threads = items > shift ? (items - shift + L - 1) / L : 1;
        p.levels.push_back({items, L, threads});
        if (threads <= 1) break;
        items = threads * 2;
        // shallow slot levels: every sequential addition costs ~17 us of latency.  For scalars that are not
        // adversarial almost every slot run is one (tail_t, head_t+1) pair, so the first slot level can be
        // made a single parallel addition per thread
        L = lvl == 0 ? ZKP_SLOT_L1 : ZKP_SLOT_LN;
    }
    // reduction plan
    if (p.B > REDUCE_DIRECT_MAX) {
        p.rowcol = true;
        p.log_cols = (p.c - 1 + 1) / 2;
        p.log_rows = p.c - 1 - p.log_cols;
        p.bits_c = p.log_cols + 1;  // column weights lo + 1 in [1, 2^log_cols]
        p.bits_r = p.log_rows;      // row weights hi in [0, 2^log_rows)
    } else {
        p.bits_c = p.c;             // bucket weights b + 1 in [1, 2^(c-1)]
        p.bits_r = 0;
    }
    p.out_per_window = p.bits_c + p.bits_r;
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
// precomp: every digit position shares the bucket set (key = |digit| - 1) and val indexes the table
// [2^(c w)] P_i at w * win_stride + i.
// MODE 0 (DIGITS_WRITE): keys/vals written in window-major order (entry w n + i) for the library radix sort.
// MODE 1 (DIGITS_COUNT) and 2 (DIGITS_SCATTER): the two passes of the hand-written bucket sort (section 2 below) --
// the digits are recomputed in the second pass instead of being stored unsorted, written once, and read back.
enum { DIGITS_WRITE = 0, DIGITS_COUNT = 1, DIGITS_SCATTER = 2 };
constexpr uint32_t SORT_CHUNK_LOG = 10, SORT_CHUNK = 1u << SORT_CHUNK_LOG;  // keys per chunk of the two-level scan
// Several MSMs of the same length over fixed-base tables can share ONE launch set (the two MSMs of a commit+open; a
// batch of requests): group g has its own scalars and its own table (val_base = first record of the table inside
// the arena), its keys are g * B + |digit| - 1, and everything after the digits -- sort, accumulation, slot levels,
// reduction -- runs once over the union.  blockIdx.y selects the group (g0 + blockIdx.y).
constexpr int MSM_MAX_GROUPS = 64;
struct MsmGroups {
    const uint32_t* scalars[MSM_MAX_GROUPS];
    uint32_t val_base[MSM_MAX_GROUPS];
    uint8_t fmt[MSM_MAX_GROUPS];
};
template <int MODE>
__global__ void k_decompose(const __grid_constant__ MsmGroups gs, uint32_t g0, uint32_t n, uint32_t c, uint32_t W,
                            uint32_t B, uint32_t discard, int precomp, uint32_t win_stride, uint32_t neg_offset,
                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ bad,
                            uint32_t* __restrict__ counters, const uint32_t* __restrict__ chunk_base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = g0 + blockIdx.y;
    const uint32_t* __restrict__ scalars = gs.scalars[g];
    const int fmt = gs.fmt[g];
    const uint32_t val_base = gs.val_base[g];
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
    } else if (MODE != DIGITS_SCATTER) {
        bool lt = false, decided = false;
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            uint32_t mk = FrParams::mod(k);
            if (!decided && s[k] != mk) { decided = true; lt = s[k] < mk; }
        }
        if (!lt) atomicOr(bad + g, 1u);
    }
    s[8] = 0;
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
    const uint32_t lane = threadIdx.x & 31, lanes_below = (1u << lane) - 1;
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
        const uint32_t key = mag ? (precomp ? g : w) * B + mag - 1 : discard;
        // with a negated table half the sign selects the half and no kernel ever negates a y-coordinate
        const uint32_t val = neg_offset ? val_base + w * win_stride + i + (neg ? neg_offset : 0u)
                                        : ((precomp ? val_base + w * win_stride + i : i) | (neg << 31));
        if (MODE == DIGITS_WRITE) {
            size_t o = ((size_t)g * W + w) * n + i;
            keys[o] = key;
            vals[o] = val;
        } else {
            // lanes of the warp that hold the same key act as one: a constant polynomial puts all 32 lanes (and every
            // warp of the grid) on ONE counter per window, and must not turn into n serialised atomics
            // (measured: no aggregation makes a lone 2^20 MSM 0.05 ms faster on uniform scalars and 42% slower on a
            // constant polynomial; aggregating only the all-equal warp loses 25% on a two-valued polynomial; issuing
            // the matches and atomics of 8 digit positions back to back before using any result is SLOWER -- count
            // 112 -> 119 us, scatter 238 -> 290 us: the passes are bound by L2 transaction throughput, not by the
            // latency a thread sees)
            const uint32_t peers = __match_any_sync(__activemask(), key);
            const uint32_t leader = __ffs(peers) - 1, rank = __popc(peers & lanes_below);
            if (MODE == DIGITS_COUNT) {
                if (lane == leader) atomicAdd(counters + key, (uint32_t)__popc(peers));
            } else {
                uint32_t first = 0;
                if (lane == leader) first = atomicAdd(counters + key, (uint32_t)__popc(peers));
                first = __shfl_sync(peers, first, leader);
                // one 8-byte store per entry: the scatter is bound by the number of L2 transactions (an atomic and the
                // stores of an entry go to unrelated sectors), not by bytes
                const size_t o = (size_t)chunk_base[key >> SORT_CHUNK_LOG] + first + rank;
                reinterpret_cast<uint2*>(keys)[o] = make_uint2(key, val);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 2. bucket sort (hand-written; the library radix sort stays selectable for A/B runs)
// ------------------------------------------------------------------------------------------------
// The accumulation needs the entries GROUPED by key, in key order; the order inside a group is irrelevant (a
// bucket is a sum).  That is a counting sort: k_decompose<COUNT> histograms the keys, two small kernels turn the
// histogram into start positions (an exclusive scan kept as "position inside a chunk of 1024 keys" + "start of the
// chunk", so that no third pass has to add the two), k_decompose<SCATTER> recomputes the digits and places every
// entry with one atomicAdd on its key's cursor.  Traffic: the 32-byte scalars twice and the sorted (key, value)
// arrays once, against four reads and three writes of the 8-byte pairs for a 3-pass LSD radix sort.
// counters[chunk * 1024 ..] -> exclusive prefix inside the chunk (in place), chunk_sum[chunk] = total of the chunk
__global__ void __launch_bounds__(SORT_CHUNK) k_sort_scan_chunks(uint32_t* __restrict__ counters, uint32_t m, uint32_t* __restrict__ chunk_sum) {
    __shared__ uint32_t warp_tot[32];
    const uint32_t idx = blockIdx.x * SORT_CHUNK + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = idx < m ? counters[idx] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, ti, d);
            if ((int)lane >= d) ti += u;
        }
        warp_tot[lane] = ti - t;  // exclusive
        if (lane == 31) chunk_sum[blockIdx.x] = ti;
    }
    __syncthreads();
    if (idx < m) counters[idx] = warp_tot[warp] + incl - v;
}
// chunk_sum[0..chunks) -> exclusive prefix in place (one block; chunks <= a few thousand)
__global__ void __launch_bounds__(1024) k_sort_scan_sums(uint32_t* __restrict__ chunk_sum, uint32_t chunks) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t running;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < chunks; base += 1024) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < chunks ? chunk_sum[idx] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t block_total = 0;
        if (warp == 0) {
            uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, ti, d);
                if ((int)lane >= d) ti += u;
            }
            warp_tot[lane] = ti - t;
            block_total = __shfl_sync(0xffffffffu, ti, 31);
        }
        __syncthreads();
        if (idx < chunks) chunk_sum[idx] = running + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) running += block_total;
        __syncthreads();
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

// a second read of the same record that the compiler cannot fold into the first one
__device__ __forceinline__ G1Affine load_affine_again(const G1Affine* src) {
    G1Affine p;
    uint32_t* d = p.x.v;
#pragma unroll
    for (int i = 0; i < 6; i++)
        asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(d[4 * i]), "=r"(d[4 * i + 1]), "=r"(d[4 * i + 2]), "=r"(d[4 * i + 3])
                     : "l"(reinterpret_cast<const uint4*>(src) + i));
    return p;
}

// LEVEL0: items are (key, point index|sign) entries, points gathered from the affine SRS row.
// else  : items are (key|flags, XYZZ) slots written by the previous level.
template <bool LEVEL0, bool COOP = false>
__global__ void __launch_bounds__(LEVEL0 ? ZKP_ACC_THREADS : 128, LEVEL0 ? ZKP_ACC_MIN_BLOCKS : 1)
k_accumulate(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
             const G1Affine* __restrict__ points, const G1Xyzz* __restrict__ slots_in, size_t items,
             uint32_t L, uint32_t discard, G1Xyzz* __restrict__ buckets, uint32_t* __restrict__ slot_keys,
             G1Xyzz* __restrict__ slot_pts, int last_level, uint32_t S = 1, uint32_t pstride = sizeof(G1Affine)) {
    // pstride: bytes between consecutive records of `points` -- 96 for an SRS row or a list of the batched-affine
    // rounds, TABLE_STRIDE (128) for a fixed-base table, whose records are padded so that a gather touches ONE
    // 128-byte line instead of 1.75 on average
    // S: stride of the level-0 entry arrays in words -- 1 for separate key / value arrays (library sort, batched-affine
    // output), 2 for the interleaved (key, value) pairs written by the bucket sort (vals == keys + 1)
    // Level 0: one thread per slice.  Small slot levels (COOP): FOUR lanes per slice -- they run the same control flow
    // on the same slice and share every point addition (coop_add4), because these levels are pure latency: a few
    // dependent additions per slice and far fewer slices than the machine has lanes.  A slot level with more slices
    // than resident lanes is throughput-bound and keeps one thread per slice (the shared addition issues 1.37x the
    // multiplies).
    static_assert(!(LEVEL0 && COOP), "level 0 is one thread per slice");
    constexpr int LANES = COOP ? 4 : 1;
    // level 0 adds with lazy reductions (G1Xyzz::madd_lazy); its correction table k p, k < 8, lives in shared memory
    __shared__ uint32_t s_kp[LEVEL0 ? FQ_KP_ROWS * 12 : 1];
    if (LEVEL0) {
        for (uint32_t i = threadIdx.x; i < FQ_KP_ROWS * 12; i += blockDim.x) s_kp[i] = fq_kp_limb(i / 12, i % 12);
        __syncthreads();
    }
    size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int lane_id = threadIdx.x & 31, gl = lane_id & (LANES - 1), gbase = lane_id - gl;
    const uint32_t gmask = COOP ? (0xfu << gbase) : 0u;
    // Slot lists hold (head_k, tail_k) pairs and the runs worth merging join tail_k with head_{k+1}
    // (odd index, even index): slices of the slot levels therefore start on ODD indices so that a
    // slice boundary falls between head_k and tail_k, never inside such a pair.
    const size_t shift = LEVEL0 ? 0 : 1;
    size_t start = t * L + (t ? shift : 0);
    if (start >= items) return;
    size_t end = (t + 1) * L + shift < items ? (t + 1) * L + shift : items;

    auto key_at = [&](size_t i) -> uint32_t {
        uint32_t k = keys[LEVEL0 ? i * S : i];
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
        if (LEVEL0) acc.normalize();  // lazy coordinates -> canonical
        bool starts_before = is_first_run && cur == prev_key;
        if (last_level || (!starts_before && !continues_after)) {
            if (!acc.is_inf() && gl == 0) store_xyzz(buckets + cur, acc);
        } else if (is_first_run) {
            if (gl == 0) store_xyzz(slot_pts + 2 * t, acc);
            have_head = true;
        } else {
            if (gl == 0) store_xyzz(slot_pts + 2 * t + 1, acc);
            have_tail = true;
        }
    };

    for (size_t i = start; i < end; i++) {
        uint32_t raw = keys[LEVEL0 ? i * S : i];
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
            // vals == nullptr: the items ARE the points (output of the batched-affine rounds), in list order
            uint32_t v = vals ? vals[i * S] : (uint32_t)i;
            if (vals && i + 1 < end) {
                // the gather is a 96-byte random read of a table far larger than L2: pull the NEXT point
                // towards L2 while this one is being added (no registers held, unlike a software pipeline)
                const char* nx = reinterpret_cast<const char*>(points) + (size_t)(vals[(i + 1) * S] & 0x7fffffffu) * pstride;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 64));
            }
            const G1Affine* pp = reinterpret_cast<const G1Affine*>(reinterpret_cast<const char*>(points) + (size_t)(v & 0x7fffffffu) * pstride);
            G1Affine p = load_affine(pp);
            if (v >> 31) p.y = p.y.neg();  // never taken with a negated table half (-(0, 0) stays the infinity marker)
            if (!p.is_inf()) {
                const int st = acc.madd_lazy_core(p.x, p.y, s_kp);
                if (st) {  // equal or opposite operands: fetch the point again rather than keep it live
                    p = load_affine_again(pp);
                    if (v >> 31) p.y = p.y.neg();
                    acc.madd_lazy_rare(st, p.x, p.y);
                }
            }
        } else if (!(raw & KEY_EMPTY_FLAG)) {
            G1Xyzz p = load_xyzz(slots_in + i);
            if (COOP) coop_add4(acc, p, gmask, gbase, gl);
            else acc.add(p);
        }
    }
    if (cur != KEY_NONE) flush(next_key == cur);
    if (!last_level && gl == 0) {
        // every slice with at least one item publishes both keys so the next level can detect
        // run boundaries by looking at adjacent slots only
        slot_keys[2 * t] = first_key == KEY_NONE ? KEY_NONE : (have_head ? first_key : (first_key | KEY_EMPTY_FLAG));
        slot_keys[2 * t + 1] = last_key == KEY_NONE ? KEY_NONE : (have_tail ? last_key : (last_key | KEY_EMPTY_FLAG));
    }
}

// ------------------------------------------------------------------------------------------------
// 4. bucket reduction
// ------------------------------------------------------------------------------------------------
constexpr int RC_THREADS = 128;
constexpr int RC_LANES = 32;                       // threads cooperating on one sum
constexpr int RC_SUMS = RC_THREADS / RC_LANES;     // sums per block
// Bucket array of one window viewed as 2^log_rows x 2^log_cols (bucket b = hi * cols + lo).
// Sum s < cols is the column sum C_s = sum_hi X[hi][s] (strided reads), sum s >= cols the row sum
// R_(s-cols) = sum_lo X[s-cols][lo] (contiguous reads).  One warp per sum: every lane adds a strided
// share sequentially (all lanes busy), then 5 shared-memory tree steps.  grid.y = bucket window.
__global__ void __launch_bounds__(RC_THREADS)
k_rowcol_sums(const G1Xyzz* __restrict__ in, uint32_t log_rows, uint32_t log_cols, G1Xyzz* __restrict__ out_c,
              G1Xyzz* __restrict__ out_r) {
    __shared__ G1Xyzz sh[RC_THREADS];
    const uint32_t rows = 1u << log_rows, cols = 1u << log_cols, w = blockIdx.y;
    const uint32_t lane = threadIdx.x % RC_LANES;
    const uint32_t s = blockIdx.x * RC_SUMS + threadIdx.x / RC_LANES;
    const G1Xyzz* x = in + ((size_t)w << (log_rows + log_cols));
    G1Xyzz acc = G1Xyzz::infinity();
    if (s < cols) {
        for (uint32_t hi = lane; hi < rows; hi += RC_LANES) {
            G1Xyzz p = load_xyzz(x + ((size_t)hi << log_cols) + s);
            acc.add(p);
        }
    } else if (s < cols + rows) {
        const uint32_t hi = s - cols;
        for (uint32_t lo = lane; lo < cols; lo += RC_LANES) {
            G1Xyzz p = load_xyzz(x + ((size_t)hi << log_cols) + lo);
            acc.add(p);
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int st = RC_LANES / 2; st > 0; st >>= 1) {
        if ((int)lane < st) {
            G1Xyzz a = sh[threadIdx.x];
            a.add(sh[threadIdx.x + st]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (lane == 0 && s < cols + rows)
        store_xyzz(s < cols ? out_c + (size_t)w * cols + s : out_r + (size_t)w * rows + (s - cols), sh[threadIdx.x]);
}

// Balanced two-stage form of the same sums (used when the bucket array is large).  Stage 1: every sum is cut
// into Q interleaved shares (share k takes elements k, k + Q, ...), ONE THREAD per share, so that all threads
// do the same number of sequential additions and the whole stage is a single wave of resident CTAs (the
// one-warp-per-sum kernel above gave row sums twice the depth of column sums and left SMs with 2 or 3 CTAs:
// measured 1.0 ms at 2^19 buckets against 0.47 ms of multiply-issue time).  Stage 2: one warp per sum adds its
// Q partials (<= 2 per lane) and finishes with 5 tree steps.
// share index space per window: [0, cols * q_c) column shares, then rows * q_r row shares.
__global__ void __launch_bounds__(128, 3)
k_rowcol_partial(const G1Xyzz* __restrict__ in, uint32_t log_rows, uint32_t log_cols, uint32_t q_c, uint32_t q_r,
                 G1Xyzz* __restrict__ part) {
    const uint32_t rows = 1u << log_rows, cols = 1u << log_cols, w = blockIdx.y;
    const uint32_t n_c = cols * q_c, total = n_c + rows * q_r;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const G1Xyzz* x = in + ((size_t)w << (log_rows + log_cols));
    G1Xyzz acc = G1Xyzz::infinity();
    if (t < n_c) {
        // adjacent threads -> adjacent columns (contiguous 192-byte records), share k = t / cols
        const uint32_t s = t & (cols - 1), k = t >> log_cols;
        for (uint32_t hi = k; hi < rows; hi += q_c) {
            G1Xyzz p = load_xyzz(x + ((size_t)hi << log_cols) + s);
            acc.add(p);
        }
    } else {
        // row hi = u / q_r, share k = u % q_r takes lo = k, k + q_r, ... (adjacent threads -> adjacent records)
        const uint32_t u = t - n_c, hi = u / q_r, k = u - hi * q_r;
        for (uint32_t lo = k; lo < cols; lo += q_r) {
            G1Xyzz p = load_xyzz(x + ((size_t)hi << log_cols) + lo);
            acc.add(p);
        }
    }
    store_xyzz(part + (size_t)w * total + t, acc);
}
// One warp per sum, as 8 groups of 4 lanes sharing every addition (coop_add4): group j adds the partials
// j, j + 8, ... one after the other, then three pairwise steps between groups.
__global__ void __launch_bounds__(RC_THREADS)
k_rowcol_finish(const G1Xyzz* __restrict__ part, uint32_t log_rows, uint32_t log_cols, uint32_t q_c, uint32_t q_r,
                G1Xyzz* __restrict__ out_c, G1Xyzz* __restrict__ out_r) {
    const uint32_t rows = 1u << log_rows, cols = 1u << log_cols, w = blockIdx.y;
    const uint32_t n_c = cols * q_c, total = n_c + rows * q_r;
    const int lane = threadIdx.x & 31, gl = lane & 3, gbase = lane - gl, grp = lane >> 2;
    const uint32_t gmask = 0xfu << gbase;
    const uint32_t s = blockIdx.x * RC_SUMS + threadIdx.x / 32;
    if (s >= cols + rows) return;  // whole warp
    const G1Xyzz* p = part + (size_t)w * total;
    const bool is_col = s < cols;
    const uint32_t q = is_col ? q_c : q_r;
    G1Xyzz acc = G1Xyzz::infinity();
    for (uint32_t k = grp; k < q; k += 8) {
        G1Xyzz v = load_xyzz(is_col ? p + (size_t)k * cols + s : p + n_c + (size_t)(s - cols) * q_r + k);
        coop_add4(acc, v, gmask, gbase, gl);
    }
    for (int st = 4; st > 0; st >>= 1) {
        G1Xyzz o;
        uint32_t* ov = o.x.v;
        const uint32_t* av = acc.x.v;
#pragma unroll
        for (int i = 0; i < 48; i++) ov[i] = __shfl_down_sync(0xffffffffu, av[i], 4 * st);
        if (grp < st) coop_add4(acc, o, gmask, gbase, gl);
    }
    if (lane == 0) store_xyzz(is_col ? out_c + (size_t)w * cols + s : out_r + (size_t)w * rows + (s - cols), acc);
}

// The same two stages for SMALL bucket arrays (one request at the mainnet row size: 2^15 buckets), where both are pure
// latency -- a lone lane needs ~17 us per dependent addition, four lanes sharing it (coop_add4) ~7.5 us:
//   k_rowcol_partial_coop   share indexing of k_rowcol_partial, four lanes per share;
//   k_rowcol_finish_wide    TWO warps per sum (16 groups of 4 lanes): a sum of q partials is q/16 shared additions per
//                           group, three steps inside each warp, one between the two warps.
// Measured at 2^16 (B200): the pair 78 + 82 us -> see profiles/r2_rowcol_coop_ab.txt.
__global__ void __launch_bounds__(128, 3)
k_rowcol_partial_coop(const G1Xyzz* __restrict__ in, uint32_t log_rows, uint32_t log_cols, uint32_t q_c, uint32_t q_r,
                      G1Xyzz* __restrict__ part) {
    const uint32_t rows = 1u << log_rows, cols = 1u << log_cols, w = blockIdx.y;
    const uint32_t n_c = cols * q_c, total = n_c + rows * q_r;
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int lane = threadIdx.x & 31, gl = lane & 3, gbase = lane - gl;
    const uint32_t gmask = 0xfu << gbase;
    if (t >= total) return;  // whole groups
    const G1Xyzz* x = in + ((size_t)w << (log_rows + log_cols));
    G1Xyzz acc = G1Xyzz::infinity();
    if (t < n_c) {
        const uint32_t s = t & (cols - 1), k = t >> log_cols;
        for (uint32_t hi = k; hi < rows; hi += q_c) {
            G1Xyzz p = load_xyzz(x + ((size_t)hi << log_cols) + s);
            coop_add4(acc, p, gmask, gbase, gl);
        }
    } else {
        const uint32_t u = t - n_c, hi = u / q_r, k = u - hi * q_r;
        for (uint32_t lo = k; lo < cols; lo += q_r) {
            G1Xyzz p = load_xyzz(x + ((size_t)hi << log_cols) + lo);
            coop_add4(acc, p, gmask, gbase, gl);
        }
    }
    if (gl == 0) store_xyzz(part + (size_t)w * total + t, acc);
}
constexpr int RCW_THREADS = 128, RCW_WARPS = 2, RCW_SUMS = RCW_THREADS / (32 * RCW_WARPS);  // 2 sums per block
__global__ void __launch_bounds__(RCW_THREADS)
k_rowcol_finish_wide(const G1Xyzz* __restrict__ part, uint32_t log_rows, uint32_t log_cols, uint32_t q_c, uint32_t q_r,
                     G1Xyzz* __restrict__ out_c, G1Xyzz* __restrict__ out_r) {
    __shared__ G1Xyzz sh[RCW_SUMS];
    const uint32_t rows = 1u << log_rows, cols = 1u << log_cols, w = blockIdx.y;
    const uint32_t n_c = cols * q_c, total = n_c + rows * q_r;
    const int lane = threadIdx.x & 31, gl = lane & 3, gbase = lane - gl, warp = threadIdx.x >> 5;
    const int sub = warp % RCW_WARPS, local = warp / RCW_WARPS;  // which half of the sum, which sum of the block
    const uint32_t grp = (uint32_t)(lane >> 2) + 8u * (uint32_t)sub;  // 0..15
    const uint32_t gmask = 0xfu << gbase;
    const uint32_t s = blockIdx.x * RCW_SUMS + local;
    const bool live = s < cols + rows;  // uniform per warp pair; no early return: the block meets at the barrier
    const G1Xyzz* p = part + (size_t)w * total;
    const bool is_col = s < cols;
    const uint32_t q = is_col ? q_c : q_r;
    G1Xyzz acc = G1Xyzz::infinity();
    if (live) {
        for (uint32_t k = grp; k < q; k += 8 * RCW_WARPS) {
            G1Xyzz v = load_xyzz(is_col ? p + (size_t)k * cols + s : p + n_c + (size_t)(s - cols) * q_r + k);
            coop_add4(acc, v, gmask, gbase, gl);
        }
        for (int st = 4; st > 0; st >>= 1) {
            G1Xyzz o;
            uint32_t* ov = o.x.v;
            const uint32_t* av = acc.x.v;
#pragma unroll
            for (int i = 0; i < 48; i++) ov[i] = __shfl_down_sync(0xffffffffu, av[i], 4 * st);
            if ((lane >> 2) < st) coop_add4(acc, o, gmask, gbase, gl);
        }
    }
    if (live && sub == 1 && lane == 0) sh[local] = acc;
    __syncthreads();
    if (live && sub == 0 && lane < 4) {
        G1Xyzz o = sh[local];
        coop_add4(acc, o, gmask, gbase, gl);
        if (lane == 0) store_xyzz(is_col ? out_c + (size_t)w * cols + s : out_r + (size_t)w * rows + (s - cols), acc);
    }
}

// Bit planes: block (j, w) computes P[w][j] = sum of the inputs whose weight has bit j set.  Planes
// j < bits_a come from array a (n_a inputs per window, weight k + 1), planes j >= bits_a from array b
// (n_b inputs per window, weight k).  The host finishes with Horner passes (sum_j 2^j P_j).
constexpr int TAIL_THREADS = 256;
// 8 warps x 8 groups of 4 lanes (coop_add4): group g adds the selected inputs g, g + 64, ... one after the other,
// three pairwise steps inside the warp, the 8 warp results through shared memory, three more steps in warp 0.
__global__ void __launch_bounds__(TAIL_THREADS)
k_bit_sums(const G1Xyzz* __restrict__ a, uint32_t n_a, uint32_t bits_a, const G1Xyzz* __restrict__ b, uint32_t n_b,
           G1Xyzz* __restrict__ out, uint32_t out_stride) {
    __shared__ G1Xyzz sh[TAIL_THREADS / 32];
    const uint32_t w = blockIdx.y;
    const bool first = blockIdx.x < bits_a;
    const uint32_t j = first ? blockIdx.x : blockIdx.x - bits_a;
    const uint32_t n_in = first ? n_a : n_b, one = first ? 1u : 0u;
    const G1Xyzz* in = (first ? a : b) + (size_t)w * n_in;
    const int lane = threadIdx.x & 31, gl = lane & 3, gbase = lane - gl, warp = threadIdx.x >> 5;
    const uint32_t gmask = 0xfu << gbase, grp_global = threadIdx.x >> 2, grp = lane >> 2;
    constexpr uint32_t GROUPS = TAIL_THREADS / 4;
    G1Xyzz acc = G1Xyzz::infinity();
    for (uint32_t k = grp_global; k < n_in; k += GROUPS) {
        if (((k + one) >> j) & 1) {  // uniform inside a group
            G1Xyzz p = load_xyzz(in + k);
            coop_add4(acc, p, gmask, gbase, gl);
        }
    }
    auto fold_warp = [&]() {
        for (int st = 4; st > 0; st >>= 1) {
            G1Xyzz o;
            uint32_t* ov = o.x.v;
            const uint32_t* av = acc.x.v;
#pragma unroll
            for (int i = 0; i < 48; i++) ov[i] = __shfl_down_sync(0xffffffffu, av[i], 4 * st);
            if (grp < st) coop_add4(acc, o, gmask, gbase, gl);
        }
    };
    fold_warp();
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = grp < TAIL_THREADS / 32 ? sh[grp] : G1Xyzz::infinity();
        fold_warp();
        if (lane == 0) store_xyzz(out + (size_t)w * out_stride + blockIdx.x, acc);
    }
}

}  // namespace zkp
