// Batched-affine pre-reduction of the sorted digit list (the front half of the bucket accumulation for large
// MSMs).
//
// The XYZZ accumulation of msm.cuh pays 10 Fq products per entry.  An AFFINE addition costs 3 products plus one
// field inversion; with Montgomery's trick a thread shares ONE inversion between all the additions of a batch
// (3 more products per addition), i.e. 6 products per addition -- provided the additions of a batch are
// independent.  Entries of one bucket are not (they form a chain), but pairs of them are: round r adds the
// entries of every bucket two by two, so a bucket of m entries shrinks to ceil(m/2), and every addition of a
// round is independent of every other.  After R rounds (R = 3: buckets of ~26 entries shrink to ~4) the
// remaining list -- (bucket, affine point) items, still grouped by bucket -- goes through the balanced XYZZ
// accumulation unchanged (k_accumulate<level0> with the round output as its point table).
//
// Layout of a round: the list of round r is described only by start_r[b] (first position of bucket b; B + 1
// entries), because the order inside a bucket never matters.  len_{r+1}[b] = ceil(len_r[b] / 2), start_{r+1} =
// exclusive scan.  Output position o of bucket b, j = o - start_{r+1}[b], is the sum of inputs
// start_r[b] + 2j and + 2j + 1 (or a copy of the first when the bucket length is odd and j is last).  Thread t
// owns outputs [t K, (t + 1) K): the same number of additions for every thread whatever the scalar
// distribution (a constant polynomial -- every digit in one bucket -- is just a long bucket).
//
// Exactness: P + Q with P = Q (doubling), P = -Q (result infinity) and infinity operands ((0,0)) are all
// handled; such pairs contribute a unit denominator to the batch.
#pragma once
#include "fq_inv.cuh"
#include "msm.cuh"

namespace zkp {

#ifndef ZKP_AFF_KB
#define ZKP_AFF_KB 64
#endif
#ifndef ZKP_AFF_MIN_BLOCKS
#define ZKP_AFF_MIN_BLOCKS 4
#endif
static_assert(ZKP_AFF_KB <= 64, "the rare-pair mask of a batch is one 64-bit word");
constexpr int AFF_KB = ZKP_AFF_KB;  // additions sharing one inversion (per thread)

// start[b] = first position i with keys[i] >= b, for b in [0, nb]; keys sorted, keys >= nb are discards
__global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, size_t n, uint32_t nb, uint32_t* __restrict__ start) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    // buckets in (prev, cur] start at i
    int64_t prev = i > 0 ? (int64_t)min(keys[i - 1], nb) : -1;
    int64_t cur = i < n ? (int64_t)min(keys[i], nb) : (int64_t)nb;
    for (int64_t b = prev + 1; b <= cur; b++) start[b] = (uint32_t)i;
}
// len_next[b] = ceil(len[b] / 2) for b < nb, 0 for b == nb (the exclusive scan of this array is start_next)
__global__ void k_next_len(const uint32_t* __restrict__ start, uint32_t nb, uint32_t* __restrict__ len_next) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    len_next[b] = b < nb ? (start[b + 1] - start[b] + 1) / 2 : 0;
}

__device__ __forceinline__ void store_affine(G1Affine* dst, const G1Affine& p) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint32_t* s = p.x.v;
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
__device__ __forceinline__ Fq load_fq(const Fq* src) {
    Fq r;
    const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 t = s[i];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
__device__ __forceinline__ void store_fq(Fq* dst, const Fq& a) {
    uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

enum { AFF_ADD = 0, AFF_DBL = 1, AFF_TAKE_P = 2, AFF_TAKE_Q = 3, AFF_INF = 4 };

// what P + Q needs: the kind of operation and the denominator that has to be inverted (1 when none)
__device__ __forceinline__ int affine_classify(const G1Affine& p, const G1Affine& q, Fq& den) {
    den = Fq::one();
    if (q.is_inf()) return AFF_TAKE_P;
    if (p.is_inf()) return AFF_TAKE_Q;
    Fq dx = q.x - p.x;
    if (!dx.is_zero()) {
        den = dx;
        return AFF_ADD;
    }
    if (q.y == p.y && !p.y.is_zero()) {
        den = p.y.dbl();
        return AFF_DBL;
    }
    return AFF_INF;
}

__device__ __forceinline__ Fq load_fq_ldg(const Fq* src) {
    Fq r;
    const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 t = __ldg(s + i);
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
// One round of pairwise additions.  FIRST: inputs are the (vals, fixed-base table) entries of the sorted digit
// list (point index | sign << 31; with a negated table half the sign is folded into the index and the bit is
// never set); else affine points of the previous round, in list order.
// Per batch of AFF_KB outputs a thread (1) walks the bucket bounds to find its input pairs, (2) forward pass:
// x-coordinates only, denominators, running product (prefix products parked in `scratch`), (3) one inversion,
// (4) backward pass: both coordinates, finish every addition.
// Measured dead ends (B200, 2^20, kept out of the code): an L2 prefetch AFF_AHEAD items ahead in both passes
// (+0.8 ms per MSM), cp.async staging of the next operands in shared memory (+0.85 ms), 5-6 CTAs/SM with
// spills (+0.5 ms).  The kernel is bound by instruction issue at 16 warps/SM, not by the gathers.
template <bool FIRST>
__global__ void __launch_bounds__(128, ZKP_AFF_MIN_BLOCKS)
k_affine_round(const uint32_t* __restrict__ start_in, const uint32_t* __restrict__ start_out, uint32_t nb,
               const uint32_t* __restrict__ vals, const G1Affine* __restrict__ in_pts, G1Affine* __restrict__ out_pts,
               uint32_t* __restrict__ out_keys, uint32_t K, Fq* __restrict__ scratch, uint32_t total_threads, uint32_t sm_count,
               uint32_t first_stride = sizeof(G1Affine)) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_out = start_out[nb];
    const uint64_t o_begin64 = (uint64_t)t * K;
    if (o_begin64 >= n_out) return;
    const uint32_t o_begin = (uint32_t)o_begin64;
    const uint32_t o_end = o_begin64 + K < n_out ? o_begin + K : n_out;

    // bucket of the first output: largest b with start_out[b] <= o_begin
    uint32_t b;
    {
        uint32_t lo = 0, hi = nb;
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo + 1) / 2;
            if (start_out[mid] <= o_begin) lo = mid;
            else hi = mid - 1;
        }
        b = lo;
    }
    uint32_t out_lo = start_out[b], out_hi = start_out[b + 1];
    uint32_t in_lo = start_in[b], in_len = start_in[b + 1] - in_lo;

    // per output of the batch: index of the first operand (table index | sign << 31 in round 0, list position
    // otherwise) and of the second operand, or PAIR_NONE when the output is a plain copy
    constexpr uint32_t PAIR_NONE = 0xffffffffu;
    uint32_t ia[AFF_KB], ib[AFF_KB];
    // FIRST: records of the SRS row (96 bytes) or of a fixed-base table (TABLE_STRIDE bytes)
    auto point_of = [&](uint32_t v) -> const G1Affine* {
        if (FIRST) return reinterpret_cast<const G1Affine*>(reinterpret_cast<const char*>(in_pts) + (size_t)(v & 0x7fffffffu) * first_stride);
        return in_pts + v;
    };
    auto load_x = [&](uint32_t v) -> Fq { return load_fq_ldg(&point_of(v)->x); };
    auto load_y = [&](uint32_t v) -> Fq {
        Fq y = load_fq_ldg(&point_of(v)->y);
        if (FIRST && (v >> 31)) y = y.neg();  // -0 = 0: an infinity entry (0, 0) stays (0, 0)
        return y;
    };

    // Stagger: the CTAs that share an SM start with first batches of 1/4, 2/4, 3/4 and 4/4 of AFF_KB outputs, so
    // that their inversions (tens of thousands of ALU instructions, no multiplies) do not all fall in the
    // same interval and leave the multiply pipe idle
    const uint32_t phase = (blockIdx.x + blockIdx.x / sm_count) & 3;
    uint32_t batch = (uint32_t)AFF_KB * (phase + 1) / 4;
    for (uint32_t base = o_begin; base < o_end; base += batch, batch = AFF_KB) {
        const uint32_t cnt = o_end - base < batch ? o_end - base : batch;
        // ---- (1) operands of every output of the batch
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t o = base + k;
            while (o >= out_hi) {  // next non-empty bucket
                b++;
                out_lo = out_hi;
                out_hi = start_out[b + 1];
                in_lo = start_in[b];
                in_len = start_in[b + 1] - in_lo;
            }
            const uint32_t j = o - out_lo, i0 = in_lo + 2 * j;
            const bool pair = 2 * j + 1 < in_len;
            ia[k] = FIRST ? vals[i0] : i0;
            ib[k] = pair ? (FIRST ? vals[i0 + 1] : i0 + 1) : PAIR_NONE;
            if (out_keys) out_keys[o] = b;
        }
        // ---- (2) forward: denominators from the x-coordinates, running product.  `rare` marks the pairs that
        //      are not a plain addition of two distinct finite points (doubling, P = -Q, infinity operand)
        Fq run = Fq::one();
        uint64_t rare = 0;
        for (uint32_t k = 0; k < cnt; k++) {
            if (ib[k] == PAIR_NONE) continue;
            store_fq(scratch + (size_t)k * total_threads + t, run);
            const Fq px = load_x(ia[k]), qx = load_x(ib[k]);
            Fq den = qx - px;
            if (den.is_zero() || px.is_zero() || qx.is_zero()) {
                G1Affine p, q;
                p.x = px; q.x = qx;
                p.y = load_y(ia[k]); q.y = load_y(ib[k]);
                affine_classify(p, q, den);
                rare |= 1ull << k;
            }
            run = run * den;
        }
        // ---- (3) one inversion for the whole batch
        Fq inv = fq_inverse(run);
        // ---- (4) backward: peel the denominators off, finish every addition
        for (int k = (int)cnt - 1; k >= 0; k--) {
            const uint32_t o = base + (uint32_t)k;
            G1Affine p;
            p.x = load_x(ia[k]);
            p.y = load_y(ia[k]);
            if (ib[k] == PAIR_NONE) {
                store_affine(out_pts + o, p);
                continue;
            }
            G1Affine q;
            q.x = load_x(ib[k]);
            q.y = load_y(ib[k]);
            const Fq pre = load_fq(scratch + (size_t)k * total_threads + t);
            G1Affine r;
            if (!((rare >> k) & 1)) {
                // the common case: two distinct finite points
                const Fq inv_den = inv * pre;
                inv = inv * (q.x - p.x);
                const Fq lam = (q.y - p.y) * inv_den;
                r.x = lam.sqr() - p.x - q.x;
                r.y = lam * (p.x - r.x) - p.y;
            } else {
                Fq den;
                const int kind = affine_classify(p, q, den);
                const Fq inv_den = inv * pre;
                inv = inv * den;
                if (kind == AFF_ADD) {  // (not reached: an addition is never marked rare)
                    const Fq lam = (q.y - p.y) * inv_den;
                    r.x = lam.sqr() - p.x - q.x;
                    r.y = lam * (p.x - r.x) - p.y;
                } else if (kind == AFF_DBL) {
                    const Fq xx = p.x.sqr();
                    const Fq lam = (xx.dbl() + xx) * inv_den;
                    r.x = lam.sqr() - p.x - q.x;
                    r.y = lam * (p.x - r.x) - p.y;
                } else if (kind == AFF_TAKE_P) {
                    r = p;
                } else if (kind == AFF_TAKE_Q) {
                    r = q;
                } else {
                    r.x = Fq::zero();
                    r.y = Fq::zero();
                }
            }
            store_affine(out_pts + o, r);
        }
    }
}

}  // namespace zkp
