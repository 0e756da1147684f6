// Host driver of the MSM pipeline described in msm.cuh.
#pragma once
#include <cub/device/device_scan.cuh>

#include "context.cuh"

namespace zkp {

inline uint64_t msm_fq_muls(const MsmPlan& p) {
    // algorithmic Fq multiplications: 10 per mixed addition into buckets (n * W of them),
    // 14 per full addition in the reduction (2 per bucket: one row sum + one column sum)
    return 10ull * p.n * p.W + 14ull * 2ull * p.B * p.Wb;
}

// The device pipeline of one MSM on a lane's stream, in two halves so that a commit+open can enqueue the cheap
// front halves (digits + sort) of BOTH lanes before either bucket accumulation: an accumulation grid keeps every
// SM busy for milliseconds and starves whatever small kernels another stream launches after it.
// Front half: workspaces, signed digits grouped by bucket (bucket sort, or the library radix sort -- see below).
// the bucket sort writes interleaved pairs, which the batched-affine rounds (separate key / value arrays) do not read
// Measured on B200 (tools/sort_ab.py, lone MSMs): the two sorts are within 0.3% of each other from 2^14 to 2^23 points
// (<= 109 M entries); at 2^24 (201 M entries, 1.6 GB of scattered pairs, 2 M cursors) the library's radix sort is 4%
// faster (79.3 vs 82.7 ms).  Inside a commit+open, where the sort of one lane runs beside the other lane's kernels, the
// counting sort is the faster one (12.0 vs 12.5 ms at 2^20, 1.49 vs 1.71 ms at 2^16): it is bound by L2 transactions
// and leaves the SMs' registers to whatever else is resident.
#ifndef ZKP_BUCKET_SORT_MAX_ENTRIES
#define ZKP_BUCKET_SORT_MAX_ENTRIES (1ull << 27)
#endif
inline bool msm_uses_bucket_sort(const zkp_ctx* ctx, const MsmPlan& plan) {
    if (plan.affine_rounds != 0 || ctx->bucket_sort == 0) return false;
    return ctx->bucket_sort == 1 || plan.N <= ZKP_BUCKET_SORT_MAX_ENTRIES;
}

inline MsmGroups msm_single_group(const uint32_t* d_scalars, int fmt, uint32_t val_base = 0) {
    MsmGroups gs;
    memset(&gs, 0, sizeof(gs));
    gs.scalars[0] = d_scalars;
    gs.fmt[0] = (uint8_t)fmt;
    gs.val_base[0] = val_base;
    return gs;
}

// Front half in three steps so that a fused commit+open can count the digits of the commitment group while the
// opening's field kernels (which produce the scalars of the proof group) are still running on the other stream:
//   msm_prep_begin   workspaces, cleared counters and flags
//   msm_prep_count   digits of groups [g0, g0 + gcount): histogram (bucket sort) or window-major write (library sort)
//   msm_prep_finish  scan + scatter of ALL groups (bucket sort) or the radix sort
inline int msm_prep_begin(zkp_ctx* ctx, int lane, const MsmPlan& plan) {
    MsmWorkspace& ws = lane ? ctx->ws2 : ctx->ws;
    cudaStream_t st = lane ? ctx->stream2 : ctx->stream;
    const size_t N = plan.N;
    if (msm_uses_bucket_sort(ctx, plan)) {
        ZKP_CUDA(ws.keys_b.ensure(N * 8));  // interleaved (key, value) pairs in bucket order
    } else {
        ZKP_CUDA(ws.keys_a.ensure(N * 4));
        ZKP_CUDA(ws.keys_b.ensure(N * 4));
        ZKP_CUDA(ws.vals_a.ensure(N * 4));
        ZKP_CUDA(ws.vals_b.ensure(N * 4));
    }
    const size_t nb = (size_t)plan.Wb * plan.B;
    ZKP_CUDA(ws.buckets.ensure(nb * sizeof(G1Xyzz)));
    if (ws.slot_keys.size() < plan.levels.size()) {
        ws.slot_keys.resize(plan.levels.size());
        ws.slot_pts.resize(plan.levels.size());
    }
    for (size_t l = 0; l + 1 < plan.levels.size(); l++) {
        ZKP_CUDA(ws.slot_keys[l].ensure(plan.levels[l].threads * 2 * 4));
        ZKP_CUDA(ws.slot_pts[l].ensure(plan.levels[l].threads * 2 * sizeof(G1Xyzz)));
    }
    const size_t out_records = (size_t)plan.Wb * plan.out_per_window;
    if (ws.h_window_cap < out_records) {
        if (ws.h_window) cudaFreeHost(ws.h_window);
        ws.h_window = nullptr;
        ws.h_window_cap = 0;
        ZKP_CUDA(cudaMallocHost(&ws.h_window, sizeof(G1Xyzz) * out_records));
        ws.h_window_cap = out_records;
    }
    if (!ws.h_bad) ZKP_CUDA(cudaMallocHost(&ws.h_bad, 4 * MSM_MAX_GROUPS));
    trace_mark(ctx, lane, st, "msm_begin");
    ZKP_CUDA(ws.bad.ensure(4 * MSM_MAX_GROUPS));
    ZKP_CUDA(cudaMemsetAsync(ws.bad.p, 0, 4 * MSM_MAX_GROUPS, st));
    if (msm_uses_bucket_sort(ctx, plan)) {
        const uint32_t m = plan.discard + 1, chunks = (m + SORT_CHUNK - 1) / SORT_CHUNK;
        ZKP_CUDA(ws.sort_counters.ensure((size_t)chunks * SORT_CHUNK * 4));
        ZKP_CUDA(ws.sort_chunks.ensure((size_t)chunks * 4));
        ZKP_CUDA(cudaMemsetAsync(ws.sort_counters.p, 0, (size_t)chunks * SORT_CHUNK * 4, st));
    }
    return ZKP_OK;
}

inline int msm_prep_count(zkp_ctx* ctx, int lane, const MsmPlan& plan, const MsmGroups& gs, uint32_t g0, uint32_t gcount) {
    MsmWorkspace& ws = lane ? ctx->ws2 : ctx->ws;
    cudaStream_t st = lane ? ctx->stream2 : ctx->stream;
    const dim3 grid((plan.n + 255) / 256, gcount);
    if (!msm_uses_bucket_sort(ctx, plan))
        k_decompose<DIGITS_WRITE><<<grid, 256, 0, st>>>(gs, g0, plan.n, plan.c, plan.W, plan.B, plan.discard, plan.precomp ? 1 : 0,
                                                        plan.win_stride, plan.neg_offset, ws.keys_a.as<uint32_t>(),
                                                        ws.vals_a.as<uint32_t>(), ws.bad.as<uint32_t>(), nullptr, nullptr);
    else
        k_decompose<DIGITS_COUNT><<<grid, 256, 0, st>>>(gs, g0, plan.n, plan.c, plan.W, plan.B, plan.discard, plan.precomp ? 1 : 0,
                                                        plan.win_stride, plan.neg_offset, nullptr, nullptr, ws.bad.as<uint32_t>(),
                                                        ws.sort_counters.as<uint32_t>(), nullptr);
    ctx->launches++;
    ZKP_CUDA(cudaGetLastError());
    return ZKP_OK;
}

inline int msm_prep_finish(zkp_ctx* ctx, int lane, const MsmPlan& plan, const MsmGroups& gs) {
    MsmWorkspace& ws = lane ? ctx->ws2 : ctx->ws;
    cudaStream_t st = lane ? ctx->stream2 : ctx->stream;
    const size_t N = plan.N;
    trace_mark(ctx, lane, st, "decompose");
    if (!msm_uses_bucket_sort(ctx, plan)) {
        // 2. sort by (window, bucket): library radix sort
        size_t temp_bytes = 0;
        ZKP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, ws.keys_a.as<uint32_t>(), ws.keys_b.as<uint32_t>(),
                                                 ws.vals_a.as<uint32_t>(), ws.vals_b.as<uint32_t>(), (int64_t)N, 0,
                                                 (int)plan.key_bits, st));
        ZKP_CUDA(ws.cub_temp.ensure(temp_bytes));
        ZKP_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_temp.p, temp_bytes, ws.keys_a.as<uint32_t>(), ws.keys_b.as<uint32_t>(),
                                                 ws.vals_a.as<uint32_t>(), ws.vals_b.as<uint32_t>(), (int64_t)N, 0,
                                                 (int)plan.key_bits, st));
    } else {
        // 1 + 2. hand-written bucket sort: count (done), scan, scatter (msm.cuh section 2)
        const uint32_t m = plan.discard + 1, chunks = (m + SORT_CHUNK - 1) / SORT_CHUNK;
        k_sort_scan_chunks<<<chunks, SORT_CHUNK, 0, st>>>(ws.sort_counters.as<uint32_t>(), m, ws.sort_chunks.as<uint32_t>());
        k_sort_scan_sums<<<1, 1024, 0, st>>>(ws.sort_chunks.as<uint32_t>(), chunks);
        k_decompose<DIGITS_SCATTER><<<dim3((plan.n + 255) / 256, plan.groups), 256, 0, st>>>(
            gs, 0, plan.n, plan.c, plan.W, plan.B, plan.discard, plan.precomp ? 1 : 0, plan.win_stride, plan.neg_offset,
            ws.keys_b.as<uint32_t>(), nullptr, ws.bad.as<uint32_t>(), ws.sort_counters.as<uint32_t>(), ws.sort_chunks.as<uint32_t>());
        ctx->launches += 3;
    }
    trace_mark(ctx, lane, st, "sort");
    ZKP_CUDA(cudaGetLastError());
    return ZKP_OK;
}

inline int msm_enqueue_prep(zkp_ctx* ctx, int lane, const MsmPlan& plan, const MsmGroups& gs) {
    int rc = msm_prep_begin(ctx, lane, plan);
    if (rc) return rc;
    rc = msm_prep_count(ctx, lane, plan, gs, 0, plan.groups);
    if (rc) return rc;
    return msm_prep_finish(ctx, lane, plan, gs);
}

// Back half: (optional batched-affine rounds,) balanced accumulation, reduction, copy of the bit-plane sums to pinned
// host memory.  d_points: the SRS row (classic) or the fixed-base table of the row (plan.precomp).
inline int msm_enqueue_main(zkp_ctx* ctx, int lane, const MsmPlan& plan, const G1Affine* d_points) {
    MsmWorkspace& ws = lane ? ctx->ws2 : ctx->ws;
    cudaStream_t st = lane ? ctx->stream2 : ctx->stream;
    cudaEvent_t ev0 = lane ? ctx->ev_acc2_0 : ctx->ev_acc0, ev1 = lane ? ctx->ev_acc2_1 : ctx->ev_acc1;
    const size_t N = plan.N;
    const size_t nb = (size_t)plan.Wb * plan.B;
    const size_t out_records = (size_t)plan.Wb * plan.out_per_window;
    const size_t temp_bytes = ws.cub_temp.cap;
    // 2b. batched-affine rounds: pairwise additions inside every bucket, 6 Fq products each instead of 10
    const bool packed = msm_uses_bucket_sort(ctx, plan);
    const uint32_t* acc_keys = ws.keys_b.as<uint32_t>();
    const uint32_t* acc_vals = packed ? acc_keys + 1 : ws.vals_b.as<uint32_t>();
    const G1Affine* acc_points = d_points;
    uint32_t acc_stride = plan.precomp ? TABLE_STRIDE : (uint32_t)sizeof(G1Affine);  // records of d_points
    if (plan.affine_rounds) {
        const uint32_t nbk = plan.discard, R = plan.affine_rounds;
        for (uint32_t r = 0; r <= R; r++) ZKP_CUDA(ws.aff_start[r].ensure(((size_t)nbk + 1) * 4));
        ZKP_CUDA(ws.aff_len.ensure(((size_t)nbk + 1) * 4));
        size_t scan_bytes = 0;
        ZKP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ws.aff_len.as<uint32_t>(), ws.aff_start[1].as<uint32_t>(), (int)(nbk + 1), st));
        if (scan_bytes > temp_bytes) ZKP_CUDA(ws.cub_temp.ensure(scan_bytes));
        k_bucket_bounds<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(ws.keys_b.as<uint32_t>(), N, nbk, ws.aff_start[0].as<uint32_t>());
        ctx->launches++;
        const size_t resident = (size_t)ctx->sm_count * 128 * ZKP_AFF_MIN_BLOCKS;
        for (uint32_t r = 0; r < R; r++) {
            const size_t bound_out = plan.bound[r + 1];
            k_next_len<<<(nbk + 1 + 255) / 256, 256, 0, st>>>(ws.aff_start[r].as<uint32_t>(), nbk, ws.aff_len.as<uint32_t>());
            ZKP_CUDA(cub::DeviceScan::ExclusiveSum(ws.cub_temp.p, scan_bytes, ws.aff_len.as<uint32_t>(), ws.aff_start[r + 1].as<uint32_t>(),
                                                   (int)(nbk + 1), st));
            // outputs per thread: one wave of resident threads, but never fewer than a full inversion batch
            uint32_t K = (uint32_t)((bound_out + resident - 1) / resident);
            if (K < (uint32_t)AFF_KB) K = AFF_KB;
            const size_t threads = (bound_out + K - 1) / K;
            const unsigned blocks = (unsigned)((threads + 127) / 128);
            const uint32_t total_threads = blocks * 128;
            ZKP_CUDA(ws.aff_scratch.ensure((size_t)AFF_KB * total_threads * sizeof(Fq)));
            ZKP_CUDA(ws.aff_pts[r & 1].ensure(bound_out * sizeof(G1Affine)));
            const bool last = r + 1 == R;
            uint32_t* out_keys = nullptr;
            if (last) {
                ZKP_CUDA(ws.aff_keys.ensure(bound_out * 4));
                ZKP_CUDA(cudaMemsetAsync(ws.aff_keys.p, 0xff, bound_out * 4, st));  // unused tail = discard keys
                out_keys = ws.aff_keys.as<uint32_t>();
            }
            if (r == 0)
                k_affine_round<true><<<blocks, 128, 0, st>>>(ws.aff_start[0].as<uint32_t>(), ws.aff_start[1].as<uint32_t>(), nbk,
                                                             ws.vals_b.as<uint32_t>(), d_points, ws.aff_pts[0].as<G1Affine>(), out_keys, K,
                                                             ws.aff_scratch.as<Fq>(), total_threads, (uint32_t)ctx->sm_count, acc_stride);
            else
                k_affine_round<false><<<blocks, 128, 0, st>>>(ws.aff_start[r].as<uint32_t>(), ws.aff_start[r + 1].as<uint32_t>(), nbk, nullptr,
                                                              ws.aff_pts[(r - 1) & 1].as<G1Affine>(), ws.aff_pts[r & 1].as<G1Affine>(), out_keys,
                                                              K, ws.aff_scratch.as<Fq>(), total_threads, (uint32_t)ctx->sm_count);
            ctx->launches += 3;
        }
        acc_keys = ws.aff_keys.as<uint32_t>();
        acc_vals = nullptr;
        acc_points = ws.aff_pts[(R - 1) & 1].as<G1Affine>();
        acc_stride = (uint32_t)sizeof(G1Affine);
        trace_mark(ctx, lane, st, "affine_rounds");
    }
    // 3. balanced accumulation, level by level
    ZKP_CUDA(cudaMemsetAsync(ws.buckets.p, 0, nb * sizeof(G1Xyzz), st));
    for (size_t l = 0; l < plan.levels.size(); l++) {
        const auto& lv = plan.levels[l];
        int last = l + 1 == plan.levels.size();
        // slot levels whose slices (x 4 lanes) fit the machine once are latency-bound: 4 lanes per slice
        const bool coop = l > 0 && lv.threads * 4 <= (size_t)ctx->sm_count * 512;
        unsigned blocks = (unsigned)((lv.threads * (coop ? 4 : 1) + 127) / 128);
        if (l == 0) {
            blocks = (unsigned)((lv.threads + ZKP_ACC_THREADS - 1) / ZKP_ACC_THREADS);
            if (ctx->time_acc) cudaEventRecord(ev0, st);
            k_accumulate<true><<<blocks, ZKP_ACC_THREADS, 0, st>>>(acc_keys, acc_vals, acc_points, nullptr,
                                                       lv.items, lv.L, plan.discard, ws.buckets.as<G1Xyzz>(),
                                                       last ? nullptr : ws.slot_keys[0].as<uint32_t>(),
                                                       last ? nullptr : ws.slot_pts[0].as<G1Xyzz>(), last, packed ? 2u : 1u,
                                                                   acc_stride);
            if (ctx->time_acc) cudaEventRecord(ev1, st);
            trace_mark(ctx, lane, st, "accumulate_l0");
        } else if (coop) {
            k_accumulate<false, true><<<blocks, 128, 0, st>>>(ws.slot_keys[l - 1].as<uint32_t>(), nullptr, nullptr,
                                                              ws.slot_pts[l - 1].as<G1Xyzz>(), lv.items, lv.L, plan.discard,
                                                              ws.buckets.as<G1Xyzz>(),
                                                              last ? nullptr : ws.slot_keys[l].as<uint32_t>(),
                                                              last ? nullptr : ws.slot_pts[l].as<G1Xyzz>(), last);
        } else {
            k_accumulate<false, false><<<blocks, 128, 0, st>>>(ws.slot_keys[l - 1].as<uint32_t>(), nullptr, nullptr,
                                                               ws.slot_pts[l - 1].as<G1Xyzz>(), lv.items, lv.L, plan.discard,
                                                               ws.buckets.as<G1Xyzz>(),
                                                               last ? nullptr : ws.slot_keys[l].as<uint32_t>(),
                                                               last ? nullptr : ws.slot_pts[l].as<G1Xyzz>(), last);
        }
        ctx->launches++;
    }
    trace_mark(ctx, lane, st, "accumulate_slots");
    // 4. reduction -> per bucket window: bits_c column bit planes, then bits_r row bit planes
    const uint32_t opw = plan.out_per_window;
    ZKP_CUDA(ws.sums_out.ensure(out_records * sizeof(G1Xyzz)));
    G1Xyzz* d_out = ws.sums_out.as<G1Xyzz>();
    if (plan.rowcol) {
        const uint32_t rows = 1u << plan.log_rows, cols = 1u << plan.log_cols;
        ZKP_CUDA(ws.sums_a.ensure((size_t)plan.Wb * cols * sizeof(G1Xyzz)));
        ZKP_CUDA(ws.sums_b.ensure((size_t)plan.Wb * rows * sizeof(G1Xyzz)));
        // shares per sum so that one wave of resident threads (3 CTAs of 128 per SM) covers all of them evenly
        const size_t elems = 2ull * plan.Wb * rows * cols, resident = (size_t)ctx->sm_count * 384;
        const uint32_t per_thread = (uint32_t)((elems + resident - 1) / resident);
        uint32_t q_c = per_thread ? (rows + per_thread - 1) / per_thread : 1, q_r = per_thread ? (cols + per_thread - 1) / per_thread : 1;
        if (q_c > 2 * RC_LANES) q_c = 2 * RC_LANES;
        if (q_r > 2 * RC_LANES) q_r = 2 * RC_LANES;
        // (>= 2 shares per sum: with many bucket windows -- a batch of requests -- two shares already give one balanced
        // wave; the one-warp-per-sum kernel below ran the 64-window reduction of a 32-request batch at 15.8 G Fq-mul/s
        // against ~25 for this pair, profiles/r2_launches_summary.txt)
        // small bucket arrays (a single request at the mainnet row size): four lanes per share and two warps per sum --
        // both stages are latency there.  Not inside a two-lane commit+open: its two tails run side by side, the cooperative
        // form issues 1.4x the multiplies and measured no gain there (and -3% for four contexts serving such requests at once).
        const bool small = ctx->rowcol_coop && !ctx->two_lanes_busy && plan.Wb == 1 && elems <= (size_t)ctx->sm_count * 512 && rows >= 32 && cols >= 32;
        if (small) {
            // e buckets per share, the same for rows and columns, as few as one wave of groups allows (>= 4)
            uint32_t e = (uint32_t)((elems * 4 + (size_t)ctx->sm_count * 384 - 1) / ((size_t)ctx->sm_count * 384));
            if (e < 4) e = 4;
            const uint32_t qc = (rows + e - 1) / e, qr = (cols + e - 1) / e;
            const uint32_t shares = cols * qc + rows * qr;
            ZKP_CUDA(ws.pool.ensure((size_t)shares * sizeof(G1Xyzz)));
            k_rowcol_partial_coop<<<dim3((shares * 4 + 127) / 128, 1), 128, 0, st>>>(ws.buckets.as<G1Xyzz>(), plan.log_rows, plan.log_cols, qc, qr,
                                                                                ws.pool.as<G1Xyzz>());
            k_rowcol_finish_wide<<<dim3((cols + rows + RCW_SUMS - 1) / RCW_SUMS, 1), RCW_THREADS, 0, st>>>(
                ws.pool.as<G1Xyzz>(), plan.log_rows, plan.log_cols, qc, qr, ws.sums_a.as<G1Xyzz>(), ws.sums_b.as<G1Xyzz>());
            ctx->launches++;
        } else if (q_c >= 2 && q_r >= 2) {
            const uint32_t shares = cols * q_c + rows * q_r;
            ZKP_CUDA(ws.pool.ensure((size_t)plan.Wb * shares * sizeof(G1Xyzz)));
            k_rowcol_partial<<<dim3((shares + 127) / 128, plan.Wb), 128, 0, st>>>(ws.buckets.as<G1Xyzz>(), plan.log_rows, plan.log_cols,
                                                                               q_c, q_r, ws.pool.as<G1Xyzz>());
            k_rowcol_finish<<<dim3((cols + rows + RC_SUMS - 1) / RC_SUMS, plan.Wb), RC_THREADS, 0, st>>>(
                ws.pool.as<G1Xyzz>(), plan.log_rows, plan.log_cols, q_c, q_r, ws.sums_a.as<G1Xyzz>(), ws.sums_b.as<G1Xyzz>());
            ctx->launches++;
        } else {
            k_rowcol_sums<<<dim3((cols + rows + RC_SUMS - 1) / RC_SUMS, plan.Wb), RC_THREADS, 0, st>>>(
                ws.buckets.as<G1Xyzz>(), plan.log_rows, plan.log_cols, ws.sums_a.as<G1Xyzz>(), ws.sums_b.as<G1Xyzz>());
        }
        trace_mark(ctx, lane, st, "rowcol_sums");
        k_bit_sums<<<dim3(plan.bits_c + plan.bits_r, plan.Wb), TAIL_THREADS, 0, st>>>(
            ws.sums_a.as<G1Xyzz>(), cols, plan.bits_c, ws.sums_b.as<G1Xyzz>(), rows, d_out, opw);
        ctx->launches += 2;
    } else {
        k_bit_sums<<<dim3(plan.bits_c, plan.Wb), TAIL_THREADS, 0, st>>>(ws.buckets.as<G1Xyzz>(), plan.B, plan.bits_c, nullptr, 0,
                                                                        d_out, opw);
        ctx->launches++;
    }
    trace_mark(ctx, lane, st, "bit_sums");
    ZKP_CUDA(cudaMemcpyAsync(ws.h_window, d_out, sizeof(G1Xyzz) * out_records, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaMemcpyAsync(ws.h_bad, ws.bad.p, 4 * plan.groups, cudaMemcpyDeviceToHost, st));
    trace_mark(ctx, lane, st, "d2h");
    return ZKP_OK;
}

// waits for the lane's pipeline; afterwards ws.h_window holds the bit-plane sums
inline int msm_wait(zkp_ctx* ctx, int lane, uint32_t groups = 1, bool check_bad = true) {
    MsmWorkspace& ws = lane ? ctx->ws2 : ctx->ws;
    cudaStream_t st = lane ? ctx->stream2 : ctx->stream;
    cudaEvent_t ev0 = lane ? ctx->ev_acc2_0 : ctx->ev_acc0, ev1 = lane ? ctx->ev_acc2_1 : ctx->ev_acc1;
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    if (ctx->time_acc) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) {
            ctx->acc_ms_total += ms;
            ctx->acc_count++;
        }
    }
    if (check_bad)
        for (uint32_t g = 0; g < groups; g++)
            if (ws.h_bad[g]) return fail(ZKP_ERR_ENCODING, "scalar is not a canonical field element (>= r)");
    return ZKP_OK;
}

// 5. host: Horner over the bit planes of every bucket window, then the window fold
inline host::G1J xyzz_to_jac(const G1Xyzz& p) {
    using namespace host;
    Fq64 x, y, zz, zzz;
    memcpy(x.v, p.x.v, 48);
    memcpy(y.v, p.y.v, 48);
    memcpy(zz.v, p.zz.v, 48);
    memcpy(zzz.v, p.zzz.v, 48);
    if (zz.is_zero()) return G1J::infinity();
    // XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2)  ->  Jacobian with Z = ZZ
    return G1J{x * zz, y * zzz, zz};
}
inline host::G1J horner_bits(const G1Xyzz* planes, uint32_t bits) {
    host::G1J acc = host::G1J::infinity();
    for (int j = (int)bits - 1; j >= 0; j--) acc = acc.dbl().add(xyzz_to_jac(planes[j]));
    return acc;
}
// result of group g of a grouped (fixed-base) launch set: one bucket window, no window fold
inline host::G1J msm_fold_group(const MsmPlan& plan, const G1Xyzz* h_window, uint32_t g) {
    using namespace host;
    const G1Xyzz* rec = h_window + (size_t)g * plan.out_per_window;
    G1J win = horner_bits(rec, plan.bits_c);
    if (plan.bits_r) {
        G1J r = horner_bits(rec + plan.bits_c, plan.bits_r);
        for (uint32_t d = 0; d < plan.log_cols; d++) r = r.dbl();
        win = win.add(r);
    }
    return win;
}
inline host::G1J msm_fold(const MsmPlan& plan, const G1Xyzz* h_window) {
    using namespace host;
    if (plan.precomp) return msm_fold_group(plan, h_window, 0);
    G1J acc = G1J::infinity();
    for (int w = (int)plan.Wb - 1; w >= 0; w--) {
        for (uint32_t d = 0; d < plan.c && !acc.is_inf(); d++) acc = acc.dbl();
        const G1Xyzz* rec = h_window + (size_t)w * plan.out_per_window;
        G1J win = horner_bits(rec, plan.bits_c);  // sum_lo (lo + 1) C_lo   (or sum_b (b + 1) B_b)
        if (plan.bits_r) {
            G1J r = horner_bits(rec + plan.bits_c, plan.bits_r);  // sum_hi hi R_hi
            for (uint32_t d = 0; d < plan.log_cols; d++) r = r.dbl();
            win = win.add(r);
        }
        acc = acc.add(win);
    }
    return acc;
}

}  // namespace zkp
