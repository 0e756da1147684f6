// Host driver of the MSM pipeline described in msm.cuh.
#pragma once
#include "context.cuh"

namespace zkp {

inline uint64_t msm_fq_muls(const MsmPlan& p) {
    // algorithmic Fq multiplications: 10 per mixed add into buckets, 14 per full add in
    // the running-sum reduction (2 adds per bucket)
    return 10ull * p.n * p.W + 14ull * 2ull * p.B * p.W;
}

// Launches the whole device pipeline on ctx->stream and copies the W window sums to pinned host
// memory; synchronises the stream before returning.
inline int msm_run(zkp_ctx* ctx, const MsmPlan& plan, const uint32_t* d_scalars, int fmt,
                   const G1Affine* d_points) {
    MsmWorkspace& ws = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t N = plan.N;
    ZKP_CUDA(ws.keys_a.ensure(N * 4));
    ZKP_CUDA(ws.keys_b.ensure(N * 4));
    ZKP_CUDA(ws.vals_a.ensure(N * 4));
    ZKP_CUDA(ws.vals_b.ensure(N * 4));
    const size_t nb = (size_t)plan.W * plan.B;
    ZKP_CUDA(ws.buckets.ensure(nb * sizeof(G1Xyzz)));
    if (ws.slot_keys.size() < plan.levels.size()) {
        ws.slot_keys.resize(plan.levels.size());
        ws.slot_pts.resize(plan.levels.size());
    }
    for (size_t l = 0; l + 1 < plan.levels.size(); l++) {
        ZKP_CUDA(ws.slot_keys[l].ensure(plan.levels[l].threads * 2 * 4));
        ZKP_CUDA(ws.slot_pts[l].ensure(plan.levels[l].threads * 2 * sizeof(G1Xyzz)));
    }
    if (ws.h_window_cap < (size_t)plan.W * plan.out_per_window) {
        if (ws.h_window) cudaFreeHost(ws.h_window);
        ZKP_CUDA(cudaMallocHost(&ws.h_window, sizeof(G1Xyzz) * plan.W * plan.out_per_window));
        ws.h_window_cap = (size_t)plan.W * plan.out_per_window;
    }
    if (!ws.h_bad) ZKP_CUDA(cudaMallocHost(&ws.h_bad, 8));

    // 1. digits
    ZKP_CUDA(ws.bad.ensure(8));
    ZKP_CUDA(cudaMemsetAsync(ws.bad.p, 0, 4, st));
    k_decompose<<<(plan.n + 255) / 256, 256, 0, st>>>(d_scalars, plan.n, plan.c, plan.W, plan.B, plan.discard,
                                                      fmt, ws.keys_a.as<uint32_t>(), ws.vals_a.as<uint32_t>(), ws.bad.as<uint32_t>());
    ctx->launches++;
    // 2. sort by (window, bucket)
    size_t temp_bytes = 0;
    ZKP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, ws.keys_a.as<uint32_t>(), ws.keys_b.as<uint32_t>(),
                                             ws.vals_a.as<uint32_t>(), ws.vals_b.as<uint32_t>(), (int64_t)N, 0,
                                             (int)plan.key_bits, st));
    ZKP_CUDA(ws.cub_temp.ensure(temp_bytes));
    ZKP_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_temp.p, temp_bytes, ws.keys_a.as<uint32_t>(), ws.keys_b.as<uint32_t>(),
                                             ws.vals_a.as<uint32_t>(), ws.vals_b.as<uint32_t>(), (int64_t)N, 0,
                                             (int)plan.key_bits, st));
    // 3. balanced accumulation, level by level
    ZKP_CUDA(cudaMemsetAsync(ws.buckets.p, 0, nb * sizeof(G1Xyzz), st));
    for (size_t l = 0; l < plan.levels.size(); l++) {
        const auto& lv = plan.levels[l];
        int last = l + 1 == plan.levels.size();
        unsigned blocks = (unsigned)((lv.threads + 127) / 128);
        if (l == 0) {
            if (ctx->time_acc) cudaEventRecord(ctx->ev_acc0, st);
            k_accumulate<true><<<blocks, 128, 0, st>>>(ws.keys_b.as<uint32_t>(), ws.vals_b.as<uint32_t>(), d_points, nullptr,
                                                       lv.items, lv.L, plan.discard, ws.buckets.as<G1Xyzz>(),
                                                       last ? nullptr : ws.slot_keys[0].as<uint32_t>(),
                                                       last ? nullptr : ws.slot_pts[0].as<G1Xyzz>(), last);
            if (ctx->time_acc) cudaEventRecord(ctx->ev_acc1, st);
        } else {
            k_accumulate<false><<<blocks, 128, 0, st>>>(ws.slot_keys[l - 1].as<uint32_t>(), nullptr, nullptr,
                                                        ws.slot_pts[l - 1].as<G1Xyzz>(), lv.items, lv.L, plan.discard,
                                                        ws.buckets.as<G1Xyzz>(),
                                                        last ? nullptr : ws.slot_keys[l].as<uint32_t>(),
                                                        last ? nullptr : ws.slot_pts[l].as<G1Xyzz>(), last);
        }
        ctx->launches++;
    }
    // 4. bucket reduction: running-sum levels -> pool of plain addends; bit-plane sums for the tail.
    //    Device output per window: [0] = sum of the pool, [1 + j] = P_j (weight 2^j).
    const uint32_t opw = plan.out_per_window;
    ZKP_CUDA(ws.sums_out.ensure((size_t)plan.W * opw * sizeof(G1Xyzz)));
    G1Xyzz* d_out = ws.sums_out.as<G1Xyzz>();
    const G1Xyzz* in = ws.buckets.as<G1Xyzz>();
    uint32_t in_stride = plan.B;
    if (!plan.rlevels.empty()) {
        ZKP_CUDA(ws.pool.ensure((size_t)plan.W * plan.pool_per_window * sizeof(G1Xyzz)));
        size_t next_cap = (size_t)plan.W * plan.rlevels[0].chunks * sizeof(G1Xyzz);
        ZKP_CUDA(ws.next_a.ensure(next_cap));
        ZKP_CUDA(ws.next_b.ensure(next_cap));
        uint32_t pool_off = 0;
        for (size_t l = 0; l < plan.rlevels.size(); l++) {
            const auto& rl = plan.rlevels[l];
            G1Xyzz* next = l % 2 == 0 ? ws.next_a.as<G1Xyzz>() : ws.next_b.as<G1Xyzz>();
            uint32_t log_m = 0;
            while ((1u << log_m) < rl.m) log_m++;
            uint32_t total = rl.chunks * plan.W;
            k_bucket_reduce<<<(total + 127) / 128, 128, 0, st>>>(in, rl.n_in, in_stride, rl.m, log_m, rl.chunks, l == 0, next,
                                                                 rl.chunks, ws.pool.as<G1Xyzz>(), plan.pool_per_window,
                                                                 pool_off, plan.W);
            ctx->launches++;
            pool_off += rl.chunks;
            in = next;
            in_stride = rl.chunks;
        }
    }
    k_bit_sums<<<dim3(plan.tail_bits, plan.W), TAIL_THREADS, 0, st>>>(in, plan.tail_n, in_stride, plan.tail_one_based, d_out, opw, 1);
    ctx->launches++;
    if (!plan.rlevels.empty()) {
        // plain sum of the pool -> d_out[w][0]
        uint32_t count = plan.pool_per_window;
        const G1Xyzz* sin = ws.pool.as<G1Xyzz>();
        uint32_t sstride = plan.pool_per_window;
        uint32_t parts0 = (count + SUM_PART - 1) / SUM_PART;
        ZKP_CUDA(ws.sums_a.ensure((size_t)plan.W * parts0 * sizeof(G1Xyzz)));
        ZKP_CUDA(ws.sums_b.ensure((size_t)plan.W * parts0 * sizeof(G1Xyzz)));
        int flip = 0;
        for (;;) {
            uint32_t parts = (count + SUM_PART - 1) / SUM_PART;
            bool fin = parts == 1;
            G1Xyzz* sout = fin ? d_out : (flip ? ws.sums_b.as<G1Xyzz>() : ws.sums_a.as<G1Xyzz>());
            k_sum_segments<<<dim3(parts, plan.W), SUM_THREADS, 0, st>>>(sin, count, sstride, sout, fin ? opw : parts);
            ctx->launches++;
            if (fin) break;
            sin = sout;
            sstride = parts;
            count = parts;
            flip ^= 1;
        }
    } else {
        ZKP_CUDA(cudaMemset2DAsync(d_out, (size_t)opw * sizeof(G1Xyzz), 0, sizeof(G1Xyzz), plan.W, st));
    }
    const G1Xyzz* sin = d_out;
    ZKP_CUDA(cudaMemcpyAsync(ws.h_window, sin, sizeof(G1Xyzz) * plan.W * opw, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaMemcpyAsync(ws.h_bad, ws.bad.p, 4, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    if (ctx->time_acc) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_acc0, ctx->ev_acc1) == cudaSuccess) {
            ctx->acc_ms_total += ms;
            ctx->acc_count++;
        }
    }
    if (*ws.h_bad) return fail(ZKP_ERR_ENCODING, "scalar is not a canonical field element (>= r)");
    return ZKP_OK;
}

// 5. host: per window  pool + sum_j 2^j P_j  (Horner over the bit planes), then fold the windows
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
inline host::G1J msm_fold(const MsmPlan& plan, const G1Xyzz* h_window) {
    using namespace host;
    G1J acc = G1J::infinity();
    const uint32_t opw = plan.out_per_window;
    for (int w = (int)plan.W - 1; w >= 0; w--) {
        for (uint32_t d = 0; d < plan.c; d++) acc = acc.dbl();
        const G1Xyzz* rec = h_window + (size_t)w * opw;
        G1J win = G1J::infinity();
        for (int j = (int)plan.tail_bits - 1; j >= 0; j--) win = win.dbl().add(xyzz_to_jac(rec[1 + j]));
        win = win.add(xyzz_to_jac(rec[0]));
        acc = acc.add(win);
    }
    return acc;
}

}  // namespace zkp
