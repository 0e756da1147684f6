// In-library multi-GPU entries (include/zkp_b200.h "multi-GPU"): one process, one context + one persistent host thread
// per device, no PyTorch, no NCCL, no collective inside any kernel.  The path shards with no data exchange on the inner
// loop: what crosses devices is one 32-byte partial sum and two Jacobian points per GPU, over
// pinned host memory, added on the calling thread (G - 1 point additions, ONE field inversion for the compression --
// no square roots, no decompression).
//
//   ZKP_LAYOUT_ROWS         every device holds the whole SRS; request / row i is served by device i mod G.  This is the
//                           Pianist split of the reference (neurons/validator.py:41-42,212-222: miner i gets row i)
//                           run on the GPUs of one box: zkp_mgpu_pianist_commit_open.
//   ZKP_LAYOUT_POINT_RANGE  device g holds points [g n/G, (g+1) n/G) of every row (zkp_srs_generate_shard); ONE
//                           polynomial is split by point range: zkp_mgpu_msm_g1, zkp_mgpu_commit_open.
// Included at the end of zkp_b200.cu (uses the internals of capi_rest.cuh).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <thread>

struct zkp_mgpu {
    std::vector<zkp_ctx*> ctxs;
    std::vector<int> devices;
    int layout = 0;
    uint32_t log_n = 0, log_m = 0, log_shards = 0;  // log_n = FULL row length
    std::mutex mu;                                   // one multi-device call at a time
    // Job handoff: a mutex + condition variable for the idle case, and an atomic sequence pair that both sides SPIN on
    // for a bounded time first -- a futex wake of a sleeping thread costs 50-100 us each way, which is most of what a
    // multi-device call adds to a 12 ms step; a worker that has just finished a job spins ~200 us for the next one (a
    // prover in a loop) before it goes to sleep, the caller spins for completion (it has nothing else to do).
    struct Worker {
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        std::function<int()> job;
        std::atomic<uint64_t> posted{0}, finished{0};  // jobs handed over / completed
        bool quit = false;
        int rc = 0;
        std::string err;
    };
    std::vector<std::unique_ptr<Worker>> workers;
    // per-device state carried between the two phases of a point-range opening
    struct Phase {
        zkp::MsmPlan plan_c;
        zkp::host::Fr64 s;
        zkp::host::G1J com, proof;
        std::unique_lock<std::mutex> lock;
    };
    std::vector<Phase> phase;
};

namespace {

void mgpu_worker_main(zkp_mgpu::Worker* w, int device) {
    cudaSetDevice(device);
    uint64_t seen = 0;
    for (;;) {
        // wait for job number seen + 1: spin briefly, then sleep
        const auto spin_until = std::chrono::steady_clock::now() + std::chrono::microseconds(200);
        bool got = false;
        while (std::chrono::steady_clock::now() < spin_until) {
            if (w->posted.load(std::memory_order_acquire) > seen) { got = true; break; }
        }
        if (!got) {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->posted.load(std::memory_order_acquire) > seen || w->quit; });
            if (w->quit) return;
        }
        seen++;
        int rc = w->job();
        w->rc = rc;
        w->err = rc ? tls_error() : std::string();
        w->finished.store(seen, std::memory_order_release);
    }
}

// run fn(k) on the worker of device k for k < count; the first failure (message included) is returned to the caller
int mgpu_run(zkp_mgpu* mg, uint32_t count, const std::function<int(uint32_t)>& fn) {
    std::vector<uint64_t> want(count);
    for (uint32_t k = 0; k < count; k++) {
        zkp_mgpu::Worker& w = *mg->workers[k];
        {
            std::lock_guard<std::mutex> lk(w.m);  // orders the job object before the sequence number for a sleeping worker
            w.job = [&fn, k] { return fn(k); };
            want[k] = w.posted.load(std::memory_order_relaxed) + 1;
            w.posted.store(want[k], std::memory_order_release);
        }
        w.cv.notify_all();
    }
    int rc = ZKP_OK;
    for (uint32_t k = 0; k < count; k++) {
        zkp_mgpu::Worker& w = *mg->workers[k];
        while (w.finished.load(std::memory_order_acquire) < want[k]) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (w.rc && !rc) rc = fail(w.rc, "device " + std::to_string(mg->devices[k]) + ": " + w.err);
    }
    return rc;
}

uint32_t mgpu_active(const zkp_mgpu* mg) { return mg->layout == ZKP_LAYOUT_POINT_RANGE ? 1u << mg->log_shards : (uint32_t)mg->ctxs.size(); }

int mgpu_check(zkp_mgpu* mg, int layout) {
    if (!mg) return fail(ZKP_ERR_ARG, "null multi-GPU handle");
    if (mg->layout != layout)
        return fail(ZKP_ERR_STATE, layout == ZKP_LAYOUT_ROWS ? "needs an SRS in the ZKP_LAYOUT_ROWS layout (zkp_mgpu_srs_generate / load)"
                                                             : "needs an SRS in the ZKP_LAYOUT_POINT_RANGE layout (zkp_mgpu_srs_generate / load)");
    return ZKP_OK;
}

// point-range opening, phase 1 on one device: upload the slice, start the commitment MSM on lane 0, form the partial
// barycentric sum S_g on lane 1 and bring it to the host (the commitment's accumulation keeps running meanwhile)
int shard_phase1(zkp_ctx* ctx, uint32_t row, const uint8_t* slice_be, size_t nl, const Fr64& x, bool resident, zkp_mgpu::Phase* ph) {
    ph->lock = std::unique_lock<std::mutex>(ctx->mu);
    DeviceGuard g(ctx->device);
    if (ctx->coeff_form) return fail(ZKP_ERR_STATE, "coefficient form is not available on point-range shards (a distributed NTT)");
    int rc;
    if (resident) {
        if (ctx->resident_n != nl) return fail(ZKP_ERR_STATE, "no polynomial slice resident on this device");
        rc = convert_poly(ctx, nl);
    } else {
        rc = upload_poly(ctx, slice_be, nl);
    }
    if (rc) return rc;
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream2;
    ZKP_CUDA(cudaEventRecord(ctx->ev_ready, s0));
    ZKP_CUDA(cudaStreamWaitEvent(s1, ctx->ev_ready, 0));
    const G1Affine* pts_c = nullptr;
    rc = msm_device_prep(ctx, 0, row, ctx->scalars.as<uint32_t>(), SCALAR_BE, nl, &ph->plan_c, &pts_c);
    if (rc) return rc;
    zkp_ctx::Domain* dom;
    rc = get_domain(ctx, ctx->shard_domain_log, false, &dom);
    if (rc) return rc;
    const uint32_t n = (uint32_t)nl;
    rc = open_buffers(ctx, n, 1);
    if (rc) return rc;
    uint32_t E = n >> 16;
    if (E < 4) E = 4;
    if (E > 16) E = 16;
    const uint32_t threads = (n + E - 1) / E, blocks = (threads + 127) / 128;
    ZKP_CUDA(ctx->partials.ensure((size_t)blocks * 32));
    ZKP_CUDA(cudaMemsetAsync(small_at<uint32_t>(ctx, SM_HIT), 0xff, 4, s1));
    k_open_pass1<<<blocks, 128, 0, s1>>>(ctx->fr_a.as<Fr>(), n, E, to_dev(x), dom->wt.as<Fr>(), to_dev(dom->w_inv), ctx->fr_b.as<Fr>(),
                                         ctx->partials.as<Fr>(), small_at<uint32_t>(ctx, SM_HIT), (uint64_t)ctx->shard_index << ctx->log_n);
    k_fr_reduce<<<1, 256, 0, s1>>>(ctx->partials.as<Fr>(), blocks, small_at<Fr>(ctx, SM_S1));
    ctx->launches += 2;
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small, small_at<uint8_t>(ctx, SM_S1), 32, cudaMemcpyDeviceToHost, s1));
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 96, small_at<uint8_t>(ctx, SM_HIT), 4, cudaMemcpyDeviceToHost, s1));
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 64, small_at<uint8_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost, s1));
    rc = msm_enqueue_main(ctx, 0, ph->plan_c, pts_c);
    if (rc) return rc;
    ZKP_CUDA(cudaStreamSynchronize(s1));
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 64)) return fail(ZKP_ERR_ENCODING, "polynomial holds a non-canonical field element");
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 96) != HIT_NONE)
        return fail(ZKP_ERR_ARG, "evaluation point lies inside the domain: not supported on point-range shards");
    memcpy(ph->s.v, ctx->h_small, 32);  // Montgomery limbs
    return ZKP_OK;
}
// phase 2: y is known -- quotient of the slice and its MSM on lane 1; wait for both lanes, fold on this thread
int shard_phase2(zkp_ctx* ctx, uint32_t row, size_t nl, const Fr64& y, zkp_mgpu::Phase* ph) {
    DeviceGuard g(ctx->device);
    cudaStream_t s1 = ctx->stream2;
    const uint32_t n = (uint32_t)nl;
    memcpy(ctx->h_small + 128, y.v, 32);
    ZKP_CUDA(cudaMemcpyAsync(small_at<uint8_t>(ctx, SM_Y), ctx->h_small + 128, 32, cudaMemcpyHostToDevice, s1));
    k_open_pass2<<<(n + 255) / 256, 256, 0, s1>>>(ctx->fr_a.as<Fr>(), ctx->fr_b.as<Fr>(), n, small_at<Fr>(ctx, SM_Y), ctx->fr_c.as<Fr>());
    ctx->launches++;
    MsmPlan plan_o;
    int rc = msm_device_enqueue(ctx, 1, row, ctx->fr_c.as<uint32_t>(), SCALAR_MONT, nl, &plan_o);
    if (!rc) rc = msm_device_finish(ctx, 0, ph->plan_c, nullptr, &ph->com);
    if (!rc) rc = msm_device_finish(ctx, 1, plan_o, nullptr, &ph->proof);
    if (rc == ZKP_OK) ctx->resident_n = nl;
    return rc;
}
void shard_phase_end(zkp_ctx* ctx, zkp_mgpu::Phase* ph) {
    if (!ph->lock.owns_lock()) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
    msm_unpin_all(ctx);
    ph->lock.unlock();
    ph->lock = std::unique_lock<std::mutex>();
}

}  // namespace

extern "C" {

int zkp_mgpu_create(const int* devices, int count, zkp_mgpu** out) {
    if (!out || count < 1 || count > 64) return fail(ZKP_ERR_ARG, "bad argument");
    *out = nullptr;
    std::unique_ptr<zkp_mgpu> mg(new zkp_mgpu());
    int rc = ZKP_OK;
    for (int k = 0; k < count && !rc; k++) {
        const int dev = devices ? devices[k] : k;
        for (int j = 0; j < k; j++)
            if (mg->devices[j] == dev) rc = fail(ZKP_ERR_ARG, "device listed twice");
        zkp_ctx* c = nullptr;
        if (!rc) rc = zkp_ctx_create(dev, &c);
        if (!rc) {
            mg->ctxs.push_back(c);
            mg->devices.push_back(dev);
        }
    }
    if (rc) {
        for (zkp_ctx* c : mg->ctxs) zkp_ctx_destroy(c);
        return rc;
    }
    mg->phase.resize(count);
    for (int k = 0; k < count; k++) {
        mg->workers.emplace_back(new zkp_mgpu::Worker());
        zkp_mgpu::Worker* w = mg->workers.back().get();
        w->th = std::thread(mgpu_worker_main, w, mg->devices[k]);
    }
    *out = mg.release();
    return ZKP_OK;
}

void zkp_mgpu_destroy(zkp_mgpu* mg) {
    if (!mg) return;
    for (auto& w : mg->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        w->th.join();
    }
    for (zkp_ctx* c : mg->ctxs) zkp_ctx_destroy(c);
    delete mg;
}

int zkp_mgpu_device_count(zkp_mgpu* mg) { return mg ? (int)mg->ctxs.size() : 0; }

// the per-device context (borrowed: for zkp_srs_import_row, zkp_worker_verify, the tuning knobs ...); it must not be
// used while a zkp_mgpu_* call is running
zkp_ctx* zkp_mgpu_ctx(zkp_mgpu* mg, int k) { return mg && k >= 0 && (size_t)k < mg->ctxs.size() ? mg->ctxs[k] : nullptr; }

// record the layout after the per-device contexts have been filled through zkp_mgpu_ctx (import / load paths)
int zkp_mgpu_set_layout(zkp_mgpu* mg, int layout, uint32_t log_n, uint32_t log_machines) {
    if (!mg || (layout != ZKP_LAYOUT_ROWS && layout != ZKP_LAYOUT_POINT_RANGE)) return fail(ZKP_ERR_ARG, "bad argument");
    uint32_t ls = 0;
    if (layout == ZKP_LAYOUT_POINT_RANGE) {
        while ((2u << ls) <= mg->ctxs.size()) ls++;
        if (ls > log_n) ls = log_n;
    }
    const uint32_t active = layout == ZKP_LAYOUT_POINT_RANGE ? 1u << ls : (uint32_t)mg->ctxs.size();
    for (uint32_t k = 0; k < active; k++) {
        zkp_ctx* c = mg->ctxs[k];
        if (!c->shaped || c->log_m != log_machines || c->log_n != log_n - ls)
            return fail(ZKP_ERR_STATE, "device " + std::to_string(mg->devices[k]) + " does not hold an SRS of that layout");
    }
    std::lock_guard<std::mutex> lk(mg->mu);
    mg->layout = layout;
    mg->log_n = log_n;
    mg->log_m = log_machines;
    mg->log_shards = ls;
    return ZKP_OK;
}

int zkp_mgpu_srs_generate(zkp_mgpu* mg, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32], uint32_t log_n, uint32_t log_machines,
                          int layout) {
    if (!mg || !tau_x_be || !tau_y_be) return fail(ZKP_ERR_ARG, "null argument");
    if (layout != ZKP_LAYOUT_ROWS && layout != ZKP_LAYOUT_POINT_RANGE) return fail(ZKP_ERR_ARG, "unknown layout");
    uint32_t ls = 0;
    if (layout == ZKP_LAYOUT_POINT_RANGE) {
        while ((2u << ls) <= mg->ctxs.size()) ls++;  // the largest power of two of devices
        if (ls > log_n) ls = log_n;
    }
    {
        std::lock_guard<std::mutex> lk(mg->mu);
        const uint32_t active = layout == ZKP_LAYOUT_POINT_RANGE ? 1u << ls : (uint32_t)mg->ctxs.size();
        int rc = mgpu_run(mg, active, [&](uint32_t k) {
            return layout == ZKP_LAYOUT_ROWS ? zkp_srs_generate(mg->ctxs[k], tau_x_be, tau_y_be, log_n, log_machines)
                                             : zkp_srs_generate_shard(mg->ctxs[k], tau_x_be, tau_y_be, log_n, log_machines, k, ls);
        });
        if (rc) return rc;
    }
    return zkp_mgpu_set_layout(mg, layout, log_n, log_machines);
}

// eager fixed-base tables on every device: in the ROWS layout device g builds the tables of the rows it serves
// (i mod G == g), in the POINT_RANGE layout every device builds the tables of its shard of every row
int zkp_mgpu_prebuild_tables(zkp_mgpu* mg) {
    if (!mg || !mg->layout) return fail(ZKP_ERR_STATE, "SRS not loaded");
    std::lock_guard<std::mutex> lk(mg->mu);
    const uint32_t active = mgpu_active(mg), rows = 1u << mg->log_m;
    return mgpu_run(mg, active, [&](uint32_t k) {
        if (mg->layout == ZKP_LAYOUT_POINT_RANGE) return zkp_srs_prebuild_tables(mg->ctxs[k], 0, rows, nullptr);
        for (uint32_t r = k; r < rows; r += active) {
            int rc = zkp_srs_prebuild_tables(mg->ctxs[k], r, 1, nullptr);
            if (rc) return rc;
        }
        return (int)ZKP_OK;
    });
}

// ONE G1 MSM split by point range: device g multiplies its points [g n/G, (g+1) n/G) of row `row` by the matching
// slice of `scalars_be` (the FULL vector of n <= 2^log_n scalars in host memory; each device uploads only its slice)
// and folds its own result; the caller's thread adds the G Jacobian partials and compresses once.
// flags & ZKP_MGPU_RESIDENT: the slices uploaded by the previous call are still on the devices -- no upload (benchmarks
// of the device-resident path).
int zkp_mgpu_msm_g1(zkp_mgpu* mg, uint32_t row, const uint8_t* scalars_be, size_t n, int flags, uint8_t out48[48]) {
    int rc = mgpu_check(mg, ZKP_LAYOUT_POINT_RANGE);
    if (rc) return rc;
    if (!scalars_be || !out48 || !n || n > ((size_t)1 << mg->log_n)) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(mg->mu);
    const uint32_t active = mgpu_active(mg);
    const size_t nl = (size_t)1 << (mg->log_n - mg->log_shards);
    std::vector<host::G1J> part(active, host::G1J::infinity());
    rc = mgpu_run(mg, active, [&](uint32_t k) -> int {
        const size_t lo = (size_t)k * nl;
        if (lo >= n) return ZKP_OK;
        const size_t cnt = n - lo < nl ? n - lo : nl;
        zkp_ctx* ctx = mg->ctxs[k];
        int r = check_row(ctx, row, cnt);
        if (r) return r;
        std::lock_guard<std::mutex> lk2(ctx->mu);
        DeviceGuard g(ctx->device);
        if (flags & ZKP_MGPU_RESIDENT) {
            if (ctx->resident_n != cnt) return fail(ZKP_ERR_STATE, "no scalar slice resident on this device");
        } else {
            r = upload_scalars(ctx, scalars_be + 32 * lo, cnt, ctx->scalars);
            if (r) return r;
        }
        r = msm_device(ctx, row, ctx->scalars.as<uint32_t>(), SCALAR_BE, cnt, nullptr, &part[k]);
        if (r == ZKP_OK) ctx->resident_n = cnt;
        return r;
    });
    if (rc) return rc;
    host::G1J acc = part[0];
    for (uint32_t k = 1; k < active; k++) acc = acc.add(part[k]);
    host::g1_compress(out48, acc);
    return ZKP_OK;
}

// commit + open of ONE polynomial of exactly 2^log_n evaluations split by point range.  Phase 1 on every device:
// upload of its slice, commitment MSM started, partial barycentric sum S_g (32 bytes) to the host.  The calling thread
// forms y = -(x^n - 1)/n * sum_g S_g.  Phase 2: quotient of the slice, its MSM, both results folded on the device's
// own host thread.  x inside the evaluation domain is refused (ZKP_ERR_ARG), as for zkp_shard_eval_partial.
int zkp_mgpu_commit_open(zkp_mgpu* mg, uint32_t row, const uint8_t* poly_be, size_t n, const uint8_t x_be[32], int flags,
                         uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]) {
    int rc = mgpu_check(mg, ZKP_LAYOUT_POINT_RANGE);
    if (rc) return rc;
    if (!poly_be || !x_be || !commitment48 || !eval_be || !proof48) return fail(ZKP_ERR_ARG, "null argument");
    if (n != ((size_t)1 << mg->log_n)) return fail(ZKP_ERR_ARG, "opening needs exactly one SRS row of evaluations");
    Fr64 x;
    if (!Fr64::from_be(x, x_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    std::lock_guard<std::mutex> lk(mg->mu);
    const uint32_t active = mgpu_active(mg);
    const size_t nl = n >> mg->log_shards;
    for (uint32_t k = 0; k < active; k++) {
        rc = check_row(mg->ctxs[k], row, nl);
        if (rc) return rc;
    }
    rc = mgpu_run(mg, active, [&](uint32_t k) {
        return shard_phase1(mg->ctxs[k], row, poly_be + 32 * nl * k, nl, x, (flags & ZKP_MGPU_RESIDENT) != 0, &mg->phase[k]);
    });
    Fr64 y = Fr64::zero();
    if (!rc) {
        Fr64 acc = Fr64::zero();
        for (uint32_t k = 0; k < active; k++) acc = acc + mg->phase[k].s;
        Fr64 xn = x;
        for (uint32_t k = 0; k < mg->log_n; k++) xn = xn.sqr();
        y = ((xn - Fr64::one()) * Fr64::from_u64(1ull << mg->log_n).inverse() * acc).neg();
        rc = mgpu_run(mg, active, [&](uint32_t k) { return shard_phase2(mg->ctxs[k], row, nl, y, &mg->phase[k]); });
    }
    // always: drain the lanes, drop the table pins, release the per-device locks (on the threads that took them)
    mgpu_run(mg, active, [&](uint32_t k) {
        shard_phase_end(mg->ctxs[k], &mg->phase[k]);
        return (int)ZKP_OK;
    });
    if (rc) return rc;
    host::G1J com = mg->phase[0].com, proof = mg->phase[0].proof;
    for (uint32_t k = 1; k < active; k++) {
        com = com.add(mg->phase[k].com);
        proof = proof.add(mg->phase[k].proof);
    }
    host::g1_compress(commitment48, com);
    host::g1_compress(proof48, proof);
    y.to_be(eval_be);
    return ZKP_OK;
}

// Pianist-style distributed commitment + batch opening (BASELINE configs[4]; reference neurons/validator.py:212-222
// gives miner k the row rows[k] of the bivariate polynomial and checks every answer on its own, :168-170): `count`
// sub-polynomials of n evaluations each (polys_be: count x n x 32 bytes, row-major), opened at the common alpha.
// Sub-polynomial k runs on device k mod G (whole commit+open there); outputs per worker (what each miner would
// answer) plus the aggregated com = sum com_k and pi_X = sum pi_k of the master node, summed as Jacobian points on
// the calling thread and compressed once each.  zkp_master_open_y / zkp_master_verify (any of the contexts) finish
// the bivariate opening.  flags & ZKP_MGPU_RESIDENT (count <= G only): reuse the uploads of the previous call.
int zkp_mgpu_pianist_commit_open(zkp_mgpu* mg, const uint32_t* rows, size_t count, const uint8_t* polys_be, size_t n,
                                 const uint8_t alpha_be[32], int flags, uint8_t* commitments48, uint8_t* evals_be, uint8_t* proofs48,
                                 uint8_t agg_commitment48[48], uint8_t agg_proof48[48]) {
    int rc = mgpu_check(mg, ZKP_LAYOUT_ROWS);
    if (rc) return rc;
    if (!rows || !count || !polys_be || !alpha_be || !commitments48 || !evals_be || !proofs48) return fail(ZKP_ERR_ARG, "null argument");
    if (n != ((size_t)1 << mg->log_n)) return fail(ZKP_ERR_ARG, "every sub-polynomial needs exactly one SRS row of evaluations");
    const uint32_t G = (uint32_t)mg->ctxs.size();
    const bool resident = (flags & ZKP_MGPU_RESIDENT) != 0;
    if (resident && count > G) return fail(ZKP_ERR_ARG, "ZKP_MGPU_RESIDENT needs at most one sub-polynomial per device");
    Fr64 x;
    if (!Fr64::from_be(x, alpha_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    for (size_t k = 0; k < count; k++) {
        rc = check_row(mg->ctxs[k % G], rows[k], n);
        if (rc) return rc;
    }
    std::lock_guard<std::mutex> lk(mg->mu);
    const uint32_t active = count < G ? (uint32_t)count : G;
    std::vector<host::G1J> com(active, host::G1J::infinity()), prf(active, host::G1J::infinity());
    static const bool trace = getenv("ZKP_MGPU_TRACE") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    std::vector<double> t_start(active, 0), t_conv(active, 0), t_done(active, 0);
    auto since = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_call).count(); };
    rc = mgpu_run(mg, active, [&](uint32_t g) -> int {
        zkp_ctx* ctx = mg->ctxs[g];
        t_start[g] = since();
        std::lock_guard<std::mutex> lk2(ctx->mu);
        DeviceGuard dg(ctx->device);
        for (size_t k = g; k < count; k += G) {
            int r;
            if (resident) {
                if (ctx->resident_n != n) return fail(ZKP_ERR_STATE, "no polynomial resident on this device");
                r = convert_poly(ctx, n, true);
            } else {
                r = upload_poly(ctx, polys_be + 32 * n * k, n, true);
            }
            host::G1J cj, pj;
            t_conv[g] = since();
            if (!r) r = commit_open_resident(ctx, rows[k], n, x, commitments48 + 48 * k, evals_be + 32 * k, proofs48 + 48 * k, &cj, &pj);
            t_done[g] = since();
            if (r) return r;
            ctx->resident_n = n;
            com[g] = com[g].add(cj);
            prf[g] = prf[g].add(pj);
        }
        return ZKP_OK;
    });
    if (rc) return rc;
    host::G1J c = com[0], p = prf[0];
    for (uint32_t g = 1; g < active; g++) {
        c = c.add(com[g]);
        p = p.add(prf[g]);
    }
    const double t_joined = since();
    if (agg_commitment48) host::g1_compress(agg_commitment48, c);
    if (agg_proof48) host::g1_compress(agg_proof48, p);
    if (trace) {
        fprintf(stderr, "zkp_mgpu_pianist: joined %.0f us, total %.0f us;", t_joined, since());
        for (uint32_t g = 0; g < active; g++) fprintf(stderr, " dev%u start %.0f conv %.0f done %.0f;", g, t_start[g], t_conv[g], t_done[g]);
        fprintf(stderr, "\n");
    }
    return ZKP_OK;
}

}  // extern "C"
