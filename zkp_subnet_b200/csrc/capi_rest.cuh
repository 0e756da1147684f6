// C-ABI entries beyond the raw MSM: opening, fused commit+open, verify, fft, eval, RNG, SRS
// generation / files, wire codec and the benchmark entries.  Included at the end of zkp_b200.cu.
#pragma once
#include <chrono>
#include <random>
#include <thread>

namespace {

using host::Fr64;
using host::Fq64;

// ---- per-size domain tables ------------------------------------------------------------------
Fr64 fr_root_of_unity(uint32_t log_n) {
    // 7^((r-1) >> log_n)
    uint64_t e[4];
    memcpy(e, host::FR_MOD64, sizeof(e));
    e[0] -= 1;
    for (uint32_t s = 0; s < log_n; s++)
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i + 1 < 4 ? e[i + 1] << 63 : 0);
    return Fr64::from_u64(7).pow(e, 4);
}
Fr to_dev(const Fr64& a) {
    Fr r;
    memcpy(r.v, a.v, 32);
    return r;
}

int get_domain(zkp_ctx* ctx, uint32_t log_n, bool need_tw, zkp_ctx::Domain** out) {
    if (log_n >= ctx->domains.size()) return fail(ZKP_ERR_ARG, "domain too large");
    zkp_ctx::Domain& d = ctx->domains[log_n];
    if (d.ready && (!need_tw || d.have_tw || log_n < 1)) { *out = &d; return ZKP_OK; }  // tables are immutable once built
    std::lock_guard<std::recursive_mutex> lk(ctx->S.mu);  // shared with the forks of this context
    if (!d.ready) {
        d.w = fr_root_of_unity(log_n);
        d.w_inv = d.w.inverse();
        d.n_inv = Fr64::from_u64(1ull << log_n).inverse();
        std::vector<Fr64> wt(log_n + 1);
        wt[0] = d.w;
        for (uint32_t k = 1; k <= log_n; k++) wt[k] = wt[k - 1].sqr();
        ZKP_CUDA(d.wt.ensure(32 * (log_n + 1)));
        ZKP_CUDA(cudaMemcpyAsync(d.wt.p, wt.data(), 32 * (log_n + 1), cudaMemcpyHostToDevice, ctx->stream));
        ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
        d.ready = true;
    }
    if (need_tw && !d.have_tw && log_n >= 1) {
        uint32_t half = 1u << (log_n - 1);
        ZKP_CUDA(d.tw.ensure((size_t)half * 32));
        uint32_t threads = (half + 15) / 16;
        k_build_twiddles<<<(threads + 127) / 128, 128, 0, ctx->stream>>>(d.tw.as<Fr>(), half, d.wt.as<Fr>());
        ctx->launches++;
        ZKP_CUDA(cudaStreamSynchronize(ctx->stream));  // other contexts of the store use the table from their own streams
        d.have_tw = true;
    }
    *out = &d;
    return ZKP_OK;
}

// small device scratch layout (ctx->small): one request record, see kzg.cuh (SM_BAD, SM_HIT, SM_Y, SM_S1, SM_S2, SM_EVAL)

int ensure_small(zkp_ctx* ctx) {
    ZKP_CUDA(ctx->small.ensure(SM_BYTES));
    return ZKP_OK;
}
template <class T> T* small_at(zkp_ctx* ctx, size_t off) { return reinterpret_cast<T*>(ctx->small.as<uint8_t>() + off); }

// ---- opening on device: `count` polynomials f (Montgomery, n elements each, back to back) -> per request y (record
// r of `records`, offset SM_Y) and q (Montgomery, fr_c).  Evaluation points: x by value (count == 1) or d_xs[r].
int open_buffers(zkp_ctx* ctx, uint32_t n, uint32_t count) {
    ZKP_CUDA(ctx->fr_b.ensure((size_t)n * count * 32));
    ZKP_CUDA(ctx->fr_c.ensure((size_t)n * count * 32));
    return ZKP_OK;
}
int open_device(zkp_ctx* ctx, cudaStream_t st, const Fr* d_f, uint32_t n, const Fr64& x, uint32_t count = 1, const Fr* d_xs = nullptr,
                uint8_t* records = nullptr, const uint8_t* h_xs = nullptr, uint8_t* h_stage = nullptr) {
    uint32_t log_n = ilog2(n);
    zkp_ctx::Domain* dom;
    int rc = get_domain(ctx, log_n, false, &dom);
    if (rc) return rc;
    rc = open_buffers(ctx, n, count);
    if (rc) return rc;
    if (!records) records = ctx->small.as<uint8_t>();
    auto rec = [&](size_t off) { return records + off; };
    // elements per thread: with one inversion per BLOCK (k_open_pass1) short runs cost nothing extra
    uint32_t E = n >> 16;
    if (E < 4) E = 4;
    if (E > 16) E = 16;
    uint32_t threads = (n + E - 1) / E, blocks = (threads + 127) / 128;
    uint32_t blocks2 = (n + 255) / 256;
    ZKP_CUDA(ctx->partials.ensure((size_t)(blocks > blocks2 ? blocks : blocks2) * count * 32));
    uint32_t* hit = reinterpret_cast<uint32_t*>(rec(SM_HIT));
    if (count == 1) ZKP_CUDA(cudaMemsetAsync(hit, 0xff, 4, st));  // a batch initialises its records itself (k_batch_init)
    if (n >= 128 * E && ctx->open_coset && (count == 1 ? !d_xs : h_xs != nullptr)) {
        // The host knows every x, so it can tell whether one lies in the domain (x^n = 1: the general kernels below)
        // and, if none does, supply 1/(x^n - 1) -- pass 1 then runs on cosets with no inversion on the device
        // (k_open_pass1_coset), y needs no squarings on the device, and the three kernels of the in-domain fix are
        // not launched at all.  A batch hands its evaluation points over as h_xs (Montgomery, 32 bytes each) and
        // a page-locked staging area for the per-request constants.
        const uint32_t log_m = ilog2(128 * E), log_S = log_n - log_m, S = 1u << log_S;
        std::vector<Fr64> hv(3 * (size_t)count);
        bool outside = true;
        for (uint32_t r = 0; r < count && outside; r++) {
            Fr64 xm;
            if (count == 1) xm = x;
            else memcpy(xm.v, h_xs + 32 * (size_t)r, 32);
            for (uint32_t k = 0; k < log_m; k++) xm = xm.sqr();
            Fr64 xn = xm;
            for (uint32_t k = 0; k < log_S; k++) xn = xn.sqr();
            if (xn == Fr64::one()) { outside = false; break; }
            const Fr64 xn1 = xn - Fr64::one();
            hv[3 * r] = xm;
            hv[3 * r + 1] = xn1.inverse();
            hv[3 * r + 2] = xn1 * dom->n_inv;
        }
        if (outside) {
            Fr64 g_inv = dom->w_inv;
            for (uint32_t k = 0; k < log_S; k++) g_inv = g_inv.sqr();
            const size_t need = ((size_t)2 * S * count + 3 * (size_t)count) * 32, need2 = (size_t)blocks2 * count * 32;
            ZKP_CUDA(ctx->partials.ensure(need > need2 ? need : need2));
            Fr* part = ctx->partials.as<Fr>();
            Fr* inv_blocks = part + (size_t)S * count;
            Fr* d_hv = nullptr;
            if (count > 1) {
                d_hv = inv_blocks + (size_t)S * count;
                memcpy(h_stage, hv.data(), hv.size() * 32);
                ZKP_CUDA(cudaMemcpyAsync(d_hv, h_stage, hv.size() * 32, cudaMemcpyHostToDevice, st));
            }
            k_open_coset_inv<<<dim3(1, count), S < COSET_INV_THREADS ? S : COSET_INV_THREADS, 0, st>>>(
                to_dev(hv[0]), to_dev(hv[1]), dom->wt.as<Fr>(), log_m, S, inv_blocks, d_hv);
            k_open_pass1_coset<<<dim3(S, count), 128, 0, st>>>(d_f, E, log_S, to_dev(x), dom->wt.as<Fr>(), to_dev(g_inv), inv_blocks,
                                                              ctx->fr_b.as<Fr>(), part, count > 1 ? d_xs : nullptr);
            trace_mark(ctx, 1, st, "open_pass1");
            k_open_reduce_y<<<dim3(1, count), 256, 0, st>>>(part, S, to_dev(hv[2]), reinterpret_cast<Fr*>(rec(SM_S1)),
                                                            reinterpret_cast<Fr*>(rec(SM_Y)), d_hv);
            trace_mark(ctx, 1, st, "open_y");
            k_open_pass2<<<dim3(blocks2, count), 256, 0, st>>>(d_f, ctx->fr_b.as<Fr>(), n, reinterpret_cast<Fr*>(rec(SM_Y)), ctx->fr_c.as<Fr>());
            trace_mark(ctx, 1, st, "open_pass2");
            ctx->launches += 4;
            ZKP_CUDA(cudaGetLastError());
            return ZKP_OK;
        }
    }
    k_open_pass1<<<dim3(blocks, count), 128, 0, st>>>(d_f, n, E, to_dev(x), dom->wt.as<Fr>(), to_dev(dom->w_inv), ctx->fr_b.as<Fr>(),
                                                      ctx->partials.as<Fr>(), hit, 0, d_xs);
    trace_mark(ctx, 1, st, "open_pass1");
    k_fr_reduce<<<dim3(1, count), 256, 0, st>>>(ctx->partials.as<Fr>(), blocks, reinterpret_cast<Fr*>(rec(SM_S1)));
    k_open_y<<<dim3(1, count), 32, 0, st>>>(d_f, log_n, to_dev(x), to_dev(dom->n_inv), reinterpret_cast<Fr*>(rec(SM_S1)), hit,
                                            reinterpret_cast<Fr*>(rec(SM_Y)), d_xs, n);
    trace_mark(ctx, 1, st, "open_y");
    k_open_pass2<<<dim3(blocks2, count), 256, 0, st>>>(d_f, ctx->fr_b.as<Fr>(), n, reinterpret_cast<Fr*>(rec(SM_Y)), ctx->fr_c.as<Fr>());
    trace_mark(ctx, 1, st, "open_pass2");
    // x in the domain (rare): q_m = -sum_{j != m} q_j w^(j-m); the kernels are no-ops otherwise
    k_open_fix_partial<<<dim3(blocks2, count), 256, 0, st>>>(ctx->fr_c.as<Fr>(), n, dom->wt.as<Fr>(), hit, ctx->partials.as<Fr>());
    k_fr_reduce<<<dim3(1, count), 256, 0, st>>>(ctx->partials.as<Fr>(), blocks2, reinterpret_cast<Fr*>(rec(SM_S2)));
    k_open_fix_apply<<<dim3(1, count), 32, 0, st>>>(ctx->fr_c.as<Fr>(), hit, reinterpret_cast<Fr*>(rec(SM_S2)), n);
    ctx->launches += 7;
    ZKP_CUDA(cudaGetLastError());
    return ZKP_OK;
}

// upload poly (big-endian), convert to Montgomery in fr_a; leaves the raw bytes in ctx->scalars
// worker_poly: the bytes are the `poly` of a worker_commit / worker_open call, i.e. subject to zkp_set_poly_form;
// everything else (fft, eval, challenge rows, shard slices) is taken as it comes
int convert_poly(zkp_ctx* ctx, size_t n, bool worker_poly = false);
int ntt_device(zkp_ctx* ctx, const Fr* in, Fr* out, uint32_t log_n, int inverse);
int upload_poly(zkp_ctx* ctx, const uint8_t* poly_be, size_t n, bool worker_poly = false) {
    int rc = upload_scalars(ctx, poly_be, n, ctx->scalars);
    if (rc) return rc;
    return convert_poly(ctx, n, worker_poly);
}
// raw big-endian bytes in ctx->scalars -> Montgomery form in fr_a (flags non-canonical elements).  With
// zkp_set_poly_form(ctx, 1) the bytes are COEFFICIENTS and fr_a receives their evaluations (one forward NTT): everything
// downstream works on evaluations either way.
int convert_poly(zkp_ctx* ctx, size_t n, bool worker_poly) {
    int rc = ensure_small(ctx);
    if (rc) return rc;
    ZKP_CUDA(ctx->fr_a.ensure(n * 32));
    ZKP_CUDA(cudaMemsetAsync(small_at<uint32_t>(ctx, SM_BAD), 0, 4, ctx->stream));
    k_fr_from_be<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->scalars.as<uint32_t>(), n, ctx->fr_a.as<Fr>(),
                                                                       small_at<uint32_t>(ctx, SM_BAD));
    ctx->launches++;
    if (ctx->coeff_form && worker_poly) {
        if (!is_pow2(n)) return fail(ZKP_ERR_ARG, "a polynomial in coefficient form must have a power-of-two length");
        return ntt_device(ctx, ctx->fr_a.as<Fr>(), ctx->fr_a.as<Fr>(), ilog2(n), 0);
    }
    return ZKP_OK;
}
// the scalars of the commitment MSM: the raw bytes as uploaded (evaluation form) or the NTT output (coefficient form)
const uint32_t* commit_scalars(zkp_ctx* ctx, int* fmt) {
    *fmt = ctx->coeff_form ? SCALAR_MONT : SCALAR_BE;
    return ctx->coeff_form ? ctx->fr_a.as<uint32_t>() : ctx->scalars.as<uint32_t>();
}

// read back y (big-endian) and the bad-encoding flag
int fetch_y_enqueue(zkp_ctx* ctx, cudaStream_t st) {
    k_fr_to_be<<<1, 32, 0, st>>>(small_at<Fr>(ctx, SM_Y), 1, small_at<uint32_t>(ctx, SM_EVAL));
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small, small_at<uint8_t>(ctx, SM_EVAL), 32, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 64, small_at<uint8_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost, st));
    return ZKP_OK;
}
int fetch_y_finish(zkp_ctx* ctx, uint8_t eval_be[32]) {  // the stream has been synchronised by the caller
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 64)) return fail(ZKP_ERR_ENCODING, "polynomial holds a non-canonical field element");
    memcpy(eval_be, ctx->h_small, 32);
    return ZKP_OK;
}
int fetch_y(zkp_ctx* ctx, uint8_t eval_be[32]) {
    int rc = fetch_y_enqueue(ctx, ctx->stream);
    if (rc) return rc;
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    return fetch_y_finish(ctx, eval_be);
}

int open_checks(zkp_ctx* ctx, uint32_t i, const void* poly, size_t n, const uint8_t* x_be, Fr64* x) {
    int rc = check_row(ctx, i, n);
    if (rc) return rc;
    if (!poly || !x_be) return fail(ZKP_ERR_ARG, "null argument");
    if (n != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "opening needs exactly one SRS row of evaluations");
    if (ctx->shard_domain_log != ctx->log_n)
        return fail(ZKP_ERR_STATE, "opening is not available on a point-range shard (commit partials only)");
    if (!Fr64::from_be(*x, x_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    return ZKP_OK;
}

// The full device part of commit (optional) + open with the polynomial already resident (raw bytes in
// ctx->scalars, Montgomery form in fr_a, both produced on lane 0).
//
// Two shapes, same bytes out:
//  * two lanes (large n): the commitment MSM runs on lane 0, the opening (field kernels + its MSM) on lane 1; the host
//    folds the commitment while lane 1 is still busy.  Each accumulation fills the machine for milliseconds, and what
//    the second stream buys is the tail of one MSM hidden under the accumulation of the other.
//  * fused (small n; zkp_set_fuse): the two MSMs are two GROUPS of one launch set -- one sort, one accumulation grid,
//    one chain of slot levels, one row/column reduction, one bit-plane launch, one read-back.  At the mainnet row size
//    (2^16) half of a lone MSM is latency-bound tail (ten dependent slot levels, reduction launches over few buckets),
//    and a commit+open pays for that tail once instead of twice.  The commitment's digits are counted on lane 0 while
//    the opening's field kernels (which produce the scalars of the proof group) run on lane 1.
bool fuse_wanted(zkp_ctx* ctx, size_t n) {
    if (ctx->fuse_mode >= 0) return ctx->fuse_mode != 0;
    return n <= ((size_t)1 << ZKP_FUSE_MAX_LOG);
}
int commit_open_fused(zkp_ctx* ctx, uint32_t i, size_t n, const Fr64& x, uint8_t* commitment48, uint8_t eval_be[32], uint8_t proof48[48],
                      host::G1J* com_jac, host::G1J* proof_jac, bool* done) {
    *done = false;
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream2;
    int rc = open_buffers(ctx, (uint32_t)n, 1);
    if (rc) return rc;
    int cfmt;
    const uint32_t* csc = commit_scalars(ctx, &cfmt);
    MsmJob jobs[2] = {{i, csc, cfmt}, {i, ctx->fr_c.as<uint32_t>(), SCALAR_MONT}};
    MsmPlan plan;
    MsmGroups gs;
    const G1Affine* pts = nullptr;
    bool grouped = false;
    rc = msm_plan_jobs(ctx, jobs, 2, n, &plan, &gs, &pts, &grouped);
    if (rc) return rc;
    if (!grouped) return ZKP_OK;  // no tables: the caller takes the two-lane path
    *done = true;
    ZKP_CUDA(cudaEventRecord(ctx->ev_ready, s0));
    ZKP_CUDA(cudaStreamWaitEvent(s1, ctx->ev_ready, 0));
    rc = msm_prep_begin(ctx, 0, plan);
    if (!rc) rc = msm_prep_count(ctx, 0, plan, gs, 0, 1);
    trace_mark(ctx, 1, s1, "open_begin");
    if (!rc) rc = open_device(ctx, s1, ctx->fr_a.as<Fr>(), (uint32_t)n, x);
    trace_mark(ctx, 1, s1, "open_field_kernels");
    if (!rc) rc = fetch_y_enqueue(ctx, s1);  // ordered before ev_join, i.e. before everything lane 0 does from here on
    if (!rc) {
        ZKP_CUDA(cudaEventRecord(ctx->ev_join, s1));
        ZKP_CUDA(cudaStreamWaitEvent(s0, ctx->ev_join, 0));
        rc = msm_prep_count(ctx, 0, plan, gs, 1, 1);
    }
    if (!rc) rc = msm_prep_finish(ctx, 0, plan, gs);
    if (!rc) rc = msm_enqueue_main(ctx, 0, plan, pts);
    if (!rc) rc = msm_wait(ctx, 0, 2);
    else { cudaStreamSynchronize(s0); cudaStreamSynchronize(s1); }
    msm_unpin_all(ctx);
    if (rc) return rc;
    const G1Xyzz* hw = ctx->ws.h_window;
    host::G1J c = msm_fold_group(plan, hw, 0), p = msm_fold_group(plan, hw, 1);
    if (commitment48) host::g1_compress(commitment48, c);
    if (proof48) host::g1_compress(proof48, p);
    if (com_jac) *com_jac = c;
    if (proof_jac) *proof_jac = p;
    ctx->last_com = c;
    ctx->last_proof = p;
    ctx->have_last = true;
    return fetch_y_finish(ctx, eval_be);
}

int commit_open_resident(zkp_ctx* ctx, uint32_t i, size_t n, const Fr64& x, uint8_t* commitment48, uint8_t eval_be[32],
                         uint8_t proof48[48], host::G1J* com_jac = nullptr, host::G1J* proof_jac = nullptr) {
    int rc;
    const bool want_com = commitment48 || com_jac;
    if (want_com && fuse_wanted(ctx, n) && tables_wanted(ctx, n)) {
        bool done = false;
        rc = commit_open_fused(ctx, i, n, x, commitment48, eval_be, proof48, com_jac, proof_jac, &done);
        if (rc || done) return rc;
    }
    MsmPlan plan_c, plan_o;
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream2;
    ZKP_CUDA(cudaEventRecord(ctx->ev_ready, s0));
    ZKP_CUDA(cudaStreamWaitEvent(s1, ctx->ev_ready, 0));
    // Front halves of both lanes are enqueued first (digits + sort of the commitment; opening field kernels, digits +
    // sort of the proof), then the two accumulation/reduction back halves.  Measured alternatives at 2^20
    // (tools/variant_bench.py): holding the first accumulation until the proof's front half is done 13.14 ms,
    // a higher stream priority for the lane that accumulates first 13.26 ms, both 13.53 ms, neither 12.95 ms --
    // the front half of lane 1 and the reduction tail of lane 0 are real work; hiding them under an accumulation
    // slows that accumulation by about as much as running them in the open would cost.
    const G1Affine *pts_c = nullptr, *pts_o = nullptr;
    auto bail = [&](int code) {
        cudaStreamSynchronize(s0);
        cudaStreamSynchronize(s1);
        msm_unpin_all(ctx);
        return code;
    };
    if (want_com) {
        int cfmt;
        const uint32_t* csc = commit_scalars(ctx, &cfmt);
        rc = msm_device_prep(ctx, 0, i, csc, cfmt, n, &plan_c, &pts_c);
        if (rc) return bail(rc);
    }
    trace_mark(ctx, 1, s1, "open_begin");
    rc = open_device(ctx, s1, ctx->fr_a.as<Fr>(), (uint32_t)n, x);
    if (rc) return bail(rc);
    trace_mark(ctx, 1, s1, "open_field_kernels");
    // y is final here: its conversion and copy go in front of the proof MSM, not behind its last kernel
    rc = fetch_y_enqueue(ctx, s1);
    if (rc) return bail(rc);
    rc = msm_device_prep(ctx, 1, i, ctx->fr_c.as<uint32_t>(), SCALAR_MONT, n, &plan_o, &pts_o);
    if (rc) return bail(rc);
    ctx->two_lanes_busy = want_com;
    if (want_com) {
        rc = msm_enqueue_main(ctx, 0, plan_c, pts_c);
        if (rc) { ctx->two_lanes_busy = false; return bail(rc); }
    }
    rc = msm_enqueue_main(ctx, 1, plan_o, pts_o);
    ctx->two_lanes_busy = false;
    if (rc) return bail(rc);
    host::G1J cj = host::G1J::infinity(), pj;
    if (want_com) {
        rc = msm_device_finish(ctx, 0, plan_c, commitment48, &cj);
        if (rc) return bail(rc);
    }
    rc = msm_device_finish(ctx, 1, plan_o, proof48, &pj);
    msm_unpin_all(ctx);
    if (rc) return rc;
    if (com_jac) *com_jac = cj;
    if (proof_jac) *proof_jac = pj;
    ctx->last_com = cj;
    ctx->last_proof = pj;
    ctx->have_last = want_com;
    return fetch_y_finish(ctx, eval_be);
}

void flush_l2(zkp_ctx* ctx) {
    const size_t bytes = 256ull << 20;  // > 126 MB L2
    if (ctx->flush.ensure(bytes) == cudaSuccess) cudaMemsetAsync(ctx->flush.p, 0x5a, bytes, ctx->stream);
}

// NTT on device buffers (Montgomery form).  out may alias in.
int ntt_device(zkp_ctx* ctx, const Fr* in, Fr* out, uint32_t log_n, int inverse) {
    if (log_n == 0) {
        if (in != out) ZKP_CUDA(cudaMemcpyAsync(out, in, 32, cudaMemcpyDeviceToDevice, ctx->stream));
        return ZKP_OK;
    }
    if (log_n > 2 * NTT_MAX_TILE_LOG) return fail(ZKP_ERR_ARG, "NTT size above 2^24 not supported");
    zkp_ctx::Domain* dom;
    int rc = get_domain(ctx, log_n, true, &dom);
    if (rc) return rc;
    const Fr n_inv = to_dev(dom->n_inv);
    // shared memory of a pass: the tile (32 B per element) + the butterfly twiddles of the sub-transform (m/2 x 32 B)
    auto smem_for = [](uint32_t log_tile, uint32_t log_m, bool tma = false) -> size_t {
        return ((size_t)32 << log_tile) + ((size_t)16 << log_m) + (tma ? (size_t)32 << log_tile : 0);
    };
    auto sub_table = [&](uint32_t log_m, const Fr** out_tw) -> int {
        if (log_m == 0) { *out_tw = dom->tw.as<Fr>(); return ZKP_OK; }
        zkp_ctx::Domain* d;
        int r = get_domain(ctx, log_m, true, &d);
        if (r) return r;
        *out_tw = d->tw.as<Fr>();
        return ZKP_OK;
    };
    auto threads_for = [](uint32_t log_tile) -> unsigned {
        unsigned t = log_tile > 2 ? 1u << (log_tile - 2) : 1u;  // one radix-4 group per thread and stage pair
        if (t < 32) t = 32;
        if (t > (unsigned)NTT_MAX_THREADS) t = NTT_MAX_THREADS;
        return t;
    };
    // columns per CTA: aim at 1024-element tiles (32 KB: 7-8 CTAs per SM) but keep >= 512 CTAs in flight
    auto cols_for = [](uint32_t log_m, uint32_t log_ncols) -> uint32_t {
        uint32_t lc = log_m < 10 ? 10 - log_m : 0;
        if (lc > log_ncols) lc = log_ncols;
        while (lc > 0 && log_ncols - lc < 9) lc--;
        return lc;
    };
    if (log_n <= NTT_MAX_TILE_LOG) {
        NttPass p = {log_n, 0, 1, 1, 0, 1, 0, 1, log_n, 0, (uint32_t)inverse, (uint32_t)inverse, 0};
        k_ntt_pass<<<1, threads_for(log_n), smem_for(log_n, log_n), ctx->stream>>>(in, out, dom->tw.as<Fr>(), dom->tw.as<Fr>(), p, n_inv);
        ctx->launches++;
        return ZKP_OK;
    }
    uint32_t l1 = log_n / 2, l2 = log_n - l1;
    uint64_t n1 = 1ull << l1, n2 = 1ull << l2;
    const Fr *tw1, *tw2;
    rc = sub_table(l1, &tw1);
    if (rc) return rc;
    rc = sub_table(l2, &tw2);
    if (rc) return rc;
    ZKP_CUDA(ctx->ntt_tmp.ensure(32ull << log_n));
    Fr* tmp = ctx->ntt_tmp.as<Fr>();
    uint32_t lc1 = cols_for(l1, l2), lc2 = cols_for(l2, l1);
    // pass 1: columns i2 (n2 of them), rows i1; element (r, c) at r*n2 + c; twiddle w^(c*k)
    NttPass p1 = {l1, lc1, (uint32_t)n2, n2, 1, n2, 1, 0, log_n, 1, (uint32_t)inverse, 0, 0};
    k_ntt_pass<<<(unsigned)(n2 >> lc1), threads_for(l1 + lc1), smem_for(l1 + lc1, l1), ctx->stream>>>(in, tmp, dom->tw.as<Fr>(), tw1, p1, n_inv);
    // pass 2: columns k1 (n1 of them), rows i2; element (r, c) at c*n2 + r; output (k2, c) at k2*n1 + c
    // pass 2 reads one contiguous column per CTA when lc2 == 0: the only tile of this transform that a 1-D bulk copy
    // (TMA) can fetch; off by default (zkp_set_ntt_tma, measured slower: DESIGN.md section 4)
    const bool tma2 = ctx->ntt_tma && lc2 == 0 && l2 <= 11;
    NttPass p2 = {l2, lc2, (uint32_t)n1, 1, n2, n1, 1, 1, log_n, 0, (uint32_t)inverse, (uint32_t)inverse, tma2 ? 1u : 0u};
    k_ntt_pass<<<(unsigned)(n1 >> lc2), threads_for(l2 + lc2), smem_for(l2 + lc2, l2, tma2), ctx->stream>>>(tmp, out, dom->tw.as<Fr>(), tw2, p2, n_inv);
    ctx->launches += 2;
    return ZKP_OK;
}

// fixed-base table [d * 256^w]G, d < 256, w < 32 (affine, Montgomery), built on the host once
int ensure_fixed_base(zkp_ctx* ctx) {
    if (ctx->fixed_base.p) return ZKP_OK;
    using namespace host;
    std::vector<G1J> jac(32 * 256);
    G1J base = g1_generator();
    for (int w = 0; w < 32; w++) {
        jac[w * 256] = G1J::infinity();
        for (int d = 1; d < 256; d++) jac[w * 256 + d] = jac[w * 256 + d - 1].add(base);
        base = jac[w * 256 + 255].add(base);
    }
    // batch to affine
    std::vector<Fq64> pre(jac.size());
    Fq64 run = Fq64::one();
    for (size_t k = 0; k < jac.size(); k++) {
        pre[k] = run;
        if (!jac[k].is_inf()) run = run * jac[k].z;
    }
    Fq64 inv = run.inverse();
    std::vector<uint8_t> tab(jac.size() * 96, 0);
    for (size_t k = jac.size(); k-- > 0;) {
        if (jac[k].is_inf()) continue;
        Fq64 zi = inv * pre[k];
        inv = inv * jac[k].z;
        Fq64 zi2 = zi.sqr();
        Fq64 ax = jac[k].x * zi2, ay = jac[k].y * zi2 * zi;
        memcpy(&tab[k * 96], ax.v, 48);
        memcpy(&tab[k * 96 + 48], ay.v, 48);
    }
    ZKP_CUDA(ctx->fixed_base.ensure(tab.size()));
    ZKP_CUDA(cudaMemcpy(ctx->fixed_base.p, tab.data(), tab.size(), cudaMemcpyHostToDevice));
    return ZKP_OK;
}

void set_pairing_lines(zkp_ctx* ctx) {
    ctx->lines_g2 = host::g2_precompute(host::g2_generator());
    ctx->lines_tau = host::g2_precompute(ctx->g2_tau);
    if (ctx->have_g2_tau_y) ctx->lines_tau_y = host::g2_precompute(ctx->g2_tau_y);
    ctx->have_lines = true;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------- SRS
int zkp_srs_generate(zkp_ctx* ctx, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32], uint32_t log_n,
                     uint32_t log_machines) {
    return zkp_srs_generate_shard(ctx, tau_x_be, tau_y_be, log_n, log_machines, 0, 0);
}

// Point-range shard `shard` of 2^log_shards: every row holds the points j in
// [shard * n/S, (shard+1) * n/S) of the size-n Lagrange SRS (n = 2^log_n), so an MSM over the row with
// the matching slice of scalars is this GPU's partial commitment.
int zkp_srs_generate_shard(zkp_ctx* ctx, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32], uint32_t log_n,
                           uint32_t log_machines, uint32_t shard, uint32_t log_shards) {
    if (!ctx || !tau_x_be || !tau_y_be) return fail(ZKP_ERR_ARG, "null argument");
    if (log_shards > log_n || shard >= (1u << log_shards)) return fail(ZKP_ERR_ARG, "bad shard");
    Fr64 tx, ty;
    if (!Fr64::from_be(tx, tau_x_be) || !Fr64::from_be(ty, tau_y_be)) return fail(ZKP_ERR_ENCODING, "trapdoor not canonical");
    const uint32_t log_local = log_n - log_shards;
    int rc = zkp_srs_set_shape(ctx, log_local, log_machines);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = ensure_small(ctx);
    if (rc) return rc;
    rc = ensure_fixed_base(ctx);
    if (rc) return rc;
    const uint32_t n = 1u << log_local, M = 1u << log_machines;
    const uint64_t j0 = (uint64_t)shard << log_local;
    cudaStream_t st = ctx->stream;
    zkp_ctx::Domain* dom;
    rc = get_domain(ctx, log_n, false, &dom);
    if (rc) return rc;
    // inv_d[j] = 1/(w^(j0+j) - tau_x) in fr_b
    ZKP_CUDA(ctx->fr_b.ensure((size_t)n * 32));
    ZKP_CUDA(ctx->fr_c.ensure((size_t)n * 32));
    uint32_t E = 16, threads = (n + E - 1) / E, blocks = (threads + 127) / 128;
    ZKP_CUDA(ctx->partials.ensure((size_t)blocks * 32));
    ZKP_CUDA(cudaMemsetAsync(small_at<uint32_t>(ctx, SM_HIT), 0xff, 4, st));
    k_open_pass1<<<blocks, 128, 0, st>>>(nullptr, n, E, to_dev(tx), dom->wt.as<Fr>(), to_dev(dom->w_inv), ctx->fr_b.as<Fr>(),
                                         ctx->partials.as<Fr>(), small_at<uint32_t>(ctx, SM_HIT), j0);
    ctx->launches++;
    uint32_t hit = 0;
    ZKP_CUDA(cudaMemcpyAsync(&hit, small_at<uint32_t>(ctx, SM_HIT), 4, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaStreamSynchronize(st));
    if (hit != HIT_NONE) return fail(ZKP_ERR_ARG, "tau_x lies in the evaluation domain");
    // zn = (tau_x^n - 1)/n over the FULL domain
    Fr64 zn = tx;
    for (uint32_t k = 0; k < log_n; k++) zn = zn.sqr();
    zn = (zn - Fr64::one()) * dom->n_inv;
    // R_i(tau_y) over the size-M domain
    std::vector<Fr64> R(M);
    if (M == 1) {
        R[0] = Fr64::one();
    } else {
        Fr64 wM = fr_root_of_unity(log_machines), zm = ty;
        for (uint32_t k = 0; k < log_machines; k++) zm = zm.sqr();
        zm = (zm - Fr64::one()) * Fr64::from_u64(M).inverse();
        Fr64 wi = Fr64::one();
        for (uint32_t i = 0; i < M; i++) {
            Fr64 d = ty - wi;
            if (d.is_zero()) return fail(ZKP_ERR_ARG, "tau_y lies in the machine domain");
            R[i] = zm * wi * d.inverse();
            wi = wi * wM;
        }
    }
    // rows
    size_t xyzz_bytes = (size_t)n * sizeof(G1Xyzz);
    ZKP_CUDA(ctx->ws.buckets.ensure(xyzz_bytes));
    ZKP_CUDA(ctx->ws.pool.ensure((size_t)n * sizeof(Fq)));
    for (uint32_t i = 0; i < M; i++) {
        Fr64 coef = (zn * R[i]).neg();
        k_lagrange_scalars<<<((n + 7) / 8 + 127) / 128, 128, 0, st>>>(ctx->fr_b.as<Fr>(), n, dom->wt.as<Fr>(), to_dev(coef),
                                                                      ctx->fr_c.as<Fr>(), j0);
        k_fixed_base_mul<<<(n + 127) / 128, 128, 0, st>>>(ctx->fr_c.as<Fr>(), n, ctx->fixed_base.as<G1Affine>(),
                                                          ctx->ws.buckets.as<G1Xyzz>());
        uint32_t EA = 16, ta = (n + EA - 1) / EA;
        k_xyzz_to_affine<<<(ta + 127) / 128, 128, 0, st>>>(ctx->ws.buckets.as<G1Xyzz>(), n, EA, ctx->ws.pool.as<Fq>(),
                                                           ctx->srs.as<G1Affine>() + ((size_t)i << log_local));
        ctx->launches += 3;
        Fr64 rc_canon = R[i].from_mont();
        ctx->scale_points[i] = host::g1_generator().mul(rc_canon.v, 4);
        ctx->row_loaded[i] = 1;
    }
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    Fr64 txc = tx.from_mont();
    ctx->g2_tau = host::g2_generator().mul(txc.v, 4);
    ctx->have_g2_tau = true;
    Fr64 tyc = ty.from_mont();
    ctx->g2_tau_y = host::g2_generator().mul(tyc.v, 4);
    ctx->have_g2_tau_y = true;
    set_pairing_lines(ctx);
    ctx->shard_domain_log = log_n;
    ctx->shard_index = shard;
    return ZKP_OK;
}

// Monomial SRS [tau^j]_1, j < 2^log_n, as the single row of the context (BASELINE configs[2] "path B": iNTT of
// the evaluations, then an MSM over the monomial basis, must give the bytes of the Lagrange-basis commitment).
int zkp_srs_generate_monomial(zkp_ctx* ctx, const uint8_t tau_x_be[32], uint32_t log_n) {
    if (!ctx || !tau_x_be) return fail(ZKP_ERR_ARG, "null argument");
    Fr64 tx;
    if (!Fr64::from_be(tx, tau_x_be)) return fail(ZKP_ERR_ENCODING, "trapdoor not canonical");
    int rc = zkp_srs_set_shape(ctx, log_n, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = ensure_fixed_base(ctx);
    if (rc) return rc;
    const uint32_t n = 1u << log_n;
    cudaStream_t st = ctx->stream;
    std::vector<Fr64> tt(32);
    tt[0] = tx;
    for (size_t k = 1; k < tt.size(); k++) tt[k] = tt[k - 1].sqr();
    ZKP_CUDA(ctx->partials.ensure(tt.size() * 32));
    ZKP_CUDA(cudaMemcpyAsync(ctx->partials.p, tt.data(), tt.size() * 32, cudaMemcpyHostToDevice, st));
    ZKP_CUDA(cudaStreamSynchronize(st));  // tt is a host stack object
    ZKP_CUDA(ctx->fr_c.ensure((size_t)n * 32));
    ZKP_CUDA(ctx->ws.buckets.ensure((size_t)n * sizeof(G1Xyzz)));
    ZKP_CUDA(ctx->ws.pool.ensure((size_t)n * sizeof(Fq)));
    k_power_scalars<<<((n + 7) / 8 + 127) / 128, 128, 0, st>>>(ctx->partials.as<Fr>(), n, ctx->fr_c.as<Fr>(), to_dev(Fr64::one()));
    k_fixed_base_mul<<<(n + 127) / 128, 128, 0, st>>>(ctx->fr_c.as<Fr>(), n, ctx->fixed_base.as<G1Affine>(), ctx->ws.buckets.as<G1Xyzz>());
    uint32_t EA = 16, ta = (n + EA - 1) / EA;
    k_xyzz_to_affine<<<(ta + 127) / 128, 128, 0, st>>>(ctx->ws.buckets.as<G1Xyzz>(), n, EA, ctx->ws.pool.as<Fq>(), ctx->srs.as<G1Affine>());
    ctx->launches += 3;
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    ctx->scale_points[0] = host::g1_generator();
    ctx->row_loaded[0] = 1;
    Fr64 txc = tx.from_mont();
    ctx->g2_tau = host::g2_generator().mul(txc.v, 4);
    ctx->have_g2_tau = true;
    ctx->have_g2_tau_y = false;
    set_pairing_lines(ctx);
    ctx->shard_domain_log = log_n;
    ctx->shard_index = 0;
    return ZKP_OK;
}

// sum of compressed G1 points (the only cross-GPU step of a sharded / Pianist commitment: N partial
// points, 48 bytes each, combined on the host of rank 0)
int zkp_g1_sum(const uint8_t* points48, size_t count, uint8_t out48[48]) {
    if (!points48 || !out48) return fail(ZKP_ERR_ARG, "null argument");
    host::G1J acc = host::G1J::infinity();
    for (size_t k = 0; k < count; k++) {
        host::G1J p;
        if (!host::g1_decompress(p, points48 + 48 * k, false)) return fail(ZKP_ERR_ENCODING, "bad G1 point");
        acc = acc.add(p);
    }
    host::g1_compress(out48, acc);
    return ZKP_OK;
}

// The same sum for inputs that come from OTHER parties (Client.master_commit / master_open aggregate what workers sent):
// every point is checked to lie in the prime-order subgroup before it enters the aggregate.  zkp_g1_sum stays the
// unchecked variant for partials produced by this process.
int zkp_g1_sum_checked(const uint8_t* points48, size_t count, uint8_t out48[48]) {
    if (!points48 || !out48) return fail(ZKP_ERR_ARG, "null argument");
    // decompression + subgroup check (~80 us per point) on the codec's host threads, ~4 points each; partial sums per thread
    unsigned cores = std::thread::hardware_concurrency();
    if (cores == 0) cores = 1;
    unsigned nth = (unsigned)(count / 4);
    if (nth > cores) nth = cores;
    if (nth > 16) nth = 16;
    if (nth < 1) nth = 1;
    std::vector<host::G1J> part(nth, host::G1J::infinity());
    std::atomic<size_t> first_bad(count);
    codec::parallel_ranges_n(nth, nth, [&](size_t tlo, size_t thi) {
        for (size_t t = tlo; t < thi; t++) {
            host::G1J acc = host::G1J::infinity();
            for (size_t k = count * t / nth; k < count * (t + 1) / nth; k++) {
                host::G1J p;
                if (!host::g1_decompress(p, points48 + 48 * k, true)) {
                    size_t cur = first_bad.load();
                    while (k < cur && !first_bad.compare_exchange_weak(cur, k)) {}
                    break;
                }
                acc = acc.add(p);
            }
            part[t] = acc;
        }
    });
    if (first_bad.load() != count)
        return fail(ZKP_ERR_ENCODING, "point " + std::to_string(first_bad.load()) + " is malformed, off the curve or outside the subgroup");
    host::G1J acc = host::G1J::infinity();
    for (const auto& pt : part) acc = acc.add(pt);
    host::g1_compress(out48, acc);
    return ZKP_OK;
}

// The same combine without the square roots: every rank expands its own partial to the 96-byte ZCash
// uncompressed form (one sqrt each, in parallel on the ranks), rank 0 adds affine points (8 x 2 decompressions
// were 0.8 ms of a 0.95 ms combine at 8 GPUs).
int zkp_g1_uncompress(const uint8_t in48[48], uint8_t out96[96]) {
    if (!in48 || !out96) return fail(ZKP_ERR_ARG, "null argument");
    host::G1J p;
    if (!host::g1_decompress(p, in48, false)) return fail(ZKP_ERR_ENCODING, "bad G1 point");
    host::g1_serialize96(out96, p);
    return ZKP_OK;
}
int zkp_g1_sum_uncompressed(const uint8_t* points96, size_t count, uint8_t out48[48]) {
    if (!points96 || !out48) return fail(ZKP_ERR_ARG, "null argument");
    host::G1J acc = host::G1J::infinity();
    for (size_t k = 0; k < count; k++) {
        host::G1J p;
        if (!host::g1_deserialize96(p, points96 + 96 * k, true)) return fail(ZKP_ERR_ENCODING, "bad G1 point");
        acc = acc.add(p);
    }
    host::g1_compress(out48, acc);
    return ZKP_OK;
}

int zkp_srs_import_g2_tau(zkp_ctx* ctx, const uint8_t tau_x_be[32]) {
    if (!ctx || !tau_x_be) return fail(ZKP_ERR_ARG, "null argument");
    Fr64 t;
    if (!Fr64::from_be(t, tau_x_be)) return fail(ZKP_ERR_ENCODING, "tau not canonical");
    Fr64 c = t.from_mont();
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->g2_tau = host::g2_generator().mul(c.v, 4);
    ctx->have_g2_tau = true;
    set_pairing_lines(ctx);
    return ZKP_OK;
}

int zkp_srs_import_g2_tau_y(zkp_ctx* ctx, const uint8_t tau_y_be[32]) {
    if (!ctx || !tau_y_be) return fail(ZKP_ERR_ARG, "null argument");
    Fr64 t;
    if (!Fr64::from_be(t, tau_y_be)) return fail(ZKP_ERR_ENCODING, "tau not canonical");
    Fr64 c = t.from_mont();
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->g2_tau_y = host::g2_generator().mul(c.v, 4);
    ctx->have_g2_tau_y = true;
    if (ctx->have_g2_tau) set_pairing_lines(ctx);
    return ZKP_OK;
}

// file: "ZKPB200S" | u32 version | u32 log_n | u32 log_m | [tau_x]_2 affine (4 x 48 B BE) |
//       version 2 only: [tau_y]_2 affine (4 x 48 B BE) |
//       M x 48 B compressed scale points | M*n x 96 B uncompressed points
int zkp_srs_save(zkp_ctx* ctx, const char* path) {
    if (!ctx || !path) return fail(ZKP_ERR_ARG, "null argument");
    if (!ctx->shaped || !ctx->have_g2_tau) return fail(ZKP_ERR_STATE, "SRS incomplete");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ZKP_ERR_IO, std::string("cannot open ") + path);
    uint32_t hdr[3] = {ctx->have_g2_tau_y ? 2u : 1u, ctx->log_n, ctx->log_m};
    bool ok = fwrite("ZKPB200S", 1, 8, f) == 8 && fwrite(hdr, 4, 3, f) == 3;
    for (int which = 0; which < (ctx->have_g2_tau_y ? 2 : 1); which++) {
        host::Fq2 gx, gy;
        (which ? ctx->g2_tau_y : ctx->g2_tau).to_affine(gx, gy);
        uint8_t g2[192];
        gx.c0.to_be(g2); gx.c1.to_be(g2 + 48); gy.c0.to_be(g2 + 96); gy.c1.to_be(g2 + 144);
        ok = ok && fwrite(g2, 1, 192, f) == 192;
    }
    size_t M = (size_t)1 << ctx->log_m, n = (size_t)1 << ctx->log_n;
    for (size_t i = 0; i < M && ok; i++) {
        uint8_t sp[48];
        host::g1_compress(sp, ctx->scale_points[i]);
        ok = fwrite(sp, 1, 48, f) == 48;
    }
    std::vector<uint8_t> row(n * 96);
    for (size_t i = 0; i < M && ok; i++) {
        int rc = zkp_srs_export_row(ctx, (uint32_t)i, row.data(), n);
        if (rc) { fclose(f); return rc; }
        ok = fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    fclose(f);
    return ok ? ZKP_OK : fail(ZKP_ERR_IO, "short write");
}

int zkp_srs_load(zkp_ctx* ctx, const char* path) {
    if (!ctx || !path) return fail(ZKP_ERR_ARG, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ZKP_ERR_IO, std::string("cannot open ") + path);
    char magic[8];
    uint32_t hdr[3];
    uint8_t g2[192], g2y[192];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "ZKPB200S", 8) || fread(hdr, 4, 3, f) != 3 || (hdr[0] != 1 && hdr[0] != 2) ||
        fread(g2, 1, 192, f) != 192 || (hdr[0] == 2 && fread(g2y, 1, 192, f) != 192)) {
        fclose(f);
        return fail(ZKP_ERR_IO, "not a zkp_b200 SRS file");
    }
    int rc = zkp_srs_set_shape(ctx, hdr[1], hdr[2]);
    if (rc) { fclose(f); return rc; }
    auto parse_g2 = [](const uint8_t* b, host::Fq2& x, host::Fq2& y) {
        return Fq64::from_be(x.c0, b) && Fq64::from_be(x.c1, b + 48) && Fq64::from_be(y.c0, b + 96) && Fq64::from_be(y.c1, b + 144) &&
               host::g2_on_curve(x, y);
    };
    host::Fq2 gx, gy, hx, hy;
    if (!parse_g2(g2, gx, gy) || (hdr[0] == 2 && !parse_g2(g2y, hx, hy))) {
        fclose(f);
        return fail(ZKP_ERR_ENCODING, "bad G2 point in SRS file");
    }
    size_t M = (size_t)1 << hdr[2], n = (size_t)1 << hdr[1];
    std::vector<uint8_t> sp(M * 48), row(n * 96);
    if (fread(sp.data(), 1, sp.size(), f) != sp.size()) { fclose(f); return fail(ZKP_ERR_IO, "short read"); }
    for (size_t i = 0; i < M; i++) {
        if (fread(row.data(), 1, row.size(), f) != row.size()) { fclose(f); return fail(ZKP_ERR_IO, "short read"); }
        rc = zkp_srs_import_row(ctx, (uint32_t)i, row.data(), n, &sp[i * 48]);
        if (rc) { fclose(f); return rc; }
    }
    fclose(f);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->g2_tau = host::G2J::from_affine(gx, gy);
    ctx->have_g2_tau = true;
    ctx->have_g2_tau_y = hdr[0] == 2;
    if (ctx->have_g2_tau_y) ctx->g2_tau_y = host::G2J::from_affine(hx, hy);
    set_pairing_lines(ctx);
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- open
int zkp_worker_open(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, const uint8_t x_be[32], uint8_t eval_be[32],
                    uint8_t proof48[48]) {
    Fr64 x;
    int rc = open_checks(ctx, i, poly_be, n, x_be, &x);
    if (rc) return rc;
    if (!eval_be || !proof48) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_poly(ctx, poly_be, n, true);
    if (rc) return rc;
    rc = commit_open_resident(ctx, i, n, x, nullptr, eval_be, proof48);
    if (rc == ZKP_OK) ctx->resident_n = n;
    return rc;
}

int zkp_worker_open_resident(zkp_ctx* ctx, uint32_t i, size_t n, const uint8_t x_be[32], uint8_t eval_be[32], uint8_t proof48[48]) {
    Fr64 x;
    int rc = open_checks(ctx, i, x_be /* any non-null pointer */, n, x_be, &x);
    if (rc) return rc;
    if (!eval_be || !proof48) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (ctx->resident_n != n) return fail(ZKP_ERR_STATE, "no polynomial of this size is resident on the device");
    rc = convert_poly(ctx, n, true);
    if (rc) return rc;
    return commit_open_resident(ctx, i, n, x, nullptr, eval_be, proof48);
}

// The same, bound to ONE upload: `generation` is the value zkp_resident_generation returned right after the caller's
// own zkp_worker_commit / zkp_worker_open / zkp_worker_commit_open.  Any later call that rewrites the staged scalars --
// from another client of a shared context, another thread, a raw zkp_msm_g1 -- changes the generation, and this
// entry then fails with ZKP_ERR_STATE instead of opening somebody else's polynomial.
int zkp_worker_open_resident_gen(zkp_ctx* ctx, uint32_t i, size_t n, uint64_t generation, const uint8_t x_be[32], uint8_t eval_be[32],
                                 uint8_t proof48[48]) {
    Fr64 x;
    int rc = open_checks(ctx, i, x_be /* any non-null pointer */, n, x_be, &x);
    if (rc) return rc;
    if (!eval_be || !proof48) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (ctx->resident_n != n || ctx->resident_gen != generation)
        return fail(ZKP_ERR_STATE, "the polynomial of that upload is no longer resident on the device");
    rc = convert_poly(ctx, n, true);
    if (rc) return rc;
    return commit_open_resident(ctx, i, n, x, nullptr, eval_be, proof48);
}
// The commitment and the proof of the last zkp_worker_commit_open / zkp_bench_commit_open on this context as 2 x 96
// bytes ZCash-UNCOMPRESSED: what a rank of a multi-process job contributes to the cross-GPU sum, so that the combining
// rank adds affine points (zkp_g1_sum_uncompressed) and nobody takes a square root.
int zkp_last_points_uncompressed(zkp_ctx* ctx, uint8_t out192[192]) {
    if (!ctx || !out192) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->have_last) return fail(ZKP_ERR_STATE, "no commit+open has completed on this context");
    host::g1_serialize96(out192, ctx->last_com);
    host::g1_serialize96(out192 + 96, ctx->last_proof);
    return ZKP_OK;
}
// ---- staged upload: the polynomial reaches the device in CHUNKS while the host is still producing it (the shim decodes a
// List[str] of base64 strings chunk by chunk: every finished chunk is already on its way over PCIe while the next one is
// being decoded).  zkp_stage_begin names the upload (generation), zkp_stage_chunk enqueues one asynchronous copy of
// elements [first, first + count) from base + 32 first (its argument order makes it usable as the per-chunk callback of
// the wire decoder), zkp_stage_end marks the polynomial resident; the *_resident entries then work on it.
int zkp_stage_begin(zkp_ctx* ctx, size_t n, uint64_t* generation) {
    if (!ctx || !n || !generation) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ctx->resident_n = 0;
    ctx->resident_gen++;
    ZKP_CUDA(ctx->scalars.ensure(n * 32));
    ctx->staging_n = n;
    ctx->staging_gen = ctx->resident_gen;
    *generation = ctx->resident_gen;
    return ZKP_OK;
}
int zkp_stage_chunk(zkp_ctx* ctx, size_t first, const uint8_t* base, size_t count) {
    if (!ctx || !base) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->staging_n || ctx->staging_gen != ctx->resident_gen) return fail(ZKP_ERR_STATE, "no staged upload in progress on this context");
    if (first + count > ctx->staging_n) return fail(ZKP_ERR_ARG, "chunk outside the staged polynomial");
    DeviceGuard g(ctx->device);
    ZKP_CUDA(cudaMemcpyAsync(ctx->scalars.as<uint8_t>() + 32 * first, base + 32 * first, 32 * count, cudaMemcpyHostToDevice, ctx->stream));
    return ZKP_OK;
}
int zkp_stage_end(zkp_ctx* ctx, uint64_t generation) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->staging_n || ctx->staging_gen != generation || ctx->resident_gen != generation)
        return fail(ZKP_ERR_STATE, "the staged upload was superseded by another call on this context");
    ctx->resident_n = ctx->staging_n;  // the copies are ordered before everything enqueued on the context's stream afterwards
    ctx->staging_n = 0;
    return ZKP_OK;
}
// worker_commit / fused commit+open on the polynomial of upload `generation` (staged, or left by an earlier call)
int zkp_worker_commit_resident(zkp_ctx* ctx, uint32_t i, size_t n, uint64_t generation, uint8_t commitment48[48]) {
    int rc = check_row(ctx, i, n);
    if (rc) return rc;
    if (!commitment48) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (ctx->resident_n != n || ctx->resident_gen != generation)
        return fail(ZKP_ERR_STATE, "the polynomial of that upload is no longer resident on the device");
    if (!ctx->coeff_form) return msm_device(ctx, i, ctx->scalars.as<uint32_t>(), SCALAR_BE, n, commitment48);
    if (n != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "coefficient form needs exactly one SRS row of coefficients");
    rc = convert_poly(ctx, n, true);
    if (rc) return rc;
    rc = msm_device(ctx, i, ctx->fr_a.as<uint32_t>(), SCALAR_MONT, n, commitment48);
    if (rc) return rc;
    uint32_t bad = 0;
    ZKP_CUDA(cudaMemcpy(&bad, small_at<uint32_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost));
    if (bad) return fail(ZKP_ERR_ENCODING, "polynomial holds a non-canonical field element");
    return ZKP_OK;
}
int zkp_worker_commit_open_resident(zkp_ctx* ctx, uint32_t i, size_t n, uint64_t generation, const uint8_t x_be[32],
                                    uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]) {
    Fr64 x;
    int rc = open_checks(ctx, i, x_be /* any non-null pointer */, n, x_be, &x);
    if (rc) return rc;
    if (!commitment48 || !eval_be || !proof48) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (ctx->resident_n != n || ctx->resident_gen != generation)
        return fail(ZKP_ERR_STATE, "the polynomial of that upload is no longer resident on the device");
    rc = convert_poly(ctx, n, true);
    if (rc) return rc;
    return commit_open_resident(ctx, i, n, x, commitment48, eval_be, proof48);
}
// The same two results as JACOBIAN points (X, Y, Z as 3 x 48 bytes of Montgomery limbs each, Z = 0 for infinity; 288
// bytes): a rank of a multi-process job hands these over without ANY field inversion, and zkp_g1_sum_jacobian on the
// combining rank adds them and compresses once.  The bytes are this library's internal representation: only for
// exchange between processes running the same build on the same box.
int zkp_last_points_jacobian(zkp_ctx* ctx, uint8_t out288[288]) {
    if (!ctx || !out288) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->have_last) return fail(ZKP_ERR_STATE, "no commit+open has completed on this context");
    const host::G1J* pts[2] = {&ctx->last_com, &ctx->last_proof};
    for (int k = 0; k < 2; k++) {
        memcpy(out288 + 144 * k, pts[k]->x.v, 48);
        memcpy(out288 + 144 * k + 48, pts[k]->y.v, 48);
        memcpy(out288 + 144 * k + 96, pts[k]->z.v, 48);
    }
    return ZKP_OK;
}
// sum of `count` Jacobian points of `stride` bytes each (the first 144 bytes of every record are read) -> compressed
int zkp_g1_sum_jacobian(const uint8_t* points, size_t count, size_t stride, uint8_t out48[48]) {
    if (!points || !out48 || stride < 144) return fail(ZKP_ERR_ARG, "bad argument");
    host::G1J acc = host::G1J::infinity();
    for (size_t k = 0; k < count; k++) {
        host::G1J p;
        memcpy(p.x.v, points + stride * k, 48);
        memcpy(p.y.v, points + stride * k + 48, 48);
        memcpy(p.z.v, points + stride * k + 96, 48);
        for (int i = 0; i < 3; i++) {
            const uint64_t* v = i == 0 ? p.x.v : (i == 1 ? p.y.v : p.z.v);
            if (Fq64::geq_mod(v)) return fail(ZKP_ERR_ENCODING, "coordinate out of range");
        }
        if (!p.is_inf()) {  // on the curve:  Y^2 = X^3 + 4 Z^6
            Fq64 z2 = p.z.sqr(), z6 = z2.sqr() * z2;
            if (p.y.sqr() != p.x.sqr() * p.x + host::fq_b4() * z6) return fail(ZKP_ERR_ENCODING, "point " + std::to_string(k) + " is not on the curve");
        }
        acc = acc.add(p);
    }
    host::g1_compress(out48, acc);
    return ZKP_OK;
}
int zkp_resident_generation(zkp_ctx* ctx, uint64_t* generation, size_t* n) {
    if (!ctx || !generation) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    *generation = ctx->resident_gen;
    if (n) *n = ctx->resident_n;
    return ZKP_OK;
}

int zkp_worker_commit_open(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, const uint8_t x_be[32],
                           uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]) {
    Fr64 x;
    int rc = open_checks(ctx, i, poly_be, n, x_be, &x);
    if (rc) return rc;
    if (!commitment48 || !eval_be || !proof48) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_poly(ctx, poly_be, n, true);
    if (rc) return rc;
    rc = commit_open_resident(ctx, i, n, x, commitment48, eval_be, proof48);
    if (rc == ZKP_OK) ctx->resident_n = n;
    return rc;
}

// ---------------------------------------------------------------------------------------------- batch
// `count` independent commit+open requests of one row length in ONE launch set (the live workload: 2^16-element rows,
// reference Makefile:64-74, several validators' requests in flight, base/miner.py:66-70).  Request r: row rows[r],
// evaluations polys_be[r] (n x 32 bytes), point xs_be + 32 r.  The 2 count MSMs are groups of one grouped pipeline
// (msm.cuh MsmGroups): a single sort, accumulation grid, slot chain and reduction for the whole batch, so the
// latency-bound tail of a 2^16 MSM is paid once per batch and the accumulation grid is large enough to fill the machine.
// status[r] = ZKP_OK or the error of request r (a non-canonical element fails that request only); the return value
// is the first hard failure (CUDA, arguments).  Outputs of failed requests are zeroed.
namespace {
__global__ void k_batch_init(uint8_t* __restrict__ records, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * (uint32_t)(SM_BYTES / 4)) return;
    reinterpret_cast<uint32_t*>(records)[i] = (i % (SM_BYTES / 4)) == SM_HIT / 4 ? HIT_NONE : 0u;
}
constexpr size_t BATCH_MAX = MSM_MAX_GROUPS / 2;

int batch_chunk(zkp_ctx* ctx, size_t count, const uint32_t* rows, const uint8_t* const* polys_be, size_t n, const uint8_t* xs_be,
                uint8_t* commitments48, uint8_t* evals_be, uint8_t* proofs48, int* status, bool* done) {
    *done = false;
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream2;
    const uint32_t nn = (uint32_t)n, cnt = (uint32_t)count;
    // evaluation points (Montgomery) and host staging for the records
    // records, evaluation points, and the per-request constants of the coset opening (open_device)
    const size_t hb = BATCH_MAX * (SM_BYTES + 32 + 96);
    if (ctx->h_batch_cap < hb) {
        if (ctx->h_batch) cudaFreeHost(ctx->h_batch);
        ctx->h_batch = nullptr;
        ctx->h_batch_cap = 0;
        ZKP_CUDA(cudaMallocHost(&ctx->h_batch, hb));
        ctx->h_batch_cap = hb;
    }
    uint8_t* h_x = ctx->h_batch + BATCH_MAX * SM_BYTES;
    for (size_t r = 0; r < count; r++) {
        Fr64 x;
        if (!Fr64::from_be(x, xs_be + 32 * r)) { status[r] = ZKP_ERR_ENCODING; x = Fr64::zero(); }
        memcpy(h_x + 32 * r, x.v, 32);
    }
    ctx->resident_n = 0;
    ctx->resident_gen++;
    ZKP_CUDA(ctx->scalars.ensure(n * count * 32));
    ZKP_CUDA(ctx->fr_a.ensure(n * count * 32));
    ZKP_CUDA(ctx->batch_small.ensure(BATCH_MAX * SM_BYTES));
    ZKP_CUDA(ctx->batch_x.ensure(BATCH_MAX * 32));
    int rc = open_buffers(ctx, nn, cnt);
    if (rc) return rc;
    // groups 0..count-1: commitments (raw big-endian scalars); count..2 count-1: proofs (Montgomery quotients)
    MsmJob jobs[MSM_MAX_GROUPS];
    for (uint32_t r = 0; r < cnt; r++) {
        jobs[r] = {rows[r], ctx->scalars.as<uint32_t>() + (size_t)r * n * 8, SCALAR_BE};
        jobs[cnt + r] = {rows[r], ctx->fr_c.as<uint32_t>() + (size_t)r * n * 8, SCALAR_MONT};
    }
    MsmPlan plan;
    MsmGroups gs;
    const G1Affine* pts = nullptr;
    bool grouped = false;
    rc = msm_plan_jobs(ctx, jobs, 2 * cnt, n, &plan, &gs, &pts, &grouped);
    if (rc) return rc;
    if (!grouped) return ZKP_OK;
    *done = true;
    uint8_t* rec = ctx->batch_small.as<uint8_t>();
    for (uint32_t r = 0; r < cnt; r++)
        ZKP_CUDA(cudaMemcpyAsync(ctx->scalars.as<uint8_t>() + (size_t)r * n * 32, polys_be[r], n * 32, cudaMemcpyHostToDevice, s0));
    ZKP_CUDA(cudaMemcpyAsync(ctx->batch_x.p, h_x, 32 * count, cudaMemcpyHostToDevice, s0));
    k_batch_init<<<(cnt * (unsigned)(SM_BYTES / 4) + 255) / 256, 256, 0, s0>>>(rec, cnt);
    k_fr_from_be<<<dim3((nn + 255) / 256, cnt), 256, 0, s0>>>(ctx->scalars.as<uint32_t>(), n, ctx->fr_a.as<Fr>(),
                                                             reinterpret_cast<uint32_t*>(rec + SM_BAD));
    ctx->launches += 2;
    ZKP_CUDA(cudaEventRecord(ctx->ev_ready, s0));
    ZKP_CUDA(cudaStreamWaitEvent(s1, ctx->ev_ready, 0));
    rc = msm_prep_begin(ctx, 0, plan);
    if (!rc) rc = msm_prep_count(ctx, 0, plan, gs, 0, cnt);
    if (!rc) rc = open_device(ctx, s1, ctx->fr_a.as<Fr>(), nn, Fr64::zero(), cnt, ctx->batch_x.as<Fr>(), rec, h_x, h_x + 32 * BATCH_MAX);
    if (!rc) {
        ZKP_CUDA(cudaEventRecord(ctx->ev_join, s1));
        ZKP_CUDA(cudaStreamWaitEvent(s0, ctx->ev_join, 0));
        rc = msm_prep_count(ctx, 0, plan, gs, cnt, cnt);
    }
    if (!rc) rc = msm_prep_finish(ctx, 0, plan, gs);
    if (!rc) rc = msm_enqueue_main(ctx, 0, plan, pts);
    if (!rc) {
        k_fr_to_be_records<<<(cnt + 63) / 64, 64, 0, s0>>>(rec, cnt);
        ctx->launches++;
        cudaMemcpyAsync(ctx->h_batch, rec, count * SM_BYTES, cudaMemcpyDeviceToHost, s0);
        rc = msm_wait(ctx, 0, 2 * cnt, false);
    } else {
        cudaStreamSynchronize(s0);
        cudaStreamSynchronize(s1);
    }
    msm_unpin_all(ctx);
    if (rc) return rc;
    // host: fold 2 count group results; a few threads when the batch is large (each fold is ~40 point operations
    // and one field inversion)
    const G1Xyzz* hw = ctx->ws.h_window;
    const uint32_t* h_bad = ctx->ws.h_bad;
    auto fold_range = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t r = lo; r < hi; r++) {
            const uint8_t* hr = ctx->h_batch + (size_t)r * SM_BYTES;
            const bool bad = h_bad[r] || *reinterpret_cast<const uint32_t*>(hr + SM_BAD);
            if (status[r] == ZKP_OK && bad) status[r] = ZKP_ERR_ENCODING;
            if (status[r] != ZKP_OK) {
                memset(commitments48 + 48 * r, 0, 48);
                memset(proofs48 + 48 * r, 0, 48);
                memset(evals_be + 32 * r, 0, 32);
                continue;
            }
            host::g1_compress(commitments48 + 48 * r, msm_fold_group(plan, hw, r));
            host::g1_compress(proofs48 + 48 * r, msm_fold_group(plan, hw, cnt + r));
            memcpy(evals_be + 32 * r, hr + SM_EVAL, 32);
        }
    };
    // (the codec's persistent host threads: ~2 requests each, one thread below 4 requests)
    unsigned cores = std::thread::hardware_concurrency();
    if (cores == 0) cores = 1;
    unsigned nth = cnt / 2;
    if (nth > cores) nth = cores;
    if (nth > 16) nth = 16;
    codec::parallel_ranges_n(cnt, nth, [&](size_t lo, size_t hi) { fold_range((uint32_t)lo, (uint32_t)hi); });
    return ZKP_OK;
}
}  // namespace

int zkp_worker_commit_open_batch(zkp_ctx* ctx, size_t count, const uint32_t* rows, const uint8_t* const* polys_be, size_t n,
                                 const uint8_t* xs_be, uint8_t* commitments48, uint8_t* evals_be, uint8_t* proofs48, int* status) {
    if (!ctx || !count || !rows || !polys_be || !xs_be || !commitments48 || !evals_be || !proofs48 || !status)
        return fail(ZKP_ERR_ARG, "null argument");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "SRS not loaded");
    if (n != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "every request needs exactly one SRS row of evaluations");
    if (ctx->shard_domain_log != ctx->log_n) return fail(ZKP_ERR_STATE, "not available on a point-range shard");
    for (size_t r = 0; r < count; r++) {
        if (!polys_be[r]) return fail(ZKP_ERR_ARG, "null polynomial");
        if (rows[r] >= (1u << ctx->log_m) || !ctx->row_loaded[rows[r]]) return fail(ZKP_ERR_ARG, "worker index out of range");
        status[r] = ZKP_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    for (size_t lo = 0; lo < count; lo += BATCH_MAX) {
        const size_t k = count - lo < BATCH_MAX ? count - lo : BATCH_MAX;
        bool done = false;
        int rc = ZKP_OK;
        if (tables_wanted(ctx, n) && !ctx->coeff_form)
            rc = batch_chunk(ctx, k, rows + lo, polys_be + lo, n, xs_be + 32 * lo, commitments48 + 48 * lo, evals_be + 32 * lo,
                             proofs48 + 48 * lo, status + lo, &done);
        if (rc) return rc;
        if (done) continue;
        // no tables (disabled, or they do not fit): the requests one after the other on the two-lane path
        for (size_t r = lo; r < lo + k; r++) {
            Fr64 x;
            if (!Fr64::from_be(x, xs_be + 32 * r)) { status[r] = ZKP_ERR_ENCODING; continue; }
            rc = upload_poly(ctx, polys_be[r], n, true);
            if (!rc) rc = commit_open_resident(ctx, rows[r], n, x, commitments48 + 48 * r, evals_be + 32 * r, proofs48 + 48 * r);
            if (rc == ZKP_ERR_ENCODING) {
                status[r] = rc;
                memset(commitments48 + 48 * r, 0, 48);
                memset(proofs48 + 48 * r, 0, 48);
                memset(evals_be + 32 * r, 0, 32);
            } else if (rc) {
                return rc;
            }
        }
    }
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- sharded open
// Opening of a polynomial whose evaluations (and SRS row) are split by point range over G GPUs
// (zkp_srs_generate_shard).  The barycentric sum splits by point range like the MSM does:
//   f(x) = -(x^n - 1)/n * sum_g S_g,   S_g = sum_{j in shard g} f_j w^j / (w^j - x)
// so the ranks exchange ONE 32-byte partial sum (all-gather), form y on the host, and then each computes the
// quotient evaluations of its own slice and the MSM over its own points; pi = sum_g pi_g (zkp_g1_sum).
namespace {
int shard_checks(zkp_ctx* ctx, uint32_t i, const void* slice, size_t n_local, const uint8_t* x_be, Fr64* x) {
    int rc = check_row(ctx, i, n_local);
    if (rc) return rc;
    if (!slice || !x_be) return fail(ZKP_ERR_ARG, "null argument");
    if (n_local != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "need exactly this shard's slice of evaluations");
    if (!Fr64::from_be(*x, x_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    return ZKP_OK;
}
// f slice -> fr_a (Montgomery), 1/(w^j - x) -> fr_b, S_g -> small[SM_S1]; fails if x lies in this shard's slice
int shard_pass1(zkp_ctx* ctx, const uint8_t* slice_be, size_t n_local, const Fr64& x) {
    int rc = upload_poly(ctx, slice_be, n_local);
    if (rc) return rc;
    zkp_ctx::Domain* dom;
    rc = get_domain(ctx, ctx->shard_domain_log, false, &dom);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const uint32_t n = (uint32_t)n_local;
    ZKP_CUDA(ctx->fr_b.ensure((size_t)n * 32));
    uint32_t E = n >> 16;
    if (E < 4) E = 4;
    if (E > 16) E = 16;
    uint32_t threads = (n + E - 1) / E, blocks = (threads + 127) / 128;
    ZKP_CUDA(ctx->partials.ensure((size_t)blocks * 32));
    ZKP_CUDA(cudaMemsetAsync(small_at<uint32_t>(ctx, SM_HIT), 0xff, 4, st));
    k_open_pass1<<<blocks, 128, 0, st>>>(ctx->fr_a.as<Fr>(), n, E, to_dev(x), dom->wt.as<Fr>(), to_dev(dom->w_inv), ctx->fr_b.as<Fr>(),
                                         ctx->partials.as<Fr>(), small_at<uint32_t>(ctx, SM_HIT), (uint64_t)ctx->shard_index << ctx->log_n);
    k_fr_reduce<<<1, 256, 0, st>>>(ctx->partials.as<Fr>(), blocks, small_at<Fr>(ctx, SM_S1));
    ctx->launches += 2;
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 96, small_at<uint8_t>(ctx, SM_HIT), 4, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 64, small_at<uint8_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaStreamSynchronize(st));
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 64)) return fail(ZKP_ERR_ENCODING, "polynomial holds a non-canonical field element");
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 96) != HIT_NONE)
        return fail(ZKP_ERR_ARG, "evaluation point lies inside the domain: not supported on point-range shards "
                                 "(use zkp_worker_open on an unsharded row)");
    return ZKP_OK;
}
}  // namespace

int zkp_shard_eval_partial(zkp_ctx* ctx, uint32_t i, const uint8_t* slice_be, size_t n_local, const uint8_t x_be[32],
                           uint8_t partial_be[32]) {
    Fr64 x;
    int rc = shard_checks(ctx, i, slice_be, n_local, x_be, &x);
    if (rc) return rc;
    if (!partial_be) return fail(ZKP_ERR_ARG, "null output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = shard_pass1(ctx, slice_be, n_local, x);
    if (rc) return rc;
    k_fr_to_be<<<1, 32, 0, ctx->stream>>>(small_at<Fr>(ctx, SM_S1), 1, small_at<uint32_t>(ctx, SM_EVAL));
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small, small_at<uint8_t>(ctx, SM_EVAL), 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(partial_be, ctx->h_small, 32);
    return ZKP_OK;
}

// y = -(x^n - 1)/n * sum_g S_g   (host; n = 2^log_n is the FULL domain size)
int zkp_shard_eval_combine(const uint8_t* partials_be, size_t count, uint32_t log_n, const uint8_t x_be[32], uint8_t y_be[32]) {
    if (!partials_be || !x_be || !y_be || !count || log_n > 32) return fail(ZKP_ERR_ARG, "bad argument");
    Fr64 x, acc = Fr64::zero();
    if (!Fr64::from_be(x, x_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    for (size_t g = 0; g < count; g++) {
        Fr64 s;
        if (!Fr64::from_be(s, partials_be + 32 * g)) return fail(ZKP_ERR_ENCODING, "partial sum is not canonical");
        acc = acc + s;
    }
    Fr64 xn = x;
    for (uint32_t k = 0; k < log_n; k++) xn = xn.sqr();
    Fr64 y = ((xn - Fr64::one()) * Fr64::from_u64(1ull << log_n).inverse() * acc).neg();
    y.to_be(y_be);
    return ZKP_OK;
}

int zkp_shard_open_partial(zkp_ctx* ctx, uint32_t i, const uint8_t* slice_be, size_t n_local, const uint8_t x_be[32],
                           const uint8_t y_be[32], uint8_t proof_partial48[48]) {
    Fr64 x, y;
    int rc = shard_checks(ctx, i, slice_be, n_local, x_be, &x);
    if (rc) return rc;
    if (!y_be || !proof_partial48) return fail(ZKP_ERR_ARG, "null argument");
    if (!Fr64::from_be(y, y_be)) return fail(ZKP_ERR_ENCODING, "evaluation is not canonical");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = shard_pass1(ctx, slice_be, n_local, x);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const uint32_t n = (uint32_t)n_local;
    ZKP_CUDA(ctx->fr_c.ensure((size_t)n * 32));
    memcpy(ctx->h_small + 128, y.v, 32);
    ZKP_CUDA(cudaMemcpyAsync(small_at<uint8_t>(ctx, SM_Y), ctx->h_small + 128, 32, cudaMemcpyHostToDevice, st));
    k_open_pass2<<<(n + 255) / 256, 256, 0, st>>>(ctx->fr_a.as<Fr>(), ctx->fr_b.as<Fr>(), n, small_at<Fr>(ctx, SM_Y), ctx->fr_c.as<Fr>());
    ctx->launches++;
    return msm_device(ctx, i, ctx->fr_c.as<uint32_t>(), SCALAR_MONT, n, proof_partial48);
}

// ---------------------------------------------------------------------------------------------- verify
int zkp_worker_verify(zkp_ctx* ctx, uint32_t i, const uint8_t proof48[48], const uint8_t alpha_be[32], const uint8_t eval_be[32],
                      const uint8_t commitment48[48], int* valid) {
    if (!ctx || !proof48 || !alpha_be || !eval_be || !commitment48 || !valid) return fail(ZKP_ERR_ARG, "null argument");
    *valid = 0;
    if (!ctx->shaped || !ctx->have_lines) return fail(ZKP_ERR_STATE, "SRS (G2 part) not loaded");
    if (i >= (1u << ctx->log_m)) return fail(ZKP_ERR_ARG, "worker index out of range");
    using namespace host;
    G1J proof, com;
    Fr64 alpha, y;
    // malformed inputs are a failed verification, not an error (reference tests/test_validator.py:66,79-86)
    if (!Fr64::from_be(alpha, alpha_be) || !Fr64::from_be(y, eval_be)) return ZKP_OK;
    Fr64 ac = alpha.from_mont(), yc = y.from_mont();
    G1J scale;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        scale = ctx->scale_points[i];
    }
    // A = C - y*S_i + alpha*pi ;  e(A, g2) * e(-pi, [tau]_2) == 1
    // Two independent chains of host work (decompression with its subgroup check, then a 255-bit scalar multiplication,
    // ~0.4 ms each) run side by side: the proof's on the calling thread, the commitment's on one of the codec's threads.
    bool ok_proof = false, ok_com = false;
    G1J a_pi, c_ys;
    codec::parallel_ranges_n(2, 2, [&](size_t lo, size_t hi) {
        for (size_t t = lo; t < hi; t++) {
            if (t == 0) {
                ok_proof = g1_decompress(proof, proof48);
                if (ok_proof) a_pi = proof.mul_w4(ac.v, 4);
            } else {
                ok_com = g1_decompress(com, commitment48);
                if (ok_com) c_ys = com.add(scale.mul_w4(yc.v, 4).neg());
            }
        }
    });
    if (!ok_proof || !ok_com) return ZKP_OK;
    G1J a = c_ys.add(a_pi);
    std::vector<G1AffineHost> ps = {g1_affine_host(a), g1_affine_host(proof.neg())};
    std::vector<const G2Lines*> qs = {&ctx->lines_g2, &ctx->lines_tau};
    *valid = pairing_product_is_one(ps, qs) ? 1 : 0;
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- Pianist master
int zkp_master_open_y(zkp_ctx* ctx, const uint8_t* worker_evals_be, size_t m, const uint8_t beta_be[32], uint8_t z_be[32],
                      uint8_t proof_y48[48]) {
    if (!ctx || !worker_evals_be || !beta_be || !z_be || !proof_y48) return fail(ZKP_ERR_ARG, "null argument");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "SRS not loaded");
    const uint32_t log_m = ctx->log_m;
    if (m != ((size_t)1 << log_m)) return fail(ZKP_ERR_ARG, "need exactly one evaluation per worker (2^log_machines)");
    std::vector<Fr64> y(m);
    Fr64 beta;
    if (!Fr64::from_be(beta, beta_be)) return fail(ZKP_ERR_ENCODING, "beta is not canonical");
    for (size_t i = 0; i < m; i++)
        if (!Fr64::from_be(y[i], worker_evals_be + 32 * i)) return fail(ZKP_ERR_ENCODING, "worker evaluation is not canonical");
    std::vector<host::G1J> scale;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (size_t i = 0; i < m; i++)
            if (!ctx->row_loaded[i]) return fail(ZKP_ERR_STATE, "SRS row not loaded");
        scale = ctx->scale_points;
    }
    // d_i = w^i - beta over the size-m domain; beta inside the domain is handled exactly
    const Fr64 w = fr_root_of_unity(log_m);
    std::vector<Fr64> wi(m), d(m);
    size_t hit = m;
    Fr64 cur = Fr64::one();
    for (size_t i = 0; i < m; i++) {
        wi[i] = cur;
        d[i] = cur - beta;
        if (d[i].is_zero()) hit = i;
        cur = cur * w;
    }
    // one batch inversion of the non-zero d_i
    std::vector<Fr64> pre(m);
    Fr64 run = Fr64::one();
    for (size_t i = 0; i < m; i++) {
        pre[i] = run;
        if (i != hit) run = run * d[i];
    }
    Fr64 inv = run.inverse();
    std::vector<Fr64> dinv(m, Fr64::zero());
    for (size_t i = m; i-- > 0;) {
        if (i == hit) continue;
        dinv[i] = inv * pre[i];
        inv = inv * d[i];
    }
    Fr64 z;
    if (hit < m) {
        z = y[hit];
    } else {
        // barycentric: g(beta) = (beta^m - 1)/m * sum_i y_i w^i / (beta - w^i)
        Fr64 acc = Fr64::zero();
        for (size_t i = 0; i < m; i++) acc = acc - y[i] * wi[i] * dinv[i];
        Fr64 bm = beta;
        for (uint32_t k = 0; k < log_m; k++) bm = bm.sqr();
        z = acc * (bm - Fr64::one()) * Fr64::from_u64(m).inverse();
    }
    std::vector<Fr64> q(m, Fr64::zero());
    for (size_t i = 0; i < m; i++)
        if (i != hit) q[i] = (y[i] - z) * dinv[i];
    if (hit < m) {
        // q_hit = g'(w^hit) = -sum_{i != hit} q_i w^(i - hit)
        Fr64 acc = Fr64::zero(), whi = wi[hit].inverse();
        for (size_t i = 0; i < m; i++)
            if (i != hit) acc = acc + q[i] * wi[i] * whi;
        q[hit] = acc.neg();
    }
    host::G1J pi = host::G1J::infinity();
    for (size_t i = 0; i < m; i++) {
        if (q[i].is_zero()) continue;
        Fr64 qc = q[i].from_mont();
        pi = pi.add(scale[i].mul(qc.v, 4));
    }
    z.to_be(z_be);
    host::g1_compress(proof_y48, pi);
    return ZKP_OK;
}

int zkp_master_verify(zkp_ctx* ctx, const uint8_t commitment48[48], const uint8_t proof_x48[48], const uint8_t proof_y48[48],
                      const uint8_t alpha_be[32], const uint8_t beta_be[32], const uint8_t z_be[32], int* valid) {
    if (!ctx || !commitment48 || !proof_x48 || !proof_y48 || !alpha_be || !beta_be || !z_be || !valid)
        return fail(ZKP_ERR_ARG, "null argument");
    *valid = 0;
    if (!ctx->shaped || !ctx->have_lines || !ctx->have_g2_tau_y) return fail(ZKP_ERR_STATE, "SRS (G2 part, tau_x and tau_y) not loaded");
    using namespace host;
    G1J com, px, py;
    Fr64 alpha, beta, z;
    if (!g1_decompress(com, commitment48) || !g1_decompress(px, proof_x48) || !g1_decompress(py, proof_y48)) return ZKP_OK;
    if (!Fr64::from_be(alpha, alpha_be) || !Fr64::from_be(beta, beta_be) || !Fr64::from_be(z, z_be)) return ZKP_OK;
    Fr64 ac = alpha.from_mont(), bc = beta.from_mont(), zc = z.from_mont();
    // A = com - z G + alpha pi_X + beta pi_Y ;  e(A, g2) e(-pi_X, [tau_x]_2) e(-pi_Y, [tau_y]_2) == 1
    G1J a = com.add(g1_generator().mul(zc.v, 4).neg()).add(px.mul(ac.v, 4)).add(py.mul(bc.v, 4));
    std::vector<G1AffineHost> ps = {g1_affine_host(a), g1_affine_host(px.neg()), g1_affine_host(py.neg())};
    std::vector<const G2Lines*> qs = {&ctx->lines_g2, &ctx->lines_tau, &ctx->lines_tau_y};
    *valid = pairing_product_is_one(ps, qs) ? 1 : 0;
    return ZKP_OK;
}

// Batched form of zkp_worker_verify for the validator's scoring loop (reference neurons/validator.py:168-170,
// 178-192 verifies the responses of one challenge one by one; they share alpha).  One random linear combination,
// two Miller loops and one final exponentiation for the whole batch:
//   e( sum_k r_k (C_k - y_k S_{i_k}) + alpha * sum_k r_k pi_k , g2 ) * e( -sum_k r_k pi_k , [tau_x]_2 ) == 1
// with fresh 128-bit r_k from the OS entropy source (an invalid proof survives with probability 2^-128).  If the
// combined check fails, every well-formed item is verified on its own so that valid[] still says which ones are
// bad.  Malformed encodings are valid = 0, never an error.
int zkp_worker_verify_batch(zkp_ctx* ctx, size_t count, const uint32_t* indices, const uint8_t* proofs48, const uint8_t alpha_be[32],
                            const uint8_t* evals_be, const uint8_t* commitments48, int* valid) {
    if (!ctx || (count && (!indices || !proofs48 || !evals_be || !commitments48 || !valid)) || !alpha_be)
        return fail(ZKP_ERR_ARG, "null argument");
    if (!ctx->shaped || !ctx->have_lines) return fail(ZKP_ERR_STATE, "SRS (G2 part) not loaded");
    using namespace host;
    for (size_t k = 0; k < count; k++) {
        valid[k] = 0;
        if (indices[k] >= (1u << ctx->log_m)) return fail(ZKP_ERR_ARG, "worker index out of range");
    }
    Fr64 alpha;
    if (!Fr64::from_be(alpha, alpha_be)) return ZKP_OK;  // malformed challenge point: nothing verifies
    std::vector<G1J> scale;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        scale = ctx->scale_points;
    }
    struct Item { size_t k; G1J proof, com; Fr64 y; };
    // decompression (a square root and a subgroup check per point) dominates a large batch: items are independent, so
    // they are spread over the codec's host threads, ~4 per thread
    unsigned cores = std::thread::hardware_concurrency();
    if (cores == 0) cores = 1;
    unsigned nth = (unsigned)(count / 4);
    if (nth > cores) nth = cores;
    if (nth > 16) nth = 16;
    if (nth < 1) nth = 1;
    std::vector<Item> all(count);
    std::vector<uint8_t> ok(count, 0);
    codec::parallel_ranges_n(count, nth, [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++) {
            Item& it = all[k];
            it.k = k;
            ok[k] = g1_decompress(it.proof, proofs48 + 48 * k) && g1_decompress(it.com, commitments48 + 48 * k) &&
                    Fr64::from_be(it.y, evals_be + 32 * k);
        }
    });
    std::vector<Item> items;
    items.reserve(count);
    for (size_t k = 0; k < count; k++)
        if (ok[k]) items.push_back(all[k]);
    if (items.empty()) return ZKP_OK;
    const Fr64 ac = alpha.from_mont();
    auto single = [&](const Item& it) {
        Fr64 yc = it.y.from_mont();
        G1J a = it.com.add(scale[indices[it.k]].mul_w4(yc.v, 4).neg()).add(it.proof.mul_w4(ac.v, 4));
        std::vector<G1AffineHost> ps = {g1_affine_host(a), g1_affine_host(it.proof.neg())};
        std::vector<const G2Lines*> qs = {&ctx->lines_g2, &ctx->lines_tau};
        return pairing_product_is_one(ps, qs);
    };
    if (items.size() == 1) {
        valid[items[0].k] = single(items[0]) ? 1 : 0;
        return ZKP_OK;
    }
    std::random_device rd;
    std::vector<uint64_t> rs(2 * items.size());
    for (size_t j = 0; j < items.size(); j++) {
        rs[2 * j] = ((uint64_t)rd() << 32) | rd();
        rs[2 * j + 1] = ((uint64_t)rd() << 32) | rd();
        if (!(rs[2 * j] | rs[2 * j + 1])) rs[2 * j] = 1;
    }
    // random linear combination: per-thread partial sums (two 128-bit scalar multiplications per item), merged below
    const size_t n_items = items.size();
    unsigned nth2 = (unsigned)(n_items / 4);
    if (nth2 > nth) nth2 = nth;
    if (nth2 < 1) nth2 = 1;
    struct Partial { G1J sum_c, sum_pi; std::vector<Fr64> t; std::vector<uint8_t> used; };
    std::vector<Partial> parts(nth2);
    for (auto& pt : parts) {
        pt.sum_c = G1J::infinity();
        pt.sum_pi = G1J::infinity();
        pt.t.assign(scale.size(), Fr64::zero());  // t_i = sum of r_k y_k over the items of row i
        pt.used.assign(scale.size(), 0);
    }
    codec::parallel_ranges_n(nth2, nth2, [&](size_t tlo, size_t thi) {
        for (size_t tix = tlo; tix < thi; tix++) {
            Partial& pt = parts[tix];
            for (size_t j = n_items * tix / nth2; j < n_items * (tix + 1) / nth2; j++) {
                const Item& it = items[j];
                uint64_t r[4] = {rs[2 * j], rs[2 * j + 1], 0, 0};
                pt.sum_c = pt.sum_c.add(it.com.mul_w4(r, 2));
                pt.sum_pi = pt.sum_pi.add(it.proof.mul_w4(r, 2));
                Fr64 rk;
                memcpy(rk.v, r, 32);
                rk = rk.to_mont();
                pt.t[indices[it.k]] = pt.t[indices[it.k]] + rk * it.y;
                pt.used[indices[it.k]] = 1;
            }
        }
    });
    G1J sum_c = G1J::infinity(), sum_pi = G1J::infinity();
    std::vector<Fr64> t(scale.size(), Fr64::zero());
    std::vector<uint8_t> row_used(scale.size(), 0);
    for (const auto& pt : parts) {
        sum_c = sum_c.add(pt.sum_c);
        sum_pi = sum_pi.add(pt.sum_pi);
        for (size_t i = 0; i < scale.size(); i++)
            if (pt.used[i]) { t[i] = t[i] + pt.t[i]; row_used[i] = 1; }
    }
    G1J a = sum_c.add(sum_pi.mul_w4(ac.v, 4));
    for (size_t i = 0; i < scale.size(); i++)
        if (row_used[i]) {
            Fr64 tc = t[i].from_mont();
            a = a.add(scale[i].mul_w4(tc.v, 4).neg());
        }
    std::vector<G1AffineHost> ps = {g1_affine_host(a), g1_affine_host(sum_pi.neg())};
    std::vector<const G2Lines*> qs = {&ctx->lines_g2, &ctx->lines_tau};
    if (pairing_product_is_one(ps, qs)) {
        for (const Item& it : items) valid[it.k] = 1;
        return ZKP_OK;
    }
    // the combined check failed: at least one response is wrong -- every one is checked on its own (in parallel)
    codec::parallel_ranges_n(n_items, nth, [&](size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; j++) valid[items[j].k] = single(items[j]) ? 1 : 0;
    });
    return ZKP_OK;
}

int zkp_pairing_check(const uint8_t* g1_48, const uint8_t* g2_192, size_t pairs, int* is_one) {
    if (!g1_48 || !g2_192 || !is_one) return fail(ZKP_ERR_ARG, "null argument");
    using namespace host;
    std::vector<G1AffineHost> ps;
    std::vector<G2Lines> lines(pairs);
    std::vector<const G2Lines*> qs;
    for (size_t k = 0; k < pairs; k++) {
        G1J p;
        if (!g1_decompress(p, g1_48 + 48 * k)) return fail(ZKP_ERR_ENCODING, "bad G1 point");
        Fq2 x, y;
        const uint8_t* q = g2_192 + 192 * k;
        if (!Fq64::from_be(x.c0, q) || !Fq64::from_be(x.c1, q + 48) || !Fq64::from_be(y.c0, q + 96) || !Fq64::from_be(y.c1, q + 144) ||
            !g2_on_curve(x, y))
            return fail(ZKP_ERR_ENCODING, "bad G2 point");
        const G2J q2 = G2J::from_affine(x, y);
        if (!q2.mul(host::FR_MOD64, 4).is_inf()) return fail(ZKP_ERR_ENCODING, "G2 point outside the prime-order subgroup");
        lines[k] = g2_precompute(q2);
        ps.push_back(g1_affine_host(p));
    }
    for (size_t k = 0; k < pairs; k++) qs.push_back(&lines[k]);
    *is_one = pairing_product_is_one(ps, qs) ? 1 : 0;
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- fft / eval / rng
int zkp_fft(zkp_ctx* ctx, const uint8_t* in_be, size_t n, int left, int inverse, uint8_t* out_be) {
    (void)left;  // a 1-D transform of length n uses w_n whichever axis of the bivariate grid it is
    if (!ctx || !in_be || !out_be) return fail(ZKP_ERR_ARG, "null argument");
    if (!is_pow2(n)) return fail(ZKP_ERR_ARG, "fft length must be a power of two");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc = upload_poly(ctx, in_be, n);
    if (rc) return rc;
    rc = ntt_device(ctx, ctx->fr_a.as<Fr>(), ctx->fr_a.as<Fr>(), ilog2(n), inverse);
    if (rc) return rc;
    k_fr_to_be<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->fr_a.as<Fr>(), n, ctx->scalars.as<uint32_t>());
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(out_be, ctx->scalars.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 64, small_at<uint8_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    ZKP_CUDA(cudaGetLastError());
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 64)) return fail(ZKP_ERR_ENCODING, "non-canonical field element");
    return ZKP_OK;
}

int zkp_eval(zkp_ctx* ctx, const uint8_t* coeffs_be, size_t n, const uint8_t x_be[32], uint8_t y_be[32]) {
    if (!ctx || !coeffs_be || !x_be || !y_be || !n) return fail(ZKP_ERR_ARG, "bad argument");
    Fr64 x;
    if (!Fr64::from_be(x, x_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    int rc = upload_poly(ctx, coeffs_be, n);
    if (rc) return rc;
    // x^(2^k) table
    std::vector<Fr64> xt(34);
    xt[0] = x;
    for (size_t k = 1; k < xt.size(); k++) xt[k] = xt[k - 1].sqr();
    ZKP_CUDA(ctx->fr_b.ensure(xt.size() * 32));
    memcpy(ctx->h_small + 128, xt.data(), xt.size() * 32);
    ZKP_CUDA(cudaMemcpyAsync(ctx->fr_b.p, ctx->h_small + 128, xt.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t E = 16, threads = (uint32_t)((n + E - 1) / E), blocks = (threads + 127) / 128;
    ZKP_CUDA(ctx->partials.ensure((size_t)blocks * 32));
    k_eval_partial<<<blocks, 128, 0, ctx->stream>>>(ctx->fr_a.as<Fr>(), (uint32_t)n, E, to_dev(x), ctx->fr_b.as<Fr>(), ctx->partials.as<Fr>());
    k_fr_reduce<<<1, 256, 0, ctx->stream>>>(ctx->partials.as<Fr>(), blocks, small_at<Fr>(ctx, SM_Y));
    ctx->launches += 2;
    return fetch_y(ctx, y_be);
}

// Validator.generate_challenge in one call (reference neurons/validator.py:106-120 does, per row, an inverse fft
// and a Horner evaluation through two RPCs): f_i(alpha) for every row of `rows` x n evaluations, by the barycentric
// formula with the weights w^j / (w^j - alpha) shared by all rows.
int zkp_challenge_evals(zkp_ctx* ctx, const uint8_t* polys_be, size_t rows, size_t n, const uint8_t alpha_be[32], uint8_t* evals_be) {
    if (!ctx || !polys_be || !alpha_be || !evals_be || !rows) return fail(ZKP_ERR_ARG, "bad argument");
    if (!is_pow2(n) || n > ((size_t)1 << 28) || rows > 65535) return fail(ZKP_ERR_ARG, "rows of a power-of-two length (at most 65535 rows)");
    Fr64 x;
    if (!Fr64::from_be(x, alpha_be)) return fail(ZKP_ERR_ENCODING, "evaluation point is not canonical");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const uint32_t log_n = ilog2(n);
    int rc = upload_poly(ctx, polys_be, rows * n);  // raw bytes -> ctx->scalars, Montgomery form -> fr_a, sets SM_BAD
    if (rc) return rc;
    zkp_ctx::Domain* dom;
    rc = get_domain(ctx, log_n, false, &dom);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const uint32_t nn = (uint32_t)n;
    ZKP_CUDA(ctx->fr_b.ensure(n * 32));
    uint32_t E = 16, threads = (nn + E - 1) / E, blocks = (threads + 127) / 128;
    const uint32_t per_block = 4096, parts = (nn + per_block - 1) / per_block;
    ZKP_CUDA(ctx->partials.ensure(((size_t)(blocks > parts * rows ? blocks : parts * rows)) * 32));
    ZKP_CUDA(cudaMemsetAsync(small_at<uint32_t>(ctx, SM_HIT), 0xff, 4, st));
    k_open_pass1<<<blocks, 128, 0, st>>>(nullptr, nn, E, to_dev(x), dom->wt.as<Fr>(), to_dev(dom->w_inv), ctx->fr_b.as<Fr>(),
                                         ctx->partials.as<Fr>(), small_at<uint32_t>(ctx, SM_HIT), 0);
    k_bary_weights<<<(threads + 127) / 128, 128, 0, st>>>(ctx->fr_b.as<Fr>(), nn, dom->wt.as<Fr>());
    k_bary_rows<<<dim3(parts, (unsigned)rows), 256, 0, st>>>(ctx->fr_a.as<Fr>(), ctx->fr_b.as<Fr>(), nn, per_block, ctx->partials.as<Fr>());
    Fr64 zn = x;
    for (uint32_t k = 0; k < log_n; k++) zn = zn.sqr();
    zn = (zn - Fr64::one()) * dom->n_inv;
    ZKP_CUDA(ctx->fr_c.ensure(rows * 32));
    k_bary_finish<<<(unsigned)rows, 32, 0, st>>>(ctx->fr_a.as<Fr>(), nn, ctx->partials.as<Fr>(), parts, to_dev(zn),
                                                 small_at<uint32_t>(ctx, SM_HIT), ctx->fr_c.as<uint32_t>());
    ctx->launches += 4;
    ZKP_CUDA(cudaMemcpyAsync(evals_be, ctx->fr_c.p, rows * 32, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaMemcpyAsync(ctx->h_small + 64, small_at<uint8_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    if (*reinterpret_cast<uint32_t*>(ctx->h_small + 64)) return fail(ZKP_ERR_ENCODING, "polynomial holds a non-canonical field element");
    return ZKP_OK;
}

int zkp_random_poly(zkp_ctx* ctx, uint64_t seed, uint8_t* out_be, size_t count) {
    if (!ctx || !out_be || !count) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ctx->resident_n = 0;
    ctx->resident_gen++;
    ZKP_CUDA(ctx->scalars.ensure(count * 32));
    k_random_fr<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(seed, count, ctx->scalars.as<uint32_t>());
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(out_be, ctx->scalars.p, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

int zkp_random_point(zkp_ctx* ctx, uint64_t seed, uint8_t out_be[32]) { return zkp_random_poly(ctx, seed ^ 0x706f696e74ull, out_be, 1); }

// elements [first, first + count) of the stream zkp_random_poly(seed, ...) produces: lets every GPU of a sharded job
// generate exactly its own slice of ONE global vector
int zkp_random_poly_range(zkp_ctx* ctx, uint64_t seed, uint64_t first, uint8_t* out_be, size_t count) {
    if (!ctx || !out_be || !count) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ctx->resident_n = 0;
    ctx->resident_gen++;
    ZKP_CUDA(ctx->scalars.ensure(count * 32));
    k_random_fr<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(seed, count, ctx->scalars.as<uint32_t>(), first);
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(out_be, ctx->scalars.p, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- tables / tuning
// Build the fixed-base tables of rows [first_row, first_row + count) now instead of inside the first request that
// needs them (Client.start(precompute="eager")).  Rows beyond the arena's capacity are skipped; *built = tables resident.
int zkp_srs_prebuild_tables(zkp_ctx* ctx, uint32_t first_row, uint32_t count, uint32_t* built) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "SRS not loaded");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    uint32_t ok = 0;
    const uint32_t rows = 1u << ctx->log_m;
    if (ctx->use_precomp && ctx->log_n >= 8) {
        for (uint32_t r = first_row; r < rows && r - first_row < count; r++) {
            if (!ctx->row_loaded[r]) continue;
            if (ok >= ctx->S.arena.row_of_slot.size() && ctx->S.arena.buf.p) break;  // arena full: leave the rest to LRU
            int slot = -1;
            int rc = acquire_table(ctx, r, &slot);
            if (rc) return rc;
            if (slot < 0) break;
            release_table(ctx, slot);
            ok++;
        }
    }
    if (built) *built = ok;
    return ZKP_OK;
}
// resident tables, slots of the arena, bytes of the arena, table builds, evictions, classic-path fallbacks so far
int zkp_srs_table_stats(zkp_ctx* ctx, uint64_t out[6]) {
    if (!ctx || !out) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->S.mu);
    const TableArena& ar = ctx->S.arena;
    uint64_t resident = 0;
    for (int r : ar.row_of_slot) resident += r >= 0;
    out[0] = resident;
    out[1] = ar.row_of_slot.size();
    out[2] = ar.buf.cap;
    out[3] = ar.builds;
    out[4] = ar.evictions;
    out[5] = ar.fallbacks;
    return ZKP_OK;
}
int zkp_set_table_budget(zkp_ctx* ctx, size_t bytes) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::recursive_mutex> lk(ctx->S.mu);
    ctx->S.table_budget = bytes;
    return ZKP_OK;
}
// 0 (default): `poly` arguments of zkp_worker_commit / open / commit_open are EVALUATIONS on the natural-order domain
// (what the reference's flow implies); 1: they are COEFFICIENTS (the reading of the comment at
// reference neurons/validator.py:67) -- the library then evaluates them first (one forward NTT) and proceeds identically.
int zkp_set_poly_form(zkp_ctx* ctx, int coefficients) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->coeff_form = coefficients != 0;
    ctx->resident_n = 0;
    ctx->resident_gen++;
    return ZKP_OK;
}
// Experiment switch: fetch the pass-2 tile of the NTT with one bulk async copy (TMA, cp.async.bulk + mbarrier) instead of
// per-thread 16-byte loads.  Same results; measured slower on B200, hence off by default.
int zkp_set_ntt_tma(zkp_ctx* ctx, int on) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->ntt_tma = on != 0;
    return ZKP_OK;
}
int zkp_set_fuse(zkp_ctx* ctx, int mode) {
    if (!ctx || mode < -1 || mode > 1) return fail(ZKP_ERR_ARG, "fuse mode must be -1 (by size), 0 or 1");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->fuse_mode = mode;
    return ZKP_OK;
}

int zkp_set_rowcol_coop(zkp_ctx* ctx, int on) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->rowcol_coop = on != 0;
    return ZKP_OK;
}
int zkp_set_open_coset(zkp_ctx* ctx, int on) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->open_coset = on != 0;
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- codec
int zkp_b64_decode_fr(const char* strs, size_t stride, size_t count, uint8_t* out_be) {
    if (!strs || !out_be || stride < 43) return fail(ZKP_ERR_ARG, "bad argument");
    size_t bad = codec::b64_decode_batch(strs, stride, count, out_be);
    if (bad != count) return fail(ZKP_ERR_ENCODING, "invalid base64 field element at index " + std::to_string(bad));
    return ZKP_OK;
}
int zkp_b64_encode_fr(const uint8_t* in_be, size_t count, char* out_strs) {
    if (!in_be || !out_strs) return fail(ZKP_ERR_ARG, "null argument");
    codec::b64_encode_batch(in_be, count, out_strs);
    return ZKP_OK;
}

// ---------------------------------------------------------------------------------------------- bench
int zkp_bench_msm(zkp_ctx* ctx, uint32_t row, const uint8_t* scalars_be, size_t n, int reps, int do_flush, float* ms_per_msm,
                  uint8_t out48[48]) {
    int rc = check_row(ctx, row, n);
    if (rc) return rc;
    if (!scalars_be || !ms_per_msm || !out48 || reps < 1) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_scalars(ctx, scalars_be, n, ctx->scalars);
    if (rc) return rc;
    // one untimed run: builds the fixed-base table on first use and grows the workspaces (cudaMalloc) so that no
    // allocation falls inside the timed repetitions
    rc = msm_device(ctx, row, ctx->scalars.as<uint32_t>(), SCALAR_BE, n, out48);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    ZKP_CUDA(cudaEventCreate(&e0));
    ZKP_CUDA(cudaEventCreate(&e1));
    double total = 0;
    ctx->time_acc = true;
    ctx->acc_ms_total = 0;
    ctx->acc_count = 0;
    for (int r = 0; r < reps && !rc; r++) {
        if (do_flush) flush_l2(ctx);
        cudaEventRecord(e0, ctx->stream);
        rc = msm_device(ctx, row, ctx->scalars.as<uint32_t>(), SCALAR_BE, n, out48);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->time_acc = false;
    ctx->last_acc_ms = ctx->acc_count ? (float)(ctx->acc_ms_total / ctx->acc_count) : 0.f;
    *ms_per_msm = (float)(total / reps);
    return rc;
}

int zkp_bench_commit_open(zkp_ctx* ctx, uint32_t row, const uint8_t* poly_be, size_t n, const uint8_t x_be[32], int reps, int do_flush,
                          float* ms_per_iter, float* ms_msm_kernel, uint32_t* launches, uint8_t commitment48[48], uint8_t eval_be[32],
                          uint8_t proof48[48]) {
    Fr64 x;
    int rc = open_checks(ctx, row, poly_be, n, x_be, &x);
    if (rc) return rc;
    if (!ms_per_iter || !commitment48 || !eval_be || !proof48 || reps < 1) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_poly(ctx, poly_be, n, true);  // resident in HBM before the timed region
    if (rc) return rc;
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEvent_t e0, e1;
    ZKP_CUDA(cudaEventCreate(&e0));
    ZKP_CUDA(cudaEventCreate(&e1));
    double total = 0;
    uint64_t launches0 = ctx->launches;
    ctx->time_acc = true;
    ctx->acc_ms_total = 0;
    ctx->acc_count = 0;
    for (int r = 0; r < reps && !rc; r++) {
        if (do_flush) flush_l2(ctx);
        cudaEventRecord(e0, ctx->stream);
        rc = commit_open_resident(ctx, row, n, x, commitment48, eval_be, proof48);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
    }
    ctx->time_acc = false;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_iter = (float)(total / reps);
    if (ms_msm_kernel) *ms_msm_kernel = ctx->acc_count ? (float)(ctx->acc_ms_total / ctx->acc_count) : 0.f;
    if (launches) *launches = (uint32_t)((ctx->launches - launches0) / reps);
    return rc;
}

// One traced commit+open (polynomial resident, after `warm` untraced runs): writes "lane stage t_ms" lines, t
// relative to the start of the request, where each line says when everything up to and including that stage
// had finished on that lane's stream; last line "host total <ms>" is the host wall time of the request.
int zkp_bench_trace(zkp_ctx* ctx, uint32_t row, const uint8_t* poly_be, size_t n, const uint8_t x_be[32], int warm, char* out,
                    size_t out_cap) {
    Fr64 x;
    int rc = open_checks(ctx, row, poly_be, n, x_be, &x);
    if (rc) return rc;
    if (!out || out_cap < 64) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_poly(ctx, poly_be, n, true);
    if (rc) return rc;
    uint8_t c48[48], y32[32], p48[48];
    for (int r = 0; r < warm && !rc; r++) rc = commit_open_resident(ctx, row, n, x, c48, y32, p48);
    if (rc) return rc;
    // ZKP_TRACE_FLUSH: "none" = no flush, "sync" = flush and wait, default = flush enqueued right in front (as zkp_bench_*)
    const char* fm = getenv("ZKP_TRACE_FLUSH");
    const bool f_none = fm && !strcmp(fm, "none"), f_sync = fm && !strcmp(fm, "sync");
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream2));
    if (!f_none) flush_l2(ctx);
    if (f_sync) ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEvent_t base;
    ZKP_CUDA(cudaEventCreate(&base));
    ctx->trace.clear();
    ctx->trace_on = true;
    auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(base, ctx->stream);
    rc = commit_open_resident(ctx, row, n, x, c48, y32, p48);
    double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ctx->trace_on = false;
    std::string text;
    for (auto& tp : ctx->trace) {
        float ms = 0;
        cudaEventSynchronize(tp.ev);
        cudaEventElapsedTime(&ms, base, tp.ev);
        char line[128];
        snprintf(line, sizeof(line), "%d %s %.4f\n", tp.lane, tp.name, ms);
        text += line;
        cudaEventDestroy(tp.ev);
    }
    ctx->trace.clear();
    cudaEventDestroy(base);
    char line[64];
    snprintf(line, sizeof(line), "host total %.4f\n", host_ms);
    text += line;
    if (text.size() + 1 > out_cap) return fail(ZKP_ERR_ARG, "trace buffer too small");
    memcpy(out, text.c_str(), text.size() + 1);
    return rc;
}

// write a buffer larger than the L2 (256 MiB) and wait: what the bench entries do between their timed iterations, exposed
// for timing loops that live outside the library (bench.py around the zkp_mgpu_* entries)
int zkp_bench_flush_l2(zkp_ctx* ctx) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    flush_l2(ctx);
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

// measured issue peaks on this device: IMAD.WIDE.U32 per second (whole chip) and dependent-chain Fq
// products per second at full occupancy
int zkp_bench_peaks(zkp_ctx* ctx, double* imad_wide_per_s, double* fq_mul_per_s) {
    if (!ctx || !imad_wide_per_s || !fq_mul_per_s) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const int threads = 256, blocks = ctx->sm_count * 8;  // 64 warps per SM
    ZKP_CUDA(ctx->flush.ensure((size_t)threads * blocks * sizeof(Fq)));
    cudaEvent_t e0, e1;
    ZKP_CUDA(cudaEventCreate(&e0));
    ZKP_CUDA(cudaEventCreate(&e1));
    float best_w = 1e30f, best_f = 1e30f;
    const int fq_iters = 512;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0, ctx->stream);
        k_peak_imad_wide<<<blocks, threads, 0, ctx->stream>>>(ctx->flush.as<uint32_t>(), 0x9e3779b1u);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best_w) best_w = ms;
        cudaEventRecord(e0, ctx->stream);
        k_peak_fq_mul<<<blocks / 2, threads, 0, ctx->stream>>>(ctx->flush.as<Fq>(), fq_iters);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best_f) best_f = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ZKP_CUDA(cudaGetLastError());
    ctx->launches += 8;
    *imad_wide_per_s = (double)threads * blocks * 4.0 * PEAK_ITERS / (best_w * 1e-3);
    *fq_mul_per_s = (double)threads * (blocks / 2) * fq_iters / (best_f * 1e-3);
    return ZKP_OK;
}

// mean duration of the dominant kernel (level-0 bucket accumulation) over the last zkp_bench_msm call,
// i.e. measured with nothing else running on the device
int zkp_bench_last_kernel_ms(zkp_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ZKP_ERR_ARG, "null argument");
    *ms = ctx->last_acc_ms;
    return ZKP_OK;
}

int zkp_bench_ntt(zkp_ctx* ctx, size_t n, int reps, int inverse, float* ms_per_ntt) {
    if (!ctx || !ms_per_ntt || !is_pow2(n) || reps < 1) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ZKP_CUDA(ctx->fr_a.ensure(n * 32));
    ZKP_CUDA(ctx->fr_b.ensure(n * 32));
    ctx->resident_n = 0;
    ctx->resident_gen++;
    ZKP_CUDA(ctx->scalars.ensure(n * 32));
    k_random_fr<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(0xB200, n, ctx->scalars.as<uint32_t>());
    int rc = ensure_small(ctx);
    if (rc) return rc;
    k_fr_from_be<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->scalars.as<uint32_t>(), n, ctx->fr_a.as<Fr>(),
                                                                       small_at<uint32_t>(ctx, SM_BAD));
    rc = ntt_device(ctx, ctx->fr_a.as<Fr>(), ctx->fr_b.as<Fr>(), ilog2(n), inverse);  // warm-up, builds tables
    if (rc) return rc;
    cudaEvent_t e0, e1;
    ZKP_CUDA(cudaEventCreate(&e0));
    ZKP_CUDA(cudaEventCreate(&e1));
    double total = 0;
    for (int r = 0; r < reps && !rc; r++) {
        flush_l2(ctx);
        cudaEventRecord(e0, ctx->stream);
        rc = ntt_device(ctx, ctx->fr_a.as<Fr>(), ctx->fr_b.as<Fr>(), ilog2(n), inverse);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ZKP_CUDA(cudaGetLastError());
    *ms_per_ntt = (float)(total / reps);
    return rc;
}

}  // extern "C"
