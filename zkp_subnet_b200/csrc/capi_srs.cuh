// C-ABI entries of the SRS tooling (srs_tools.cuh): compressed-point import / export, G2 and scale-point import /
// export by POINT (a ceremony SRS has no known tau), point-range shard marking, the bivariate monomial test setup and
// the group inverse FFT monomial -> Lagrange.  Included by zkp_b200.cu after capi_rest.cuh.
#pragma once
#include "srs_tools.cuh"

namespace {

// ZCash uncompressed G2: x.c1 || x.c0 || y.c1 || y.c0, 48 bytes big-endian each; infinity not accepted here
bool g2_from_zcash192(const uint8_t* b, host::G2J* out) {
    host::Fq2 x, y;
    if (b[0] & 0xe0) return false;
    if (!Fq64::from_be(x.c1, b) || !Fq64::from_be(x.c0, b + 48) || !Fq64::from_be(y.c1, b + 96) || !Fq64::from_be(y.c0, b + 144)) return false;
    if (!host::g2_on_curve(x, y)) return false;
    host::G2J q = host::G2J::from_affine(x, y);
    if (!q.mul(host::FR_MOD64, 4).is_inf()) return false;  // prime-order subgroup
    *out = q;
    return true;
}
void g2_to_zcash192(const host::G2J& q, uint8_t* b) {
    host::Fq2 x, y;
    q.to_affine(x, y);
    x.c1.to_be(b);
    x.c0.to_be(b + 48);
    y.c1.to_be(b + 96);
    y.c0.to_be(b + 144);
}

// in-place inverse group FFT of `batch` transforms inside the XYZZ array `data` (see k_gfft_stage for the addressing);
// `tmp` is a second array of the same size (bit reversal is out of place); the result ends up in `data`
int gfft_inverse(zkp_ctx* ctx, G1Xyzz* data, G1Xyzz* tmp, size_t total, uint32_t log_len, size_t batch, size_t tstride, size_t estride) {
    if (log_len == 0) return ZKP_OK;
    zkp_ctx::Domain* dom;
    int rc = get_domain(ctx, log_len, false, &dom);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const size_t bf = batch << (log_len - 1);
    for (int lh = (int)log_len - 1; lh >= 0; lh--) {
        k_gfft_stage<<<(unsigned)((bf + 127) / 128), 128, 0, st>>>(data, log_len, (uint32_t)lh, batch, tstride, estride, dom->wt.as<Fr>());
        ctx->launches++;
    }
    const Fr scale = to_dev(dom->n_inv.from_mont());
    const size_t pts = batch << log_len;
    k_gfft_finish<<<(unsigned)((pts + 127) / 128), 128, 0, st>>>(data, tmp, log_len, batch, tstride, estride, scale);
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(data, tmp, total * sizeof(G1Xyzz), cudaMemcpyDeviceToDevice, st));
    ZKP_CUDA(cudaGetLastError());
    return ZKP_OK;
}

int xyzz_to_affine_rows(zkp_ctx* ctx, const G1Xyzz* src, size_t count, G1Affine* dst) {
    const uint32_t EA = 16;
    const size_t ta = (count + EA - 1) / EA;
    ZKP_CUDA(ctx->scratch_fq.ensure(count * sizeof(Fq)));
    k_xyzz_to_affine<<<(unsigned)((ta + 127) / 128), 128, 0, ctx->stream>>>(src, count, EA, ctx->scratch_fq.as<Fq>(), dst);
    ctx->launches++;
    return ZKP_OK;
}

}  // namespace

extern "C" {

// compressed (48-byte) counterpart of zkp_srs_import_row: the `.compressed` files of the reference's setup
// (tests/conftest.py:28-29, --uncompressed false); every point is decompressed and checked on the device
int zkp_srs_import_row_compressed(zkp_ctx* ctx, uint32_t row, const uint8_t* points48, size_t n, const uint8_t scale_point48[48]) {
    if (!ctx || !points48) return fail(ZKP_ERR_ARG, "null argument");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "call zkp_srs_set_shape first");
    if (row >= (1u << ctx->log_m) || n != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "row/size mismatch");
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    DeviceGuard g(ctx->device);
    if (scale_point48) {
        host::G1J s;
        if (!host::g1_decompress(s, scale_point48)) return fail(ZKP_ERR_ENCODING, "bad scale point");
        ctx->scale_points[row] = s;
    } else {
        ctx->scale_points[row] = host::g1_generator();
    }
    ZKP_CUDA(ctx->fr_a.ensure(n * 48));
    ZKP_CUDA(ctx->fr_b.ensure(4));
    ZKP_CUDA(cudaMemcpyAsync(ctx->fr_a.p, points48, n * 48, cudaMemcpyHostToDevice, ctx->stream));
    ZKP_CUDA(cudaMemsetAsync(ctx->fr_b.p, 0, 4, ctx->stream));
    G1Affine* dst = ctx->srs.as<G1Affine>() + ((size_t)row << ctx->log_n);
    k_points_from_be48<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->fr_a.as<uint8_t>(), n, dst, ctx->fr_b.as<uint32_t>());
    ctx->launches++;
    uint32_t bad = 0;
    ZKP_CUDA(cudaMemcpyAsync(&bad, ctx->fr_b.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ZKP_ERR_ENCODING, "SRS row holds a malformed compressed point");
    {
        TableArena& ar = ctx->S.arena;
        if (row < ar.slot_of_row.size() && ar.slot_of_row[row] >= 0) {
            ar.row_of_slot[ar.slot_of_row[row]] = -1;
            ar.slot_of_row[row] = -1;
        }
    }
    ctx->row_loaded[row] = 1;
    return ZKP_OK;
}

int zkp_srs_export_row_compressed(zkp_ctx* ctx, uint32_t row, uint8_t* points48, size_t n) {
    int rc = check_row(ctx, row, n);
    if (rc) return rc;
    if (!points48) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ZKP_CUDA(ctx->fr_a.ensure(n * 48));
    k_points_to_be48<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(row_ptr(ctx, row), n, ctx->fr_a.as<uint8_t>());
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(points48, ctx->fr_a.p, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

// [tau_x]_2 (which = 0) / [tau_y]_2 (which = 1) as POINTS, ZCash uncompressed G2 (192 bytes: x.c1, x.c0, y.c1, y.c0):
// what a ceremony SRS provides.  Checked on the curve and in the prime-order subgroup.
int zkp_srs_import_g2(zkp_ctx* ctx, int which, const uint8_t g2_192[192]) {
    if (!ctx || !g2_192 || which < 0 || which > 1) return fail(ZKP_ERR_ARG, "bad argument");
    host::G2J q;
    if (!g2_from_zcash192(g2_192, &q)) return fail(ZKP_ERR_ENCODING, "bad G2 point (encoding, curve or subgroup)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    if (which == 0) {
        ctx->g2_tau = q;
        ctx->have_g2_tau = true;
    } else {
        ctx->g2_tau_y = q;
        ctx->have_g2_tau_y = true;
    }
    if (ctx->have_g2_tau) set_pairing_lines(ctx);
    return ZKP_OK;
}
int zkp_srs_export_g2(zkp_ctx* ctx, int which, uint8_t g2_192[192]) {
    if (!ctx || !g2_192 || which < 0 || which > 1) return fail(ZKP_ERR_ARG, "bad argument");
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    if (which == 0 ? !ctx->have_g2_tau : !ctx->have_g2_tau_y) return fail(ZKP_ERR_STATE, "that G2 point is not loaded");
    g2_to_zcash192(which == 0 ? ctx->g2_tau : ctx->g2_tau_y, g2_192);
    return ZKP_OK;
}
int zkp_srs_export_scale_point(zkp_ctx* ctx, uint32_t row, uint8_t out48[48]) {
    if (!ctx || !out48) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    if (!ctx->shaped || row >= (1u << ctx->log_m)) return fail(ZKP_ERR_ARG, "row out of range");
    host::g1_compress(out48, ctx->scale_points[row]);
    return ZKP_OK;
}

// mark the rows of this context as point-range shard `shard` of a domain of 2^log_domain points (after
// zkp_srs_set_shape with the LOCAL row length and the imports of the slices): what zkp_srs_generate_shard records
int zkp_srs_set_shard(zkp_ctx* ctx, uint32_t log_domain, uint32_t shard) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    if (!ctx->shaped || log_domain < ctx->log_n || log_domain > 32 || (uint64_t)shard >= (1ull << (log_domain - ctx->log_n)))
        return fail(ZKP_ERR_ARG, "bad shard");
    ctx->shard_domain_log = log_domain;
    ctx->shard_index = shard;
    return ZKP_OK;
}

// The bivariate monomial SRS [tau_x^j tau_y^i]_1 (row i, column j) from a trapdoor: the `--generate-setup` half of the
// reference's setup command (tests/conftest.py:50-65) for tests and local networks.  zkp_srs_monomial_to_lagrange then
// produces the rows the worker calls use -- without looking at the trapdoor again.
int zkp_srs_generate_monomial2(zkp_ctx* ctx, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32], uint32_t log_n,
                               uint32_t log_machines) {
    if (!ctx || !tau_x_be || !tau_y_be) return fail(ZKP_ERR_ARG, "null argument");
    Fr64 tx, ty;
    if (!Fr64::from_be(tx, tau_x_be) || !Fr64::from_be(ty, tau_y_be)) return fail(ZKP_ERR_ENCODING, "trapdoor not canonical");
    int rc = zkp_srs_set_shape(ctx, log_n, log_machines);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    DeviceGuard g(ctx->device);
    rc = ensure_fixed_base(ctx);
    if (rc) return rc;
    const uint32_t n = 1u << log_n, M = 1u << log_machines;
    cudaStream_t st = ctx->stream;
    std::vector<Fr64> tt(32);
    tt[0] = tx;
    for (size_t k = 1; k < tt.size(); k++) tt[k] = tt[k - 1].sqr();
    ZKP_CUDA(ctx->partials.ensure(tt.size() * 32));
    ZKP_CUDA(cudaMemcpyAsync(ctx->partials.p, tt.data(), tt.size() * 32, cudaMemcpyHostToDevice, st));
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(ctx->fr_c.ensure((size_t)n * 32));
    ZKP_CUDA(ctx->ws.buckets.ensure((size_t)n * sizeof(G1Xyzz)));
    ZKP_CUDA(ctx->ws.pool.ensure((size_t)n * sizeof(Fq)));
    Fr64 ypow = Fr64::one();
    for (uint32_t i = 0; i < M; i++) {
        k_power_scalars<<<((n + 7) / 8 + 127) / 128, 128, 0, st>>>(ctx->partials.as<Fr>(), n, ctx->fr_c.as<Fr>(), to_dev(ypow));
        k_fixed_base_mul<<<(n + 127) / 128, 128, 0, st>>>(ctx->fr_c.as<Fr>(), n, ctx->fixed_base.as<G1Affine>(), ctx->ws.buckets.as<G1Xyzz>());
        uint32_t EA = 16, ta = (n + EA - 1) / EA;
        k_xyzz_to_affine<<<(ta + 127) / 128, 128, 0, st>>>(ctx->ws.buckets.as<G1Xyzz>(), n, EA, ctx->ws.pool.as<Fq>(),
                                                           ctx->srs.as<G1Affine>() + ((size_t)i << log_n));
        ctx->launches += 3;
        Fr64 yc = ypow.from_mont();
        ctx->scale_points[i] = host::g1_generator().mul(yc.v, 4);  // [tau_y^i]_1 = column 0 of row i
        ctx->row_loaded[i] = 1;
        ypow = ypow * ty;
    }
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    Fr64 txc = tx.from_mont(), tyc = ty.from_mont();
    ctx->g2_tau = host::g2_generator().mul(txc.v, 4);
    ctx->have_g2_tau = true;
    ctx->g2_tau_y = host::g2_generator().mul(tyc.v, 4);
    ctx->have_g2_tau_y = true;
    set_pairing_lines(ctx);
    return ZKP_OK;
}

// Monomial -> Lagrange, in place, no trapdoor: the resident rows [tau_x^j tau_y^i]_1 become U[i][j] =
// [R_i(tau_y) L_j(tau_x)]_1 and the row scale points become [R_i(tau_y)]_1 (the `--generate-precompute` half of the
// reference's setup command; the only way to derive the worker rows from a ceremony SRS).
int zkp_srs_monomial_to_lagrange(zkp_ctx* ctx) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "SRS not loaded");
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    DeviceGuard g(ctx->device);
    const uint32_t log_n = ctx->log_n, log_m = ctx->log_m;
    const size_t n = (size_t)1 << log_n, M = (size_t)1 << log_m, total = n * M;
    for (size_t i = 0; i < M; i++)
        if (!ctx->row_loaded[i]) return fail(ZKP_ERR_STATE, "every row of the monomial SRS must be loaded");
    if (ctx->shard_domain_log != log_n) return fail(ZKP_ERR_STATE, "not available on a point-range shard");
    drop_tables(ctx);
    cudaStream_t st = ctx->stream;
    DevBuf work, tmp;
    cudaError_t e1 = work.ensure(total * sizeof(G1Xyzz)), e2 = tmp.ensure(total * sizeof(G1Xyzz));
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        work.release();
        tmp.release();
        cudaGetLastError();
        return fail(ZKP_ERR_CUDA, "not enough device memory for the group FFT work arrays");
    }
    struct Cleanup {
        DevBuf &a, &b;
        cudaStream_t st;
        ~Cleanup() {
            cudaStreamSynchronize(st);
            a.release();
            b.release();
        }
    } cleanup{work, tmp, st};
    auto done = [&](int code) { return code; };
    k_gfft_load<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->srs.as<G1Affine>(), total, work.as<G1Xyzz>());
    ctx->launches++;
    // Y direction: n transforms of length M, element (i, j) at i * n + j
    int rc = gfft_inverse(ctx, work.as<G1Xyzz>(), tmp.as<G1Xyzz>(), total, log_m, n, 1, n);
    if (rc) return done(rc);
    // scale points: column 0
    {
        std::vector<G1Xyzz> col(M);
        k_gather_stride<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(work.as<G1Xyzz>(), n, M, tmp.as<G1Xyzz>());
        ctx->launches++;
        ZKP_CUDA(cudaMemcpyAsync(col.data(), tmp.p, M * sizeof(G1Xyzz), cudaMemcpyDeviceToHost, st));
        ZKP_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i < M; i++) ctx->scale_points[i] = xyzz_to_jac(col[i]);
    }
    // X direction: M transforms of length n
    rc = gfft_inverse(ctx, work.as<G1Xyzz>(), tmp.as<G1Xyzz>(), total, log_n, M, n, 1);
    if (rc) return done(rc);
    rc = xyzz_to_affine_rows(ctx, work.as<G1Xyzz>(), total, ctx->srs.as<G1Affine>());
    if (rc) return done(rc);
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    return done(ZKP_OK);
}

}  // extern "C"
