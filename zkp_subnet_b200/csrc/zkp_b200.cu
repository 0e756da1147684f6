// C-ABI of the B200-native KZG prover backend (see include/zkp_b200.h for the contract and the
// reference call site each entry replaces).  One translation unit: nvcc -gencode
// arch=compute_100a,code=sm_100a.  No PyTorch, no CPU fallback: every entry needs a CUDA device.
#include <cstdio>
#include <cstring>
#include <memory>

#include "codec.hpp"
#include "context.cuh"
#include "kzg.cuh"
#include "msm_driver.cuh"
#include "ntt.cuh"
#include "host/pairing.hpp"
#include "srs.cuh"
#include "bench_kernels.cuh"

using namespace zkp;

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
uint32_t ilog2(size_t n) {
    uint32_t l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

int check_row(zkp_ctx* ctx, uint32_t row, size_t n) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "SRS not loaded (call zkp_srs_generate / zkp_srs_load / zkp_srs_import_row)");
    if (row >= (1u << ctx->log_m)) return fail(ZKP_ERR_ARG, "worker index out of range");
    if (!ctx->row_loaded[row]) return fail(ZKP_ERR_STATE, "SRS row not loaded");
    if (n == 0 || n > ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "polynomial longer than the SRS row");
    return ZKP_OK;
}

const G1Affine* row_ptr(zkp_ctx* ctx, uint32_t row) { return ctx->srs.as<G1Affine>() + ((size_t)row << ctx->log_n); }

// upload big-endian scalars and validate them (< r) on the device
int upload_scalars(zkp_ctx* ctx, const uint8_t* be, size_t n, DevBuf& dst) {
    ctx->resident_n = 0;
    ZKP_CUDA(dst.ensure(n * 32));
    ZKP_CUDA(cudaMemcpyAsync(dst.p, be, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return ZKP_OK;
}

int finish_point(const MsmPlan& plan, zkp_ctx* ctx, int lane, uint8_t out48[48]) {
    host::G1J r = msm_fold(plan, (lane ? ctx->ws2 : ctx->ws).h_window);
    host::g1_compress(out48, r);
    return ZKP_OK;
}

// Build (once per row and window size) the fixed-base table [2^(c w)] P_i, w < W.  Returns false in
// *ok if the table would not fit comfortably in free HBM (the caller then uses the classic path).
int ensure_precomp(zkp_ctx* ctx, uint32_t row, uint32_t c, bool* ok) {
    *ok = false;
    if (ctx->precomp.size() != ctx->row_loaded.size()) ctx->precomp.assign(ctx->row_loaded.size(), zkp_ctx::Precomp());
    zkp_ctx::Precomp& pc = ctx->precomp[row];
    const uint32_t W = 255 / c + 1;
    if (pc.table.p && pc.c == c && pc.W == W) { *ok = true; return ZKP_OK; }
    const size_t n_row = (size_t)1 << ctx->log_n;
    // W slices [2^(c w)] P_i followed by the same W slices negated: the sign of a digit picks the half
    const size_t bytes = 2 * (size_t)W * n_row * TABLE_STRIDE;
    pc.table.release();
    pc.c = 0;
    size_t free_b = 0, total_b = 0;
    ZKP_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (bytes > free_b / 2) return ZKP_OK;  // keep at least half of the free HBM for workspaces
    ZKP_CUDA(pc.table.ensure(bytes));
    ZKP_CUDA(ctx->scratch_xyzz.ensure(n_row * sizeof(G1Xyzz)));
    ZKP_CUDA(ctx->scratch_fq.ensure(n_row * sizeof(Fq)));
    cudaStream_t st = ctx->stream;
    // records padded to TABLE_STRIDE bytes (msm.cuh); every slice is converted to affine in a 96-byte-stride
    // scratch slice and then placed
    char* tab = pc.table.as<char>();
    const size_t slice = n_row * TABLE_STRIDE;
    ZKP_CUDA(ctx->scratch_aff.ensure(n_row * sizeof(G1Affine)));
    ZKP_CUDA(cudaMemsetAsync(tab, 0, bytes, st));
    const unsigned gp = (unsigned)((n_row + 255) / 256);
    k_pad_points<<<gp, 256, 0, st>>>(row_ptr(ctx, row), n_row, tab);
    ctx->launches++;
    const uint32_t EA = 16;
    const unsigned ta = (unsigned)((n_row + EA - 1) / EA);
    for (uint32_t w = 1; w < W; w++) {
        k_table_next<<<(unsigned)((n_row + 127) / 128), 128, 0, st>>>(tab + (size_t)(w - 1) * slice, n_row, c, ctx->scratch_xyzz.as<G1Xyzz>());
        k_xyzz_to_affine<<<(ta + 127) / 128, 128, 0, st>>>(ctx->scratch_xyzz.as<G1Xyzz>(), n_row, EA, ctx->scratch_fq.as<Fq>(),
                                                           ctx->scratch_aff.as<G1Affine>());
        k_pad_points<<<gp, 256, 0, st>>>(ctx->scratch_aff.as<G1Affine>(), n_row, tab + (size_t)w * slice);
        ctx->launches += 3;
    }
    k_negate_points<<<(unsigned)(((size_t)W * n_row + 255) / 256), 256, 0, st>>>(tab, (size_t)W * n_row, tab + (size_t)W * slice);
    ctx->launches++;
    ZKP_CUDA(cudaStreamSynchronize(st));
    ZKP_CUDA(cudaGetLastError());
    pc.c = c;
    pc.W = W;
    *ok = true;
    return ZKP_OK;
}

void drop_precomp(zkp_ctx* ctx, int row /* -1 = all */) {
    for (size_t r = 0; r < ctx->precomp.size(); r++)
        if (row < 0 || (size_t)row == r) { ctx->precomp[r].table.release(); ctx->precomp[r].c = 0; }
}

MsmPlan plan_for(zkp_ctx* ctx, size_t n, bool precomp) {
    uint32_t c = ctx->c_override ? ctx->c_override : msm_window_bits((uint32_t)n, precomp);
    // Batched-affine rounds (msm_affine.cuh) are OFF unless forced with zkp_set_msm_affine_rounds.  Measured on
    // B200 at 2^20: a round costs about as much per addition as the XYZZ kernel (6 instead of
    // 10 products, but it keeps the multiply pipe only ~55% busy against ~87%), so a lone MSM is 9% slower with
    // 3 rounds and the two-stream commit+open only 2.7% faster with 2 rounds (12.69 vs 13.04 ms) -- not worth
    // a second set of point lists in HBM and a dominant kernel split three ways.
    uint32_t rounds = ctx->affine_rounds_override > 0 ? (uint32_t)ctx->affine_rounds_override : 0;
    MsmPlan plan = msm_make_plan((uint32_t)n, ctx->sm_count, c, precomp, 1u << ctx->log_n, rounds);
    if (precomp) plan.neg_offset = plan.W << ctx->log_n;
    return plan;
}

// Enqueue an MSM over row `row` with device-resident scalars on a lane (no host sync), in two halves
// (msm_driver.cuh): digits + sort, then accumulation + reduction.
int msm_device_prep(zkp_ctx* ctx, int lane, uint32_t row, const uint32_t* d_scalars, int fmt, size_t n, MsmPlan* plan_out,
                    const G1Affine** pts_out) {
    bool precomp = false;
    MsmPlan plan;
    if (ctx->use_precomp && n >= 256) {
        plan = plan_for(ctx, n, true);
        int rc = ensure_precomp(ctx, row, plan.c, &precomp);
        if (rc) return rc;
    }
    if (!precomp) plan = plan_for(ctx, n, false);
    *pts_out = precomp ? ctx->precomp[row].table.as<G1Affine>() : row_ptr(ctx, row);
    *plan_out = plan;
    return msm_enqueue_prep(ctx, lane, plan, d_scalars, fmt);
}
int msm_device_enqueue(zkp_ctx* ctx, int lane, uint32_t row, const uint32_t* d_scalars, int fmt, size_t n, MsmPlan* plan_out) {
    const G1Affine* pts;
    int rc = msm_device_prep(ctx, lane, row, d_scalars, fmt, n, plan_out, &pts);
    if (rc) return rc;
    return msm_enqueue_main(ctx, lane, *plan_out, pts);
}
int msm_device_finish(zkp_ctx* ctx, int lane, const MsmPlan& plan, uint8_t out48[48]) {
    int rc = msm_wait(ctx, lane);
    if (rc) return rc;
    return finish_point(plan, ctx, lane, out48);
}
// MSM over row `row` with device-resident scalars (lane 0, synchronous)
int msm_device(zkp_ctx* ctx, uint32_t row, const uint32_t* d_scalars, int fmt, size_t n, uint8_t out48[48]) {
    MsmPlan plan;
    int rc = msm_device_enqueue(ctx, 0, row, d_scalars, fmt, n, &plan);
    if (rc) return rc;
    return msm_device_finish(ctx, 0, plan, out48);
}

}  // namespace

extern "C" {

const char* zkp_last_error(void) { return tls_error().c_str(); }

int zkp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int zkp_host_alloc(size_t bytes, void** out) {
    if (!out || !bytes) return fail(ZKP_ERR_ARG, "bad argument");
    *out = nullptr;
    ZKP_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return ZKP_OK;
}
int zkp_host_free(void* p) {
    if (!p) return ZKP_OK;
    ZKP_CUDA(cudaFreeHost(p));
    return ZKP_OK;
}

int zkp_ctx_create(int device, zkp_ctx** out) {
    if (!out) return fail(ZKP_ERR_ARG, "null out pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ZKP_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this backend has no CPU fallback)");
    if (device < 0 || device >= count) return fail(ZKP_ERR_ARG, "device index out of range");
    DeviceGuard g(device);
    std::unique_ptr<zkp_ctx> ctx(new zkp_ctx());
    ctx->device = device;
    cudaDeviceProp prop;
    ZKP_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ZKP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ZKP_CUDA(cudaMallocHost(&ctx->h_small, 4096));
    ZKP_CUDA(cudaEventCreate(&ctx->ev_acc0));
    ZKP_CUDA(cudaEventCreate(&ctx->ev_acc1));
    ZKP_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    ZKP_CUDA(cudaEventCreate(&ctx->ev_acc2_0));
    ZKP_CUDA(cudaEventCreate(&ctx->ev_acc2_1));
    ZKP_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    ZKP_CUDA(cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16 * (1 << NTT_MAX_TILE_LOG)));
    *out = ctx.release();
    return ZKP_OK;
}

void zkp_ctx_destroy(zkp_ctx* ctx) {
    if (!ctx) return;
    {
        DeviceGuard g(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream2);
        ctx->ws.release();
        ctx->ws2.release();
        cudaEventDestroy(ctx->ev_acc2_0);
        cudaEventDestroy(ctx->ev_acc2_1);
        cudaEventDestroy(ctx->ev_ready);
        cudaStreamDestroy(ctx->stream2);
        drop_precomp(ctx, -1);
        ctx->scratch_xyzz.release();
        ctx->scratch_fq.release();
        ctx->scratch_aff.release();
        for (DevBuf* b : {&ctx->srs, &ctx->scalars, &ctx->fr_a, &ctx->fr_b, &ctx->fr_c, &ctx->flush, &ctx->small, &ctx->partials,
                          &ctx->ntt_tmp, &ctx->fixed_base})
            b->release();
        for (auto& d : ctx->domains) { d.wt.release(); d.tw.release(); }
        if (ctx->h_small) cudaFreeHost(ctx->h_small);
        cudaEventDestroy(ctx->ev_acc0);
        cudaEventDestroy(ctx->ev_acc1);
        cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

int zkp_set_msm_window(zkp_ctx* ctx, uint32_t c) {
    if (!ctx || c > 24 || (c && c < 2)) return fail(ZKP_ERR_ARG, "window bits must be 0 (auto) or 2..24");
    ctx->c_override = c;
    return ZKP_OK;
}

int zkp_set_msm_sort(zkp_ctx* ctx, int bucket_sort) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->bucket_sort = bucket_sort < 0 || bucket_sort > 2 ? 2 : bucket_sort;
    return ZKP_OK;
}

int zkp_set_msm_mode(zkp_ctx* ctx, int fixed_base_tables) {
    if (!ctx) return fail(ZKP_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->use_precomp = fixed_base_tables != 0;
    return ZKP_OK;
}

int zkp_set_msm_affine_rounds(zkp_ctx* ctx, int rounds) {
    if (!ctx || rounds < -1 || rounds > (int)AFFINE_MAX_ROUNDS) return fail(ZKP_ERR_ARG, "affine rounds must be -1 (auto) or 0..6");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->affine_rounds_override = rounds;
    return ZKP_OK;
}

int zkp_msm_info(zkp_ctx* ctx, size_t n, uint32_t* c, uint32_t* windows, uint64_t* fq_muls) {
    if (!ctx || !n) return fail(ZKP_ERR_ARG, "bad argument");
    bool pre = ctx->use_precomp && n >= 256;
    MsmPlan plan = plan_for(ctx, n, pre);
    if (c) *c = plan.c;
    if (windows) *windows = plan.W;
    if (fq_muls) *fq_muls = msm_fq_muls(plan);
    return ZKP_OK;
}

// ------------------------------------------------------------------------------------------ SRS
int zkp_srs_set_shape(zkp_ctx* ctx, uint32_t log_n, uint32_t log_machines) {
    if (!ctx || log_n > 28 || log_machines > 16) return fail(ZKP_ERR_ARG, "bad SRS shape");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    size_t total = (size_t)1 << (log_n + log_machines);
    ZKP_CUDA(ctx->srs.ensure(total * sizeof(G1Affine)));
    ctx->log_n = log_n;
    ctx->log_m = log_machines;
    ctx->shard_domain_log = log_n;
    ctx->shard_index = 0;
    drop_precomp(ctx, -1);
    ctx->precomp.clear();
    ctx->row_loaded.assign((size_t)1 << log_machines, 0);
    ctx->scale_points.assign((size_t)1 << log_machines, host::G1J::infinity());
    ctx->shaped = true;
    return ZKP_OK;
}

int zkp_srs_shape(zkp_ctx* ctx, uint32_t* log_n, uint32_t* log_machines) {
    if (!ctx || !ctx->shaped) return fail(ZKP_ERR_STATE, "SRS not loaded");
    if (log_n) *log_n = ctx->log_n;
    if (log_machines) *log_machines = ctx->log_m;
    return ZKP_OK;
}

int zkp_srs_import_row(zkp_ctx* ctx, uint32_t row, const uint8_t* points96, size_t n, const uint8_t scale_point48[48]) {
    if (!ctx || !points96) return fail(ZKP_ERR_ARG, "null argument");
    if (!ctx->shaped) return fail(ZKP_ERR_STATE, "call zkp_srs_set_shape first");
    if (row >= (1u << ctx->log_m) || n != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "row/size mismatch");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (scale_point48) {
        host::G1J s;
        if (!host::g1_decompress(s, scale_point48)) return fail(ZKP_ERR_ENCODING, "bad scale point");
        ctx->scale_points[row] = s;
    } else {
        ctx->scale_points[row] = host::g1_generator();
    }
    ZKP_CUDA(ctx->fr_a.ensure(n * 96));
    ZKP_CUDA(ctx->fr_b.ensure(4));
    ZKP_CUDA(cudaMemcpyAsync(ctx->fr_a.p, points96, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    ZKP_CUDA(cudaMemsetAsync(ctx->fr_b.p, 0, 4, ctx->stream));
    G1Affine* dst = ctx->srs.as<G1Affine>() + ((size_t)row << ctx->log_n);
    k_points_from_be96<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->fr_a.as<uint8_t>(), n, dst, ctx->fr_b.as<uint32_t>());
    ctx->launches++;
    uint32_t bad = 0;
    ZKP_CUDA(cudaMemcpyAsync(&bad, ctx->fr_b.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ZKP_ERR_ENCODING, "SRS row holds a malformed or off-curve point");
    drop_precomp(ctx, (int)row);
    ctx->row_loaded[row] = 1;
    return ZKP_OK;
}

int zkp_srs_export_row(zkp_ctx* ctx, uint32_t row, uint8_t* points96, size_t n) {
    int rc = check_row(ctx, row, n);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ZKP_CUDA(ctx->fr_a.ensure(n * 96));
    k_points_to_be96<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(row_ptr(ctx, row), n, ctx->fr_a.as<uint8_t>());
    ctx->launches++;
    ZKP_CUDA(cudaMemcpyAsync(points96, ctx->fr_a.p, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

// ------------------------------------------------------------------------------------------ hot path
int zkp_msm_g1(zkp_ctx* ctx, uint32_t row, const uint8_t* scalars_be, size_t n, uint8_t out48[48]) {
    int rc = check_row(ctx, row, n);
    if (rc) return rc;
    if (!scalars_be || !out48) return fail(ZKP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_scalars(ctx, scalars_be, n, ctx->scalars);
    if (rc) return rc;
    rc = msm_device(ctx, row, ctx->scalars.as<uint32_t>(), SCALAR_BE, n, out48);
    if (rc == ZKP_OK) ctx->resident_n = n;  // canonical (the MSM validated every scalar) and still on the device
    return rc;
}

int zkp_worker_commit(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, uint8_t commitment48[48]) {
    return zkp_msm_g1(ctx, i, poly_be, n, commitment48);
}

}  // extern "C"

#include "capi_rest.cuh"
