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
    ctx->resident_gen++;
    ZKP_CUDA(dst.ensure(n * 32));
    ZKP_CUDA(cudaMemcpyAsync(dst.p, be, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return ZKP_OK;
}

// ---- fixed-base tables (TableArena, context.cuh) ----------------------------------------------------------------
// window width of the tables: a function of the ROW length only, so that every table of an SRS has the same size and
// calls with different n on the same row never rebuild anything
uint32_t table_window_bits(zkp_ctx* ctx) {
    return ctx->c_override ? ctx->c_override : msm_window_bits(1u << ctx->log_n, true);
}

void drop_tables(zkp_ctx* ctx) { ctx->S.arena.release(); }

// Build the table of `row` in slot `slot`: W slices [2^(c w)] P_i followed by the same W slices negated.
int build_table(zkp_ctx* ctx, uint32_t row, int slot) {
    TableArena& ar = ctx->S.arena;
    const uint32_t c = ar.c, W = ar.W;
    const size_t n_row = (size_t)1 << ctx->log_n;
    ZKP_CUDA(ctx->scratch_xyzz.ensure(n_row * sizeof(G1Xyzz)));
    ZKP_CUDA(ctx->scratch_fq.ensure(n_row * sizeof(Fq)));
    ZKP_CUDA(ctx->scratch_aff.ensure(n_row * sizeof(G1Affine)));
    cudaStream_t st = ctx->stream;
    // records padded to TABLE_STRIDE bytes (msm.cuh); every slice is converted to affine in a 96-byte-stride
    // scratch slice and then placed
    char* tab = ar.buf.as<char>() + (size_t)slot * ar.slot_records * TABLE_STRIDE;
    const size_t slice = n_row * TABLE_STRIDE;
    ZKP_CUDA(cudaMemsetAsync(tab, 0, ar.slot_records * TABLE_STRIDE, st));
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
    ar.builds++;
    return ZKP_OK;
}

// Pin the table of `row` (building it, and the arena, on first use).  *slot = -1 when no table can be had -- the
// arena does not fit in the HBM budget, or every slot is pinned by MSMs in flight -- and the caller takes the classic
// per-window path; that is logged once per SRS and counted (zkp_srs_table_stats), never silent.
int acquire_table(zkp_ctx* ctx, uint32_t row, int* slot) {
    *slot = -1;
    std::lock_guard<std::recursive_mutex> lk(ctx->S.mu);
    TableArena& ar = ctx->S.arena;
    const uint32_t c = table_window_bits(ctx), W = 255 / c + 1;
    const size_t n_row = (size_t)1 << ctx->log_n, rows = (size_t)1 << ctx->log_m;
    if (ar.buf.p && ar.c != c) {
        bool pinned = false;
        for (uint32_t p : ar.pins) pinned |= p != 0;
        if (pinned) { ar.fallbacks++; return ZKP_OK; }
        ar.release();
    }
    if (!ar.buf.p) {
        const size_t slot_records = 2 * (size_t)W * n_row, slot_bytes = slot_records * TABLE_STRIDE;
        size_t free_b = 0, total_b = 0;
        ZKP_CUDA(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = ctx->S.table_budget ? ctx->S.table_budget : free_b / 10 * 6;
        if (const char* e = getenv("ZKP_B200_TABLE_BYTES")) budget = (size_t)strtoull(e, nullptr, 10);
        if (budget > free_b / 10 * 9) budget = free_b / 10 * 9;
        size_t nslots = budget / slot_bytes;
        if (nslots > rows) nslots = rows;
        while (nslots && nslots * slot_records >= ((size_t)1 << 31)) nslots--;  // 31-bit record indices
        if (nslots == 0 || ar.buf.ensure(nslots * slot_bytes) != cudaSuccess) {
            cudaGetLastError();
            if (!ar.warned) {
                fprintf(stderr, "zkp_b200: fixed-base tables need %.2f GiB per row and do not fit in the HBM budget (%.2f GiB of %.2f GiB "
                                "free): falling back to the classic per-window MSM (about 20%% more field multiplications)\n",
                        slot_bytes / 1073741824.0, budget / 1073741824.0, free_b / 1073741824.0);
                ar.warned = true;
            }
            ar.fallbacks++;
            return ZKP_OK;
        }
        ar.c = c;
        ar.W = W;
        ar.slot_records = slot_records;
        ar.slot_of_row.assign(rows, -1);
        ar.row_of_slot.assign(nslots, -1);
        ar.pins.assign(nslots, 0);
        ar.last_use.assign(nslots, 0);
        if (nslots < rows && !ar.warned) {
            fprintf(stderr, "zkp_b200: fixed-base table arena holds %zu of %zu rows (%.2f GiB each); tables of other rows are built on "
                            "demand and evicted least-recently-used first\n", nslots, rows, slot_bytes / 1073741824.0);
            ar.warned = true;
        }
    }
    int s = ar.slot_of_row[row];
    if (s < 0) {
        uint64_t best = ~0ull;
        for (size_t k = 0; k < ar.row_of_slot.size(); k++) {
            if (ar.row_of_slot[k] < 0) { s = (int)k; break; }
            if (!ar.pins[k] && ar.last_use[k] < best) { best = ar.last_use[k]; s = (int)k; }
        }
        if (s < 0) { ar.fallbacks++; return ZKP_OK; }  // every slot pinned
        if (ar.row_of_slot[s] >= 0) {
            ar.slot_of_row[ar.row_of_slot[s]] = -1;
            ar.evictions++;
        }
        ar.row_of_slot[s] = -1;
        int rc = build_table(ctx, row, s);
        if (rc) return rc;
        ar.row_of_slot[s] = (int)row;
        ar.slot_of_row[row] = s;
    }
    ar.pins[s]++;
    ar.last_use[s] = ++ar.clock;
    *slot = s;
    return ZKP_OK;
}
void release_table(zkp_ctx* ctx, int slot) {
    if (slot < 0) return;
    std::lock_guard<std::recursive_mutex> lk(ctx->S.mu);
    TableArena& ar = ctx->S.arena;
    if ((size_t)slot < ar.pins.size() && ar.pins[slot]) ar.pins[slot]--;
}
// slots pinned by the MSMs in flight on a context are remembered in the context and dropped by msm_unpin_all
// (called wherever a lane has been waited for, and on every error path)
void remember_pin(zkp_ctx* ctx, int slot) {
    if (slot >= 0 && ctx->pinned_count < 64) ctx->pinned_slots[ctx->pinned_count++] = slot;
    else release_table(ctx, slot);
}
void msm_unpin_all(zkp_ctx* ctx) {
    for (int k = 0; k < ctx->pinned_count; k++) release_table(ctx, ctx->pinned_slots[k]);
    ctx->pinned_count = 0;
}

bool tables_wanted(zkp_ctx* ctx, size_t n) {
    // tables pay off when the call uses a fair share of the row (the bucket count is chosen for the whole row)
    return ctx->use_precomp && ctx->log_n >= 8 && n * 16 >= ((size_t)1 << ctx->log_n);
}

MsmPlan plan_for(zkp_ctx* ctx, size_t n, bool precomp, uint32_t groups = 1) {
    uint32_t c = precomp ? table_window_bits(ctx) : (ctx->c_override ? ctx->c_override : msm_window_bits((uint32_t)n, false));
    // Batched-affine rounds (msm_affine.cuh) are OFF unless forced with zkp_set_msm_affine_rounds.  Measured on
    // B200 at 2^20: a round costs about as much per addition as the XYZZ kernel (6 instead of
    // 10 products, but it keeps the multiply pipe only ~55% busy against ~87%), so a lone MSM is 9% slower with
    // 3 rounds and the two-stream commit+open only 2.7% faster with 2 rounds (12.69 vs 13.04 ms) -- not worth
    // a second set of point lists in HBM and a dominant kernel split three ways.
    uint32_t rounds = ctx->affine_rounds_override > 0 && groups == 1 ? (uint32_t)ctx->affine_rounds_override : 0;
    MsmPlan plan = msm_make_plan((uint32_t)n, ctx->sm_count, c, precomp, 1u << ctx->log_n, rounds, groups);
    if (precomp) plan.neg_offset = plan.W << ctx->log_n;
    return plan;
}

// One MSM of a (possibly grouped) launch set: scalars on the device in format fmt, over SRS row `row`
struct MsmJob { uint32_t row; const uint32_t* d_scalars; int fmt; };

// Plan a launch set of `count` MSMs of n points each (all through fixed-base tables when count > 1): pins the tables,
// fills the group descriptors.  *grouped_ok = false when count > 1 and a table is unavailable (the caller then runs
// the jobs one by one).  For count == 1 the classic path is the fallback and *pts_out says where the points are.
int msm_plan_jobs(zkp_ctx* ctx, const MsmJob* jobs, uint32_t count, size_t n, MsmPlan* plan_out, MsmGroups* gs_out,
                  const G1Affine** pts_out, bool* grouped_ok) {
    memset(gs_out, 0, sizeof(*gs_out));
    *grouped_ok = true;
    bool precomp = tables_wanted(ctx, n);
    int slots[MSM_MAX_GROUPS];
    uint32_t got = 0;
    if (precomp) {
        for (; got < count; got++) {
            int rc = acquire_table(ctx, jobs[got].row, &slots[got]);
            if (rc || slots[got] < 0) {
                for (uint32_t k = 0; k < got; k++) release_table(ctx, slots[k]);
                if (rc) return rc;
                precomp = false;
                break;
            }
        }
    }
    if (!precomp && count > 1) { *grouped_ok = false; return ZKP_OK; }
    MsmPlan plan = plan_for(ctx, n, precomp, count);
    if (precomp && plan.N >= ((size_t)1 << 31)) {  // positions of the bucket sort are 32-bit
        for (uint32_t k = 0; k < got; k++) release_table(ctx, slots[k]);
        if (count > 1) { *grouped_ok = false; return ZKP_OK; }
        precomp = false;
        plan = plan_for(ctx, n, false, 1);
    }
    for (uint32_t g = 0; g < count; g++) {
        gs_out->scalars[g] = jobs[g].d_scalars;
        gs_out->fmt[g] = (uint8_t)jobs[g].fmt;
        gs_out->val_base[g] = precomp ? (uint32_t)((size_t)slots[g] * ctx->S.arena.slot_records) : 0u;
        if (precomp) remember_pin(ctx, slots[g]);
    }
    *pts_out = precomp ? ctx->S.arena.buf.as<G1Affine>() : row_ptr(ctx, jobs[0].row);
    *plan_out = plan;
    return ZKP_OK;
}

// Enqueue an MSM over row `row` with device-resident scalars on a lane (no host sync), in two halves
// (msm_driver.cuh): digits + sort, then accumulation + reduction.
int msm_device_prep(zkp_ctx* ctx, int lane, uint32_t row, const uint32_t* d_scalars, int fmt, size_t n, MsmPlan* plan_out,
                    const G1Affine** pts_out) {
    MsmJob job = {row, d_scalars, fmt};
    MsmGroups gs;
    bool ok;
    int rc = msm_plan_jobs(ctx, &job, 1, n, plan_out, &gs, pts_out, &ok);
    if (rc) return rc;
    return msm_enqueue_prep(ctx, lane, *plan_out, gs);
}
int msm_device_enqueue(zkp_ctx* ctx, int lane, uint32_t row, const uint32_t* d_scalars, int fmt, size_t n, MsmPlan* plan_out) {
    const G1Affine* pts;
    int rc = msm_device_prep(ctx, lane, row, d_scalars, fmt, n, plan_out, &pts);
    if (rc) return rc;
    return msm_enqueue_main(ctx, lane, *plan_out, pts);
}
// wait for the lane and fold its result(s) on the host: group g of the launch set -> out48 (+ the Jacobian point)
int msm_device_finish(zkp_ctx* ctx, int lane, const MsmPlan& plan, uint8_t out48[48], host::G1J* jac = nullptr) {
    int rc = msm_wait(ctx, lane, plan.groups);
    if (rc) return rc;
    host::G1J r = msm_fold(plan, (lane ? ctx->ws2 : ctx->ws).h_window);
    if (jac) *jac = r;
    if (out48) host::g1_compress(out48, r);
    return ZKP_OK;
}
// MSM over row `row` with device-resident scalars (lane 0, synchronous)
int msm_device(zkp_ctx* ctx, uint32_t row, const uint32_t* d_scalars, int fmt, size_t n, uint8_t out48[48], host::G1J* jac = nullptr) {
    MsmPlan plan;
    host::G1J r;
    int rc = msm_device_enqueue(ctx, 0, row, d_scalars, fmt, n, &plan);
    if (rc == ZKP_OK) rc = msm_device_finish(ctx, 0, plan, out48, &r);
    else cudaStreamSynchronize(ctx->stream);
    msm_unpin_all(ctx);
    if (rc == ZKP_OK) {
        if (jac) *jac = r;
        ctx->last_com = r;  // zkp_last_points_uncompressed: the partial of a point-range shard
        ctx->last_proof = host::G1J::infinity();
        ctx->have_last = true;
    }
    return rc;
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

namespace {
int ctx_init(int device, std::shared_ptr<SrsStore> store, zkp_ctx** out) {
    DeviceGuard g(device);
    std::unique_ptr<zkp_ctx> ctx(new zkp_ctx(std::move(store)));
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
    ZKP_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    ZKP_CUDA(cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16 * (1 << NTT_MAX_TILE_LOG)));
    *out = ctx.release();
    return ZKP_OK;
}
}  // namespace

int zkp_ctx_create(int device, zkp_ctx** out) {
    if (!out) return fail(ZKP_ERR_ARG, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ZKP_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this backend has no CPU fallback)");
    if (device < 0 || device >= count) return fail(ZKP_ERR_ARG, "device index out of range");
    auto store = std::make_shared<SrsStore>();
    store->device = device;
    return ctx_init(device, store, out);
}

// A second context on the same device that SHARES the parent's SRS rows, fixed-base tables and domain tables (one
// copy in HBM) but has its own streams and workspaces: what a pool of request handlers wants (reference
// base/miner.py:66-70 hands `forward` to the axon's threads).  The SRS must not be replaced while forks are computing.
int zkp_ctx_fork(zkp_ctx* parent, zkp_ctx** out) {
    if (!parent || !out) return fail(ZKP_ERR_ARG, "null argument");
    *out = nullptr;
    int rc = ctx_init(parent->device, parent->store, out);
    if (rc) return rc;
    (*out)->bucket_sort = parent->bucket_sort;
    (*out)->use_precomp = parent->use_precomp;
    (*out)->fuse_mode = parent->fuse_mode;
    (*out)->open_coset = parent->open_coset;
    (*out)->rowcol_coop = parent->rowcol_coop;
    (*out)->coeff_form = parent->coeff_form;
    return ZKP_OK;
}

void zkp_ctx_destroy(zkp_ctx* ctx) {
    if (!ctx) return;
    {
        DeviceGuard g(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream2);
        msm_unpin_all(ctx);
        ctx->ws.release();
        ctx->ws2.release();
        cudaEventDestroy(ctx->ev_acc2_0);
        cudaEventDestroy(ctx->ev_acc2_1);
        cudaEventDestroy(ctx->ev_ready);
        cudaEventDestroy(ctx->ev_join);
        cudaStreamDestroy(ctx->stream2);
        ctx->scratch_xyzz.release();
        ctx->scratch_fq.release();
        ctx->scratch_aff.release();
        for (DevBuf* b : {&ctx->scalars, &ctx->fr_a, &ctx->fr_b, &ctx->fr_c, &ctx->flush, &ctx->small, &ctx->partials, &ctx->ntt_tmp,
                          &ctx->batch_small, &ctx->batch_x})
            b->release();
        if (ctx->h_small) cudaFreeHost(ctx->h_small);
        if (ctx->h_batch) cudaFreeHost(ctx->h_batch);
        cudaEventDestroy(ctx->ev_acc0);
        cudaEventDestroy(ctx->ev_acc1);
        cudaStreamDestroy(ctx->stream);
        ctx->store.reset();  // the last context of a store frees the SRS, the tables and the domain tables
    }
    delete ctx;
}

int zkp_set_msm_window(zkp_ctx* ctx, uint32_t c) {
    if (!ctx || c > 24 || (c && c < 2)) return fail(ZKP_ERR_ARG, "window bits must be 0 (auto) or 2..24");
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    ctx->c_override = c;  // the tables are rebuilt for the new width on their next use (acquire_table)
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
    bool pre = ctx->shaped && tables_wanted(ctx, n);
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
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
    DeviceGuard g(ctx->device);
    size_t total = (size_t)1 << (log_n + log_machines);
    drop_tables(ctx);
    ZKP_CUDA(ctx->srs.ensure(total * sizeof(G1Affine)));
    ctx->log_n = log_n;
    ctx->log_m = log_machines;
    ctx->shard_domain_log = log_n;
    ctx->shard_index = 0;
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
    std::lock_guard<std::recursive_mutex> lk2(ctx->S.mu);
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
    {  // a table built from the previous contents of the row is stale
        TableArena& ar = ctx->S.arena;
        if (row < ar.slot_of_row.size() && ar.slot_of_row[row] >= 0) {
            ar.row_of_slot[ar.slot_of_row[row]] = -1;
            ar.slot_of_row[row] = -1;
        }
    }
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

}  // extern "C"

#include "capi_rest.cuh"
#include "capi_srs.cuh"

extern "C" int zkp_worker_commit(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, uint8_t commitment48[48]) {
    if (!ctx || !ctx->coeff_form) return zkp_msm_g1(ctx, i, poly_be, n, commitment48);
    // coefficient form: evaluate first (forward NTT over the row's domain), then the same MSM over the Lagrange row
    int rc = check_row(ctx, i, n);
    if (rc) return rc;
    if (!poly_be || !commitment48) return fail(ZKP_ERR_ARG, "null argument");
    if (n != ((size_t)1 << ctx->log_n)) return fail(ZKP_ERR_ARG, "coefficient form needs exactly one SRS row of coefficients");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    rc = upload_poly(ctx, poly_be, n, true);
    if (rc) return rc;
    rc = msm_device(ctx, i, ctx->fr_a.as<uint32_t>(), SCALAR_MONT, n, commitment48);
    if (rc) return rc;
    uint32_t bad = 0;
    ZKP_CUDA(cudaMemcpy(&bad, small_at<uint32_t>(ctx, SM_BAD), 4, cudaMemcpyDeviceToHost));
    if (bad) return fail(ZKP_ERR_ENCODING, "polynomial holds a non-canonical field element");
    ctx->resident_n = n;
    return ZKP_OK;
}
#include "mgpu.cuh"
