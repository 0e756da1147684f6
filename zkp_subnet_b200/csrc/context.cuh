// Context, device buffers and the MSM driver shared by the C-ABI entry points.
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zkp_b200.h"
#include "host/pairing.hpp"
#include "msm_affine.cuh"

namespace zkp {

// ---- error plumbing -----------------------------------------------------------------------
inline std::string& tls_error() {
    static thread_local std::string e;
    return e;
}
inline int fail(int code, const std::string& msg) {
    tls_error() = msg;
    return code;
}
#define ZKP_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(ZKP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    } while (0)

// ---- growable device buffer ------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct MsmWorkspace {
    DevBuf keys_a, keys_b, vals_a, vals_b, cub_temp, buckets, next_a, next_b, pool, sums_a, sums_b, sums_out, bad;
    DevBuf sort_counters, sort_chunks;  // bucket sort: one cursor per key (+ padding to whole chunks), one start per chunk
    std::vector<DevBuf> slot_keys, slot_pts;
    // batched-affine rounds: per-round bucket starts, scan input, ping-pong point lists, keys of the last list,
    // prefix-product scratch
    DevBuf aff_start[8], aff_len, aff_pts[2], aff_keys, aff_scratch;
    G1Xyzz* h_window = nullptr;  // pinned, W records
    size_t h_window_cap = 0;
    uint32_t* h_bad = nullptr;   // pinned
    void release() {
        if (h_bad) cudaFreeHost(h_bad);
        h_bad = nullptr;
        for (DevBuf* b : {&keys_a, &keys_b, &vals_a, &vals_b, &cub_temp, &buckets, &next_a, &next_b, &pool, &sums_a, &sums_b, &sums_out, &bad,
                          &sort_counters, &sort_chunks})
            b->release();
        for (auto& b : slot_keys) b.release();
        for (auto& b : slot_pts) b.release();
        for (auto& b : aff_start) b.release();
        for (DevBuf* b : {&aff_len, &aff_pts[0], &aff_pts[1], &aff_keys, &aff_scratch}) b->release();
        if (h_window) cudaFreeHost(h_window);
        h_window = nullptr;
        h_window_cap = 0;
    }
};

}  // namespace zkp

namespace zkp {

// Fixed-base tables of the SRS rows, all in ONE device allocation cut into equal slots (every row has the same
// length and the window width is fixed per SRS, so every table has the same size).  A table entry is addressed by
// a 32-bit record index relative to the arena, which is what lets several rows -- the two MSMs of a commit+open, or a
// batch of requests on different rows -- share one launch set: the (bucket, point) entries carry arena-relative
// indices and the accumulation kernel gathers from one base pointer.  Slots are handed out on demand (LRU among the
// unpinned ones when the arena holds fewer slots than rows); a slot is pinned while an MSM that reads it is in flight.
struct TableArena {
    DevBuf buf;
    uint32_t c = 0, W = 0;
    size_t slot_records = 0;          // 2 * W * 2^log_n records of TABLE_STRIDE bytes (positive half, negated half)
    std::vector<int> slot_of_row;     // -1 = no table
    std::vector<int> row_of_slot;     // -1 = free
    std::vector<uint32_t> pins;
    std::vector<uint64_t> last_use;
    uint64_t clock = 0;
    uint64_t builds = 0, evictions = 0, fallbacks = 0;  // statistics (zkp_srs_table_stats)
    bool warned = false;
    void release() {
        buf.release();
        c = W = 0;
        slot_records = 0;
        slot_of_row.clear();
        row_of_slot.clear();
        pins.clear();
        last_use.clear();
    }
};

// Everything that describes the resident SRS: shared (by shared_ptr) between a context and the contexts forked from it
// with zkp_ctx_fork, which have their own streams and workspaces but read the same rows, tables and domain tables.
// Mutations (generate / import / load, table and twiddle builds) take `mu`; the compute paths only read.
struct SrsStore {
    std::recursive_mutex mu;
    uint32_t log_n = 0, log_m = 0;
    bool shaped = false;
    DevBuf srs;                               // 2^log_m rows x 2^log_n G1Affine (Montgomery)
    std::vector<uint8_t> row_loaded;          // per row flag
    std::vector<host::G1J> scale_points;      // [R_i(tau_y)]_1
    bool have_g2_tau = false;
    host::G2J g2_tau;                         // [tau_x]_2
    bool have_g2_tau_y = false;
    host::G2J g2_tau_y;                       // [tau_y]_2 (Pianist master verification only)
    host::G2Lines lines_g2, lines_tau, lines_tau_y;
    bool have_lines = false;
    uint32_t shard_domain_log = 0;            // log2 of the full domain when the rows are point-range shards
    uint32_t shard_index = 0;                 // which slice [shard * 2^log_n, (shard+1) * 2^log_n) of that domain
    uint32_t c_override = 0;
    size_t table_budget = 0;                  // bytes the arena may take (0 = 60% of the free HBM at first use)
    TableArena arena;
    // per-size domain tables: wt[k] = w_n^(2^k) (k <= log_n), tw[e] = w_n^e (e < n/2, built on demand)
    struct Domain {
        bool ready = false, have_tw = false;
        DevBuf wt, tw;
        host::Fr64 w, w_inv, n_inv;
    };
    std::vector<Domain> domains = std::vector<Domain>(32);
    DevBuf fixed_base;                        // [d*256^w]G table (SRS generation from a trapdoor)
    int device = 0;
    ~SrsStore() {
        int prev = -1;
        cudaGetDevice(&prev);
        if (prev != device) cudaSetDevice(device);
        arena.release();
        srs.release();
        fixed_base.release();
        for (auto& d : domains) { d.wt.release(); d.tw.release(); }
        if (prev >= 0 && prev != device) cudaSetDevice(prev);
    }
};

}  // namespace zkp

struct zkp_ctx {
    std::shared_ptr<zkp::SrsStore> store;
    zkp::SrsStore& S;                         // *store
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // SRS (in the shared store; the references keep the call sites short)
    uint32_t &log_n, &log_m;
    bool& shaped;
    zkp::DevBuf& srs;
    std::vector<uint8_t>& row_loaded;
    std::vector<zkp::host::G1J>& scale_points;
    bool& have_g2_tau;
    zkp::host::G2J& g2_tau;
    bool& have_g2_tau_y;
    zkp::host::G2J& g2_tau_y;
    zkp::host::G2Lines &lines_g2, &lines_tau, &lines_tau_y;
    bool& have_lines;
    uint32_t &shard_domain_log, &shard_index, &c_override;
    std::vector<zkp::SrsStore::Domain>& domains;
    zkp::DevBuf& fixed_base;
    typedef zkp::SrsStore::Domain Domain;
    explicit zkp_ctx(std::shared_ptr<zkp::SrsStore> st)
        : store(std::move(st)), S(*store), log_n(S.log_n), log_m(S.log_m), shaped(S.shaped), srs(S.srs), row_loaded(S.row_loaded),
          scale_points(S.scale_points), have_g2_tau(S.have_g2_tau), g2_tau(S.g2_tau), have_g2_tau_y(S.have_g2_tau_y),
          g2_tau_y(S.g2_tau_y), lines_g2(S.lines_g2), lines_tau(S.lines_tau), lines_tau_y(S.lines_tau_y), have_lines(S.have_lines),
          shard_domain_log(S.shard_domain_log), shard_index(S.shard_index), c_override(S.c_override), domains(S.domains),
          fixed_base(S.fixed_base) {}
    // scratch
    zkp::DevBuf scalars, fr_a, fr_b, fr_c, flush;
    // > 0: `scalars` still holds the n big-endian evaluations uploaded by the last worker_commit / worker_open /
    // worker_commit_open call (zkp_worker_open_resident); every other writer of `scalars` clears it.
    // resident_gen identifies that upload: it changes whenever `scalars` is rewritten, so that a caller who remembers
    // the generation of ITS upload (zkp_resident_generation) cannot be handed another caller's polynomial.
    size_t resident_n = 0;
    uint64_t resident_gen = 0;
    size_t staging_n = 0;                     // > 0: a chunked upload (zkp_stage_begin .. zkp_stage_end) is in progress
    uint64_t staging_gen = 0;
    zkp::MsmWorkspace ws;                     // lane 0 workspace (runs on `stream`)
    // lane 1: second stream + workspace so that the two MSMs of a commit+open (and their
    // latency-bound reduction tails) overlap on the device
    cudaStream_t stream2 = nullptr;
    zkp::MsmWorkspace ws2;
    cudaEvent_t ev_ready = nullptr;           // polynomial uploaded + converted (lane 0 -> lane 1)
    cudaEvent_t ev_join = nullptr;            // lane 1 -> lane 0 (fused launch sets)
    cudaEvent_t ev_acc2_0 = nullptr, ev_acc2_1 = nullptr;
    int bucket_sort = 2;                      // digits grouped by the hand-written counting sort (1), cub::DeviceRadixSort (0), by size (2)
    int affine_rounds_override = -1;          // <= 0: off (default, see plan_for); 1..6: rounds of batched-affine additions
    bool use_precomp = true;
    bool coeff_form = false;                  // worker_* polynomials arrive as coefficients (zkp_set_poly_form)
    bool ntt_tma = false;                     // NTT pass 2 tile by bulk async copy (experiment, zkp_set_ntt_tma)
    bool two_lanes_busy = false;              // set while a commit+open has both lanes' MSMs in flight (msm_enqueue_main)
    bool rowcol_coop = true;                  // small bucket arrays: cooperative row/column stages (zkp_set_rowcol_coop)
    bool open_coset = true;                   // single-request opening: coset blocks + host inversion (zkp_set_open_coset)
    int fuse_mode = -1;                       // commit+open as ONE grouped launch set: 1 always, 0 never, -1 by size
    zkp::DevBuf scratch_xyzz, scratch_fq, scratch_aff;
    uint64_t launches = 0;                    // kernels launched by this context (bench accounting)
    zkp::DevBuf small, partials, ntt_tmp;     // device scalars / block partial sums / NTT scratch
    zkp::DevBuf batch_small, batch_x;         // per-request scalars of a batch (SM_BYTES each) / evaluation points
    uint8_t* h_small = nullptr;               // pinned scratch (>= 4096 B)
    uint8_t* h_batch = nullptr;               // pinned, batch results
    size_t h_batch_cap = 0;
    zkp::host::G1J last_com, last_proof;      // results of the last commit+open as group elements (zkp_last_points_uncompressed)
    bool have_last = false;
    int pinned_slots[64];                     // arena slots pinned by the MSM(s) in flight on each lane (-1 = none)
    int pinned_count = 0;
    // kernel timing of the dominant kernel (k_accumulate level 0), enabled by the bench entries
    bool time_acc = false;
    cudaEvent_t ev_acc0 = nullptr, ev_acc1 = nullptr;
    double acc_ms_total = 0;
    uint64_t acc_count = 0;
    float last_acc_ms = 0;
    // stage timeline (zkp_bench_trace): events recorded on the lanes' streams between pipeline stages
    struct TracePoint { cudaEvent_t ev; int lane; const char* name; };
    bool trace_on = false;
    std::vector<TracePoint> trace;
};

namespace zkp {
// marks "everything enqueued so far on this lane's stream has finished" under a stage name
inline void trace_mark(zkp_ctx* ctx, int lane, cudaStream_t st, const char* name) {
    if (!ctx->trace_on) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ctx->trace.push_back({e, lane, name});
}
}  // namespace zkp
