// Context, device buffers and the MSM driver shared by the C-ABI entry points.
#pragma once
#include <cuda_runtime.h>
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

struct zkp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // SRS
    uint32_t log_n = 0, log_m = 0;
    bool shaped = false;
    zkp::DevBuf srs;                          // 2^log_m rows x 2^log_n G1Affine (Montgomery)
    std::vector<uint8_t> row_loaded;          // per row flag
    std::vector<zkp::host::G1J> scale_points; // [R_i(tau_y)]_1
    bool have_g2_tau = false;
    zkp::host::G2J g2_tau;                    // [tau_x]_2
    bool have_g2_tau_y = false;
    zkp::host::G2J g2_tau_y;                  // [tau_y]_2 (Pianist master verification only)
    // scratch
    zkp::DevBuf scalars, fr_a, fr_b, fr_c, flush;
    // > 0: `scalars` still holds the n big-endian evaluations uploaded by the last worker_commit / worker_open /
    // worker_commit_open call (zkp_worker_open_resident); every other writer of `scalars` clears it
    size_t resident_n = 0;
    zkp::MsmWorkspace ws;                     // lane 0 workspace (runs on `stream`)
    // lane 1: second stream + workspace so that the two MSMs of a commit+open (and their
    // latency-bound reduction tails) overlap on the device
    cudaStream_t stream2 = nullptr;
    zkp::MsmWorkspace ws2;
    cudaEvent_t ev_ready = nullptr;           // polynomial uploaded + converted (lane 0 -> lane 1)
    cudaEvent_t ev_acc2_0 = nullptr, ev_acc2_1 = nullptr;
    uint32_t c_override = 0;
    int bucket_sort = 2;                      // digits grouped by the hand-written counting sort (1), cub::DeviceRadixSort (0), by size (2)
    int affine_rounds_override = -1;          // <= 0: off (default, see plan_for); 1..6: rounds of batched-affine additions
    // fixed-base tables: per SRS row, [2^(c w)] P_i for w < W (slice w at w * 2^log_n); built lazily
    struct Precomp { zkp::DevBuf table; uint32_t c = 0, W = 0; };
    std::vector<Precomp> precomp;
    bool use_precomp = true;
    zkp::DevBuf scratch_xyzz, scratch_fq, scratch_aff;
    uint32_t shard_domain_log = 0;            // log2 of the full domain when the rows are point-range shards
    uint32_t shard_index = 0;                 // which slice [shard * 2^log_n, (shard+1) * 2^log_n) of that domain
    uint64_t launches = 0;                    // kernels launched by this context (bench accounting)
    // per-size domain tables: wt[k] = w_n^(2^k) (k <= log_n), tw[e] = w_n^e (e < n/2, built on demand)
    struct Domain {
        bool ready = false, have_tw = false;
        zkp::DevBuf wt, tw;
        zkp::host::Fr64 w, w_inv, n_inv;
    };
    std::vector<Domain> domains = std::vector<Domain>(32);
    zkp::DevBuf small, partials, ntt_tmp, fixed_base;  // device scalars / block partial sums / NTT scratch / [d*256^w]G table
    uint8_t* h_small = nullptr;               // pinned scratch (>= 256 B)
    // pairing data fixed per SRS
    zkp::host::G2Lines lines_g2, lines_tau, lines_tau_y;
    bool have_lines = false;
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
