// Batch base64 <-> 32-byte codec for the List[str] wire format of the Prove synapse
// (reference base/protocol.py:35-40; strings are standard-alphabet base64 of 32 big-endian bytes,
// 43 characters unpadded as in reference tests/test_miner.py:33-55, 44 with '=' also accepted).
//
// At n = 2^20 the wire form of a polynomial is 45 MB of text: decoded one character at a time it costs
// ten times the GPU's commit+open.  The decoder therefore works on 4-character groups through four
// pre-shifted lookup tables (invalid characters carry a flag bit that is OR-ed through the whole
// element and tested once) and the batch entry points split the elements over host threads.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace zkp {
namespace codec {

struct B64Tables {
    uint32_t t[4][256];  // t[k][c] = value(c) << (18 - 6k), or BAD
    static constexpr uint32_t BAD = 0x80000000u;
    B64Tables() {
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int k = 0; k < 4; k++) {
            for (int c = 0; c < 256; c++) t[k][c] = BAD;
            for (int i = 0; i < 64; i++) t[k][(uint8_t)a[i]] = (uint32_t)i << (18 - 6 * k);
        }
    }
};
inline const B64Tables& b64_tables() {
    static const B64Tables tabs;
    return tabs;
}

// one 43/44-char string -> 32 bytes; false on any invalid character or non-zero trailing bits
inline bool b64_decode32(const char* s, uint8_t out[32]) {
    const B64Tables& T = b64_tables();
    const uint8_t* u = reinterpret_cast<const uint8_t*>(s);
    uint32_t bad = 0;
#pragma GCC unroll 10
    for (int g = 0; g < 10; g++) {
        uint32_t v = T.t[0][u[4 * g]] | T.t[1][u[4 * g + 1]] | T.t[2][u[4 * g + 2]] | T.t[3][u[4 * g + 3]];
        bad |= v;
        out[3 * g] = (uint8_t)(v >> 16);
        out[3 * g + 1] = (uint8_t)(v >> 8);
        out[3 * g + 2] = (uint8_t)v;
    }
    uint32_t v = T.t[0][u[40]] | T.t[1][u[41]] | T.t[2][u[42]];
    bad |= v;
    out[30] = (uint8_t)(v >> 16);
    out[31] = (uint8_t)(v >> 8);
    // 43 * 6 = 258 bits: the last 2 must be zero
    return !(bad & B64Tables::BAD) && !(v & 0xc0u);
}

#if defined(__x86_64__) && defined(__GNUC__)
#define ZKP_CODEC_AVX2 1
}  // namespace codec
}  // namespace zkp
#include <immintrin.h>
namespace zkp {
namespace codec {
inline bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
// 32 base64 characters -> 24 bytes in the low 192 bits; *bad accumulates a non-zero byte for every invalid character.
// Validity and value come from the two nibbles of a character c: hi = c >> 4 selects one bit, lo = c & 15 selects the
// set of hi values for which (hi, lo) is in the alphabet ('A'-'O' 4x, 'P'-'Z' 5x, 'a'-'o' 6x, 'p'-'z' 7x, '0'-'9' 3x,
// '+' 2B, '/' 2F); the 6-bit value is c plus an offset that depends on hi only, except for '/'.
__attribute__((target("avx2"))) inline __m256i b64_decode_32chars(__m256i in, __m256i* bad) {
    const __m256i lo_sets = _mm256_setr_epi8(
        (char)0xa8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8,
        (char)0xf8, (char)0xf8, (char)0xf0, (char)0x54, (char)0x50, (char)0x50, (char)0x50, (char)0x54,
        (char)0xa8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8, (char)0xf8,
        (char)0xf8, (char)0xf8, (char)0xf0, (char)0x54, (char)0x50, (char)0x50, (char)0x50, (char)0x54);
    const __m256i hi_bit = _mm256_setr_epi8(1, 2, 4, 8, 16, 32, 64, (char)128, 0, 0, 0, 0, 0, 0, 0, 0,
                                            1, 2, 4, 8, 16, 32, 64, (char)128, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i offset = _mm256_setr_epi8(0, 0, 19, 4, -65, -65, -71, -71, 0, 0, 0, 0, 0, 0, 0, 0,
                                            0, 0, 19, 4, -65, -65, -71, -71, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i nib = _mm256_set1_epi8(0x0f);
    const __m256i hi = _mm256_and_si256(_mm256_srli_epi32(in, 4), nib);
    const __m256i lo = _mm256_and_si256(in, nib);
    const __m256i ok = _mm256_and_si256(_mm256_shuffle_epi8(lo_sets, lo), _mm256_shuffle_epi8(hi_bit, hi));
    *bad = _mm256_or_si256(*bad, _mm256_cmpeq_epi8(ok, _mm256_setzero_si256()));
    const __m256i is_slash = _mm256_cmpeq_epi8(in, _mm256_set1_epi8(0x2f));
    const __m256i off = _mm256_blendv_epi8(_mm256_shuffle_epi8(offset, hi), _mm256_set1_epi8(16), is_slash);
    const __m256i v = _mm256_add_epi8(in, off);
    // four 6-bit values -> one 24-bit group, then the three bytes of every group in stream order
    const __m256i t = _mm256_maddubs_epi16(v, _mm256_set1_epi32(0x01400140));
    const __m256i g = _mm256_madd_epi16(t, _mm256_set1_epi32(0x00011000));
    const __m256i bytes = _mm256_shuffle_epi8(g, _mm256_setr_epi8(2, 1, 0, 6, 5, 4, 10, 9, 8, 14, 13, 12, -1, -1, -1, -1,
                                                                 2, 1, 0, 6, 5, 4, 10, 9, 8, 14, 13, 12, -1, -1, -1, -1));
    return _mm256_permutevar8x32_epi32(bytes, _mm256_setr_epi32(0, 1, 2, 4, 5, 6, 3, 7));
}
// b64_decode32 with two overlapping 32-character blocks: characters 0..31 give bytes 0..23, characters 12..43 (the
// 44th replaced by 'A') give bytes 9..32, and byte 32 holds exactly the two trailing bits that have to be zero.
// READS s[43]: the caller guarantees 44 readable bytes (a Python str / bytes object always has its terminator).
__attribute__((target("avx2"))) inline bool b64_decode32_avx2(const char* s, uint8_t out[32]) {
    __m256i bad = _mm256_setzero_si256();
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 12));
    b = _mm256_blendv_epi8(b, _mm256_set1_epi8('A'), _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                                                       0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1));
    alignas(32) uint8_t t0[32], t1[32];
    _mm256_store_si256(reinterpret_cast<__m256i*>(t0), b64_decode_32chars(a, &bad));
    _mm256_store_si256(reinterpret_cast<__m256i*>(t1), b64_decode_32chars(b, &bad));
    memcpy(out, t0, 16);
    memcpy(out + 16, t1 + 7, 16);
    return _mm256_testz_si256(bad, bad) && t1[23] == 0;
}
// The same with non-temporal stores (out 16-byte aligned; the caller issues _mm_sfence() when its range is done): a
// polynomial decoded into page-locked memory is read next by the GPU's DMA engine, not by a core, and a host->device
// copy of 32 MiB that sixteen cores have just left dirty in their caches was measured 0.9 ms slower than the same
// copy of data at rest.
__attribute__((target("avx2"))) inline bool b64_decode32_avx2_stream(const char* s, uint8_t out[32]) {
    __m256i bad = _mm256_setzero_si256();
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 12));
    b = _mm256_blendv_epi8(b, _mm256_set1_epi8('A'), _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                                                       0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1));
    alignas(32) uint8_t t1[32];
    const __m256i d0 = b64_decode_32chars(a, &bad);
    _mm256_store_si256(reinterpret_cast<__m256i*>(t1), b64_decode_32chars(b, &bad));
    _mm_stream_si128(reinterpret_cast<__m128i*>(out), _mm256_castsi256_si128(d0));
    _mm_stream_si128(reinterpret_cast<__m128i*>(out + 16), _mm_loadu_si128(reinterpret_cast<const __m128i*>(t1 + 7)));
    return _mm256_testz_si256(bad, bad) && t1[23] == 0;
}
#endif

inline void b64_encode32(const uint8_t in[32], char out[43]) {
    static const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    int o = 0;
    for (int i = 0; i < 30; i += 3) {
        uint32_t v = (uint32_t)in[i] << 16 | (uint32_t)in[i + 1] << 8 | in[i + 2];
        out[o++] = a[v >> 18]; out[o++] = a[(v >> 12) & 63]; out[o++] = a[(v >> 6) & 63]; out[o++] = a[v & 63];
    }
    uint32_t v = (uint32_t)in[30] << 16 | (uint32_t)in[31] << 8;
    out[o++] = a[v >> 18]; out[o++] = a[(v >> 12) & 63]; out[o++] = a[(v >> 6) & 63];
}

inline unsigned codec_threads(size_t count) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    if (hw > 16) hw = 16;
    size_t by_work = count / 8192 + 1;  // below ~8k elements a thread costs more than it saves
    return (unsigned)(by_work < hw ? by_work : hw);
}

// Persistent host threads for the batch codecs.  A polynomial crosses the wire codec once or twice per request
// (reference neurons/miner.py:56-61 ships it with worker_commit AND worker_open); creating and joining fifteen
// std::threads costs ~0.13 ms per call on the GPU host -- a tenth of a whole request at the mainnet row size (2^16).
// The workers are created on first use and then sleep on a condition variable between calls; after a job they keep
// polling for ~20 us, because the second call of a request follows the first within that time.  One job at a time: a
// caller that finds the pool busy (two requests decoding at once from different threads) spawns threads the old way.
// A fork()ed child has no workers: the pool is dropped in the child (pthread_atfork) and rebuilt on first use.
class WorkerPool {
   public:
    typedef void (*RangeFn)(void* ctx, size_t lo, size_t hi);
    static WorkerPool*& instance() {
        static WorkerPool* p = nullptr;
        return p;
    }
    static std::mutex& guard() {
        static std::mutex m;
        return m;
    }
    // nullptr when the pool is busy with another caller's job
    static WorkerPool* acquire(unsigned want_workers) {
        std::mutex& g = guard();
        if (!g.try_lock()) return nullptr;
        WorkerPool*& p = instance();
        if (!p) {
            static bool hooked = false;
            if (!hooked) {
                hooked = true;
                pthread_atfork(nullptr, nullptr, [] {
                    // child: the worker threads do not exist here; forget the pool (and the lock the parent may have held)
                    new (&guard()) std::mutex();
                    instance() = nullptr;
                });
            }
            p = new WorkerPool();
        }
        p->grow(want_workers);
        return p;  // guard stays locked until release()
    }
    static void release() { guard().unlock(); }

    // runs fn(ctx, count*t/nt, count*(t+1)/nt) for t in [0, nt): t = 0 on the calling thread, the rest on workers
    void run(unsigned nt, size_t count, RangeFn fn, void* ctx) {
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = fn; ctx_ = ctx; count_ = count; nt_ = nt;
            pending_.store((int)nt - 1, std::memory_order_relaxed);
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        fn(ctx, 0, count / nt);
        // the workers' share is a few hundred microseconds at most: poll, then sleep
        for (int spin = 0; pending_.load(std::memory_order_acquire) > 0; spin++) {
            if (spin < 4096) { cpu_relax(); continue; }
            std::unique_lock<std::mutex> lk(m_);
            done_cv_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) <= 0; });
        }
    }
    unsigned workers() const { return (unsigned)th_.size(); }

   private:
    static inline void cpu_relax() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    void grow(unsigned want) {
        while (th_.size() < want) {
            const unsigned id = (unsigned)th_.size() + 1;  // share index of this worker
            const uint64_t seen = epoch_.load(std::memory_order_acquire);
            th_.emplace_back([this, id, seen] { loop(id, seen); });
            th_.back().detach();
        }
    }
    void loop(unsigned id, uint64_t seen) {
        for (;;) {
            // short poll (the next call of the same request is microseconds away), then sleep
            bool got = false;
            for (int spin = 0; spin < 1000; spin++) {
                if (epoch_.load(std::memory_order_acquire) != seen) { got = true; break; }
                cpu_relax();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
            }
            RangeFn fn; void* ctx; size_t count; unsigned nt;
            {
                std::lock_guard<std::mutex> lk(m_);
                seen = epoch_.load(std::memory_order_acquire);
                fn = fn_; ctx = ctx_; count = count_; nt = nt_;
            }
            if (id < nt) {
                fn(ctx, count * id / nt, count * (id + 1) / nt);
                if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                    std::lock_guard<std::mutex> lk(m_);
                    done_cv_.notify_one();
                }
            }
        }
    }
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    std::atomic<uint64_t> epoch_{0};
    std::atomic<int> pending_{0};
    RangeFn fn_ = nullptr;
    void* ctx_ = nullptr;
    size_t count_ = 0;
    unsigned nt_ = 0;
    std::vector<std::thread> th_;
};

// runs fn(begin, end) over [0, count) on codec_threads(count) host threads (the persistent pool, or freshly spawned
// threads when the pool is serving another caller)
// fn(begin, end) over [0, count) split evenly over exactly nt threads (the calling thread included)
template <class Fn>
inline void parallel_ranges_n(size_t count, unsigned nt, Fn fn) {
    if (nt <= 1 || count <= 1) { fn((size_t)0, count); return; }
    static const bool use_pool = [] {
        const char* e = getenv("ZKP_CODEC_POOL");  // "0": fresh threads per call (A/B runs, tools/pool_ab.py)
        return !(e && e[0] == '0');
    }();
    if (WorkerPool* pool = use_pool ? WorkerPool::acquire(nt - 1) : nullptr) {
        struct Release { ~Release() { WorkerPool::release(); } } rel;
        pool->run(nt, count, [](void* c, size_t lo, size_t hi) { (*static_cast<Fn*>(c))(lo, hi); }, &fn);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (unsigned t = 1; t < nt; t++) th.emplace_back(fn, count * t / nt, count * (t + 1) / nt);
    fn((size_t)0, count / nt);
    for (auto& x : th) x.join();
}
template <class Fn>
inline void parallel_ranges(size_t count, Fn fn, unsigned max_threads = 0) {
    unsigned nt = codec_threads(count);
    if (max_threads && nt > max_threads) nt = max_threads;
    parallel_ranges_n(count, nt, fn);
}

// strings at base + i * stride; returns the index of the first invalid element or count
inline size_t b64_decode_batch(const char* base, size_t stride, size_t count, uint8_t* out) {
    std::atomic<size_t> first_bad(count);
    parallel_ranges(count, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++)
            if (!b64_decode32(base + i * stride, out + 32 * i)) {
                size_t cur = first_bad.load();
                while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
                return;
            }
    });
    return first_bad.load();
}
// strings at ptrs[i]
inline size_t b64_decode_ptrs(const char* const* ptrs, size_t count, uint8_t* out) {
    std::atomic<size_t> first_bad(count);
    parallel_ranges(count, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++)
            if (!b64_decode32(ptrs[i], out + 32 * i)) {
                size_t cur = first_bad.load();
                while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
                return;
            }
    });
    return first_bad.load();
}
inline void b64_encode_batch(const uint8_t* in, size_t count, char* out /* 43 per element */) {
    parallel_ranges(count, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) b64_encode32(in + 32 * i, out + 43 * i);
    });
}

}  // namespace codec
}  // namespace zkp
