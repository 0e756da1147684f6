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
#include <string.h>

#include <atomic>
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

// runs fn(begin, end) over [0, count) on codec_threads(count) host threads
template <class Fn>
inline void parallel_ranges(size_t count, Fn fn, unsigned max_threads = 0) {
    unsigned nt = codec_threads(count);
    if (max_threads && nt > max_threads) nt = max_threads;
    if (nt <= 1) { fn((size_t)0, count); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (unsigned t = 1; t < nt; t++) th.emplace_back(fn, count * t / nt, count * (t + 1) / nt);
    fn((size_t)0, count / nt);
    for (auto& x : th) x.join();
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
