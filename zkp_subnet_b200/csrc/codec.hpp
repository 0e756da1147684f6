// Batch base64 <-> 32-byte codec for the List[str] wire format of the Prove synapse
// (reference base/protocol.py:35-40; strings are standard-alphabet base64 of 32 big-endian bytes,
// 43 characters unpadded as in reference tests/test_miner.py:33-55, 44 with '=' also accepted).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace zkp {
namespace codec {

inline const int8_t* b64_table() {
    static int8_t t[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) t[i] = -1;
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) t[(uint8_t)a[i]] = (int8_t)i;
        init = true;
    }
    return t;
}

// one 43/44-char string -> 32 bytes; false on any invalid character or non-zero trailing bits
inline bool b64_decode32(const char* s, uint8_t out[32]) {
    const int8_t* t = b64_table();
    uint32_t acc = 0;
    int bits = 0, o = 0;
    for (int i = 0; i < 43; i++) {
        int v = t[(uint8_t)s[i]];
        if (v < 0) return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            if (o < 32) out[o++] = (uint8_t)(acc >> bits);
            acc &= (1u << bits) - 1;
        }
    }
    return o == 32 && acc == 0;  // 43*6 = 258 bits: the last 2 must be zero
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

}  // namespace codec
}  // namespace zkp
