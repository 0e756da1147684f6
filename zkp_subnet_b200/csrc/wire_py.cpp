// CPython-side helper of the wire codec: the reference hands a polynomial to its prover as a Python
// List[str] of base64 field elements (reference base/protocol.py:35-40, neurons/miner.py:39,48).  Turning
// that list into n x 32 bytes with "".join + encode + a flat decode costs ~0.3 s at n = 2^20 -- twenty
// times the GPU's commit+open -- so this module walks the list through the C API (the UTF-8 buffer of an
// ASCII str is borrowed, not copied) and decodes straight into the caller's (page-locked) buffer on
// several host threads.  It is loaded with ctypes.PyDLL (GIL held for the whole call); the CUDA library
// itself stays free of any Python dependency.
//
// Build: g++ -O3 -shared -fPIC -I<python include> wire_py.cpp -o _zkp_wire.so -lpthread
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <atomic>

#include "codec.hpp"

using namespace zkp;

extern "C" {

// list of str/bytes (43 or 44 chars each) -> out[32 * n].  Returns n, or -1 - i when element i is not a
// valid field-element string, or a value <= -(1 << 40) for a wrong argument type / too small a buffer.
// The compare variant runs BESIDE a prover call that is launching kernels from another thread of this process: it
// takes half of the cores so that the launching thread is never waiting for a time slice.
static unsigned cmp_threads() {
    unsigned hw = std::thread::hardware_concurrency();
    return hw >= 4 ? hw / 2 : 1;
}
// ref != nullptr: also compares every decoded element with ref[32 i ..] and reports *same (the fourier.Client
// shim uses it to recognise the polynomial it was handed by the previous call, see client.py worker_open)
// one element whose buffer has at least 44 readable bytes (str and bytes objects keep a terminator)
static inline bool decode_one(const char* p, uint8_t* out, bool avx2, bool stream = false) {
#ifdef ZKP_CODEC_AVX2
    if (avx2) return stream ? codec::b64_decode32_avx2_stream(p, out) : codec::b64_decode32_avx2(p, out);
#endif
    (void)avx2;
    (void)stream;
    return codec::b64_decode32(p, out);
}
// decodes elements [first, first + count) of the list into out + 32 first (every index below is relative to `first`)
static long long decode_range_impl(PyObject** items_all, size_t first, Py_ssize_t n, uint8_t* out_all, const uint8_t* ref_all, int* same);
static long long decode_list_impl(PyObject* seq, uint8_t* out, size_t capacity, const uint8_t* ref, int* same) {
    const long long BAD_ARG = -(1ll << 40);
    if (!seq || !out || !(PyList_Check(seq) || PyTuple_Check(seq))) return BAD_ARG;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if ((size_t)n * 32 > capacity) return BAD_ARG - 1;
    return decode_range_impl(PySequence_Fast_ITEMS(seq), 0, n, out, ref, same);
}
static long long decode_range_impl(PyObject** items_all, size_t first, Py_ssize_t n, uint8_t* out_all, const uint8_t* ref_all, int* same) {
    PyObject** items = items_all + first;
    uint8_t* out = out_all + 32 * first;
    const uint8_t* ref = ref_all ? ref_all + 32 * first : nullptr;
    // Walk AND decode on the host threads.  The calling thread holds the GIL for the whole call, so no other
    // Python code runs and the list keeps every element alive; the workers only READ immutable object fields
    // through macros (type flags, length, the inline buffer of a compact ASCII str / of a bytes object) -- no
    // C-API call that could allocate or touch the error indicator.  Anything else (a non-ASCII or non-compact
    // str, a str subclass with its own buffer layout, another type) is left to the serial pass below, which
    // uses the regular API with the GIL held.  (Measured on the 16-core GPU host at n = 2^20: a fourier.Client commit+open call went from 26.9 to 16.1 ms when the
    // walk moved from the calling thread to the workers.)
    std::atomic<size_t> first_bad((size_t)n), first_slow((size_t)n);
    std::atomic<int> differs(0);
#ifdef ZKP_CODEC_AVX2
    const bool avx2 = codec::have_avx2();
#else
    const bool avx2 = false;
#endif
    // non-temporal stores for a large result that nobody on the host reads back (not the compare variant)
    const bool stream = avx2 && !ref && n >= 4096 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    auto lower = [](std::atomic<size_t>& a, size_t i) {
        size_t cur = a.load();
        while (i < cur && !a.compare_exchange_weak(cur, i)) {}
    };
    codec::parallel_ranges((size_t)n, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            PyObject* it = items[i];
            const char* p;
            Py_ssize_t len;
            if (PyUnicode_CheckExact(it) && PyUnicode_IS_COMPACT_ASCII(it)) {
                p = reinterpret_cast<const char*>(PyUnicode_1BYTE_DATA(it));
                len = PyUnicode_GET_LENGTH(it);
            } else if (PyBytes_CheckExact(it)) {
                p = PyBytes_AS_STRING(it);
                len = PyBytes_GET_SIZE(it);
            } else {
                lower(first_slow, i);
                continue;
            }
            if (!(len == 43 || (len == 44 && p[43] == '=')) || !decode_one(p, out + 32 * i, avx2, stream)) {
                lower(first_bad, i);
                break;
            }
            if (ref && memcmp(out + 32 * i, ref + 32 * i, 32) != 0) differs.store(1, std::memory_order_relaxed);
        }
#ifdef ZKP_CODEC_AVX2
        if (stream) _mm_sfence();
#endif
    }, ref ? cmp_threads() : 0u);
    // serial pass over the elements the workers skipped (none for the lists the reference sends)
    for (size_t i = first_slow.load(); i < (size_t)n && i < first_bad.load(); i++) {
        PyObject* it = items[i];
        if ((PyUnicode_CheckExact(it) && PyUnicode_IS_COMPACT_ASCII(it)) || PyBytes_CheckExact(it)) continue;  // done above
        const char* p = nullptr;
        Py_ssize_t len = 0;
        if (PyUnicode_Check(it)) {
            p = PyUnicode_AsUTF8AndSize(it, &len);
            if (!p) { PyErr_Clear(); lower(first_bad, i); break; }
        } else if (PyBytes_Check(it)) {
            char* q = nullptr;
            if (PyBytes_AsStringAndSize(it, &q, &len) < 0) { PyErr_Clear(); lower(first_bad, i); break; }
            p = q;
        } else {
            lower(first_bad, i);
            break;
        }
        if (!(len == 43 || (len == 44 && p[43] == '=')) || !decode_one(p, out + 32 * i, avx2)) {
            lower(first_bad, i);
            break;
        }
        if (ref && memcmp(out + 32 * i, ref + 32 * i, 32) != 0) differs.store(1, std::memory_order_relaxed);
    }
    if (same) *same = differs.load() ? 0 : 1;
    if (first_bad.load() != (size_t)n) return -1 - (long long)first_bad.load();
    return (long long)n;
}
long long zkp_wire_decode_list(PyObject* seq, uint8_t* out, size_t capacity) {
    return decode_list_impl(seq, out, capacity, nullptr, nullptr);
}
// The same in chunks of `chunk` elements: after each chunk, on_chunk(user, first, out, count) is called on the calling
// thread (the shim passes zkp_stage_chunk of libzkp_b200.so and its context: the chunk's host-to-device copy is enqueued
// and runs while the next chunk is being decoded).  A non-zero return of the callback aborts with BAD_ARG - 2.
typedef int (*zkp_wire_chunk_fn)(void* user, size_t first, const uint8_t* base, size_t count);
long long zkp_wire_decode_list_chunked(PyObject* seq, uint8_t* out, size_t capacity, size_t chunk, zkp_wire_chunk_fn on_chunk, void* user) {
    const long long BAD_ARG = -(1ll << 40);
    if (!seq || !out || !chunk || !on_chunk || !(PyList_Check(seq) || PyTuple_Check(seq))) return BAD_ARG;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if ((size_t)n * 32 > capacity) return BAD_ARG - 1;
    PyObject** items = PySequence_Fast_ITEMS(seq);
    for (size_t first = 0; first < (size_t)n; first += chunk) {
        const size_t count = (size_t)n - first < chunk ? (size_t)n - first : chunk;
        const long long r = decode_range_impl(items, first, (Py_ssize_t)count, out, nullptr, nullptr);
        if (r != (long long)count) return r < 0 && r > BAD_ARG ? r - (long long)first : r;  // index of the bad element, list-relative
        if (on_chunk(user, first, out, count) != 0) return BAD_ARG - 2;
    }
    return (long long)n;
}
long long zkp_wire_decode_list_cmp(PyObject* seq, uint8_t* out, size_t capacity, const uint8_t* ref, int* same) {
    if (!ref || !same) return -(1ll << 40);
    return decode_list_impl(seq, out, capacity, ref, same);
}

// in[32 * n] -> new list of n str (43 chars each, unpadded); nullptr with a Python error set on failure
PyObject* zkp_wire_encode_list(const uint8_t* in, size_t n) {
    PyObject* list = PyList_New((Py_ssize_t)n);
    if (!list) return nullptr;
    // create the str objects first (needs the GIL), then fill their buffers in parallel
    std::vector<char*> bufs(n);
    for (size_t i = 0; i < n; i++) {
        PyObject* s = PyUnicode_New(43, 127);
        if (!s) { Py_DECREF(list); return nullptr; }
        bufs[i] = reinterpret_cast<char*>(PyUnicode_1BYTE_DATA(s));
        PyList_SET_ITEM(list, (Py_ssize_t)i, s);
    }
    codec::parallel_ranges(n, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) codec::b64_encode32(in + 32 * i, bufs[i]);
    });
    return list;
}

}  // extern "C"
