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

#include "codec.hpp"

using namespace zkp;

extern "C" {

// list of str/bytes (43 or 44 chars each) -> out[32 * n].  Returns n, or -1 - i when element i is not a
// valid field-element string, or a value <= -(1 << 40) for a wrong argument type / too small a buffer.
long long zkp_wire_decode_list(PyObject* seq, uint8_t* out, size_t capacity) {
    const long long BAD_ARG = -(1ll << 40);
    if (!seq || !out || !(PyList_Check(seq) || PyTuple_Check(seq))) return BAD_ARG;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if ((size_t)n * 32 > capacity) return BAD_ARG - 1;
    PyObject** items = PySequence_Fast_ITEMS(seq);
    std::vector<const char*> ptrs((size_t)n);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject* it = items[i];
        const char* p = nullptr;
        Py_ssize_t len = 0;
        if (PyUnicode_Check(it)) {
            p = PyUnicode_AsUTF8AndSize(it, &len);
            if (!p) { PyErr_Clear(); return -1 - (long long)i; }
        } else if (PyBytes_Check(it)) {
            char* q = nullptr;
            if (PyBytes_AsStringAndSize(it, &q, &len) < 0) { PyErr_Clear(); return -1 - (long long)i; }
            p = q;
        } else {
            return -1 - (long long)i;
        }
        if (!(len == 43 || (len == 44 && p[43] == '='))) return -1 - (long long)i;
        ptrs[(size_t)i] = p;
    }
    size_t bad = codec::b64_decode_ptrs(ptrs.data(), (size_t)n, out);
    if (bad != (size_t)n) return -1 - (long long)bad;
    return (long long)n;
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
