// Carry-chain primitives for multi-limb integer arithmetic on sm_100a.
//
// Device: one PTX instruction each (add.cc / addc / mad.lo.cc / madc.hi.cc ...); ptxas fuses a
// mad.lo.cc + madc.hi.cc pair on the same operands into a single IMAD.WIDE.U32(.X) with the carry
// held in a predicate register.
// Host  : bit-exact emulation with an explicit thread-local carry flag, so the *same* limb
// algorithms (ff.cuh, g1.cuh) can be unit-tested by g++ without a GPU.  The host build of these
// templates is test scaffolding only; the product's host arithmetic is host/field64.hpp.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKP_HD __host__ __device__ __forceinline__
#define ZKP_D __device__ __forceinline__
#else
#define ZKP_HD inline
#define ZKP_D inline
#endif

namespace zkp {
namespace ptx {

#if defined(__CUDA_ARCH__)

ZKP_D uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t subc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t mul_lo(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t mul_hi(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
ZKP_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
ZKP_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
ZKP_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
ZKP_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
ZKP_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

#else  // ---------------------------------------------------------------- host emulation

inline uint32_t& cf() {
    static thread_local uint32_t flag = 0;
    return flag;
}
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool set) {
    uint64_t s = (uint64_t)a + b + cin;
    if (set) cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool set) {
    uint64_t d = (uint64_t)a - b - bin;
    if (set) cf() = (uint32_t)((d >> 32) & 1);
    return (uint32_t)d;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, cf(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3(a, b, cf(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, cf(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3(a, b, cf(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, 0, true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, cf(), true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, cf(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, cf(), false); }

#endif

}  // namespace ptx
}  // namespace zkp
