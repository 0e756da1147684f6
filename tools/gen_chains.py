#!/usr/bin/env python3
"""Generate zkp_subnet_b200/csrc/mont_chains.cuh.

Every carry chain of the Fq/Fr arithmetic is emitted as ONE inline-asm statement, so the PTX
condition-code register never lives across two asm statements: the statements are pure functions
of their operands (no `volatile` needed, the compiler may schedule/CSE them freely) and ptxas fuses
each mad.lo.cc/madc.hi.cc pair into a single IMAD.WIDE.U32.X.

The host (non-__CUDA_ARCH__) bodies run the same instruction sequences on the emulated carry flag of
ptx_chain.cuh; they exist so the limb algorithms can be unit-tested without a GPU."""
import os

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


class Block:
    """Collects PTX lines over symbolic operands and renders device asm + host emulation."""

    def __init__(self):
        self.lines = []  # (op, dst, srcs)
        self.ops = {}  # name -> (kind, cexpr) kind in {"+r","=r","r"}
        self.order = []

    def operand(self, name, kind, cexpr):
        if name not in self.ops:
            self.ops[name] = [kind, cexpr]
            self.order.append(name)
        else:
            # upgrade plain input to in/out if later written
            k = self.ops[name][0]
            if k != kind and "+r" in (k, kind):
                self.ops[name][0] = "+r"
        return name

    def emit(self, op, dst, *srcs):
        self.lines.append((op, dst, srcs))

    def render(self, indent="    "):
        outs = [n for n in self.order if self.ops[n][0] in ("+r", "=r")]
        ins = [n for n in self.order if self.ops[n][0] == "r"]
        idx = {n: i for i, n in enumerate(outs + ins)}

        def ref(s):
            if isinstance(s, int):
                return f"0x{s:08x}" if s else "0"
            return f"%{idx[s]}"

        text = []
        for op, dst, srcs in self.lines:
            text.append(f'"{op} {ref(dst)}, {", ".join(ref(s) for s in srcs)};\\n\\t"')
        dev = indent + "asm(" + ("\n" + indent + "    ").join(text) + "\n"
        dev += indent + "    : " + ", ".join(f'"{self.ops[n][0]}"({self.ops[n][1]})' for n in outs) + "\n"
        dev += indent + "    : " + ", ".join(f'"r"({self.ops[n][1]})' for n in ins) + ");\n"

        def cref(s):
            if isinstance(s, int):
                return f"0x{s:08x}u"
            return self.ops[s][1]

        hostmap = {
            "add.cc.u32": "add_cc", "addc.cc.u32": "addc_cc", "addc.u32": "addc",
            "sub.cc.u32": "sub_cc", "subc.cc.u32": "subc_cc", "subc.u32": "subc",
            "mad.lo.cc.u32": "mad_lo_cc", "madc.lo.cc.u32": "madc_lo_cc",
            "madc.hi.cc.u32": "madc_hi_cc", "madc.hi.u32": "madc_hi",
            "mul.lo.u32": "mul_lo", "mul.hi.u32": "mul_hi",
        }
        host = ""
        for op, dst, srcs in self.lines:
            host += indent + f"{cref(dst)} = ptx::{hostmap[op]}({', '.join(cref(s) for s in srcs)});\n"
        return dev, host


def func(sig, blocks_and_code):
    """blocks_and_code: list of Block or raw C strings (shared by device and host)."""
    dev = host = ""
    for item in blocks_and_code:
        if isinstance(item, Block):
            d, h = item.render()
            dev += d
            host += h
        else:
            dev += item
            host += item
    return (f"ZKP_HD void {sig} {{\n#if defined(__CUDA_ARCH__)\n{dev}#else\n{host}#endif\n}}\n\n")


def gen_field(name, mod, n):
    m = limbs(mod, n)
    inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    out = f"// ---------------------------------------------------------------- {name} (N = {n})\n"

    # ---- first Montgomery row: x = a_even*bi, y = a_odd*bi, then reduction step
    def reduction(b, x, y):
        b.operand("mi", "=r", "mi")
        b.emit("mul.lo.u32", "mi", x(0), inv)
        # y += p_odd * mi  (never carries out of y)
        for j in range(0, n, 2):
            b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", y(j), "mi", m[j + 1], y(j))
            b.emit("madc.hi.cc.u32" if j < n - 2 else "madc.hi.u32", y(j + 1), "mi", m[j + 1], y(j + 1))
        # x += p_even * mi ; carry into y[n-1]
        for j in range(0, n, 2):
            b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", x(j), "mi", m[j], x(j))
            b.emit("madc.hi.cc.u32", x(j + 1), "mi", m[j], x(j + 1))
        b.emit("addc.u32", y(n - 1), y(n - 1), 0)

    b = Block()
    # outputs only ("=r"): NVPTX gives every asm output a fresh virtual register, so an output
    # written early can never alias a still-live input
    x = lambda j: b.operand(f"x{j}", "=r", f"x[{j}]")
    y = lambda j: b.operand(f"y{j}", "=r", f"y[{j}]")
    a = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    for j in range(n):
        x(j)
    for j in range(n):
        y(j)
    b.operand("mi", "=r", "mi")
    b.operand("bi", "r", "bi")
    for j in range(0, n, 2):
        b.emit("mul.lo.u32", y(j), a(j + 1), "bi")
        b.emit("mul.hi.u32", y(j + 1), a(j + 1), "bi")
    for j in range(0, n, 2):
        b.emit("mul.lo.u32", x(j), a(j), "bi")
        b.emit("mul.hi.u32", x(j + 1), a(j), "bi")
    reduction(b, x, y)
    pre = f"    uint32_t mi;\n"
    out += func(f"{name}_row_first(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t bi)",
                [pre, b])

    # ---- generic row: x aligned on limb 0 (was y), y = previous x to be shifted by 64 bits
    b = Block()
    x = lambda j: b.operand(f"x{j}", "+r", f"x[{j}]")
    y = lambda j: b.operand(f"y{j}", "+r", f"y[{j}]")
    a = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    for j in range(n):
        x(j)
    for j in range(n):
        y(j)
    b.operand("mi", "=r", "mi")
    b.operand("bi", "r", "bi")
    b.emit("add.cc.u32", x(0), x(0), y(1))
    for j in range(0, n - 2, 2):
        b.emit("madc.lo.cc.u32", y(j), a(j + 1), "bi", y(j + 2))
        b.emit("madc.hi.cc.u32", y(j + 1), a(j + 1), "bi", y(j + 3))
    b.emit("madc.lo.cc.u32", y(n - 2), a(n - 1), "bi", 0)
    b.emit("madc.hi.u32", y(n - 1), a(n - 1), "bi", 0)
    for j in range(0, n, 2):
        b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", x(j), a(j), "bi", x(j))
        b.emit("madc.hi.cc.u32", x(j + 1), a(j), "bi", x(j + 1))
    b.emit("addc.u32", y(n - 1), y(n - 1), 0)
    reduction(b, x, y)
    out += func(f"{name}_row(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t bi)", [pre, b])

    # ---- merge: r = (x >> 32) + y   (x[0] == 0 after the last row)
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    x = lambda j: b.operand(f"x{j}", "r", f"x[{j}]")
    y = lambda j: b.operand(f"y{j}", "r", f"y[{j}]")
    for j in range(n):
        r(j)
    for j in range(n - 1):
        b.emit("add.cc.u32" if j == 0 else "addc.cc.u32", r(j), y(j), x(j + 1))
    b.emit("addc.u32", r(n - 1), y(n - 1), 0)
    out += func(f"{name}_merge(uint32_t* r, const uint32_t* x, const uint32_t* y)", [b])

    # ---- t = r - p, borrow (0 / 0xffffffff) ; carry = bit 32n of r
    b = Block()
    t = lambda j: b.operand(f"t{j}", "=r", f"t[{j}]")
    rr = lambda j: b.operand(f"r{j}", "r", f"r[{j}]")
    for j in range(n):
        t(j)
    b.operand("bw", "=r", "bw")
    b.operand("cy", "r", "carry")
    for j in range(n):
        b.emit("sub.cc.u32" if j == 0 else "subc.cc.u32", t(j), rr(j), m[j])
    b.emit("subc.u32", "bw", "cy", 0)
    out += func(f"{name}_reduce_once(uint32_t* r, uint32_t carry)",
                [f"    uint32_t t[{n}], bw;\n", b,
                 f"#pragma unroll\n    for (int i = 0; i < {n}; i++) r[i] = bw ? r[i] : t[i];\n"])

    # ---- r = a + b (raw), returns carry through *cy
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    aa = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    bb = lambda j: b.operand(f"b{j}", "r", f"b[{j}]")
    for j in range(n):
        r(j)
    b.operand("cy", "=r", "cy")
    for j in range(n):
        b.emit("add.cc.u32" if j == 0 else "addc.cc.u32", r(j), aa(j), bb(j))
    b.emit("addc.u32", "cy", 0, 0)
    out += func(f"{name}_add(uint32_t* r, const uint32_t* a, const uint32_t* b)",
                ["    uint32_t cy;\n", b, f"    {name}_reduce_once(r, cy);\n"])

    # ---- r = a - b mod p
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    aa = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    bb = lambda j: b.operand(f"b{j}", "r", f"b[{j}]")
    for j in range(n):
        r(j)
    b.operand("mk", "=r", "mk")
    for j in range(n):
        b.emit("sub.cc.u32" if j == 0 else "subc.cc.u32", r(j), aa(j), bb(j))
    b.emit("subc.u32", "mk", 0, 0)
    b2 = Block()
    r2 = lambda j: b2.operand(f"r{j}", "+r", f"r[{j}]")
    for j in range(n):
        r2(j)
    for j in range(n):
        b2.operand(f"p{j}", "r", f"(0x{m[j]:08x}u & mk)")
    for j in range(n):
        op = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < n - 1 else "addc.u32")
        b2.emit(op, r2(j), r2(j), f"p{j}")
    out += func(f"{name}_sub(uint32_t* r, const uint32_t* a, const uint32_t* b)",
                ["    uint32_t mk;\n", b, b2])
    return out



def gen_extra(name, mod, n):
    """Dedicated squaring rows and the fused two-product row (see ff.cuh: Fp::sqr, Fp::mul2)."""
    m = limbs(mod, n)
    inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    out = ""

    def reduction(b, x, y):
        b.operand("mi", "=r", "mi")
        b.emit("mul.lo.u32", "mi", x(0), inv)
        for j in range(0, n, 2):
            b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", y(j), "mi", m[j + 1], y(j))
            b.emit("madc.hi.cc.u32" if j < n - 2 else "madc.hi.u32", y(j + 1), "mi", m[j + 1], y(j + 1))
        for j in range(0, n, 2):
            b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", x(j), "mi", m[j], x(j))
            b.emit("madc.hi.cc.u32", x(j + 1), "mi", m[j], x(j + 1))
        b.emit("addc.u32", y(n - 1), y(n - 1), 0)

    pre = "    uint32_t mi;\n"
    # ---- squaring row i: multiplies a_i with the limbs of  a_i 2^(32 i) + 2 (a with limbs 0..i cleared);
    #      positions below i are plain shifts of the accumulator (no multiply issued).  d = limbs of 2a.
    for i in range(n):
        b = Block()
        x = lambda j: b.operand(f"x{j}", "+r", f"x[{j}]")
        y = lambda j: b.operand(f"y{j}", "+r", f"y[{j}]")
        for j in range(n):
            x(j)
        for j in range(n):
            y(j)
        b.operand("mi", "=r", "mi")
        b.operand("bi", "r", f"a[{i}]")

        def mult(j):
            if j < i:
                return None
            if j == i:
                return "bi"
            if j == i + 1:  # lowest limb of 2 * (a with limbs 0..i cleared): no bit carried in from a_i
                return b.operand(f"e{j}", "r", f"(a[{j}] << 1)")
            return b.operand(f"d{j}", "r", f"d[{j}]")

        b.emit("add.cc.u32", x(0), x(0), y(1))
        for j in range(0, n - 2, 2):
            op = mult(j + 1)
            if op:
                b.emit("madc.lo.cc.u32", y(j), op, "bi", y(j + 2))
                b.emit("madc.hi.cc.u32", y(j + 1), op, "bi", y(j + 3))
            else:
                b.emit("addc.cc.u32", y(j), y(j + 2), 0)
                b.emit("addc.cc.u32", y(j + 1), y(j + 3), 0)
        op = mult(n - 1)  # always present: n - 1 >= i
        b.emit("madc.lo.cc.u32", y(n - 2), op, "bi", 0)
        b.emit("madc.hi.u32", y(n - 1), op, "bi", 0)
        first = True
        for j in range(0, n, 2):
            op = mult(j)
            if not op:
                continue
            b.emit("mad.lo.cc.u32" if first else "madc.lo.cc.u32", x(j), op, "bi", x(j))
            b.emit("madc.hi.cc.u32", x(j + 1), op, "bi", x(j + 1))
            first = False
        if not first:
            b.emit("addc.u32", y(n - 1), y(n - 1), 0)
        reduction(b, x, y)
        out += func(f"{name}_sqr_row_{i}(uint32_t* x, uint32_t* y, const uint32_t* a, const uint32_t* d)", [pre, b])

    # ---- fused row: x, y += a * bi + c * di, then one reduction step (a*b + c*d with ONE Montgomery reduction)
    b = Block()
    x = lambda j: b.operand(f"x{j}", "+r", f"x[{j}]")
    y = lambda j: b.operand(f"y{j}", "+r", f"y[{j}]")
    a = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    c = lambda j: b.operand(f"c{j}", "r", f"c[{j}]")
    for j in range(n):
        x(j)
    for j in range(n):
        y(j)
    b.operand("mi", "=r", "mi")
    b.operand("bi", "r", "bi")
    b.operand("di", "r", "di")
    b.emit("add.cc.u32", x(0), x(0), y(1))
    for j in range(0, n - 2, 2):
        b.emit("madc.lo.cc.u32", y(j), a(j + 1), "bi", y(j + 2))
        b.emit("madc.hi.cc.u32", y(j + 1), a(j + 1), "bi", y(j + 3))
    b.emit("madc.lo.cc.u32", y(n - 2), a(n - 1), "bi", 0)
    b.emit("madc.hi.u32", y(n - 1), a(n - 1), "bi", 0)
    for j in range(0, n, 2):
        b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", x(j), a(j), "bi", x(j))
        b.emit("madc.hi.cc.u32", x(j + 1), a(j), "bi", x(j + 1))
    b.emit("addc.u32", y(n - 1), y(n - 1), 0)
    # second product, no shift
    for j in range(0, n, 2):
        b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", y(j), c(j + 1), "di", y(j))
        b.emit("madc.hi.cc.u32" if j < n - 2 else "madc.hi.u32", y(j + 1), c(j + 1), "di", y(j + 1))
    for j in range(0, n, 2):
        b.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", x(j), c(j), "di", x(j))
        b.emit("madc.hi.cc.u32", x(j + 1), c(j), "di", x(j + 1))
    b.emit("addc.u32", y(n - 1), y(n - 1), 0)
    reduction(b, x, y)
    out += func(f"{name}_row2(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di)", [pre, b])
    return out


def gen_lazy(name, mod, n):
    """Unreduced helpers for the lazily reduced mixed addition (g1.cuh madd_lazy; ranges in tools/lazy_bounds.py):
    all results are plain integers modulo 2^(32 n), the callers keep them below that."""
    out = ""
    m2 = limbs(2 * mod, n)

    def chain(b, op0, opc, opl, dst, s1, s2):
        for j in range(n):
            op = op0 if j == 0 else (opc if j < n - 1 else opl)
            b.emit(op, dst(j), s1(j), s2(j))

    # ---- r = a + b
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    aa = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    bb = lambda j: b.operand(f"b{j}", "r", f"b[{j}]")
    for j in range(n):
        r(j)
    chain(b, "add.cc.u32", "addc.cc.u32", "addc.u32", r, aa, bb)
    out += func(f"{name}_add_raw(uint32_t* r, const uint32_t* a, const uint32_t* b)", [b])

    # ---- r = a - b, *bw = 0xffffffff when a < b (r is then a - b + 2^(32 n))
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    aa = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    bb = lambda j: b.operand(f"b{j}", "r", f"b[{j}]")
    for j in range(n):
        r(j)
    b.operand("mk", "=r", "mk")
    chain(b, "sub.cc.u32", "subc.cc.u32", "subc.cc.u32", r, aa, bb)
    b.emit("subc.u32", "mk", 0, 0)
    out += func(f"{name}_sub_borrow(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t* bw)",
                ["    uint32_t mk;\n", b, "    *bw = mk;\n"])

    # ---- r = a - b + 2p   (a - b + 2p must lie in [0, 2^(32 n)))
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    aa = lambda j: b.operand(f"a{j}", "r", f"a[{j}]")
    bb = lambda j: b.operand(f"b{j}", "r", f"b[{j}]")
    for j in range(n):
        r(j)
    chain(b, "sub.cc.u32", "subc.cc.u32", "subc.u32", r, aa, bb)
    b2 = Block()
    r2 = lambda j: b2.operand(f"r{j}", "+r", f"r[{j}]")
    for j in range(n):
        r2(j)
    chain(b2, "add.cc.u32", "addc.cc.u32", "addc.u32", r2, r2, lambda j: m2[j])
    out += func(f"{name}_sub_p2(uint32_t* r, const uint32_t* a, const uint32_t* b)", [b, b2])

    # ---- r = 2p - b   (b <= 2p)
    b = Block()
    r = lambda j: b.operand(f"r{j}", "=r", f"r[{j}]")
    bb = lambda j: b.operand(f"b{j}", "r", f"b[{j}]")
    for j in range(n):
        r(j)
    chain(b, "sub.cc.u32", "subc.cc.u32", "subc.u32", r, lambda j: m2[j], bb)
    out += func(f"{name}_rsub_p2(uint32_t* r, const uint32_t* b)", [b])
    return out


hdr = """// GENERATED by tools/gen_chains.py -- do not edit.
// One inline-asm statement per carry chain (see the generator's docstring).
#pragma once
#include <stdint.h>
#include "ptx_chain.cuh"
namespace zkp {
namespace chains {

"""
body = gen_field("fq", P, 12) + gen_extra("fq", P, 12) + gen_lazy("fq", P, 12) + gen_field("fr", R, 8)
path = os.path.join(os.path.dirname(__file__), "..", "zkp_subnet_b200", "csrc", "mont_chains.cuh")
open(path, "w").write(hdr + body + "}  // namespace chains\n}  // namespace zkp\n")
print("wrote", os.path.abspath(path))
