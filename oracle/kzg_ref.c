/*
 * ORACLE (test infrastructure, NOT product code) -- CPU restatement of the KZG hot path in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libzkp_b200.so) never links or calls it.
 *
 * What it restates: the work the reference delegates to the external Rust prover `fourier`
 * (reference neurons/miner.py:38-61 -> Client.worker_commit / worker_open; neurons/validator.py:58-104
 * -> Client.fft / eval).  `fourier` is un-vendored and un-pinned (reference requirements.txt:3,
 * Makefile:24-28), so this file follows the published math: BLS12-381,
 * KZG10, the ZCash compressed-G1 encoding and the natural-order radix-2 domain w_n = 7^((r-1)/n).
 * PARITY UNPINNED at the commitment/proof byte level; pinned against oracle/bls12_381.py (plain
 * big-int arithmetic), which is itself pinned on the reference's TEST_POLY/TEST_POINT/TEST_EVAL
 * known answer and the SURVEY section 8c vectors (tests/test_oracle.py).
 *
 * Deliberately different from the product: 64-bit limbs with unsigned __int128, Jacobian
 * coordinates (madd-2007-bl mixed addition), unsigned-window Pippenger run as (window, point chunk) jobs on a
 * pthread queue, iNTT + synthetic division for the quotient.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -pthread)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------ */
/* generic Montgomery arithmetic on NL 64-bit limbs                                            */
/* ------------------------------------------------------------------------------------------ */
#define QL 6
#define RL 4

static const u64 Q_MOD[QL] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                              0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const u64 Q_INV = 0x89f3fffcfffcfffdull;
static const u64 Q_R2[QL] = {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull,
                             0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull};
static const u64 R_MOD[RL] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                              0x73eda753299d7d48ull};
static const u64 R_INV = 0xfffffffeffffffffull;
static const u64 R_R2[RL] = {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full,
                             0x0748d9d99f59ff11ull};

static inline __attribute__((always_inline)) int ge_n(const u64* a, const u64* b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline __attribute__((always_inline)) u64 sub_n(u64* r, const u64* a, const u64* b, int n) {
    u64 borrow = 0;
    for (int i = 0; i < n; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static inline __attribute__((always_inline)) u64 add_n(u64* r, const u64* a, const u64* b, int n) {
    u64 carry = 0;
    for (int i = 0; i < n; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    return carry;
}
static inline __attribute__((always_inline)) void mont_mul(u64* r, const u64* a, const u64* b, const u64* mod, u64 inv, int n) {
    u64 t[8] = {0};
    for (int i = 0; i < n; i++) {
        u128 c = 0;
        for (int j = 0; j < n; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[n];
        t[n] = (u64)c;
        t[n + 1] = (u64)(c >> 64);
        u64 m = t[0] * inv;
        c = ((u128)m * mod[0] + t[0]) >> 64;
        for (int j = 1; j < n; j++) {
            c += (u128)m * mod[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[n];
        t[n - 1] = (u64)c;
        t[n] = t[n + 1] + (u64)(c >> 64);
    }
    if (t[n] || ge_n(t, mod, n)) sub_n(r, t, mod, n);
    else memcpy(r, t, n * sizeof(u64));
}
static inline __attribute__((always_inline)) void mod_add(u64* r, const u64* a, const u64* b, const u64* mod, int n) {
    u64 t[6];
    u64 c = add_n(t, a, b, n);
    if (c || ge_n(t, mod, n)) sub_n(r, t, mod, n);
    else memcpy(r, t, n * sizeof(u64));
}
static inline __attribute__((always_inline)) void mod_sub(u64* r, const u64* a, const u64* b, const u64* mod, int n) {
    u64 t[6];
    if (sub_n(t, a, b, n)) add_n(r, t, mod, n);
    else memcpy(r, t, n * sizeof(u64));
}
static int is_zero_n(const u64* a, int n) {
    u64 acc = 0;
    for (int i = 0; i < n; i++) acc |= a[i];
    return acc == 0;
}

/* ---- Fq ---- */
typedef struct { u64 v[QL]; } fq;
/* fixed-size CIOS for the 6-limb base field (p < 2^383: one spare limb suffices); ~1.5x the generic loop */
static inline __attribute__((always_inline)) void mont_mul6(u64* r, const u64* a, const u64* b) {
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
#define FQ_ROW(bi) do { u128 c; u64 m; \
    c = (u128)a[0] * (bi) + t0; t0 = (u64)c; c >>= 64; \
    c += (u128)a[1] * (bi) + t1; t1 = (u64)c; c >>= 64; \
    c += (u128)a[2] * (bi) + t2; t2 = (u64)c; c >>= 64; \
    c += (u128)a[3] * (bi) + t3; t3 = (u64)c; c >>= 64; \
    c += (u128)a[4] * (bi) + t4; t4 = (u64)c; c >>= 64; \
    c += (u128)a[5] * (bi) + t5; t5 = (u64)c; c >>= 64; \
    t6 += (u64)c; \
    m = t0 * Q_INV; \
    c = ((u128)m * Q_MOD[0] + t0) >> 64; \
    c += (u128)m * Q_MOD[1] + t1; t0 = (u64)c; c >>= 64; \
    c += (u128)m * Q_MOD[2] + t2; t1 = (u64)c; c >>= 64; \
    c += (u128)m * Q_MOD[3] + t3; t2 = (u64)c; c >>= 64; \
    c += (u128)m * Q_MOD[4] + t4; t3 = (u64)c; c >>= 64; \
    c += (u128)m * Q_MOD[5] + t5; t4 = (u64)c; c >>= 64; \
    c += t6; t5 = (u64)c; t6 = (u64)(c >> 64); } while (0)
    FQ_ROW(b[0]); FQ_ROW(b[1]); FQ_ROW(b[2]); FQ_ROW(b[3]); FQ_ROW(b[4]); FQ_ROW(b[5]);
#undef FQ_ROW
    u64 t[6] = {t0, t1, t2, t3, t4, t5};
    if (t6 || ge_n(t, Q_MOD, 6)) sub_n(r, t, Q_MOD, 6);
    else memcpy(r, t, sizeof(t));
}
static void fq_mul(fq* r, const fq* a, const fq* b) { mont_mul6(r->v, a->v, b->v); }
static void fq_sqr(fq* r, const fq* a) { fq_mul(r, a, a); }
static void fq_add(fq* r, const fq* a, const fq* b) { mod_add(r->v, a->v, b->v, Q_MOD, QL); }
static void fq_sub(fq* r, const fq* a, const fq* b) { mod_sub(r->v, a->v, b->v, Q_MOD, QL); }
static int fq_is_zero(const fq* a) { return is_zero_n(a->v, QL); }
static int fq_eq(const fq* a, const fq* b) { return memcmp(a->v, b->v, sizeof(a->v)) == 0; }
static void fq_from_u64(fq* r, u64 x) {
    fq t = {{x, 0, 0, 0, 0, 0}}, r2;
    memcpy(r2.v, Q_R2, sizeof(Q_R2));
    fq_mul(r, &t, &r2);
}
static void fq_to_mont(fq* r, const fq* a) { fq r2; memcpy(r2.v, Q_R2, sizeof(Q_R2)); fq_mul(r, a, &r2); }
static void fq_from_mont(fq* r, const fq* a) { fq one = {{1, 0, 0, 0, 0, 0}}; fq_mul(r, a, &one); }
static void fq_pow(fq* r, const fq* a, const u64* e, int elimbs) {
    fq acc, base = *a;
    fq_from_u64(&acc, 1);
    for (int i = 0; i < elimbs * 64; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fq_mul(&acc, &acc, &base);
        fq_sqr(&base, &base);
    }
    *r = acc;
}
static void fq_inv(fq* r, const fq* a) {
    u64 e[QL];
    u64 two[QL] = {2, 0, 0, 0, 0, 0};
    sub_n(e, Q_MOD, two, QL);
    fq_pow(r, a, e, QL);
}

/* ---- Fr ---- */
typedef struct { u64 v[RL]; } fr;
static void fr_mul(fr* r, const fr* a, const fr* b) { mont_mul(r->v, a->v, b->v, R_MOD, R_INV, RL); }
static void fr_add(fr* r, const fr* a, const fr* b) { mod_add(r->v, a->v, b->v, R_MOD, RL); }
static void fr_sub(fr* r, const fr* a, const fr* b) { mod_sub(r->v, a->v, b->v, R_MOD, RL); }
static void fr_to_mont(fr* r, const fr* a) { fr r2; memcpy(r2.v, R_R2, sizeof(R_R2)); fr_mul(r, a, &r2); }
static void fr_from_mont(fr* r, const fr* a) { fr one = {{1, 0, 0, 0}}; fr_mul(r, a, &one); }
static void fr_from_u64(fr* r, u64 x) { fr t = {{x, 0, 0, 0}}; fr_to_mont(r, &t); }
static void fr_pow(fr* r, const fr* a, const u64* e, int elimbs) {
    fr acc, base = *a;
    fr_from_u64(&acc, 1);
    for (int i = 0; i < elimbs * 64; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fr_mul(&acc, &acc, &base);
        fr_mul(&base, &base, &base);
    }
    *r = acc;
}
static void fr_inv(fr* r, const fr* a) {
    u64 e[RL], two[RL] = {2, 0, 0, 0};
    sub_n(e, R_MOD, two, RL);
    fr_pow(r, a, e, RL);
}
/* big-endian 32 bytes <-> canonical limbs */
static int fr_from_be(fr* r, const uint8_t* be) {
    for (int i = 0; i < RL; i++) {
        u64 w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | be[(RL - 1 - i) * 8 + k];
        r->v[i] = w;
    }
    return !ge_n(r->v, R_MOD, RL);
}
static void fr_to_be(uint8_t* be, const fr* a) {
    for (int i = 0; i < RL; i++)
        for (int k = 0; k < 8; k++) be[(RL - 1 - i) * 8 + k] = (uint8_t)(a->v[i] >> (56 - 8 * k));
}
static void fq_from_be(fq* r, const uint8_t* be) {
    for (int i = 0; i < QL; i++) {
        u64 w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | be[(QL - 1 - i) * 8 + k];
        r->v[i] = w;
    }
}
static void fq_to_be(uint8_t* be, const fq* a) {
    for (int i = 0; i < QL; i++)
        for (int k = 0; k < 8; k++) be[(QL - 1 - i) * 8 + k] = (uint8_t)(a->v[i] >> (56 - 8 * k));
}
/* w_n = 7^((r-1)/n), Montgomery form */
static void fr_root_of_unity(fr* w, unsigned log_n) {
    u64 e[RL], one[RL] = {1, 0, 0, 0};
    sub_n(e, R_MOD, one, RL);
    /* e >>= log_n */
    for (unsigned s = 0; s < log_n; s++) {
        for (int i = 0; i < RL; i++) e[i] = (e[i] >> 1) | (i + 1 < RL ? e[i + 1] << 63 : 0);
    }
    fr g;
    fr_from_u64(&g, 7);
    fr_pow(w, &g, e, RL);
}

/* ------------------------------------------------------------------------------------------ */
/* G1: Jacobian (X, Y, Z), y^2 = x^3 + 4                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { fq x, y; int inf; } g1a;   /* affine, Montgomery coordinates */
typedef struct { fq x, y, z; } g1j;         /* z == 0 -> infinity */

static void g1j_set_inf(g1j* p) { memset(p, 0, sizeof(*p)); }
static int g1j_is_inf(const g1j* p) { return fq_is_zero(&p->z); }

static void g1j_double(g1j* r, const g1j* p) {
    if (g1j_is_inf(p)) { g1j_set_inf(r); return; }
    /* dbl-2009-l, a = 0 */
    fq a, b, c, d, e, f, t;
    fq_sqr(&a, &p->x);
    fq_sqr(&b, &p->y);
    fq_sqr(&c, &b);
    fq_add(&t, &p->x, &b); fq_sqr(&t, &t); fq_sub(&t, &t, &a); fq_sub(&t, &t, &c);
    fq_add(&d, &t, &t);
    fq_add(&e, &a, &a); fq_add(&e, &e, &a);
    fq_sqr(&f, &e);
    fq z3; fq_mul(&z3, &p->y, &p->z); fq_add(&z3, &z3, &z3);
    fq x3; fq_sub(&x3, &f, &d); fq_sub(&x3, &x3, &d);
    fq c8; fq_add(&c8, &c, &c); fq_add(&c8, &c8, &c8); fq_add(&c8, &c8, &c8);
    fq y3; fq_sub(&y3, &d, &x3); fq_mul(&y3, &e, &y3); fq_sub(&y3, &y3, &c8);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_add(g1j* r, const g1j* p, const g1j* q) {
    if (g1j_is_inf(p)) { *r = *q; return; }
    if (g1j_is_inf(q)) { *r = *p; return; }
    fq z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
    fq_sqr(&z1z1, &p->z); fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2); fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &q->z); fq_mul(&s1, &s1, &z2z2);
    fq_mul(&s2, &q->y, &p->z); fq_mul(&s2, &s2, &z1z1);
    fq_sub(&h, &u2, &u1); fq_sub(&rr, &s2, &s1);
    if (fq_is_zero(&h)) {
        if (fq_is_zero(&rr)) { g1j_double(r, p); return; }
        g1j_set_inf(r); return;
    }
    fq hh, hhh, v;
    fq_sqr(&hh, &h); fq_mul(&hhh, &hh, &h); fq_mul(&v, &u1, &hh);
    fq x3; fq_sqr(&x3, &rr); fq_sub(&x3, &x3, &hhh); fq_sub(&x3, &x3, &v); fq_sub(&x3, &x3, &v);
    fq y3; fq_sub(&y3, &v, &x3); fq_mul(&y3, &rr, &y3); fq_mul(&t, &s1, &hhh); fq_sub(&y3, &y3, &t);
    fq z3; fq_mul(&z3, &p->z, &q->z); fq_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_from_affine(g1j* r, const g1a* p) {
    if (p->inf) { g1j_set_inf(r); return; }
    r->x = p->x; r->y = p->y; fq_from_u64(&r->z, 1);
}
/* Jacobian + affine, madd-2007-bl (7M + 4S); every exceptional case falls back to the general formulas */
static void g1j_add_affine(g1j* r, const g1j* p, const g1a* q) {
    if (q->inf) { *r = *p; return; }
    if (g1j_is_inf(p)) { g1j_from_affine(r, q); return; }
    fq z1z1, u2, s2, h, hh, i, j, rr, v, t;
    fq_sqr(&z1z1, &p->z);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s2, &q->y, &p->z); fq_mul(&s2, &s2, &z1z1);
    fq_sub(&h, &u2, &p->x);
    fq_sub(&rr, &s2, &p->y);
    if (fq_is_zero(&h)) {
        g1j tq; g1j_from_affine(&tq, q);
        if (fq_is_zero(&rr)) g1j_double(r, &tq); else g1j_set_inf(r);
        return;
    }
    fq_add(&rr, &rr, &rr);
    fq_sqr(&hh, &h);
    fq_add(&i, &hh, &hh); fq_add(&i, &i, &i);
    fq_mul(&j, &h, &i);
    fq_mul(&v, &p->x, &i);
    fq x3; fq_sqr(&x3, &rr); fq_sub(&x3, &x3, &j); fq_sub(&x3, &x3, &v); fq_sub(&x3, &x3, &v);
    fq y3; fq_sub(&y3, &v, &x3); fq_mul(&y3, &rr, &y3); fq_mul(&t, &p->y, &j); fq_add(&t, &t, &t); fq_sub(&y3, &y3, &t);
    fq z3; fq_add(&z3, &p->z, &h); fq_sqr(&z3, &z3); fq_sub(&z3, &z3, &z1z1); fq_sub(&z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_to_affine(g1a* r, const g1j* p) {
    if (g1j_is_inf(p)) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
    fq zi, zi2, zi3;
    fq_inv(&zi, &p->z); fq_sqr(&zi2, &zi); fq_mul(&zi3, &zi2, &zi);
    fq_mul(&r->x, &p->x, &zi2); fq_mul(&r->y, &p->y, &zi3); r->inf = 0;
}
static void g1j_mul(g1j* r, const g1j* p, const fr* k_canonical) {
    g1j acc; g1j_set_inf(&acc);
    for (int i = 255; i >= 0; i--) {
        g1j_double(&acc, &acc);
        if ((k_canonical->v[i >> 6] >> (i & 63)) & 1) g1j_add(&acc, &acc, p);
    }
    *r = acc;
}
static const u64 GX_CANON[QL] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull,
                                 0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
static const u64 GY_CANON[QL] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull,
                                 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
static void g1_generator(g1a* g) {
    fq x, y;
    memcpy(x.v, GX_CANON, sizeof(GX_CANON)); memcpy(y.v, GY_CANON, sizeof(GY_CANON));
    fq_to_mont(&g->x, &x); fq_to_mont(&g->y, &y); g->inf = 0;
}
/* ZCash compressed encoding */
static void g1_compress(uint8_t out[48], const g1a* p) {
    if (p->inf) { memset(out, 0, 48); out[0] = 0xc0; return; }
    fq x, y, ny;
    fq_from_mont(&x, &p->x); fq_from_mont(&y, &p->y);
    fq_to_be(out, &x);
    out[0] |= 0x80;
    /* y > (p-1)/2  <=>  y > p - y */
    fq zero; memset(&zero, 0, sizeof(zero));
    sub_n(ny.v, Q_MOD, y.v, QL);
    if (ge_n(y.v, ny.v, QL) && !fq_eq(&y, &ny)) out[0] |= 0x20;
}
/* uncompressed ZCash 96-byte form */
static void g1_serialize96(uint8_t out[96], const g1a* p) {
    if (p->inf) { memset(out, 0, 96); out[0] = 0x40; return; }
    fq x, y;
    fq_from_mont(&x, &p->x); fq_from_mont(&y, &p->y);
    fq_to_be(out, &x); fq_to_be(out + 48, &y);
}
static void g1_deserialize96(g1a* p, const uint8_t in[96]) {
    if (in[0] & 0x40) { memset(p, 0, sizeof(*p)); p->inf = 1; return; }
    fq x, y;
    uint8_t tmp[48];
    memcpy(tmp, in, 48); tmp[0] &= 0x1f;
    fq_from_be(&x, tmp); fq_from_be(&y, in + 48);
    fq_to_mont(&p->x, &x); fq_to_mont(&p->y, &y); p->inf = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Pippenger MSM, unsigned windows.  Parallel over (window, point chunk) jobs taken from a shared queue: */
/* every job fills its own bucket array from its chunk, reduces it with the running-sum trick and      */
/* leaves one window partial; the caller folds the windows with c doublings each (Horner).             */
/* ------------------------------------------------------------------------------------------ */
static unsigned window_bits(size_t n) {
    unsigned lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    if (lg < 4) return 2;
    unsigned c = lg > 3 ? lg - 3 : 1;
    return c > 16 ? 16 : c;
}
static unsigned get_window(const fr* k, unsigned lo, unsigned c) {
    unsigned limb = lo >> 6, off = lo & 63;
    u64 w = k->v[limb] >> off;
    if (off + c > 64 && limb + 1 < RL) w |= k->v[limb + 1] << (64 - off);
    return (unsigned)(w & (((u64)1 << c) - 1));
}
/* sum over points [lo, hi) of digit_w(scalar) * point */
static void msm_window_chunk(g1j* out, const g1a* pts, const fr* scal, size_t lo, size_t hi, unsigned w, unsigned c, g1j* buckets) {
    size_t nb = ((size_t)1 << c) - 1;
    for (size_t b = 0; b < nb; b++) g1j_set_inf(&buckets[b]);
    for (size_t i = lo; i < hi; i++) {
        unsigned d = get_window(&scal[i], w * c, c);
        if (d && !pts[i].inf) g1j_add_affine(&buckets[d - 1], &buckets[d - 1], &pts[i]);
    }
    g1j run, sum;
    g1j_set_inf(&run); g1j_set_inf(&sum);
    for (size_t b = nb; b-- > 0;) {
        g1j_add(&run, &run, &buckets[b]);
        g1j_add(&sum, &sum, &run);
    }
    *out = sum;
}
typedef struct {
    const g1a* pts; const fr* scal; size_t n; unsigned c, nwin, chunks;
    g1j* partial;            /* nwin * chunks window partials */
    volatile long next;      /* job queue head */
} msm_shared;
static void* msm_worker(void* arg) {
    msm_shared* sh = (msm_shared*)arg;
    g1j* buckets = (g1j*)malloc(sizeof(g1j) * (((size_t)1 << sh->c) - 1 + 1));
    for (;;) {
        long job = __sync_fetch_and_add(&sh->next, 1);
        if (job >= (long)(sh->nwin * sh->chunks)) break;
        unsigned w = (unsigned)(job / sh->chunks), k = (unsigned)(job % sh->chunks);
        size_t lo = sh->n * k / sh->chunks, hi = sh->n * (k + 1) / sh->chunks;
        msm_window_chunk(&sh->partial[job], sh->pts, sh->scal, lo, hi, w, sh->c, buckets);
    }
    free(buckets);
    return NULL;
}
static void msm_parallel(g1j* out, const g1a* pts, const fr* scal, size_t n, int threads) {
    g1j_set_inf(out);
    if (n == 0) return;
    if (threads < 1) threads = 1;
    msm_shared sh;
    sh.pts = pts; sh.scal = scal; sh.n = n; sh.next = 0;
    /* enough jobs to keep every thread busy: split the points when there are fewer windows than ~3x threads */
    sh.chunks = 1;
    size_t per = n;
    for (;;) {
        sh.c = window_bits(per);
        sh.nwin = (255 + sh.c - 1) / sh.c;
        if (sh.nwin * sh.chunks >= 3u * (unsigned)threads || per < 1024 || threads == 1) break;
        sh.chunks *= 2;
        per = n / sh.chunks;
    }
    sh.partial = (g1j*)calloc((size_t)sh.nwin * sh.chunks, sizeof(g1j));
    pthread_t* th = (pthread_t*)calloc(threads, sizeof(pthread_t));
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, msm_worker, &sh);
    msm_worker(&sh);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    for (int w = (int)sh.nwin - 1; w >= 0; w--) {
        for (unsigned d = 0; d < sh.c; d++) g1j_double(out, out);
        for (unsigned k = 0; k < sh.chunks; k++) g1j_add(out, out, &sh.partial[(size_t)w * sh.chunks + k]);
    }
    free(sh.partial); free(th);
}

/* ------------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------ */
/* NTT (natural order in/out), Horner, quotient                                                 */
/* ------------------------------------------------------------------------------------------ */
static unsigned ilog2(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) l++; return l; }
static void ntt_inplace(fr* a, size_t n, int inverse) {
    unsigned lg = ilog2(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (unsigned b = 0; b < lg; b++) if (i & ((size_t)1 << b)) j |= (size_t)1 << (lg - 1 - b);
        if (i < j) { fr t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    fr w;
    fr_root_of_unity(&w, lg);
    if (inverse) fr_inv(&w, &w);
    /* table of w^k, k < n/2 */
    size_t half = n / 2 ? n / 2 : 1;
    fr* tw = (fr*)malloc(sizeof(fr) * half);
    fr_from_u64(&tw[0], 1);
    for (size_t k = 1; k < half; k++) fr_mul(&tw[k], &tw[k - 1], &w);
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t step = n / len;
        for (size_t s = 0; s < n; s += len) {
            for (size_t k = 0; k < len / 2; k++) {
                fr u = a[s + k], v;
                fr_mul(&v, &a[s + k + len / 2], &tw[k * step]);
                fr_add(&a[s + k], &u, &v);
                fr_sub(&a[s + k + len / 2], &u, &v);
            }
        }
    }
    if (inverse) {
        fr ninv;
        fr_from_u64(&ninv, (u64)n);
        fr_inv(&ninv, &ninv);
        for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &ninv);
    }
    free(tw);
}

/* ------------------------------------------------------------------------------------------ */
/* exported C interface (ctypes); all field elements cross as 32-byte big-endian canonical      */
/* values, all points as ZCash 96-byte uncompressed / 48-byte compressed                        */
/* ------------------------------------------------------------------------------------------ */
static fr* load_scalars_mont(const uint8_t* be, size_t n, int* ok) {
    fr* a = (fr*)malloc(sizeof(fr) * (n ? n : 1));
    *ok = 1;
    for (size_t i = 0; i < n; i++) {
        fr t;
        if (!fr_from_be(&t, be + 32 * i)) *ok = 0;
        fr_to_mont(&a[i], &t);
    }
    return a;
}
static void store_scalars_mont(uint8_t* be, const fr* a, size_t n) {
    for (size_t i = 0; i < n; i++) { fr t; fr_from_mont(&t, &a[i]); fr_to_be(be + 32 * i, &t); }
}

/* out[i] = sum_j in[j] w^(ij) (inverse: w^-1 and 1/n).  Client.fft (reference neurons/validator.py:58-65) */
int ref_ntt(const uint8_t* in_be, uint8_t* out_be, size_t n, int inverse) {
    int ok;
    if (n & (n - 1)) return -1;
    fr* a = load_scalars_mont(in_be, n, &ok);
    if (ok && n > 1) ntt_inplace(a, n, inverse);
    if (ok) store_scalars_mont(out_be, a, n);
    free(a);
    return ok ? 0 : -2;
}
/* coefficient-form Horner.  Client.eval (reference neurons/validator.py:97-104, tests/test_miner.py:33-55) */
int ref_eval(const uint8_t* coeffs_be, size_t n, const uint8_t* x_be, uint8_t* y_be) {
    int ok, okx;
    fr* a = load_scalars_mont(coeffs_be, n, &ok);
    fr* x = load_scalars_mont(x_be, 1, &okx);
    fr acc; memset(&acc, 0, sizeof(acc));
    for (size_t i = n; i-- > 0;) { fr_mul(&acc, &acc, x); fr_add(&acc, &acc, &a[i]); }
    store_scalars_mont(y_be, &acc, 1);
    free(a); free(x);
    return ok && okx ? 0 : -2;
}
/* G1 MSM over ZCash-uncompressed points; scalars canonical.  Client.worker_commit's arithmetic. */
int ref_msm(const uint8_t* points96, const uint8_t* scalars_be, size_t n, uint8_t out48[48], int threads) {
    g1a* pts = (g1a*)malloc(sizeof(g1a) * (n ? n : 1));
    fr* sc = (fr*)malloc(sizeof(fr) * (n ? n : 1));
    int ok = 1;
    for (size_t i = 0; i < n; i++) {
        g1_deserialize96(&pts[i], points96 + 96 * i);
        if (!fr_from_be(&sc[i], scalars_be + 32 * i)) ok = 0;
    }
    g1j acc; g1a res;
    msm_parallel(&acc, pts, sc, n, threads);
    g1j_to_affine(&res, &acc);
    g1_compress(out48, &res);
    free(pts); free(sc);
    return ok ? 0 : -2;
}
/* Evaluation-form opening: y = f(x) and q = (f - y)/(X - x) in evaluation form, computed the long way
 * round (iNTT -> Horner -> synthetic division -> NTT).  Client.worker_open's field arithmetic. */
int ref_quotient_evals(const uint8_t* evals_be, size_t n, const uint8_t* x_be, uint8_t* y_be, uint8_t* q_be) {
    int ok, okx;
    if (n & (n - 1) || n == 0) return -1;
    fr* a = load_scalars_mont(evals_be, n, &ok);
    fr* x = load_scalars_mont(x_be, 1, &okx);
    if (n > 1) ntt_inplace(a, n, 1);
    fr* q = (fr*)calloc(n, sizeof(fr));
    fr acc; memset(&acc, 0, sizeof(acc));
    for (size_t k = n - 1; k >= 1; k--) {
        fr_mul(&acc, &acc, x); fr_add(&acc, &acc, &a[k]);
        q[k - 1] = acc;
    }
    fr_mul(&acc, &acc, x); fr_add(&acc, &acc, &a[0]);
    store_scalars_mont(y_be, &acc, 1);
    if (n > 1) ntt_inplace(q, n, 0);
    store_scalars_mont(q_be, q, n);
    free(a); free(x); free(q);
    return ok && okx ? 0 : -2;
}
/* [k]G compressed */
int ref_g1_mul_gen(const uint8_t* k_be, uint8_t out48[48]) {
    fr k; g1a g, res; g1j gj, acc;
    if (!fr_from_be(&k, k_be)) return -2;
    g1_generator(&g); g1j_from_affine(&gj, &g);
    g1j_mul(&acc, &gj, &k);
    g1j_to_affine(&res, &acc);
    g1_compress(out48, &res);
    return 0;
}
/* SRS rows from a public test trapdoor: kind 0 -> [scale * tau^j]G, kind 1 -> [scale * L_j(tau)]G
 * (natural-order domain of size n).  Output: ZCash 96-byte uncompressed points. */
typedef struct { const fr* k; size_t lo, hi; uint8_t* out; } srs_job;
static void* srs_worker(void* arg) {
    srs_job* j = (srs_job*)arg;
    g1a g; g1j gj;
    g1_generator(&g); g1j_from_affine(&gj, &g);
    /* 4-bit fixed-base table: tab[w][d] = d * 16^w * G */
    static g1j tab[64][16];
    static pthread_once_t once = PTHREAD_ONCE_INIT;
    (void)once;
    g1j (*t)[16] = (g1j(*)[16])malloc(sizeof(g1j) * 64 * 16);
    g1j base = gj;
    for (int w = 0; w < 64; w++) {
        g1j_set_inf(&t[w][0]);
        for (int d = 1; d < 16; d++) g1j_add(&t[w][d], &t[w][d - 1], &base);
        g1j_add(&base, &t[w][15], &base);
    }
    (void)tab;
    /* chunks of CH points share one field inversion (Montgomery's trick on the Z coordinates) */
    enum { CH = 256 };
    g1j* pts = (g1j*)malloc(sizeof(g1j) * CH);
    fq* pre = (fq*)malloc(sizeof(fq) * CH);
    for (size_t lo = j->lo; lo < j->hi; lo += CH) {
        size_t cnt = j->hi - lo < CH ? j->hi - lo : CH;
        for (size_t c = 0; c < cnt; c++) {
            fr k; fr_from_mont(&k, &j->k[lo + c]);
            g1j acc; g1j_set_inf(&acc);
            for (int w = 0; w < 64; w++) {
                unsigned d = (unsigned)((k.v[w >> 4] >> ((w & 15) * 4)) & 15);
                if (d) g1j_add(&acc, &acc, &t[w][d]);
            }
            pts[c] = acc;
        }
        fq run; memset(&run, 0, sizeof(run));
        { fq one = {{1, 0, 0, 0, 0, 0}}; fq_to_mont(&run, &one); }
        for (size_t c = 0; c < cnt; c++) {
            pre[c] = run;
            if (!g1j_is_inf(&pts[c])) fq_mul(&run, &run, &pts[c].z);
        }
        fq inv; fq_inv(&inv, &run);
        for (size_t c = cnt; c-- > 0;) {
            g1a a;
            if (g1j_is_inf(&pts[c])) { memset(&a, 0, sizeof(a)); a.inf = 1; }
            else {
                fq zi, zi2, zi3;
                fq_mul(&zi, &inv, &pre[c]);
                fq_mul(&inv, &inv, &pts[c].z);
                fq_sqr(&zi2, &zi); fq_mul(&zi3, &zi2, &zi);
                fq_mul(&a.x, &pts[c].x, &zi2); fq_mul(&a.y, &pts[c].y, &zi3); a.inf = 0;
            }
            g1_serialize96(j->out + 96 * (lo + c), &a);
        }
    }
    free(pts); free(pre);
    free(t);
    return NULL;
}
int ref_srs(const uint8_t* tau_be, const uint8_t* scale_be, size_t n, int kind, uint8_t* out96, int threads) {
    int ok1, ok2;
    fr* tau = load_scalars_mont(tau_be, 1, &ok1);
    fr* scale = load_scalars_mont(scale_be, 1, &ok2);
    if (!ok1 || !ok2) { free(tau); free(scale); return -2; }
    fr* k = (fr*)malloc(sizeof(fr) * (n ? n : 1));
    if (kind == 0) {
        fr cur = *scale;
        for (size_t j = 0; j < n; j++) { k[j] = cur; fr_mul(&cur, &cur, tau); }
    } else {
        /* L_j(tau) = (tau^n - 1) w^j / (n (tau - w^j)) */
        unsigned lg = ilog2(n);
        fr w, wj, zn, ninv, one;
        fr_root_of_unity(&w, lg);
        fr_from_u64(&one, 1);
        zn = *tau;
        for (unsigned s = 0; s < lg; s++) fr_mul(&zn, &zn, &zn);
        fr_sub(&zn, &zn, &one);
        fr_from_u64(&ninv, (u64)n); fr_inv(&ninv, &ninv);
        fr_mul(&zn, &zn, &ninv);
        fr_mul(&zn, &zn, scale);
        /* one inversion for all the denominators (Montgomery's trick): pre[j] = d_0 .. d_(j-1) */
        fr* d = (fr*)malloc(sizeof(fr) * (n ? n : 1));
        fr* pre = (fr*)malloc(sizeof(fr) * (n ? n : 1));
        fr run = one;
        wj = one;
        for (size_t j = 0; j < n; j++) {
            fr_sub(&d[j], tau, &wj);
            fr_mul(&k[j], &zn, &wj);
            pre[j] = run;
            fr_mul(&run, &run, &d[j]);
            fr_mul(&wj, &wj, &w);
        }
        fr inv; fr_inv(&inv, &run);
        for (size_t j = n; j-- > 0;) {
            fr dj; fr_mul(&dj, &inv, &pre[j]);
            fr_mul(&inv, &inv, &d[j]);
            fr_mul(&k[j], &k[j], &dj);
        }
        free(d); free(pre);
    }
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    srs_job* jobs = (srs_job*)calloc(threads, sizeof(srs_job));
    pthread_t* th = (pthread_t*)calloc(threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].k = k; jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads; jobs[t].out = out96;
        if (t) pthread_create(&th[t], NULL, srs_worker, &jobs[t]);
    }
    srs_worker(&jobs[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th); free(k); free(tau); free(scale);
    return 0;
}
/* sum_j f_j * k_j mod r for canonical big-endian vectors (trapdoor check: commit == [sum f_j L_j(tau)]G) */
int ref_fr_dot(const uint8_t* a_be, const uint8_t* b_be, size_t n, uint8_t* out_be) {
    int ok1, ok2;
    fr* a = load_scalars_mont(a_be, n, &ok1);
    fr* b = load_scalars_mont(b_be, n, &ok2);
    fr acc; memset(&acc, 0, sizeof(acc));
    for (size_t i = 0; i < n; i++) { fr t; fr_mul(&t, &a[i], &b[i]); fr_add(&acc, &acc, &t); }
    store_scalars_mont(out_be, &acc, 1);
    free(a); free(b);
    return ok1 && ok2 ? 0 : -2;
}
/* Lagrange basis values L_j(tau) * scale as canonical big-endian scalars */
int ref_lagrange_scalars(const uint8_t* tau_be, const uint8_t* scale_be, size_t n, uint8_t* out_be) {
    int ok1, ok2;
    fr* tau = load_scalars_mont(tau_be, 1, &ok1);
    fr* scale = load_scalars_mont(scale_be, 1, &ok2);
    unsigned lg = ilog2(n);
    fr w, wj, zn, ninv, one;
    fr_root_of_unity(&w, lg);
    fr_from_u64(&one, 1);
    zn = *tau;
    for (unsigned s = 0; s < lg; s++) fr_mul(&zn, &zn, &zn);
    fr_sub(&zn, &zn, &one);
    fr_from_u64(&ninv, (u64)n); fr_inv(&ninv, &ninv);
    fr_mul(&zn, &zn, &ninv); fr_mul(&zn, &zn, scale);
    /* batch inversion of (tau - w^j) */
    fr* d = (fr*)malloc(sizeof(fr) * n);
    fr* pre = (fr*)malloc(sizeof(fr) * n);
    wj = one;
    fr acc = one;
    for (size_t j = 0; j < n; j++) {
        fr_sub(&d[j], tau, &wj);
        pre[j] = acc;
        fr_mul(&acc, &acc, &d[j]);
        fr_mul(&wj, &wj, &w);
    }
    fr inv; fr_inv(&inv, &acc);
    for (size_t j = n; j-- > 0;) {
        fr dj; fr_mul(&dj, &inv, &pre[j]);
        fr_mul(&inv, &inv, &d[j]);
        d[j] = dj;
    }
    wj = one;
    for (size_t j = 0; j < n; j++) {
        fr t; fr_mul(&t, &zn, &wj); fr_mul(&t, &t, &d[j]);
        fr c; fr_from_mont(&c, &t); fr_to_be(out_be + 32 * j, &c);
        fr_mul(&wj, &wj, &w);
    }
    free(d); free(pre); free(tau); free(scale);
    return ok1 && ok2 ? 0 : -2;
}
/* Elements [first, first + n) of the COUNTER-BASED stream the product's device generator produces (zkp_random_poly /
 * zkp_random_poly_range, kzg.cuh k_random_fr): candidate `attempt` of element i is four SplitMix64 outputs from the
 * state seed + GOLDEN * (4 (64 i + attempt) + 1), little-endian 64-bit words, top bit cleared, first candidate < r
 * wins.  Restated here so that the tests and the benchmark can compute expected commitments of device-generated
 * inputs without copying them back. */
void ref_random_scalars_ctr(uint64_t seed, uint64_t first, size_t n, uint8_t* out_be) {
    for (size_t t = 0; t < n; t++) {
        const u64 i = first + t;
        fr v;
        for (u64 attempt = 0;; attempt++) {
            u64 s = seed + 0x9e3779b97f4a7c15ull * (4 * (i * 64 + attempt) + 1);
            for (int k = 0; k < RL; k++) {
                s += 0x9e3779b97f4a7c15ull;
                u64 z = s;
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
                v.v[k] = z ^ (z >> 31);
            }
            v.v[RL - 1] &= 0x7fffffffffffffffull;
            if (!ge_n(v.v, R_MOD, RL)) break;
        }
        fr_to_be(out_be + 32 * t, &v);
    }
}
/* uniform-ish test scalars: SplitMix64 stream -> 256 bits -> top bit cleared twice -> < r by rejection */
void ref_random_scalars(uint64_t seed, size_t n, uint8_t* out_be) {
    u64 s = seed;
    for (size_t i = 0; i < n; i++) {
        fr v;
        do {
            for (int k = 0; k < RL; k++) {
                s += 0x9e3779b97f4a7c15ull;
                u64 z = s;
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
                v.v[k] = z ^ (z >> 31);
            }
            v.v[RL - 1] &= 0x7fffffffffffffffull;
        } while (ge_n(v.v, R_MOD, RL));
        fr_to_be(out_be + 32 * i, &v);
    }
}
