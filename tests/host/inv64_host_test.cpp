// Host-side exerciser for F64::inverse (csrc/host/field64.hpp: binary extended Euclid on 64-bit limbs) against the
// Fermat form and against x * inv(x) == 1, for Fq and Fr: random values, powers of two, small values, 0, 1, p - 1.
// Prints one "name ok|BAD" line per field for tests/test_host_limbs.py.
#include <cstdio>
#include "../../zkp_subnet_b200/csrc/host/curve.hpp"
using namespace zkp::host;
static uint64_t s = 88172645463325252ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
template <class F> int check(const char* name, int iters) {
    int bad = 0, done = 0;
    for (int it = 0; it < iters; it++) {
        F x;
        for (int i = 0; i < F::N; i++) x.v[i] = rnd();
        x.v[F::N - 1] &= 0x0fffffffffffffffull;
        if (it < 64 * F::N) { x = F::zero(); x.v[it / 64] = 1ull << (it % 64); }
        else if (it == 64 * F::N) x = F::zero();
        else if (it == 64 * F::N + 1) x = F::one();
        else if (it == 64 * F::N + 2) x = F::zero() - F::one();
        else if (it < 64 * F::N + 200) x = F::from_u64(it);
        if (F::geq_mod(x.v)) continue;
        const F a = x.inverse(), b = x.inverse_fermat();
        if (!(a == b) || F::geq_mod(a.v)) bad++;
        if (x.is_zero() ? !a.is_zero() : !((x * a) == F::one())) bad++;
        done++;
    }
    {   // a non-canonical representative of zero (the raw limbs of p itself) must terminate and answer 0
        F one_plain = F::zero();
        one_plain.v[0] = 1;
        F t = F::zero() - one_plain;  // raw limbs of p - 1
        t.v[0] += 1;                  // p is odd, so p - 1 is even: no carry
        if (!t.inverse().is_zero()) bad++;
    }
    printf("%s %s %d\n", name, bad ? "BAD" : "ok", done);
    return bad;
}
int main() { return check<Fq64>("fq_inverse", 4000) + check<Fr64>("fr_inverse", 4000); }
