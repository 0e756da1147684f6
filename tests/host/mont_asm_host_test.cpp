// Host-side exerciser for csrc/host/mont_asm.hpp (mulx/adcx/adox Montgomery products, 6 and 4 limbs) against the
// portable product of field64.hpp: random operands, 0, 1, p - 1, single-bit values, operands with all-ones limbs
// below p.  Prints "name ok|BAD|skipped count" per field for tests/test_host_limbs.py.
#include <cstdio>
#include "../../zkp_subnet_b200/csrc/host/curve.hpp"
using namespace zkp::host;
static uint64_t s = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
template <class F> void check(const char* name, int iters) {
#ifdef ZKP_HOST_MONT_ASM
    if (!have_mulx_adx()) { printf("%s skipped 0\n", name); return; }
    int bad = 0, done = 0;
    F prev = F::one();
    for (int it = 0; it < iters; it++) {
        F x, y;
        for (int i = 0; i < F::N; i++) { x.v[i] = rnd(); y.v[i] = rnd(); }
        x.v[F::N - 1] &= 0x0fffffffffffffffull; y.v[F::N - 1] &= 0x0fffffffffffffffull;
        switch (it % 11) {
            case 0: x = F::zero() - F::one(); break;                       // p - 1
            case 1: y = F::zero() - F::one(); break;
            case 2: x = F::zero() - F::one(); y = x; break;
            case 3: x = F::zero(); break;
            case 4: y = F::one(); break;
            case 5: x = F::zero(); x.v[(it / 11) % F::N] = 1ull << ((it / 7) % 60); break;
            case 6: for (int i = 0; i < F::N - 1; i++) x.v[i] = ~0ull; break;  // all-ones low limbs
            case 7: for (int i = 0; i < F::N - 1; i++) { x.v[i] = ~0ull; y.v[i] = ~0ull; } break;
            case 8: x = prev; break;                                        // chained values
            default: break;
        }
        if (F::geq_mod(x.v) || F::geq_mod(y.v)) continue;
        const F a = x * y, b = x.mul_portable(y);
        if (!(a == b) || F::geq_mod(a.v)) bad++;
        prev = a;
        done++;
    }
    printf("%s %s %d\n", name, bad ? "BAD" : "ok", done);
#else
    printf("%s skipped 0\n", name);
#endif
}
int main() {
    check<Fq64>("fq_mont_asm", 400000);
    check<Fr64>("fr_mont_asm", 400000);
    return 0;
}
