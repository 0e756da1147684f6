// Host-side exerciser for csrc/fq_inv.cuh (chunked binary-GCD inversion): prints x and fq_inverse(x) (both in
// Montgomery form) for random and adversarial x so that tests/test_host_limbs.py can check x * inv == R mod p.
#include <cstdio>
#include <cstdlib>
#include "../../zkp_subnet_b200/csrc/fq_inv.cuh"
using namespace zkp;
static uint64_t s = 0x243F6A8885A308D3ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static void pr(const char* tag, const Fq& a) {
    printf("%s ", tag);
    for (int i = 11; i >= 0; i--) printf("%08x", a.v[i]);
    printf("\n");
}
static bool lt_p(const Fq& a) {
    for (int i = 11; i >= 0; i--) {
        if (a.v[i] < FqParams::MOD[i]) return true;
        if (a.v[i] > FqParams::MOD[i]) return false;
    }
    return false;
}
int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 300;
    for (int it = 0; it < iters; it++) {
        Fq x = Fq::zero();
        int mode = it % 10;
        if (mode < 6) {
            do { for (int i = 0; i < 12; i++) x.v[i] = (uint32_t)rnd(); x.v[11] &= 0x1fffffffu; } while (!lt_p(x));
        } else if (mode == 6) {           // single bit 2^k
            int k = (int)(rnd() % 380);
            x.v[k / 32] = 1u << (k % 32);
        } else if (mode == 7) {           // p - 2^k
            int k = (int)(rnd() % 380);
            Fq t = Fq::zero();
            t.v[k / 32] = 1u << (k % 32);
            x = Fq::zero() - t;
        } else if (mode == 8) {           // small values
            x.v[0] = 1 + (uint32_t)(rnd() % 1000);
        } else {                          // short random
            int limbs = 1 + (int)(rnd() % 11);
            for (int i = 0; i < limbs; i++) x.v[i] = (uint32_t)rnd();
            if (x.is_zero()) x.v[0] = 1;
        }
        pr("x", x);
        pr("inv", fq_inverse(x));
    }
    pr("x", Fq::zero());
    pr("inv", fq_inverse(Fq::zero()));
    return 0;
}
